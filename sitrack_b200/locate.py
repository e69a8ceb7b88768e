"""Drop-in for `sitrack.locate` (reference: sitrack/locate.py) -- same names,
arguments and return shapes; the arithmetic runs on the GPU through
libsitrack_b200.so (batch-of-1 launches for the scalar calls, `*Batch` variants
for real work).  No CPU fallback.
"""
import numpy as np

from . import _lib, config
from ._lib import as_c, check, hptr


def find_ji_of_min(x):
    """locate.py:13-20 -- (j,i) of the first minimum of a 2-D array (host index arithmetic)."""
    k = int(np.argmin(x))
    return k // x.shape[1], k % x.shape[1]


def IsInsideQuadrangleBatch(yx, quads):
    """yx (n,2) [y,x]; quads (n,4,2) -> bool (n,).  locate.py:49-78 per element."""
    yx = as_c(yx, np.float64).reshape(-1, 2)
    q = as_c(quads, np.float64).reshape(-1, 4, 2)
    out = np.zeros(yx.shape[0], np.int8)
    check(_lib.lib().st_inside_quad(config.device, yx.shape[0], hptr(yx), hptr(q), hptr(out)))
    return out.astype(bool)


def IsInsideQuadrangle(y, x, quad):
    """locate.py:49: is (y,x) inside quad [[y0,x0],..,[y3,x3]]? (half-open ray-casting rules kept)."""
    if len(quad) != 4:
        print('ERROR: `len(quad) !=: 4`')
        raise SystemExit(0)
    return bool(IsInsideQuadrangleBatch([[y, x]], [np.asarray(quad, np.float64)])[0])


def _engine_for_F(pYf, pXf):
    from .engine import TrackEngine
    tm = np.ones(np.shape(pYf), np.int8)
    return TrackEngine(pYf, pXf, tmask=tm, uv_strategy=0, device=config.device)


def _vertices(jT, iT):
    return [[jT - 1, jT - 1, jT, jT], [iT - 1, iT, iT, iT - 1]]


def FindContainingCell(pyx, kjiT, pYf, pXf, iverbose=0):
    """locate.py:280-330 -> (lPin, [jT,iT], [[4 j],[4 i]]) around the nearest T-point kjiT."""
    with _engine_for_F(pYf, pXf) as eng:
        cell, found = eng.find_containing_cell([pyx], [kjiT])
    jT, iT = int(cell[0, 0]), int(cell[0, 1])
    return bool(found[0]), [jT, iT], _vertices(jT, iT)


def TheCell(pyx, kjiT, pYf, pXf, iverbose=0):
    """locate.py:81-133 -- as FindContainingCell but the vertices come back as a (4,2) array."""
    lPin, ji, v = FindContainingCell(pyx, kjiT, pYf, pXf, iverbose=iverbose)
    return lPin, ji, np.array(v, dtype=int).T.copy()


def NearestPointBatch(latlon, pLat, pLon, rd_found_km=10., resolkm=None, max_itr=5, brute=False):
    """(n,2) [lat,lon] -> (ji (n,2) with -1,-1 when not found, dist_km (n,))."""
    from .engine import TrackEngine
    Ny, Nx = pLat.shape
    dummy = np.zeros((Ny, Nx))
    with TrackEngine(dummy, dummy, tmask=np.ones((Ny, Nx), np.int8), uv_strategy=0, device=config.device) as eng:
        eng.set_locate_grid(pLat, pLon, resolkm)
        return eng.nearest_point(latlon, rd_found_km=rd_found_km, max_itr=max_itr, brute=brute)


def NearestPoint(pntGcoor, pLat, pLon, rd_found_km=10., resolkm=[], ji_prv=(), np_box_r=10, max_itr=5):
    """locate.py:222-276 -> (jy,jx) of the nearest grid point, or (-1,-1)."""
    if np.shape(pLon) != np.shape(pLat):
        print('ERROR [NearestPoint]: `pLat` & `pLon` do not have the same shape!')
        raise SystemExit(0)
    if len(ji_prv) == 2:
        raise NotImplementedError("NearestPoint: the `ji_prv` local-box variant is not used by the tracker "
                                  "and is not provided by sitrack_b200")
    res = resolkm if np.shape(resolkm) == np.shape(pLat) else None
    ji, d = NearestPointBatch([pntGcoor], pLat, pLon, rd_found_km, res, max_itr)
    jy, jx = int(ji[0, 0]), int(ji[0, 1])
    if jy < 0:
        print('    WARNING [NearestPoint()]: did not find a nearest point for target point ',
              pntGcoor[0], pntGcoor[1], ' !')
    return (jy, jx)
