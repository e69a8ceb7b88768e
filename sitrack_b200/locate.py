"""Drop-in for `sitrack.locate` (reference: sitrack/locate.py) -- same names,
arguments and return shapes; the arithmetic runs on the GPU through
libsitrack_b200.so (batch-of-1 launches for the scalar calls, `*Batch` variants
for real work).  No CPU fallback.
"""
import numpy as np

from . import _lib, config

__all__ = ['find_ji_of_min', 'NorthStereoProj', 'IsInsideQuadrangle', 'IsInsideQuadrangleBatch', 'TheCell', 'FCC', 'NearestPoint', 'NearestPointBatch', 'FindContainingCell']
from ._lib import as_c, check, hptr


def _haversine_host(plat, plon, xlat, xlon):
    """util.py:85-103 with numpy on the host, operation for operation: used only to re-decide the handful of seeds
    whose nearest point or acceptance the device decided by less than 1e-11 relative -- numpy's
    sin/cos/asin are the reference's own, CUDA's differ from them in the last ulp."""
    to_rad = np.pi / 180.
    R = 6360.
    a1 = np.sin(0.5 * ((xlat - plat) * to_rad))
    a2 = np.sin(0.5 * ((xlon - plon) * to_rad))
    a3 = np.cos(xlat * to_rad) * np.cos(plat * to_rad)
    return 2. * R * np.arcsin(np.sqrt(a1 * a1 + a3 * a2 * a2))


def _recheck_nearest(pnt, k1, k2, latT, lonT, resKM, rd_found_km=2.5, max_itr=10):
    """NearestPoint (locate.py:222-276) for ONE seed restricted to the device's two candidates (flat indices k1 and,
    when the argmin was a near-tie, k2): first-minimum argmin in flat order like numpy's, then the reference's
    acceptance ladder.  -> (jy, jx) or (-1, -1)."""
    if k1 < 0:
        return (-1, -1)
    Ni = latT.shape[1]
    cand = sorted(k for k in (k1, k2) if k >= 0)
    lat, lon = latT.reshape(-1)[cand], lonT.reshape(-1)[cand]
    d = _haversine_host(pnt[0], pnt[1], lat, lon)
    kk = int(np.argmin(d))                                     # ties: the lower flat index, as np.argmin over the grid
    k, dmin = cand[kk], d[kk]
    jy, jx = k // Ni, k % Ni
    rfnd, igo, lfound = rd_found_km, 0, False
    while (not lfound) and igo < max_itr:                      # :250-269, no `ji_prv` box
        igo = igo + 1
        if igo == 1 and resKM is not None:
            rfnd = 0.5 * resKM[jy, jx]                         # :262
        if igo == 1:
            igo = 2                                            # :264
        lfound = bool(dmin < rfnd)                             # :266
        if igo > 1 and not lfound:
            rfnd = 1.2 * rfnd                                  # :268
    if igo == max_itr:                                         # :274 a hit on the last pass is thrown away
        return (-1, -1)
    return (jy, jx) if lfound else (-1, -1)


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


def NorthStereoProj(pphi, plam, lam0=0., phi0=90.):
    """locate.py:23-43 -- spherical (R=6300 km) stereographic projection about (phi0, lam0) -> (y, x) km.
    Not on the tracker's path (only FCC uses it, on a 5x5 neighbourhood): plain numpy host helper."""
    to_rad = 3.141592653589793 / 180.
    s0, c0 = np.sin(to_rad * phi0), np.cos(to_rad * phi0)
    sp, cp = np.sin(to_rad * pphi), np.cos(to_rad * pphi)
    cl = np.cos(to_rad * (plam - lam0))
    zk = 2. * 6300. / (1. + s0 * sp + c0 * cp * cl)
    return zk * (c0 * sp - s0 * cp * cl), zk * cp * np.sin(to_rad * (plam - lam0))


def FCC(pntGcoor, pLat, pLon, pLatC, pLonC, cellType='T', rd_found_km=10., resolkm=[],
        ji_prv=(), np_box_r=10, max_itr=5, pntID=None, iverbose=0):
    """locate.py:139-218 -- geographic-coordinate variant of the containing-cell search (unused by the
    tracker): nearest `cellType` point on the device, then TheCell on a locally projected 5x5 box.
    -> [jX,iX], vertices (4,2).  Upstream quirks kept: the vertex offsets use the UPDATED jX,iX, and an
    unfound point raises (upstream returns an unbound KVRTCS)."""
    nhb = 2
    if cellType not in ('T', 'F'):
        print('ERROR [FCC]: for now we just expect the mesh to be centered on "T" or "F" points.')
        raise SystemExit(0)
    (zlat, zlon) = pntGcoor
    (jX, iX) = NearestPoint(pntGcoor, pLat, pLon, rd_found_km=rd_found_km, resolkm=resolkm,
                            ji_prv=ji_prv, np_box_r=np_box_r, max_itr=max_itr)
    if jX < 0 or iX < 0:
        raise UnboundLocalError("FCC: no nearest point found (upstream fails on an unbound `KVRTCS` here)")
    zYC, zXC = NorthStereoProj(pLatC[jX - nhb:jX + 3, iX - nhb:iX + 3], pLonC[jX - nhb:jX + 3, iX - nhb:iX + 3],
                               lam0=pLon[jX, iX], phi0=pLat[jX, iX])
    zy0, zx0 = NorthStereoProj(np.array([zlat]), np.array([zlon]), lam0=pLon[jX, iX], phi0=pLat[jX, iX])
    lPin, [jM, iM], KVRTCS = TheCell((float(zy0[0]), float(zx0[0])), (nhb, nhb), zYC, zXC)
    jX = jM + jX - nhb
    iX = iM + iX - nhb
    KVRTCS[:, 0] = KVRTCS[:, 0] + jX - nhb
    KVRTCS[:, 1] = KVRTCS[:, 1] + iX - nhb
    if not lPin:
        print('WARNING [SeedInit()]: could not find the proper F-point cell!!!')
        print('         => when lookin for point:', zlat, zlon)
    return [jX, iX], KVRTCS


def NearestPointBatch(latlon, pLat, pLon, rd_found_km=10., resolkm=None, max_itr=5, brute=False):
    """(n,2) [lat,lon] -> (ji (n,2) with -1,-1 when not found, dist_km (n,))."""
    from .engine import TrackEngine
    Ny, Nx = pLat.shape
    dummy = np.zeros((Ny, Nx))
    with TrackEngine(dummy, dummy, tmask=np.ones((Ny, Nx), np.int8), uv_strategy=0, device=config.device) as eng:
        eng.set_locate_grid(pLat, pLon, resolkm)
        return eng.nearest_point(latlon, rd_found_km=rd_found_km, max_itr=max_itr, brute=brute)


def _nearest_point_box(pntGcoor, pLat, pLon, rd_found_km, resolkm, ji_prv, np_box_r, max_itr):
    """NearestPoint with a previous guess (locate.py:241-244,255-256): pass 1 looks at the box of half-width
    `np_box_r` around `ji_prv` only -- a few hundred points, evaluated here with numpy, the reference's own
    arithmetic -- and passes 2.. fall back on the whole domain (device search).  Upstream quirks kept: the first
    radius is 0.5*resolkm[jy,jx] with (jy,jx) the BOX-LOCAL indices of the box's argmin (:262), it is not grown
    after pass 1 (:268 needs igo > 1), and a hit on pass `max_itr` is thrown away (:274)."""
    (Ny, Nx) = pLat.shape
    l2Dresol = np.shape(resolkm) == (Ny, Nx)
    (latP, lonP) = pntGcoor
    (j_prv, i_prv) = ji_prv
    j1, j2 = max(j_prv - np_box_r, 0), min(j_prv + np_box_r + 1, Ny)
    i1, i2 = max(i_prv - np_box_r, 0), min(i_prv + np_box_r + 1, Nx)
    xd = _haversine_host(latP, lonP, np.asarray(pLat)[j1:j2, i1:i2], np.asarray(pLon)[j1:j2, i1:i2])
    jy, jx = find_ji_of_min(xd)
    rfnd = 0.5 * np.asarray(resolkm)[jy, jx] if l2Dresol else rd_found_km
    if max_itr > 1 and xd[jy, jx] < rfnd:
        return (int(jy + j1), int(jx + i1))
    if max_itr <= 2:                                           # pass 2 would be the last one: its hit is discarded
        print('    WARNING [NearestPoint()]: did not find a nearest point for target point ', latP, lonP, ' !')
        return (-1, -1)
    # whole domain, radii rfnd * 1.2^(igo-2) for igo = 2 .. max_itr-1: the no-box ladder started at rfnd
    ji, d = NearestPointBatch([pntGcoor], pLat, pLon, float(rfnd), None, max_itr)
    jy, jx = int(ji[0, 0]), int(ji[0, 1])
    if jy < 0:
        print('    WARNING [NearestPoint()]: did not find a nearest point for target point ', latP, lonP, ' !')
    return (jy, jx)


def NearestPoint(pntGcoor, pLat, pLon, rd_found_km=10., resolkm=[], ji_prv=(), np_box_r=10, max_itr=5):
    """locate.py:222-276 -> (jy,jx) of the nearest grid point, or (-1,-1)."""
    if np.shape(pLon) != np.shape(pLat):
        print('ERROR [NearestPoint]: `pLat` & `pLon` do not have the same shape!')
        raise SystemExit(0)
    if len(ji_prv) == 2:
        return _nearest_point_box(pntGcoor, pLat, pLon, rd_found_km, resolkm, ji_prv, np_box_r, max_itr)
    res = resolkm if np.shape(resolkm) == np.shape(pLat) else None
    ji, d = NearestPointBatch([pntGcoor], pLat, pLon, rd_found_km, res, max_itr)
    jy, jx = int(ji[0, 0]), int(ji[0, 1])
    if jy < 0:
        print('    WARNING [NearestPoint()]: did not find a nearest point for target point ',
              pntGcoor[0], pntGcoor[1], ' !')
    return (jy, jx)
