"""Drop-in for the part of `sitrack.util` the tracker uses (reference: sitrack/util.py).

Haversine and the NorthPolarStereo(-45,70) conversions run on the GPU; the
reference reaches the latter through cartopy -> PROJ, here they are kernels of
libsitrack_b200.so (tolerance vs PROJ: 1e-9 degrees / 1e-6 km).
Time-string helpers are host Python.
"""
import numpy as np

from . import _lib, config

__all__ = ['chck4f', 'epoch2clock', 'clock2epoch', 'degE_to_degWE', 'Haversine', 'CartNPSkm2Geo1D', 'Geo2CartNPSkm1D', 'ConvertGeo2CartesianNPSkm', 'ConvertCartesianNPSkm2Geo', 'StdDev']
from ._lib import as_c, check, hptr


def chck4f(cf):
    from os.path import exists
    if not exists(cf):
        print(' ERROR [chck4f()]: file ' + cf + ' does not exist!')
        raise SystemExit(0)


_FMT = {'s': "%Y-%m-%d_%H:%M:%S", 'm': "%Y-%m-%d_%H:%M", 'h': "%Y-%m-%d_%H", 'D': "%Y-%m-%d"}


def epoch2clock(it, precision='s'):
    """util.py:18-33"""
    from datetime import datetime, timezone
    if precision not in _FMT:
        print('ERROR [epoch2clock]: unknown precision "' + precision + '" !')
        raise SystemExit(0)
    return str(datetime.fromtimestamp(int(it), timezone.utc).strftime(_FMT[precision]))


def clock2epoch(cdate, precision='s', cfrmt='advanced'):
    """util.py:35-40; `precision='D'` / cfrmt='guess' follow the un-vendored mojito helper the
    CLI uses for -e (si3_part_tracker.py:163-166): YYYY-MM-DD or YYYYMMDD at midnight UTC."""
    from datetime import datetime, timezone
    if precision == 'D':
        s = cdate.replace('-', '')[:8]
        it = datetime.strptime(s, "%Y%m%d")
    else:
        it = datetime.strptime(cdate, "%Y-%m-%d_%H:%M:%S")
    return int(it.replace(tzinfo=timezone.utc).timestamp())


def degE_to_degWE(X):
    """util.py:43-53: longitude 0..360 -> -180..180"""
    X = np.asarray(X, dtype=float)
    r = np.copysign(1., 180. - X) * np.minimum(X, np.abs(X - 360.))
    return float(r) if r.shape == () else r


def Haversine(plat, plon, xlat, xlon):
    """util.py:85-103: distance [km] (R=6360) from one point to every point of xlat/xlon."""
    la, lo = as_c(xlat, np.float64), as_c(xlon, np.float64)
    out = np.empty(la.shape, np.float64)
    check(_lib.lib().st_haversine(config.device, la.size, float(plat), float(plon), hptr(la), hptr(lo), hptr(out)))
    return out


def _xy2ll(yx, lat0, lon0):
    yx = as_c(yx, np.float64).reshape(-1, 2)
    out = np.empty_like(yx)
    check(_lib.lib().st_xy2latlon(config.device, yx.shape[0], hptr(yx), hptr(out), lat0, lon0))
    return out


def _ll2xy(ll, lat0, lon0):
    ll = as_c(ll, np.float64).reshape(-1, 2)
    out = np.empty_like(ll)
    check(_lib.lib().st_latlon2xy(config.device, ll.shape[0], hptr(ll), hptr(out), lat0, lon0))
    return out


def CartNPSkm2Geo1D(pcoorC, lat0=70., lon0=-45.):
    """util.py:413-429: (n,2) [y,x] km -> (n,2) [lat,lon] degrees."""
    if np.shape(pcoorC)[1] != 2:
        print(' ERROR [CartNPSkm2Geo1D()]: input array `pcoorC` has a wrong a shape!')
        raise SystemExit(0)
    return _xy2ll(pcoorC, lat0, lon0)


def Geo2CartNPSkm1D(pcoorG, lat0=70., lon0=-45.):
    """util.py:394-410: (n,2) [lat,lon] degrees -> (n,2) [y,x] km."""
    if np.shape(pcoorG)[1] != 2:
        print(' ERROR [Geo2CartNPSkm1D()]: input array `pcoorG` has a wrong a shape!')
        raise SystemExit(0)
    return _ll2xy(pcoorG, lat0, lon0)


def ConvertGeo2CartesianNPSkm(plat, plon, lat0=70., lon0=-45.):
    """util.py:434-451: arrays of lat, lon (any shape) -> (Y, X) km of the same shape."""
    shp = np.shape(plat)
    ll = np.stack([np.ravel(plat), np.ravel(plon)], axis=1)
    yx = _ll2xy(ll, lat0, lon0)
    return yx[:, 0].reshape(shp), yx[:, 1].reshape(shp)


def ConvertCartesianNPSkm2Geo(pY, pX, lat0=70., lon0=-45.):
    """util.py:455-472: arrays of Y, X km (any shape) -> (lat, lon) degrees of the same shape."""
    shp = np.shape(pY)
    yx = np.stack([np.ravel(pY), np.ravel(pX)], axis=1)
    ll = _xy2ll(yx, lat0, lon0)
    return ll[:, 0].reshape(shp), ll[:, 1].reshape(shp)


def StdDev(pmean, pX):
    zz = np.asarray(pX) - pmean
    return np.sqrt(np.mean(zz * zz))
