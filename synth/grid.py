"""Synthetic curvilinear Arakawa C-grid in the NorthPolarStereo(-45,70) km plane.

T at (j,i), U at (j,i+1/2), V at (j+1/2,i), F at (j+1/2,i+1/2) -- the NEMO
index convention sitrack relies on (host cell of T[j,i] = quad F[j-1,i-1],
F[j-1,i], F[j,i], F[j,i-1]; reference sitrack/locate.py:320-321).
"""
import numpy as np

GRID_PRESETS = {
    # NANUK4-shaped 0.25 deg grid (566 x 492, ~12.5 km)
    "nanuk4": dict(Nj=566, Ni=492, dx_km=12.5),
    # 1/12 deg class grid (1700 x 1475, ~4.2 km)
    "arctic12": dict(Nj=1700, Ni=1475, dx_km=4.2),
    # small grids for unit tests
    "small": dict(Nj=120, Ni=100, dx_km=12.5),
    "tiny": dict(Nj=48, Ni=40, dx_km=12.5),
}

# WGS84 / PROJ stere north-pole constants (cartopy's default globe)
_A = 6378137.0
_F = 1.0 / 298.257223563
_E2 = _F * (2.0 - _F)
_E = np.sqrt(_E2)


def _tsfn(phi):
    es = _E * np.sin(phi)
    return np.tan(0.5 * (0.5 * np.pi - phi)) / ((1.0 - es) / (1.0 + es)) ** (0.5 * _E)


def _akm1(lat_ts=70.0):
    p = np.deg2rad(lat_ts)
    return np.cos(p) / _tsfn(p) / np.sqrt(1.0 - _E2 * np.sin(p) ** 2)


def km_to_latlon(y_km, x_km, lat_ts=70.0, lon0=-45.0):
    """Inverse ellipsoidal polar stereographic, iterated to machine precision."""
    x = 1000.0 * np.asarray(x_km, np.float64) / _A
    y = 1000.0 * np.asarray(y_km, np.float64) / _A
    tp = np.hypot(x, y) / _akm1(lat_ts)
    phi = 0.5 * np.pi - 2.0 * np.arctan(tp)
    for _ in range(12):
        es = _E * np.sin(phi)
        phi = 0.5 * np.pi - 2.0 * np.arctan(tp * ((1.0 - es) / (1.0 + es)) ** (0.5 * _E))
    lam = np.arctan2(x, -y) + np.deg2rad(lon0)
    lam = (lam + np.pi) % (2.0 * np.pi) - np.pi
    return np.rad2deg(phi), np.rad2deg(lam)


def latlon_to_km(lat, lon, lat_ts=70.0, lon0=-45.0):
    phi = np.deg2rad(np.asarray(lat, np.float64))
    lam = np.deg2rad(np.asarray(lon, np.float64) - lon0)
    rho = _A * _akm1(lat_ts) * _tsfn(phi)
    return -rho * np.cos(lam) / 1000.0, rho * np.sin(lam) / 1000.0


class Warp:
    """Smooth, non-affine map from fractional grid indices (j,i) to plane km."""

    def __init__(self, Nj, Ni, dx_km, seed):
        rng = np.random.default_rng(seed)
        self.Nj, self.Ni, self.dx = Nj, Ni, dx_km
        self.Ly, self.Lx = Nj * dx_km, Ni * dx_km
        self.cy, self.cx = 0.03 * self.Ly + 41.7, -0.04 * self.Lx - 57.3   # pole is off-node
        self.rot = np.deg2rad(17.0 + 6.0 * rng.random())
        self.amp = 0.028 * min(self.Ly, self.Lx) * (0.8 + 0.4 * rng.random(4))
        self.ph = 2.0 * np.pi * rng.random(4)

    def __call__(self, jj, ii):
        jj = np.asarray(jj, np.float64)
        ii = np.asarray(ii, np.float64)
        x0 = (ii - 0.5 * (self.Ni - 1)) * self.dx
        y0 = (jj - 0.5 * (self.Nj - 1)) * self.dx
        a, p = self.amp, self.ph
        x1 = x0 + a[0] * np.sin(2 * np.pi * y0 / (0.8 * self.Ly) + p[0]) \
                + a[1] * np.sin(2 * np.pi * x0 / (0.6 * self.Lx) + p[1])
        y1 = y0 + a[2] * np.sin(2 * np.pi * x0 / (0.7 * self.Lx) + p[2]) \
                + a[3] * np.cos(2 * np.pi * y0 / (0.5 * self.Ly) + p[3])
        c, s = np.cos(self.rot), np.sin(self.rot)
        return self.cy + s * x1 + c * y1, self.cx + c * x1 - s * y1   # (y, x)


def make_grid(Nj=566, Ni=492, dx_km=12.5, seed=0, with_latlon=True):
    """-> dict: Yt,Xt,Yf,Xf,Yu,Xu,Yv,Xv (f8 km), latT,lonT (f8 deg, lon in [0,360)),
    tmask (i1), ResKM (f8 km), warp (callable), Nj, Ni."""
    w = Warp(Nj, Ni, dx_km, seed)
    jj, ii = np.meshgrid(np.arange(Nj, dtype=np.float64), np.arange(Ni, dtype=np.float64), indexing="ij")
    g = dict(Nj=Nj, Ni=Ni, dx_km=dx_km, warp=w, seed=seed)
    g["Yt"], g["Xt"] = w(jj, ii)
    g["Yu"], g["Xu"] = w(jj, ii + 0.5)
    g["Yv"], g["Xv"] = w(jj + 0.5, ii)
    g["Yf"], g["Xf"] = w(jj + 0.5, ii + 0.5)
    if with_latlon:
        lat, lon = km_to_latlon(g["Yt"], g["Xt"])
        g["latT"], g["lonT"] = lat, np.mod(lon, 360.0)      # ncio.py:48 mod 360

    # metrics e1t/e2t [m] from the neighbouring faces (ncio.py:34-35,56-57)
    e1 = np.hypot(np.diff(g["Yu"], axis=1, prepend=np.nan), np.diff(g["Xu"], axis=1, prepend=np.nan))
    e1[:, 0] = e1[:, 1]
    e2 = np.hypot(np.diff(g["Yv"], axis=0, prepend=np.nan), np.diff(g["Xv"], axis=0, prepend=np.nan))
    e2[0, :] = e2[1, :]
    g["ResKM"] = np.sqrt(e1 * e1 + e2 * e2)

    # land-sea mask: continental rim, a circular basin edge, a few islands and a peninsula
    rng = np.random.default_rng(seed + 1000)
    tm = np.ones((Nj, Ni), np.int8)
    rim = max(4, int(0.012 * min(Nj, Ni)))
    tm[:rim, :] = 0; tm[-rim:, :] = 0; tm[:, :rim] = 0; tm[:, -rim:] = 0
    rj, ri = (jj - 0.5 * Nj) / (0.5 * Nj), (ii - 0.5 * Ni) / (0.5 * Ni)
    tm[(rj * rj + ri * ri) > 0.93 ** 2 * (1.0 + 0.08 * np.sin(5 * np.arctan2(rj, ri))) ** 2] = 0
    for _ in range(7):
        cj, ci = rng.uniform(0.2, 0.8) * Nj, rng.uniform(0.2, 0.8) * Ni
        rad = rng.uniform(0.012, 0.04) * min(Nj, Ni)
        tm[(jj - cj) ** 2 + (ii - ci) ** 2 < rad * rad] = 0
    pj = int(0.62 * Nj)
    tm[pj:pj + max(2, Nj // 60), : int(0.33 * Ni)] = 0              # peninsula
    g["tmask"] = tm
    return g
