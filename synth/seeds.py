"""Synthetic seedings: HSS-n (nemoSeed semantics), scattered (SIDFEX-style) and
dense jittered clouds.  Positions pass through float32 like the seed netCDF
does (reference ncio.py:156-159 writes f4, :294-297 reads them back into f8)."""
import numpy as np

from .grid import km_to_latlon


def _through_f4(a):
    return np.asarray(a, np.float32).astype(np.float64)


def _pack(lat, lon, y, x, f4=True):
    SG = np.stack([lat, np.mod(lon, 360.0)], axis=1)          # ncio.py:286 lon mod 360
    SC = np.stack([y, x], axis=1)
    if f4:
        SG = np.stack([_through_f4(lat), np.mod(_through_f4(lon), 360.0)], axis=1)
        SC = _through_f4(SC)
    ids = np.arange(SG.shape[0], dtype=np.int64) + 1
    return ids, np.ascontiguousarray(SG), np.ascontiguousarray(SC)


def hss_seeds(grid, ic0, khss=5):
    """Every khss-th ocean T-point north of 55N with siconc>=0.9
    (reference tracking.py:375-423).  -> IDs, SG (nP,2) lat,lon, SC (nP,2) y,x km."""
    sl = (slice(None, None, khss), slice(None, None, khss))
    m = grid["tmask"][sl].astype(bool) & (grid["latT"][sl] >= 55.0) & (ic0[sl] >= 0.9)
    jj, ii = np.where(m)
    lat, lon = grid["latT"][sl][jj, ii], grid["lonT"][sl][jj, ii]
    y, x = grid["Yt"][sl][jj, ii], grid["Xt"][sl][jj, ii]
    lon = np.where(lon > 180.0, lon - 360.0, lon)            # seed files hold (-180,180]
    return _pack(lat, lon, y, x)


def scattered_seeds(grid, n, seed=2):
    """n uniformly random points over (and a little beyond) the domain's bounding
    box: some land on continents / outside the grid and must be dropped."""
    rng = np.random.default_rng(seed)
    y0, y1 = grid["Yt"].min(), grid["Yt"].max()
    x0, x1 = grid["Xt"].min(), grid["Xt"].max()
    py, px = 0.04 * (y1 - y0), 0.04 * (x1 - x0)
    y = rng.uniform(y0 - py, y1 + py, n)
    x = rng.uniform(x0 - px, x1 + px, n)
    lat, lon = km_to_latlon(y, x)
    return _pack(lat, lon, y, x)


def dense_seeds(grid, n, ic0, seed=3, jitter=0.42, f4=False, with_latlon=True, box=None, cells_out=None):
    """n buoys over the pack: random ocean T-cells with siconc>=0.9, jittered
    inside the cell, returned sorted by (j,i) (HSS1-with-replicas style)."""
    rng = np.random.default_rng(seed)
    Nj, Ni = grid["Nj"], grid["Ni"]
    m = grid["tmask"].astype(bool) & (ic0 >= 0.9)
    m[:3, :] = False; m[-3:, :] = False; m[:, :3] = False; m[:, -3:] = False
    if box:                      # diagnostic: confine the cloud to a centred box of `box` x `box` cells
        keep = np.zeros_like(m); j0, i0 = Nj // 2 - box // 2, Ni // 2 - box // 2
        keep[j0:j0 + box, i0:i0 + box] = True; m &= keep
    cells = np.flatnonzero(m)
    pick = np.sort(rng.choice(cells, size=n, replace=n > cells.size))
    jj, ii = np.divmod(pick, Ni)
    if cells_out is not None:    # the T-cell each buoy was drawn in (CPU arm of bench.py: skips a nearest-point search)
        cells_out["seed_cells"] = np.stack([jj, ii], axis=1)
    y, x = grid["warp"](jj + rng.uniform(-jitter, jitter, n), ii + rng.uniform(-jitter, jitter, n))
    if not with_latlon:          # caller derives lat/lon itself (e.g. on the device)
        return np.arange(n, dtype=np.int64) + 1, None, np.ascontiguousarray(np.stack([y, x], axis=1))
    lat, lon = km_to_latlon(y, x)
    return _pack(lat, lon, y, x, f4=f4)
