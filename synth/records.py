"""Synthetic hourly SI3 records: u_ice, v_ice [m/s] and siconc, all float32 like
the netCDF variables the reference reads (si3_part_tracker.py:372-374)."""
import numpy as np

T0_EPOCH = 850608000            # 1996-12-15 00:00:00 UTC


def time_counter(nrec, rdt=3600):
    """Centre-of-interval model times, ncio.ModelFileTimeInfo style (i4)."""
    return (T0_EPOCH + rdt // 2 + rdt * np.arange(nrec)).astype(np.int32)


def make_records(grid, nrec, seed=1, k0=0, noise=0.03, out=None):
    """-> (U, V, IC) each (nrec, Nj, Ni) float32; record k is hour k0+k.

    u/v: basin gyre (~0.1 m/s) + travelling synoptic wave (~0.1 m/s, 3-day
    period) + white noise, clipped to |.|<=1 and zeroed on land faces.
    siconc: disc of pack ice whose edge migrates over a 60-day cycle, so some
    buoys die from low concentration (reference tracking.py:84-91)."""
    Nj, Ni = grid["Nj"], grid["Ni"]
    tm = grid["tmask"].astype(bool)
    rng = np.random.default_rng(seed)
    yc, xc = float(np.mean(grid["Yt"])), float(np.mean(grid["Xt"]))
    Lr = 0.5 * min(Nj, Ni) * grid["dx_km"]

    def gyre(Y, X):
        dy, dx = Y - yc, X - xc
        r2 = (dy * dy + dx * dx) / (0.55 * Lr) ** 2
        om = 0.12 / (0.45 * Lr) * np.exp(-r2)
        return (-om * dy).astype(np.float32), (om * dx).astype(np.float32)

    ug, _ = gyre(grid["Yu"], grid["Xu"])
    _, vg = gyre(grid["Yv"], grid["Xv"])
    kw = 2.0 * np.pi / 800.0
    phu = (kw * (0.8 * grid["Xu"] + 0.6 * grid["Yu"])).astype(np.float32)
    phv = (kw * (0.6 * grid["Xv"] - 0.8 * grid["Yv"])).astype(np.float32)
    rT = np.hypot(grid["Yt"] - yc, grid["Xt"] - xc).astype(np.float32)
    landU = ~(tm & np.roll(tm, -1, axis=1))
    landV = ~(tm & np.roll(tm, -1, axis=0))

    if out is None:
        U = np.empty((nrec, Nj, Ni), np.float32)
        V = np.empty((nrec, Nj, Ni), np.float32)
        IC = np.empty((nrec, Nj, Ni), np.float32)
    else:
        U, V, IC = out
    for k in range(nrec):
        t = float(k0 + k)
        w = np.float32(2.0 * np.pi * t / 72.0)
        u = ug + np.float32(0.1) * np.sin(phu - w) + np.float32(noise) * rng.standard_normal((Nj, Ni), np.float32)
        v = vg + np.float32(0.1) * np.cos(phv - w) + np.float32(noise) * rng.standard_normal((Nj, Ni), np.float32)
        np.clip(u, -1.0, 1.0, out=u)
        np.clip(v, -1.0, 1.0, out=v)
        u[landU] = 0.0
        v[landV] = 0.0
        redge = np.float32(Lr * (0.62 + 0.06 * np.sin(2.0 * np.pi * t / 1440.0)))
        ic = np.clip(np.float32(0.5) + (redge - rT) / np.float32(0.08 * Lr), 0.0, 1.0).astype(np.float32)
        ic[~tm] = 0.0
        U[k], V[k], IC[k] = u, v, ic
    return U, V, IC
