"""The certified fast path of step variant 4 (sitrack_b200/csrc/st_cert.cuh: k_advect_cert + k_walk).

k_advect_cert decides the U/V pick (si3_part_tracker.py:430-441) and "the buoy is still inside its cell"
(sitrack/locate.py:49-78) from a 32-byte per-cell frame in f32 whenever the buoy is farther than per-cell
margins from every line involved; everything else runs the reference's own tests in k_walk.  These tests attack the
claim "certified => equal to the reference's decision" directly: st_selftest_cert evaluates, for arbitrary
(position, cell, velocities), the certified decision next to the exact one (the same device predicates that
tests/test_gpu_parity.py pins against the reference's golden vectors), with positions placed ON the margins.
The bit-exact trajectory tests of tests/test_gpu_parity.py run variant 4 next to the default kernel, and
test_cert_step_vs_default_at_size compares the two on a cloud large enough to mark tens of thousands of lanes."""
import numpy as np
import pytest

from conftest import engine_for

pytestmark = pytest.mark.gpu

PICK, STAY, CW, CS, XW, XS, XIN = 1, 2, 4, 8, 16, 32, 64


def _grid(name, seed=0):
    import synth
    return synth.make_grid(**synth.GRID_PRESETS[name], seed=seed, with_latlon=False)


def _points_in_frame(fr, cells, s, t):
    """km positions whose frame coordinates in their cell are (s,t): inverse of s = a dx + b dy + es, ..."""
    f = fr[cells[:, 0], cells[:, 1]].astype(np.float64)
    oy, ox, a, b, c, d, es, et = f.T
    det = a * d - b * c
    s0, t0 = s - es, t - et
    dx = (d * s0 - b * t0) / det
    dy = (-c * s0 + a * t0) / det
    return np.stack([oy + dy, ox + dx], axis=1)


def _check(flags):
    pick = (flags & PICK) != 0
    stay = (flags & STAY) != 0
    west_c, south_c = (flags & CW) != 0, (flags & CS) != 0
    west_x, south_x = (flags & XW) != 0, (flags & XS) != 0
    inside_x = (flags & XIN) != 0
    bad_pick = pick & ((west_c != west_x) | (south_c != south_x))
    bad_stay = pick & stay & ~inside_x
    return pick, stay, int(bad_pick.sum()), int(bad_stay.sum())


@pytest.mark.parametrize("name", ["tiny", "small", "nanuk4", "arctic12"])
def test_cert_decisions_equal_the_reference(torch, name):
    g = _grid(name)
    Nj, Ni = g["tmask"].shape
    rng = np.random.default_rng(5)
    with engine_for(g) as eng:
        adm, exa = eng.cert_stats()
        assert exa == (Nj - 1) * (Ni - 1) and adm >= 0.99 * exa          # smooth synthetic grid: every cell qualifies
        fr, hin, msep = eng.cert_frames()
        n = 400_000
        cells = np.stack([rng.integers(1, Nj, n), rng.integers(1, Ni, n)], axis=1).astype(np.int32)
        h = hin[cells[:, 0], cells[:, 1]].astype(np.float64)
        m = msep[cells[:, 0], cells[:, 1]].astype(np.float64)
        ok = h > 0
        cells, h, m = cells[ok], h[ok], m[ok]
        n = cells.shape[0]
        # frame coordinates: a quarter uniform over and around the cell, the rest ON the margins (relative
        # offsets down to 1e-7, both sides), sign and which coordinate chosen at random
        s = rng.uniform(-0.75, 0.75, n); t = rng.uniform(-0.75, 0.75, n)
        k = rng.integers(0, 8, n)
        off = 1.0 + rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-7, -1, n)
        sg = rng.choice([-1.0, 1.0], n)
        s = np.where(k == 2, sg * m * off, s); t = np.where(k == 3, sg * m * off, t)
        s = np.where(k == 4, sg * h * off, s); t = np.where(k == 5, sg * h * off, t)
        both = k == 6
        s = np.where(both, sg * m * off, s); t = np.where(both, rng.choice([-1.0, 1.0], n) * m * off[::-1], t)
        corner = k == 7
        s = np.where(corner, sg * h * off, s); t = np.where(corner, rng.choice([-1.0, 1.0], n) * h * off[::-1], t)
        yx = _points_in_frame(fr, cells, s, t)
        # velocities [m/s]: mostly realistic, some large (multi-cell jumps), a few so that the NEW position sits
        # on the stay margin, a few non-finite
        vel = rng.uniform(-0.4, 0.4, (n, 4)).astype(np.float32)
        big = rng.random(n) < 0.15
        vel[big] *= np.float32(12.0)
        vel[rng.random(n) < 0.002] = np.float32(np.nan)
        vel[rng.random(n) < 0.002] = np.float32(np.inf)
        vel[rng.random(n) < 0.002] = np.float32(1e20)
        flags = eng.selftest_cert(yx, cells, vel)
        pick, stay, bad_pick, bad_stay = _check(flags)
        assert bad_pick == 0 and bad_stay == 0
        assert pick.mean() > 0.15 and (pick & stay).mean() > 0.05           # the test does exercise the certified path
        # realistic cloud: uniform in the cell, |u| <= 0.3 m/s: most lanes take the fast path on the fine grids
        s = rng.uniform(-0.5, 0.5, n); t = rng.uniform(-0.5, 0.5, n)
        yx = _points_in_frame(fr, cells, s, t)
        vel = rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)
        flags = eng.selftest_cert(yx, cells, vel)
        pick, stay, bad_pick, bad_stay = _check(flags)
        assert bad_pick == 0 and bad_stay == 0
        inside_x = (flags & XIN) != 0
        frac_pick, frac_all = pick.mean(), (pick & stay).mean()
        print("%s: pick certified %.4f, pick+stay certified %.4f, reference says inside %.4f; median msep %.2e, 0.5-hin %.2e"
              % (name, frac_pick, frac_all, inside_x.mean(), np.median(m), np.median(0.5 - h)))
        if name in ("nanuk4", "arctic12"):
            assert frac_pick > 0.97 and frac_all > 0.9 * inside_x.mean()


def test_cert_stay_margin_from_the_new_position(torch):
    """Positions well inside, velocities chosen so that the NEW position lands on the stay margin or on the
    cell edge itself: certified stay must imply the reference's inside test."""
    g = _grid("nanuk4")
    Nj, Ni = g["tmask"].shape
    rng = np.random.default_rng(6)
    with engine_for(g) as eng:
        fr, hin, msep = eng.cert_frames()
        n = 300_000
        cells = np.stack([rng.integers(2, Nj - 1, n), rng.integers(2, Ni - 1, n)], axis=1).astype(np.int32)
        h = hin[cells[:, 0], cells[:, 1]].astype(np.float64)
        s0 = rng.uniform(0.05, 0.3, n) * rng.choice([-1, 1], n); t0 = rng.uniform(0.05, 0.3, n) * rng.choice([-1, 1], n)
        P = _points_in_frame(fr, cells, s0, t0)
        # target: |s'| = h (1 +- tiny) or 0.5 (1 +- tiny)
        tgt = np.where(rng.random(n) < 0.5, h, 0.5) * (1.0 + rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-7, -2, n))
        s1 = np.where(rng.random(n) < 0.5, np.sign(s0) * tgt, rng.uniform(-0.4, 0.4, n))
        t1 = np.where(s1 == np.sign(s0) * tgt, rng.uniform(-0.4, 0.4, n), np.sign(t0) * tgt)
        Q = _points_in_frame(fr, cells, s1, t1)
        d = (Q - P) * 1000.0 / 3600.0                                   # m/s that carries P to Q in one record
        zv, zu = d[:, 0].astype(np.float32), d[:, 1].astype(np.float32)
        vel = np.stack([zu, zu, zv, zv], axis=1)
        flags = eng.selftest_cert(P, cells, vel)
        pick, stay, bad_pick, bad_stay = _check(flags)
        assert bad_pick == 0 and bad_stay == 0
        assert pick.mean() > 0.95 and 0.2 < (pick & stay).mean() < 0.8


def test_cert_rejects_bad_cells(torch, gold_track):
    """Folded / clockwise cells and cells whose U/V points are far from the face centres get no frame
    (hin = -1) and go through the reference's tests; a grid beyond 2^17 km gets no frames at all."""
    T, g0 = gold_track
    g = {k: np.array(v, copy=True) for k, v in g0.items()}
    rng = np.random.default_rng(11)
    Nj, Ni = g["tmask"].shape
    jj, ii = rng.integers(3, Nj - 3, 40), rng.integers(3, Ni - 3, 40)
    g["Yf"][jj, ii] += rng.uniform(-1.2, 1.2, 40) * 12.5
    g["Xf"][jj, ii] += rng.uniform(-1.2, 1.2, 40) * 12.5
    ju, iu = rng.integers(3, Nj - 3, 40), rng.integers(3, Ni - 3, 40)
    g["Xu"][ju, iu] += 5.0                                             # U-points dragged 40 % of a cell east
    with engine_for(g) as eng:
        adm, exa = eng.cert_stats()
        assert 0 < adm < exa
        fr, hin, msep = eng.cert_frames()
        big = np.hypot(g["Yf"][jj, ii] - g0["Yf"][jj, ii], g["Xf"][jj, ii] - g0["Xf"][jj, ii]) > 2.0
        assert big.sum() > 20
        for dj, di in ((0, 0), (1, 0), (0, 1), (1, 1)):                        # the four cells around a dragged corner
            assert (hin[jj[big] + dj, ii[big] + di] < 0).all()
        assert (hin[ju, iu] < 0).all() and (hin[ju, iu + 1] < 0).all()         # both cells of a dragged U-point
        n = 200_000
        cells = np.stack([rng.integers(1, Nj, n), rng.integers(1, Ni, n)], axis=1).astype(np.int32)
        yx = np.stack([g["Yt"][cells[:, 0], cells[:, 1]], g["Xt"][cells[:, 0], cells[:, 1]]], axis=1) if "Yt" in g else None
        if yx is None:
            yx = 0.25 * np.stack([g["Yf"][cells[:, 0], cells[:, 1]] + g["Yf"][cells[:, 0] - 1, cells[:, 1] - 1]
                                  + g["Yf"][cells[:, 0] - 1, cells[:, 1]] + g["Yf"][cells[:, 0], cells[:, 1] - 1],
                                  g["Xf"][cells[:, 0], cells[:, 1]] + g["Xf"][cells[:, 0] - 1, cells[:, 1] - 1]
                                  + g["Xf"][cells[:, 0] - 1, cells[:, 1]] + g["Xf"][cells[:, 0], cells[:, 1] - 1]], axis=1)
        yx = yx + rng.uniform(-7, 7, (n, 2))
        vel = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        flags = eng.selftest_cert(yx, cells, vel)
        pick, stay, bad_pick, bad_stay = _check(flags)
        assert bad_pick == 0 and bad_stay == 0
        assert not pick[hin[cells[:, 0], cells[:, 1]] < 0].any()
    far = {k: np.array(v, copy=True) for k, v in g0.items()}
    for k in ("Xf", "Xu", "Xv", "Xt"):
        if k in far:
            far[k] = far[k] + 262144.0
    with engine_for(far) as eng:
        assert eng.cert_stats() == (0, 0)


def test_cert_step_vs_default_at_size(torch):
    """400 k buoys x 12 records on the NANUK4-shaped grid, 3x faster ice: the certified two-kernel step (variant 4)
    against the default kernel, record by record: rows, masks, alive counts and final state bit-identical."""
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS["nanuk4"], seed=0)
    nrec = 12
    U, V, IC = synth.make_records(g, nrec, seed=3)
    U *= np.float32(3.0); V *= np.float32(3.0)
    ids, SG, SC = synth.hss_seeds(g, IC[0], khss=1)
    rng = np.random.default_rng(11)
    rep = int(np.ceil(400_000 / SC.shape[0]))
    res = {}
    for variant in (0, 4, 5):
        with engine_for(g) as eng:
            eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
            cell, near, keep = eng.seed_locate(SG, SC, IC[0])
            ik = np.flatnonzero(keep)
            pos0 = np.repeat(SC[ik], rep, axis=0)
            cell0 = np.repeat(cell[ik], rep, axis=0)
            # jitter inside the cell: a few hundred metres (stays inside on the 12.5 km grid)
            pos0 = pos0 + np.random.default_rng(12).uniform(-1.5, 1.5, pos0.shape)
            eng.set_kernel_variant(variant)
            eng.set_buoys(pos0, cell0)
            r = eng.track((U, V, IC), nrec, pos0=pos0)
            res[variant] = (r, eng.get_state())
    r0, s0 = res[0]; r4, s4 = res[4]
    for k in ("posC", "mask", "n_alive"):
        assert np.array_equal(r0[k], res[5][0][k]), k
    for a, b in zip(s0, res[5][1]):
        assert np.array_equal(a, b)
    assert np.array_equal(r0["posC"], r4["posC"]) and np.array_equal(r0["mask"], r4["mask"])
    assert np.array_equal(r0["n_alive"], r4["n_alive"])
    assert np.abs(r0["posG"][1:] - r4["posG"][1:]).max() == 0.0
    for a, b in zip(s0, s4):
        assert np.array_equal(a, b)
    assert r0["n_alive"][-1] < r0["n_alive"][0]                  # the run does kill buoys
