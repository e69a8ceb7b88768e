"""GPU: the optional physics modes of st_step_ext (Runge-Kutta stepping, C-grid linear interpolation,
multi-hop cell walk).  The reference has no such modes (it is Euler + face pick + one hop,
si3_part_tracker.py:423-484), so these are checked against closed forms, not against an oracle:
  * on a field that is linear in space the amplification factor of each scheme is known exactly
    (Euler 1+z, midpoint 1+z+z^2/2, RK4 1+z+z^2/2+z^3/6+z^4/24 per step, z = a*h), and C-grid linear
    interpolation reproduces the field exactly;
  * the multi-hop walk must end in the cell that a direct index computation gives on a regular grid;
  * with (Euler, face pick, one hop) the mode must agree with the bit-exact reference step wherever the two
    definitions coincide."""
import numpy as np
import pytest

from conftest import engine_for

pytestmark = pytest.mark.gpu

D = 10.0                      # km, regular grid spacing
NJ, NI = 64, 96
X0, Y0 = -300.0, -250.0       # km, T[0,0]


def regular_grid():
    jj, ii = np.meshgrid(np.arange(NJ, dtype=np.float64), np.arange(NI, dtype=np.float64), indexing="ij")
    Xt, Yt = X0 + D * ii, Y0 + D * jj
    g = dict(Xt=Xt, Yt=Yt, Xu=Xt + D / 2, Yu=Yt.copy(), Xv=Xt.copy(), Yv=Yt + D / 2, Xf=Xt + D / 2, Yf=Yt + D / 2,
             tmask=np.ones((NJ, NI), np.int8))
    return g


def cell_of(y, x):
    """T index of the cell containing (y,x): quad F[j-1,i-1]..F[j,i] spans (x_i - D/2, x_i + D/2]."""
    i = np.ceil((x - X0 - D / 2) / D).astype(int)
    j = np.ceil((y - Y0 - D / 2) / D).astype(int)
    return np.stack([j, i], axis=1)


def run(torch, g, U, V, pos0, nrec, scheme, interp, hops, uv_strategy=1):
    dev = torch.device("cuda", 0)
    nP = pos0.shape[0]
    IC = np.ones((NJ, NI), np.float32)
    with engine_for(g, uv_strategy=uv_strategy) as eng:
        eng.set_buoys(pos0, cell_of(pos0[:, 0], pos0[:, 1]).astype(np.int32))
        eng.record_slots(1)
        st = eng.staging(0)
        st[0], st[1], st[2] = U, V, IC
        eng.submit_record(0)
        yx = torch.empty((nP, 2), dtype=torch.float64, device=dev)
        ll = torch.empty((nP, 2), dtype=torch.float64, device=dev)
        mk = torch.empty((nP,), dtype=torch.int8, device=dev)
        na = torch.zeros((nrec,), dtype=torch.int64, device=dev)
        for k in range(nrec):
            eng.step_ext(0, k, scheme, interp, hops, yx, ll, mk, na[k:k + 1])
        torch.cuda.synchronize()
        pos, cell, alive = eng.get_state()
        return pos, cell, alive, yx.cpu().numpy(), mk.cpu().numpy(), na.cpu().numpy()


@pytest.mark.parametrize("scheme", [1, 2, 4])
def test_amplification_factor_on_a_linear_field(torch, scheme):
    """u = a x, v = b y sampled at the U/V points: linear C-grid interpolation is exact, so n steps multiply
    x by R(a h)^n with the scheme's stability polynomial R -- to the rounding of the f4 face values."""
    g = regular_grid()
    h = 3.6                                                    # km per (m/s) per record (rdt = 3600 s)
    za, zb = 0.02, -0.015                                      # a*h, b*h
    U = ((za / h) * g["Xu"]).astype(np.float32)
    V = ((zb / h) * g["Yv"]).astype(np.float32)
    rng = np.random.default_rng(3)
    nP, nrec = 500, 40
    pos0 = np.stack([rng.uniform(-120, 120, nP), rng.uniform(20, 150, nP)], axis=1)       # [y, x]
    pos, cell, alive, yx, mk, na = run(torch, g, U, V, pos0, nrec, scheme, 1, 4)
    R = {1: lambda z: 1 + z, 2: lambda z: 1 + z + z * z / 2, 4: lambda z: 1 + z + z * z / 2 + z ** 3 / 6 + z ** 4 / 24}[scheme]
    want = np.stack([pos0[:, 0] * R(zb) ** nrec, pos0[:, 1] * R(za) ** nrec], axis=1)
    assert alive.all() and mk.all() and (na == nP).all()
    assert np.abs(pos - want).max() < 2e-4                     # km; f4 face values: ~6e-8 relative per step
    assert np.array_equal(yx, pos)                             # the row is the new state
    assert np.array_equal(cell, cell_of(pos[:, 0], pos[:, 1]))  # warm-started walk kept every host cell right
    exact = np.stack([pos0[:, 0] * np.exp(zb * nrec), pos0[:, 1] * np.exp(za * nrec)], axis=1)
    err = np.abs(pos - exact).max()
    # order of accuracy against the exact flow: Euler ~ n z^2/2, midpoint ~ n z^3/6, RK4 ~ n z^5/120 (relative)
    bound = {1: 3.0, 2: 0.03, 4: 2e-4}[scheme]
    assert err < bound and (scheme == 4 or err > bound / 50)


def test_multi_hop_walk_finds_the_host_cell(torch):
    """Uniform fast ice (2.9 and 1.8 cells per record): with enough hops the state's cell is the containing
    cell after every record; with one hop (the reference's rule) it falls behind."""
    g = regular_grid()
    U = np.full((NJ, NI), 8.0, np.float32)                    # 28.8 km per record
    V = np.full((NJ, NI), -5.0, np.float32)                   # -18 km per record
    rng = np.random.default_rng(4)
    nP, nrec = 300, 6
    pos0 = np.stack([rng.uniform(50, 250, nP), rng.uniform(-250, -50, nP)], axis=1)
    pos, cell, alive, yx, mk, na = run(torch, g, U, V, pos0, nrec, 1, 0, 8)
    want = pos0 + nrec * 3.6 * np.array([-5.0, 8.0])
    assert np.abs(pos - want).max() < 1e-9 and alive.all()
    assert np.array_equal(cell, cell_of(pos[:, 0], pos[:, 1]))
    pos1, cell1, alive1, *_ = run(torch, g, U, V, pos0, nrec, 1, 0, 1)
    assert np.abs(pos1 - want).max() < 1e-9
    lag = np.abs(cell1 - cell_of(pos1[:, 0], pos1[:, 1])).max(axis=1)
    assert (lag > 0).mean() > 0.9                              # one diagonal hop per record is not enough here


def test_land_and_domain_rim_kill_in_ext_mode(torch):
    """Survive (tracking.py:62-93) applies to every cell entered: a land block and the 2-cell rim."""
    g = regular_grid()
    g["tmask"][20:30, 60:64] = 0                               # a wall east of the cloud
    U = np.full((NJ, NI), 6.0, np.float32)
    V = np.zeros((NJ, NI), np.float32)
    yy, xx = np.meshgrid(Y0 + D * np.arange(10, 50), np.array([100.0, 103.0]), indexing="ij")
    pos0 = np.stack([yy.ravel() + 1.0, xx.ravel()], axis=1)
    pos, cell, alive, yx, mk, na = run(torch, g, U, V, pos0, 40, 2, 1, 4)
    rows = cell_of(pos0[:, 0], pos0[:, 1])[:, 0]
    hit_wall = (rows >= 19) & (rows <= 30)                     # the 5-point stencil reaches one row further
    assert not alive[hit_wall].any()
    assert (pos[hit_wall][:, 1] < X0 + D * 62).all()           # stopped at the wall ...
    assert not alive[~hit_wall].any()                          # ... the others ran into the eastern rim
    assert (pos[~hit_wall][:, 1] > X0 + D * (NI - 4)).all()
    assert na[0] == pos0.shape[0] and na[-1] == 0


def test_ext_reference_settings_track_the_bit_exact_step(torch, gold_track):
    """(Euler, face pick, one hop): the same trajectories as the reference step up to the rounding of h*u
    versus (u*rdt)/1000, except where the two cell-search rules differ (corner exits, points on an edge)."""
    T, g = gold_track
    nrec, nP = T["U"].shape[0], T["pos0"].shape[0]
    dev = torch.device("cuda", 0)
    with engine_for(g) as eng:
        eng.set_buoys(T["pos0"], T["jiT0"])
        eng.record_slots(1)
        yx = torch.empty((nP, 2), dtype=torch.float64, device=dev)
        mk = torch.empty((nP,), dtype=torch.int8, device=dev)
        same = np.ones(nP, bool)
        for k in range(nrec):
            st = eng.staging(0)
            torch.cuda.synchronize()
            st[0], st[1], st[2] = T["U"][k], T["V"][k], T["IC"][k]
            eng.submit_record(0)
            eng.step_ext(0, k, 1, 0, 1, yx, None, mk, None)
            torch.cuda.synchronize()
            p, c, a = eng.get_state()
            same &= (c == T["uv1_jiT"][k + 1]).all(axis=1) & (a == T["uv1_alive"][k + 1])
            got, want = yx.cpu().numpy(), T["uv1_posC"][k + 1]
            live = same & (T["uv1_mask"][k + 1] == 1)
            assert np.abs(got[live] - want[live]).max() < 1e-9
        assert same.mean() > 0.9, same.mean()


@pytest.mark.parametrize("scheme,interp,hops,uv", [(2, 1, 4, 1), (4, 1, 4, 1), (1, 0, 3, 1), (4, 0, 2, 1), (2, 0, 2, 0)])
def test_ext_modes_on_a_curvilinear_grid_vs_numpy_restatement(torch, gold_track, scheme, interp, hops, uv):
    """On the warped golden grid (no closed form) the kernel must follow oracle/ext_ref.py -- an independent
    numpy restatement of the same definitions -- record by record: positions to 1e-9 km, cells and alive
    flags exactly.  3x faster ice than the golden run, so multi-cell moves, stage walks and kills all occur."""
    from oracle import ext_ref
    T, g = gold_track
    nrec, nP = 20, T["pos0"].shape[0]
    U, V, IC = 3 * T["U"][:nrec], 3 * T["V"][:nrec], T["IC"][:nrec]
    pos, cell, alive = T["pos0"].copy(), T["jiT0"].astype(np.int64).copy(), np.ones(nP, np.int8)
    dev = torch.device("cuda", 0)
    moved = 0
    with engine_for(g, uv_strategy=uv) as eng:
        eng.set_buoys(T["pos0"], T["jiT0"])
        eng.record_slots(1)
        yx = torch.empty((nP, 2), dtype=torch.float64, device=dev)
        mk = torch.empty((nP,), dtype=torch.int8, device=dev)
        for k in range(nrec):
            st = eng.staging(0)
            torch.cuda.synchronize()
            st[0], st[1], st[2] = U[k], V[k], IC[k]
            eng.submit_record(0)
            eng.step_ext(0, k, scheme, interp, hops, yx, None, mk, None)
            torch.cuda.synchronize()
            c_before = cell.copy()
            want_yx, want_mk = ext_ref.step(g, U[k], V[k], IC[k], pos, cell, alive, scheme, interp, hops, uv_strategy=uv)
            moved += int((c_before != cell).any(axis=1).sum())
            got = yx.cpu().numpy()
            assert np.array_equal(mk.cpu().numpy(), want_mk), k
            live = want_mk == 1
            assert np.abs(got[live] - want_yx[live]).max() < 1e-9, k
            assert np.array_equal(got[~live], want_yx[~live])
            p, c, a = eng.get_state()
            assert np.array_equal(c, cell) and np.array_equal(a, alive), k
    assert moved > 50 and alive.sum() < nP
