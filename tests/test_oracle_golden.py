"""CPU: the C oracle (oracle/st_oracle.c) against fixtures produced by the reference's
own Python functions (oracle/make_golden.py).  Bit-exact for everything built from
+ - * / and comparisons; Haversine within 1e-9 km (libm vs numpy transcendentals)."""
import numpy as np
import pytest

from conftest import TRACK_CASES


def test_known_answer_inside_quad(gold_pred, corc):
    # the only known-answer vector upstream (tools/tests/test_pnt_inside_quad.py:16-24)
    G = gold_pred
    assert list(G["kat_inside"]) == [True, False, False, True]
    got = [corc.inside_quad(p[0], p[1], G["kat_quad"]) for p in G["kat_pts"]]
    assert got == list(G["kat_inside"])


def test_unit_square_edge_rules(gold_pred, corc):
    G = gold_pred
    got = [corc.inside_quad(p[0], p[1], G["sq_quad"]) for p in G["sq_pts"]]
    assert got == list(G["sq_inside"])


def test_random_quads(gold_pred, corc):
    G = gold_pred
    got = np.array([corc.inside_quad(p[0], p[1], q) for p, q in zip(G["rq_pts"], G["rq_quads"])])
    assert np.array_equal(got, G["rq_inside"])


def test_intersect2seg(gold_pred, corc):
    G = gold_pred
    got = np.array([corc.intersect2seg(*s) for s in G["seg_pts"]])
    assert np.array_equal(got, G["seg_hit"])


def test_cell_walk(gold_pred, corc):
    G = gold_pred
    Yf, Xf = G["g_Yf"], G["g_Xf"]
    ce = np.array([corc.crossed_edge(a, b, j, i, Yf, Xf)
                   for a, b, j, i in zip(G["walk_p1"], G["walk_p2"], G["walk_jT"], G["walk_iT"])])
    assert np.array_equal(ce, G["walk_cross"])
    nh = np.array([corc.new_host_cell(c, a, b, j, i, Yf, Xf)
                   for c, a, b, j, i in zip(ce, G["walk_p1"], G["walk_p2"], G["walk_jT"], G["walk_iT"])])
    assert np.array_equal(nh, G["walk_newcell"])
    assert set(np.unique(nh)) == set(range(1, 9))          # all 8 directions are exercised


def test_survive(gold_pred, gold_track, corc):
    G = gold_pred
    ic = gold_track[0]["IC"][7]
    got = np.array([corc.survive(j, i, G["g_tmask"], ic) for j, i in zip(G["sv_jT"], G["sv_iT"])])
    assert np.array_equal(got, G["sv_kill"])


def test_survive_quirks(corc):
    tm = np.ones((9, 9), np.int8)
    ic = np.full((9, 9), 0.1, np.float32)
    # 0.2*(5 x f4(0.1)) is not < 0.1 -> survives; land at [jT-1,iT] does not kill, [jT-1,iT-1] does
    assert corc.survive(4, 4, tm, ic) == 0
    tm2 = tm.copy(); tm2[3, 4] = 0
    assert corc.survive(4, 4, tm2, ic) == 0
    tm3 = tm.copy(); tm3[3, 3] = 0
    assert corc.survive(4, 4, tm3, ic) == 1


def test_haversine(gold_pred, corc):
    G = gold_pred
    for p, want in zip(G["hav_pts"], G["hav_d"]):
        got = np.array([[corc.lib().orc_haversine(p[0], p[1], a, b) for a, b in zip(ra, rb)]
                        for ra, rb in zip(G["hav_glat"], G["hav_glon"])])
        assert np.allclose(got, want, rtol=0, atol=1e-9)


def test_nearest_ladder(gold_pred):
    lad = gold_pred["np_ladder"]
    assert lad[0] == 20.0 and lad[2] == 28.799999999999997 and abs(lad[7] - 71.66361599999998) < 1e-12


@pytest.mark.parametrize("name", list(TRACK_CASES))
def test_track_cases(gold_track, corc, name):
    T, g = gold_track
    c = TRACK_CASES[name]
    kw = dict(uv_strategy=c["uv_strategy"], kstrt=c["kstrt"])
    if c["win"]:
        kw.update(rec_first=T["win_first"], rec_last=T["win_last"])
    s = np.float32(c["scale"])
    r = corc.track(g, s * T["U"], s * T["V"], T["IC"], T["pos0"], T["jiT0"], **kw)
    assert np.array_equal(r["posC"], T[name + "_posC"])             # bit-exact f8 positions
    assert np.array_equal(r["mask"], T[name + "_mask"])
    assert np.array_equal(r["jiT_hist"], T[name + "_jiT"])
    assert np.array_equal(r["alive_hist"], T[name + "_alive"])
    assert np.array_equal(r["nalive"], T[name + "_nalive"])


def test_seed_init(gold_seed, corc):
    S, g = gold_seed
    jiT, keep, near = corc.seed_init(S["SG"], S["SC"], g["latT"], g["lonT"], g["Yf"], g["Xf"], g["ResKM"],
                                     g["tmask"], S["ic0"])
    ik = np.flatnonzero(keep)
    assert np.array_equal(near, S["out_nearest"])
    assert np.array_equal(ik, S["out_iKeep"])
    assert np.array_equal(jiT[ik], S["out_jiT"])
    assert 0 < ik.size < S["SG"].shape[0]                           # some seeds are dropped


def test_projection_roundtrip_and_series(corc):
    """orc_inv_stere (PROJ-style iteration) vs an independent numpy iteration to machine
    precision, and the forward/inverse round trip.  Parity with PROJ itself is UNPINNED."""
    import synth.grid as sg
    rng = np.random.default_rng(3)
    yx = rng.uniform(-4000, 4000, (2000, 2))
    ll = corc.inv_stere(yx)
    lat, lon = sg.km_to_latlon(yx[:, 0], yx[:, 1])
    assert np.abs(ll[:, 0] - lat).max() < 1e-9 and np.abs(ll[:, 1] - lon).max() < 1e-9
    back = corc.fwd_stere(ll)
    assert np.abs(back - yx).max() < 1e-6


@pytest.mark.parametrize("name", ["uv1", "uv0", "fast", "win"])
def test_pyport_track_cases(gold_track, name):
    """The pure-Python port (timed as the as-shipped CPU baseline) is pinned by the same fixtures."""
    from oracle import pyport
    T, g = gold_track
    c = TRACK_CASES[name]
    kw = dict(uv_strategy=c["uv_strategy"], kstrt=c["kstrt"])
    if c["win"]:
        kw.update(rec_first=T["win_first"], rec_last=T["win_last"])
    sc = np.float32(c["scale"])
    posC, mask, jiT, alive, nsteps = pyport.track(g, T["U"] * sc, T["V"] * sc, T["IC"], T["pos0"], T["jiT0"], **kw)
    assert np.array_equal(posC, T[name + "_posC"]) and np.array_equal(mask, T[name + "_mask"])
    assert np.array_equal(jiT, T[name + "_jiT"][-1]) and np.array_equal(alive, T[name + "_alive"][-1])
    assert nsteps == int(T[name + "_mask"][1:].sum()) - (int((T["win_first"] > c["kstrt"]).sum()) if c["win"] else 0)
