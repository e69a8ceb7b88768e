"""GPU: the CUDA path (through the C ABI) against (1) the committed golden fixtures made by
the reference's own functions and (2) the C oracle on larger seeded inputs.

Bars: cell indices, alive flags, masks, alive counts and f8 positions bit-exact (every
operation is an un-fused IEEE double op in the reference's order); lat/lon within
1e-9 degrees of the PROJ-style iterative inverse (PROJ itself iterates to 1e-10 rad).
"""
import os

import numpy as np
import pytest

from conftest import TRACK_CASES, engine_for

pytestmark = pytest.mark.gpu

LATLON_TOL_DEG = 1e-9


# ---- scalar / batched predicates against the reference's golden answers -------------------

def test_inside_quad_golden(sit, gold_pred):
    G = gold_pred
    assert [sit.IsInsideQuadrangle(p[0], p[1], G["kat_quad"]) for p in G["kat_pts"]] == [True, False, False, True]
    got = sit.IsInsideQuadrangleBatch(G["sq_pts"], np.repeat(G["sq_quad"][None], len(G["sq_pts"]), 0))
    assert np.array_equal(got, G["sq_inside"])
    assert np.array_equal(sit.IsInsideQuadrangleBatch(G["rq_pts"], G["rq_quads"]), G["rq_inside"])


def test_intersect_golden(sit, gold_pred):
    S = gold_pred["seg_pts"]
    got = sit.intersect2SegBatch(S[:, 0], S[:, 1], S[:, 2], S[:, 3])
    assert np.array_equal(got, gold_pred["seg_hit"])
    assert sit.intersect2Seg(*S[0]) == bool(gold_pred["seg_hit"][0])


def test_cell_walk_golden(sit, gold_pred):
    G = gold_pred
    Yf, Xf = G["g_Yf"], G["g_Xf"]
    for k in range(0, len(G["walk_jT"]), 7):
        jT, iT = int(G["walk_jT"][k]), int(G["walk_iT"][k])
        V = np.array([[jT - 1, jT - 1, jT, jT], [iT - 1, iT, iT, iT - 1]])
        kc = sit.CrossedEdge(G["walk_p1"][k], G["walk_p2"][k], V, Yf, Xf)
        assert kc == G["walk_cross"][k]
        assert sit.NewHostCell(kc, G["walk_p1"][k], G["walk_p2"][k], V, Yf, Xf) == G["walk_newcell"][k]


def test_survive_golden(sit, gold_pred, gold_track):
    G = gold_pred
    ic = np.zeros(G["g_tmask"].shape); ic[:, :] = gold_track[0]["IC"][7]
    got = sit.SurviveBatch(np.stack([G["sv_jT"], G["sv_iT"]], 1), G["g_tmask"], ic)
    assert np.array_equal(got, G["sv_kill"])
    assert sit.Survive(1, [int(G["sv_jT"][20]), int(G["sv_iT"][20])], G["g_tmask"], pIceC=ic) == G["sv_kill"][20]
    with pytest.raises(UnboundLocalError):
        sit.Survive(1, [5, 5], G["g_tmask"])


def test_haversine_golden(sit, gold_pred):
    G = gold_pred
    for p, want in zip(G["hav_pts"][:16], G["hav_d"][:16]):
        assert np.allclose(sit.Haversine(p[0], p[1], G["hav_glat"], G["hav_glon"]), want, rtol=0, atol=1e-9)


# ---- the record loop against the reference's golden trajectories ----------------------------

def _run_engine(torch, g, U, V, IC, pos0, cell0, kstrt=0, first=None, last=None, uv_strategy=1, multi=False,
                variant=0):
    nrec = U.shape[0]
    nP = pos0.shape[0]
    dev = torch.device("cuda", 0)
    with engine_for(g, uv_strategy=uv_strategy) as eng:
        eng.set_kernel_variant(variant)
        eng.set_buoys(pos0, cell0, first, last)
        yx = torch.empty((nrec, nP, 2), dtype=torch.float64, device=dev)
        ll = torch.empty((nrec, nP, 2), dtype=torch.float64, device=dev)
        mk = torch.empty((nrec, nP), dtype=torch.int8, device=dev)
        na = torch.zeros((nrec,), dtype=torch.int64, device=dev)
        cells = np.zeros((nrec + 1, nP, 2), np.int32); alive = np.zeros((nrec + 1, nP), np.int8)
        cells[0] = cell0; alive[0] = 1
        if multi:
            rec = torch.from_numpy(np.stack([U, V, IC], axis=1).astype(np.float32)).to(dev).contiguous()
            eng.step_multi(rec, kstrt, yx, ll, mk, na)
            torch.cuda.synchronize()
            _, c, a = eng.get_state()
            cells[-1], alive[-1] = c, a
        else:
            eng.record_slots(1)
            for k in range(nrec):
                st = eng.staging(0)
                st[0], st[1], st[2] = U[k], V[k], IC[k]
                eng.submit_record(0)
                eng.step(0, k + kstrt, yx[k], ll[k], mk[k], na[k:k + 1])
                torch.cuda.synchronize()
                _, c, a = eng.get_state()
                cells[k + 1], alive[k + 1] = c, a
        return yx.cpu().numpy(), ll.cpu().numpy(), mk.cpu().numpy(), na.cpu().numpy(), cells, alive


@pytest.mark.parametrize("multi", [False, True])
@pytest.mark.parametrize("name", list(TRACK_CASES))
def test_track_golden(torch, corc, gold_track, name, multi):
    T, g = gold_track
    c = TRACK_CASES[name]
    s = np.float32(c["scale"])
    first = T["win_first"] if c["win"] else None
    last = T["win_last"] if c["win"] else None
    yx, ll, mk, na, cells, alive = _run_engine(torch, g, s * T["U"], s * T["V"], T["IC"], T["pos0"],
                                               T["jiT0"], c["kstrt"], first, last, c["uv_strategy"], multi)
    assert np.array_equal(yx, T[name + "_posC"][1:])                 # bit-exact positions, fill rows included
    assert np.array_equal(mk, T[name + "_mask"][1:])
    assert np.array_equal(na, T[name + "_nalive"])
    if multi:
        assert np.array_equal(cells[-1], T[name + "_jiT"][-1]) and np.array_equal(alive[-1], T[name + "_alive"][-1])
    else:
        assert np.array_equal(cells, T[name + "_jiT"])               # cell index per buoy per record
        assert np.array_equal(alive, T[name + "_alive"])
    want = corc.inv_stere(T[name + "_posC"][1:].reshape(-1, 2)).reshape(ll.shape)
    assert np.abs(ll - want).max() < LATLON_TOL_DEG                  # also for the -9999 rows (:493)


@pytest.mark.parametrize("variant", [1, 3, 4, 5])
@pytest.mark.parametrize("name", list(TRACK_CASES))
def test_track_golden_other_kernels(torch, corc, gold_track, name, variant):
    """v1 (straightforward) and the round-1 default (k_advect_warp with and without the orientation filter)
    reproduce the reference's golden trajectories bit for bit too; the default (0, k_advect_cert) is what
    test_track_golden runs.  Round-1 experiment kernels live behind -DST_EXPERIMENTS."""
    T, g = gold_track
    c = TRACK_CASES[name]
    s = np.float32(c["scale"])
    first = T["win_first"] if c["win"] else None
    last = T["win_last"] if c["win"] else None
    yx, ll, mk, na, cells, alive = _run_engine(torch, g, s * T["U"], s * T["V"], T["IC"], T["pos0"], T["jiT0"],
                                               c["kstrt"], first, last, c["uv_strategy"], False, variant)
    assert np.array_equal(yx, T[name + "_posC"][1:]) and np.array_equal(mk, T[name + "_mask"][1:])
    assert np.array_equal(na, T[name + "_nalive"])
    assert np.array_equal(cells, T[name + "_jiT"]) and np.array_equal(alive, T[name + "_alive"])
    want = corc.inv_stere(T[name + "_posC"][1:].reshape(-1, 2)).reshape(ll.shape)
    assert np.abs(ll - want).max() < LATLON_TOL_DEG                  # every launch shape writes the same xPosG row


@pytest.mark.parametrize("variant", [0, 3, 5])
@pytest.mark.parametrize("name", list(TRACK_CASES))
def test_row_chain_golden(torch, gold_track, name, variant):
    """Row chaining (st_set_row_chain): the f8 row of record k is the position input of record k+1 and the state's
    own copy is rewritten only at a buoy's death and by st_sync_state.  Rows, masks, alive counts, cells and the
    final positions (discontinued buoys included) equal the golden run / the un-chained state bit for bit, whichever
    way the caller lays out its rows: one row per record, two alternating buffers, ONE buffer stepped in place, a
    state read-back after every record, and steps that cannot chain (no yx row, f4 rows) in between.  The windowed
    case never chains (a buoy outside its window has a fill row but a live position) and must be unaffected."""
    T, g = gold_track
    c = TRACK_CASES[name]
    sc = np.float32(c["scale"])
    U, V, IC = sc * T["U"], sc * T["V"], T["IC"]
    first = T["win_first"] if c["win"] else None
    last = T["win_last"] if c["win"] else None
    nrec, nP = U.shape[0], T["pos0"].shape[0]
    dev = torch.device("cuda", 0)
    want_yx, want_mk = T[name + "_posC"][1:], T[name + "_mask"][1:]
    with engine_for(g, uv_strategy=c["uv_strategy"]) as eng:
        eng.set_kernel_variant(variant)
        eng.record_slots(1)

        def run(mode):
            eng.set_buoys(T["pos0"], T["jiT0"], first, last)
            eng.set_row_chain(mode != "off")
            nb = {"rows": nrec, "off": nrec, "sync_each": nrec, "mixed": nrec, "pingpong": 2, "inplace": 1}[mode]
            yx = torch.full((nb, nP, 2), 7.0, dtype=torch.float64, device=dev)
            yx4 = torch.empty((nP, 2), dtype=torch.float32, device=dev)
            mk = torch.empty((nrec, nP), dtype=torch.int8, device=dev)
            na = torch.zeros((nrec,), dtype=torch.int64, device=dev)
            rows = np.full((nrec, nP, 2), np.nan)
            for k in range(nrec):
                st = eng.staging(0)
                st[0], st[1], st[2] = U[k], V[k], IC[k]
                eng.submit_record(0)
                if mode == "mixed" and k % 3 == 1:
                    eng.step(0, k + c["kstrt"], None, None, mk[k], na[k:k + 1])        # no row: cannot chain
                elif mode == "mixed" and k % 3 == 2:
                    eng.step(0, k + c["kstrt"], yx4, None, mk[k], na[k:k + 1])         # f4 row: cannot chain
                    torch.cuda.synchronize()
                    rows[k] = yx4.cpu().numpy().astype(np.float64)
                else:
                    eng.step(0, k + c["kstrt"], yx[k % nb], None, mk[k], na[k:k + 1])
                    torch.cuda.synchronize()
                    rows[k] = yx[k % nb].cpu().numpy()
                if mode == "sync_each":
                    _, cc, aa = eng.get_state()
                    assert np.array_equal(cc, T[name + "_jiT"][k + 1]) and np.array_equal(aa, T[name + "_alive"][k + 1])
            state = eng.get_state()
            eng.set_row_chain(False)
            return rows, mk.cpu().numpy(), na.cpu().numpy(), state

        rows0, mk0, na0, st0 = run("off")
        assert np.array_equal(rows0, want_yx) and np.array_equal(mk0, want_mk) and np.array_equal(na0, T[name + "_nalive"])
        assert np.array_equal(st0[1], T[name + "_jiT"][-1]) and np.array_equal(st0[2], T[name + "_alive"][-1])
        for mode in ("rows", "pingpong", "inplace", "sync_each", "mixed"):
            rows, mk, na, st = run(mode)
            if mode == "mixed":
                f8 = np.arange(nrec) % 3 == 0
                f4 = np.arange(nrec) % 3 == 2
                assert np.array_equal(rows[f8], want_yx[f8]), mode
                assert np.array_equal(rows[f4], want_yx[f4].astype(np.float32).astype(np.float64)), mode
            else:
                assert np.array_equal(rows, want_yx), mode
            assert np.array_equal(mk, want_mk) and np.array_equal(na, T[name + "_nalive"]), mode
            for a, b in zip(st, st0):
                assert np.array_equal(a, b), mode                   # positions of discontinued buoys included


def test_track_host_call_and_pipeline(torch, gold_track):
    """The host-buffer entry point (what a reference-side binding calls per record) and the
    pipelined engine.track() give the same rows as the golden run."""
    T, g = gold_track
    nrec, nP = T["U"].shape[0], T["pos0"].shape[0]
    with engine_for(g) as eng:
        eng.set_buoys(T["pos0"], T["jiT0"])
        yx = np.empty((nP, 2)); ll = np.empty((nP, 2)); mk = np.empty(nP, np.int8)
        for k in range(nrec):
            na = eng.track_record_host(k, T["U"][k], T["V"][k], T["IC"][k], yx, ll, mk)
            assert na == T["uv1_nalive"][k]
            assert np.array_equal(yx, T["uv1_posC"][k + 1]) and np.array_equal(mk, T["uv1_mask"][k + 1])
    with engine_for(g) as eng:
        eng.set_buoys(T["pos0"], T["jiT0"])
        r = eng.track((T["U"], T["V"], T["IC"]), nrec, pos0=T["pos0"], posG0=T["posG0"])
        assert np.array_equal(r["posC"], T["uv1_posC"]) and np.array_equal(r["mask"], T["uv1_mask"])
        assert np.array_equal(r["n_alive"], T["uv1_nalive"])
        rows = []
        eng.set_buoys(T["pos0"], T["jiT0"])
        eng.track((T["U"], T["V"], T["IC"]), nrec, sink=lambda k, y, l, m: rows.append((k, y.copy(), m.copy())))
        assert [k for k, _, _ in rows] == list(range(nrec))
        assert all(np.array_equal(y, T["uv1_posC"][k + 1]) for k, y, _ in rows)


def test_file_dtype_rows_are_the_f4_cast_of_the_f8_rows(torch, gold_track):
    """st_step_f4 / st_step_multi_f4 / st_track_record_host_f4: rows in the output file's dtype
    (ncio.py:153-159) equal numpy's f8 -> f4 cast of the golden rows bit for bit, for every kernel family,
    and the f8 state on the device is untouched by the narrower rows."""
    T, g = gold_track
    nrec, nP = T["U"].shape[0], T["pos0"].shape[0]
    want_yx, want_mk = T["uv1_posC"].astype(np.float32), T["uv1_mask"]
    with engine_for(g) as eng:
        eng.set_buoys(T["pos0"], T["jiT0"])
        yx = np.empty((nP, 2), np.float32); ll = np.empty((nP, 2), np.float32); mk = np.empty(nP, np.int8)
        ll8 = np.empty((nP, 2))
        for k in range(nrec):
            eng.track_record_host(k, T["U"][k], T["V"][k], T["IC"][k], yx, ll, mk)
            assert np.array_equal(yx, want_yx[k + 1]) and np.array_equal(mk, want_mk[k + 1])
        with pytest.raises(TypeError):
            eng.track_record_host(0, T["U"][0], T["V"][0], T["IC"][0], yx, ll8, mk)
        pos, cell, alive = eng.get_state()
        assert np.array_equal(cell, T["uv1_jiT"][-1]) and np.array_equal(alive, T["uv1_alive"][-1])
    for variant, chunk in ((0, None), (1, None), (3, None), (4, None), (5, None), (0, 7)):
        with engine_for(g) as eng:
            eng.set_kernel_variant(variant)
            eng.set_buoys(T["pos0"], T["jiT0"])
            r = eng.track((T["U"], T["V"], T["IC"]), nrec, pos0=T["pos0"], posG0=T["posG0"], row_dtype="f4", chunk=chunk)
            r8 = None
            assert r["posC"].dtype == np.float32 and r["posG"].dtype == np.float32
            assert np.array_equal(r["posC"], want_yx) and np.array_equal(r["mask"], want_mk)
            eng.set_buoys(T["pos0"], T["jiT0"])
            r8 = eng.track((T["U"], T["V"], T["IC"]), nrec, pos0=T["pos0"], posG0=T["posG0"], chunk=chunk)
            assert np.array_equal(r["posG"][1:], r8["posG"][1:].astype(np.float32))     # lat/lon: the cast of the f8 row


@pytest.mark.parametrize("mode", [0, 1, 2], ids=["thread_stores", "copy_engines", "bulk_stores"])
@pytest.mark.parametrize("f4", [False, True])
def test_fused_position_allgather_two_ranks_one_device(torch, gold_track, f4, mode):
    """st_step_gather: two contexts on cuda:0 play two ranks (st_gather_connect_ptrs).  Each owns a shard
    of the golden cloud and its rows reach BOTH gathered arrays -- stored by the step kernel itself (per thread, or
    a tile at a time through shared memory and cp.async.bulk) or pushed by the copy engines; after every record
    both arrays must equal the unsharded golden row -- the ready/ack flag protocol included (nbuf = 2)."""
    from sitrack_b200 import dist as sdist
    T, g = gold_track
    nrec, nP = T["U"].shape[0], T["pos0"].shape[0]
    b = sdist.shard_bounds(nP, 2, tile=64)
    dev = torch.device("cuda", 0)
    engs = [engine_for(g).__enter__() for _ in range(2)]
    try:
        for r, eng in enumerate(engs):
            eng.set_buoys(T["pos0"][b[r]:b[r + 1]], T["jiT0"][b[r]:b[r + 1]])
            eng.record_slots(1)
            h = eng.gather_create(r, 2, nP, int(b[r]), f4=f4, nbuf=2)
            assert len(h) == 64
        blocks = [eng.gather_block()[0] for eng in engs]
        for eng in engs:
            eng.gather_connect_ptrs(blocks)
            eng.gather_set_mode(mode)
        streams = [torch.cuda.Stream(dev) for _ in range(2)]
        mks = [torch.empty((int(b[r + 1] - b[r]),), dtype=torch.int8, device=dev) for r in range(2)]
        for k in range(nrec):
            seq, buf = k + 1, k % 2
            for r, eng in enumerate(engs):
                st = eng.staging(0)
                st[0], st[1], st[2] = T["U"][k], T["V"][k], T["IC"][k]
                eng.submit_record(0, streams[r])
                eng.step_gather(0, k, buf, seq, None, mks[r], None, streams[r])
            for r, eng in enumerate(engs):
                eng.gather_wait(seq, streams[r])
            torch.cuda.synchronize()
            want = T["uv1_posC"][k + 1]
            for r, eng in enumerate(engs):
                got = eng.gather_buffer(buf).cpu().numpy()
                assert np.array_equal(got, want.astype(np.float32) if f4 else want), (k, r)
                assert np.array_equal(mks[r].cpu().numpy(), T["uv1_mask"][k + 1][b[r]:b[r + 1]])
                eng.gather_ack(seq, streams[r])
        torch.cuda.synchronize()
        assert not any(eng.gather_timed_out() for eng in engs)
    finally:
        for eng in engs:
            eng.close()


# ---- seeding ------------------------------------------------------------------------------------

def test_seedinit_golden(sit, gold_seed):
    S, g = gold_seed
    ic = np.zeros(g["tmask"].shape); ic[:, :] = S["ic0"]
    nP, pSG, pSC, pIDs, zjiT, zV, iKeep = sit.SeedInit(S["ids"].copy(), S["SG"], S["SC"], g["latT"], g["lonT"],
                                                       g["Yf"], g["Xf"], g["ResKM"], g["tmask"], xIceConc=ic)
    assert nP == int(S["out_nP"]) and np.array_equal(iKeep, S["out_iKeep"])
    assert np.array_equal(zjiT, S["out_jiT"]) and np.array_equal(zV, S["out_VRTCS"])
    assert np.array_equal(pSG, S["out_SG"]) and np.array_equal(pSC, S["out_SC"]) and np.array_equal(pIDs, S["out_IDs"])


def test_nearest_point_hash_vs_brute_vs_golden(sit, gold_seed):
    S, g = gold_seed
    ji_h, d_h = sit.NearestPointBatch(S["SG"], g["latT"], g["lonT"], 2.5, g["ResKM"], 10)
    assert np.array_equal(ji_h, S["out_nearest"])
    # exactness of the hash search: unconditional argmin (max_itr=1 accepts anything) == whole-grid scan
    ji_a, d_a = sit.NearestPointBatch(S["SG"], g["latT"], g["lonT"], 2.5, g["ResKM"], 1)
    ji_b, d_b = sit.NearestPointBatch(S["SG"], g["latT"], g["lonT"], 2.5, g["ResKM"], 1, brute=True)
    assert np.array_equal(ji_a, ji_b) and np.array_equal(d_a, d_b)
    k = int(np.flatnonzero(S["out_nearest"][:, 0] >= 0)[0])
    assert sit.NearestPoint(tuple(S["SG"][k]), g["latT"], g["lonT"], rd_found_km=2.5, resolkm=g["ResKM"],
                            max_itr=10) == tuple(S["out_nearest"][k])
    lPin, ji, vr = sit.FindContainingCell(tuple(S["SC"][S["out_iKeep"][0]]), tuple(S["out_nearest"][S["out_iKeep"][0]]),
                                          g["Yf"], g["Xf"])
    assert lPin and ji == list(S["out_jiT"][0]) and np.array_equal(np.array(vr), S["out_VRTCS"][0])


# ---- larger seeded inputs against the C oracle ------------------------------------------------------

@pytest.mark.parametrize("preset,khss,nrec,scale", [("small", 1, 72, 1.0), ("nanuk4", 5, 24, 1.0),
                                                     ("nanuk4", 2, 48, 3.0)])
def test_track_vs_oracle(torch, corc, preset, khss, nrec, scale):
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS[preset], seed=0)
    U, V, IC = synth.make_records(g, nrec, seed=1)
    U *= np.float32(scale); V *= np.float32(scale)
    ids, SG, SC = synth.hss_seeds(g, IC[0], khss=khss)
    with engine_for(g) as eng:
        eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
        cell, near, keep = eng.seed_locate(SG, SC, IC[0])
    # seeds sit on T-points: the nearest point is that T-point and the containing cell is its own
    jj, ii = np.where(g["tmask"][::khss, ::khss].astype(bool) & (g["latT"][::khss, ::khss] >= 55.) & (IC[0][::khss, ::khss] >= 0.9))
    assert np.array_equal(near[:, 0], jj * khss) and np.array_equal(near[:, 1], ii * khss)
    ik = np.flatnonzero(keep)
    pos0, cell0 = SC[ik], cell[ik]
    assert np.array_equal(cell0, near[ik])
    ref = corc.track(g, U, V, IC, pos0, cell0.astype(np.int64))
    for multi in (False, True):
        yx, ll, mk, na, cells, alive = _run_engine(torch, g, U, V, IC, pos0, cell0, multi=multi)
        assert np.array_equal(yx, ref["posC"][1:]) and np.array_equal(mk, ref["mask"][1:])
        assert np.array_equal(na, ref["nalive"])
        assert np.array_equal(cells[-1], ref["jiT"]) and np.array_equal(alive[-1], ref["alive"])
        if not multi:
            assert np.array_equal(cells, ref["jiT_hist"]) and np.array_equal(alive, ref["alive_hist"])
        assert np.abs(ll - ref["posG"][1:]).max() < LATLON_TOL_DEG
    assert ref["ncross"] > 0 and ref["alive"].sum() < ik.size or nrec < 48


def test_seed_locate_scattered_vs_oracle(corc):
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS["small"], seed=4)
    _, _, IC = synth.make_records(g, 1, seed=5)
    ids, SG, SC = synth.scattered_seeds(g, 3000, seed=6)
    jiT, keep, near = corc.seed_init(SG, SC, g["latT"], g["lonT"], g["Yf"], g["Xf"], g["ResKM"], g["tmask"], IC[0])
    with engine_for(g) as eng:
        eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
        cell, near_g, keep_g = eng.seed_locate(SG, SC, IC[0])
    assert np.array_equal(near_g, near)
    assert np.array_equal(keep_g, keep)
    ik = np.flatnonzero(keep)
    assert np.array_equal(cell[ik], jiT[ik]) and 0 < ik.size < 3000


def test_projection_kernels(sit, corc):
    rng = np.random.default_rng(8)
    yx = rng.uniform(-4500, 4500, (5000, 2))
    yx[0] = [0.0, 0.0]; yx[1] = [-9999.0, -9999.0]
    ll = sit.CartNPSkm2Geo1D(yx)
    assert np.abs(ll - corc.inv_stere(yx)).max() < LATLON_TOL_DEG
    back = sit.Geo2CartNPSkm1D(ll[2:])
    assert np.abs(back - yx[2:]).max() < 1e-6                        # km
    Y, X = sit.ConvertGeo2CartesianNPSkm(ll[2:, 0].reshape(2, -1), ll[2:, 1].reshape(2, -1))
    assert Y.shape == (2, 2499) and np.abs(Y.ravel() - yx[2:, 0]).max() < 1e-6


def test_empty_and_errors(sit, gold_track):
    T, g = gold_track
    with engine_for(g) as eng:
        eng.set_buoys(np.zeros((0, 2)), np.zeros((0, 2), np.int32))
        assert eng.track_record_host(0, T["U"][0], T["V"][0], T["IC"][0]) == 0
        with pytest.raises(sit.SitrackCudaError):
            eng.step(5, 0)                                          # slot out of range
        with pytest.raises(sit.SitrackCudaError):
            eng.seed_locate(np.zeros((1, 2)), np.zeros((1, 2)), T["IC"][0])   # no locate grid yet
    assert sit.IsInsideQuadrangleBatch(np.zeros((0, 2)), np.zeros((0, 4, 2))).shape == (0,)


# ---- the tuned kernel's shortcuts ---------------------------------------------------------------------

def test_div1000_is_ieee_division(sit):
    """dx/1000 (si3_part_tracker.py:457-458) via reciprocal + exact residual == IEEE division, on the
    operands velocities produce (f4 x 3600) and on adversarial near-midpoint quotients."""
    import ctypes as C
    from sitrack_b200 import _lib
    rng = np.random.default_rng(21)
    u = np.concatenate([rng.uniform(-2, 2, 2_000_000).astype(np.float32),
                        rng.standard_normal(1_000_000).astype(np.float32) * np.float32(1e-3),
                        np.float32([0, -0.0, 1e-30, 3.4e38, -3.4e38, 1.4e-45])])
    a = u.astype(np.float64) * 3600.0
    q = rng.uniform(1, 2, 1_000_000)                               # quotients; operands next to 1000*(q + ulp/2)
    mid = (q + np.spacing(q) / 2).astype(np.longdouble) * np.longdouble(1000)
    hard = np.concatenate([np.nextafter(mid.astype(np.float64), s) for s in (-np.inf, np.inf)] + [mid.astype(np.float64)])
    a = np.ascontiguousarray(np.concatenate([a, hard, rng.uniform(-1e6, 1e6, 1_000_000)]))
    qf = np.empty_like(a); qd = np.empty_like(a)
    _lib.check(_lib.lib().st_selftest_div1000(0, a.size, a.ctypes.data, qf.ctypes.data, qd.ctypes.data))
    assert np.array_equal(qd, a / 1000.0)                          # device division is IEEE
    assert np.array_equal(qf, qd)                                  # and so is the shortcut


@pytest.mark.parametrize("case", ["far", "mangled"])
def test_orientation_filter_fallbacks(torch, corc, gold_track, case):
    """The default kernel certifies "inside" with an orientation filter and leaves everything else to the
    reference's own test.  The filter must switch itself off where its error bound does not hold:
    'far'     -- coordinates beyond 2^17 km (whole grid: st_create clears filter_ok);
    'mangled' -- cells that are not convex anticlockwise quadrangles (per cell: k_cell_bits bit 2).
    Either way trajectories, cells and alive flags stay those of the C oracle, bit for bit."""
    T, g0 = gold_track
    g = {k: np.array(v, copy=True) for k, v in g0.items()}
    pos0 = T["pos0"].copy()
    if case == "far":
        for k in ("Xf", "Xu", "Xv", "Xt"):
            if k in g:
                g[k] = g[k] + 262144.0                         # 2^18 km east: exactly representable shift
        pos0[:, 1] += 262144.0
    else:
        rng = np.random.default_rng(11)
        Nj, Ni = g["tmask"].shape
        jj, ii = rng.integers(3, Nj - 3, 60), rng.integers(3, Ni - 3, 60)
        # drag corner points far enough to fold their four cells (concave, some clockwise)
        g["Yf"][jj, ii] += rng.uniform(-1.2, 1.2, 60) * 12.5
        g["Xf"][jj, ii] += rng.uniform(-1.2, 1.2, 60) * 12.5
    nrec = 30
    U, V, IC = 3 * T["U"][:nrec], 3 * T["V"][:nrec], T["IC"][:nrec]
    ref = corc.track(g, U, V, IC, pos0, T["jiT0"].astype(np.int64), do_latlon=False)
    assert ref["ncross"] > 100
    dev = torch.device("cuda", 0)
    nP = pos0.shape[0]
    for variant in (0, 3, 4, 5):
        with engine_for(g) as eng:
            eng.set_kernel_variant(variant)
            eng.set_buoys(pos0, T["jiT0"])
            eng.record_slots(1)
            yx = torch.empty((nP, 2), dtype=torch.float64, device=dev)
            mk = torch.empty((nP,), dtype=torch.int8, device=dev)
            for k in range(nrec):
                st = eng.staging(0)
                torch.cuda.synchronize()
                st[0], st[1], st[2] = U[k], V[k], IC[k]
                eng.submit_record(0)
                eng.step(0, k, yx, None, mk, None)
                torch.cuda.synchronize()
                assert np.array_equal(yx.cpu().numpy(), ref["posC"][k + 1]), (case, variant, k)
                assert np.array_equal(mk.cpu().numpy(), ref["mask"][k + 1])
            p, c, a = eng.get_state()
            assert np.array_equal(c, ref["jiT"]) and np.array_equal(a, ref["alive"])


@pytest.mark.parametrize("preset,n,nrec,scale", [("nanuk4", 200_000, 30, 2.0), ("arctic12", 400_000, 16, 1.0)])
def test_large_cloud_tuned_vs_v1_vs_oracle(torch, corc, preset, n, nrec, scale):
    """Dense clouds (many buoys per warp leaving their cell, kills, domain edges): the tuned kernel,
    the straightforward v1 kernel and the C oracle agree bit for bit."""
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS[preset], seed=0)
    U, V, IC = synth.make_records(g, nrec, seed=1)
    U *= np.float32(scale); V *= np.float32(scale)
    ids, SG, SC = synth.dense_seeds(g, n, IC[0], seed=9)
    with engine_for(g) as eng:
        eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
        cell, near, keep = eng.seed_locate(SG, SC, IC[0])
    ik = np.flatnonzero(keep)
    pos0, cell0 = SC[ik], cell[ik]
    ref = corc.track(g, U, V, IC, pos0, cell0.astype(np.int64), history=False)
    assert ref["ncross"] > 0.02 * ik.size * nrec and ref["alive"].sum() < ik.size
    dev = torch.device("cuda", 0)
    pos_end = None
    for variant, chain in ((0, False), (1, False), (3, False), (4, False), (5, False), (0, True), (5, True)):
        with engine_for(g) as eng:
            eng.set_kernel_variant(variant)
            eng.set_buoys(pos0, cell0)
            eng.set_row_chain(chain)                                # chained: the ONE row buffer below is stepped in place
            eng.record_slots(1)
            yx = torch.empty((ik.size, 2), dtype=torch.float64, device=dev)
            ll = torch.empty((ik.size, 2), dtype=torch.float64, device=dev)
            mk = torch.empty((ik.size,), dtype=torch.int8, device=dev)
            na = torch.zeros((nrec,), dtype=torch.int64, device=dev)
            for k in range(nrec):
                st = eng.staging(0)
                st[0], st[1], st[2] = U[k], V[k], IC[k]
                eng.submit_record(0)
                eng.step(0, k, yx, ll, mk, na[k:k + 1])
                torch.cuda.synchronize()
                if k in (0, nrec // 2, nrec - 1):
                    assert np.array_equal(yx.cpu().numpy(), ref["posC"][k + 1])
                    assert np.array_equal(mk.cpu().numpy(), ref["mask"][k + 1])
                    assert np.abs(ll.cpu().numpy() - ref["posG"][k + 1]).max() < LATLON_TOL_DEG
            p, c, a = eng.get_state()
            assert np.array_equal(c, ref["jiT"]) and np.array_equal(a, ref["alive"])
            assert np.array_equal(na.cpu().numpy(), ref["nalive"])
            if pos_end is None:                                     # the oracle's last recorded position of every buoy
                assert ref["mask"][0].all()
                last = ref["mask"].shape[0] - 1 - np.argmax(ref["mask"][::-1] == 1, axis=0)
                pos_end = ref["posC"][last, np.arange(ik.size)]
            assert np.array_equal(p, pos_end), (variant, chain)     # final positions, discontinued buoys included


# ---- full season (BASELINE config 2) -------------------------------------------------------------------

@pytest.mark.parametrize("path", ["season", "per_record"])
def test_full_season_config2_vs_oracle(torch, corc, path):
    """NANUK4-shaped grid, HSS5 seeding (~1k buoys), 3024 hourly records (1996-12-15 -> 1997-04-20), -F:
    the chunked season path (k_advect_multi, 126 records per launch) and the per-record pipeline (the default
    step kernel k_advect_warp, one launch per record, rows streamed back) against the C oracle, record
    by record.  Target of the north star: >= 95 % of buoys bit-exact in cell index over the season;
    measured: 100 %, with bit-identical f8 positions (accumulated divergence 0 km)."""
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS["nanuk4"], seed=0)
    nrec, chunk = 3024, 126
    _, _, IC0 = synth.make_records(g, 1, seed=1)
    ids, SG, SC = synth.hss_seeds(g, IC0[0], khss=5)
    with engine_for(g) as eng:
        eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
        cell, near, keep = eng.seed_locate(SG, SC, IC0[0])
        ik = np.flatnonzero(keep)
        pos0, cell0 = SC[ik], cell[ik]
        assert 900 < ik.size < 3000
        eng.set_buoys(pos0, cell0)
        cache = {}

        def records(k):                                   # generated chunk-wise, shared with the oracle
            c = k // chunk
            if c not in cache:
                cache.clear()
                cache[c] = synth.make_records(g, chunk, seed=1000 + c, k0=c * chunk)
            U, V, IC = cache[c]
            return U[k % chunk], V[k % chunk], IC[k % chunk]
        r = eng.track(records, nrec, pos0=pos0, chunk=chunk if path == "season" else None)
        p_end, c_end, a_end = eng.get_state()
    pos, ji, alive = pos0.copy(), cell0.astype(np.int64), np.ones(ik.size, np.int8)
    same_cell = np.ones(ik.size, bool)
    ncross = 0
    for c in range(nrec // chunk):
        U, V, IC = synth.make_records(g, chunk, seed=1000 + c, k0=c * chunk)
        ref = corc.track(g, U, V, IC, pos, ji, alive0=alive, history=False)
        rows = slice(c * chunk + 1, (c + 1) * chunk + 1)
        alive_rows = ref["mask"][1:] == 1
        assert np.array_equal(r["mask"][rows], ref["mask"][1:])
        assert np.array_equal(r["posC"][rows], ref["posC"][1:])                  # bit-exact, fill rows included
        assert np.abs(r["posG"][rows] - ref["posG"][1:]).max() < LATLON_TOL_DEG
        assert np.array_equal(r["n_alive"][c * chunk:(c + 1) * chunk], ref["nalive"])
        ncross += ref["ncross"]
        # carry the oracle's state: dead buoys keep their last recorded position out of the loop
        last = np.where(alive_rows.any(axis=0), alive_rows.shape[0] - 1 - np.argmax(alive_rows[::-1], axis=0), -1)
        for b in np.flatnonzero(last >= 0):
            pos[b] = ref["posC"][1 + last[b], b]
        ji, alive = ref["jiT"], ref["alive"]
    same_cell &= (c_end == ji).all(axis=1)
    assert np.array_equal(a_end, alive)
    assert same_cell.mean() == 1.0                                             # north star asks for >= 0.95
    assert ncross > 10 * ik.size and alive.sum() < ik.size                      # the season is eventful


def test_div_core_is_ieee_division(sit):
    """The branch-free division of the inside test (xints, locate.py:72) equals IEEE division: random
    operands over the magnitudes km coordinates produce, zeros, signs, and quotients built to sit next
    to rounding midpoints."""
    from sitrack_b200 import _lib
    rng = np.random.default_rng(33)
    n = 3_000_000
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-14, 9, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-13, 5, n)
    a[:1000] = 0.0; a[1000:2000] = -0.0
    # adversarial: choose b and a quotient q, then a = RN(b * (q + ulp/2)) and its neighbours
    q = rng.uniform(1, 2, 1_000_000) * rng.choice([-1.0, 1.0], 1_000_000)
    bb = rng.uniform(1, 2, 1_000_000) * 10.0 ** rng.integers(-6, 5, 1_000_000)
    mid = ((np.abs(q) + np.spacing(np.abs(q)) / 2) * np.sign(q)).astype(np.longdouble) * bb.astype(np.longdouble)
    mids = mid.astype(np.float64)
    aa = np.concatenate([mids, np.nextafter(mids, np.inf), np.nextafter(mids, -np.inf)])
    a = np.ascontiguousarray(np.concatenate([a, aa]))
    b = np.ascontiguousarray(np.concatenate([b, bb, bb, bb]))
    keep = b != 0
    a, b = np.ascontiguousarray(a[keep]), np.ascontiguousarray(b[keep])
    qf = np.empty_like(a); qd = np.empty_like(a)
    _lib.check(_lib.lib().st_selftest_divide(0, a.size, a.ctypes.data, b.ctypes.data, qf.ctypes.data, qd.ctypes.data))
    assert np.array_equal(qd, a / b)
    assert np.array_equal(qf, qd)
    # the only bit-level difference allowed: the sign of a zero quotient when a = -0 (the kernel's
    # numerator is (y-y1)*(x2-x1) with y-y1 > 0, never -0; and x <= xints cannot see the sign of zero)
    nz = qd != 0
    assert np.array_equal(np.signbit(qf[nz]), np.signbit(qd[nz])) and (qf[~nz] == 0).all()


# ---- full BASELINE size: size-independent properties ----------------------------------------------------

def test_properties_at_full_size(torch, sit, corc):
    """12.5 M buoys on the 1/12-degree-class grid (config 5's per-GPU share), where the oracle is too
    slow to run everything: (1) a zero-velocity record is the identity on positions and cells and
    leaves everyone alive, (2) two runs are bit-identical (no atomics or ordering in the results),
    (3) tracking the two halves of the cloud separately equals tracking it whole (buoys never
    interact: the sharding property the multi-GPU path rests on), (4) alive counts never grow and
    match the row masks, (5) lat/lon rows invert back to the km rows within 1e-6 km, (6) a random
    20 k-buoy sub-sample matches the C oracle bit for bit over all records."""
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS["arctic12"], seed=0)
    nrec = 6
    U, V, IC = synth.make_records(g, nrec, seed=1)
    ids, _, SC = synth.dense_seeds(g, 12_500_000, IC[0], seed=3, with_latlon=False)
    dev = torch.device("cuda", 0)
    SC_t = torch.from_numpy(SC).to(dev)
    SG_t = torch.empty_like(SC_t)
    with engine_for(g) as eng:
        eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
        sit._lib.check(eng.L.st_xy2latlon_dev(SC_t.shape[0], SC_t.data_ptr(), SG_t.data_ptr(), 70.0, -45.0, None))
        SG_t[:, 1] = torch.remainder(SG_t[:, 1], 360.0)
        cell_t, near_t, keep_t = eng.seed_locate_dev(SG_t, SC_t, torch.from_numpy(IC[0]).to(dev))
        kp = keep_t.bool()
        pos0, cell0 = SC_t[kp].contiguous(), cell_t[kp].contiguous()
        nP = pos0.shape[0]
        assert nP > 12_000_000
        eng.record_slots(nrec + 1)
        for k in range(nrec):
            st = eng.staging(k); st[0], st[1], st[2] = U[k], V[k], IC[k]; eng.submit_record(k)
        st = eng.staging(nrec); st[0], st[1] = 0.0, 0.0; st[2] = IC[0]; eng.submit_record(nrec)
        torch.cuda.synchronize()

        def run(p0, c0, recs):
            n = p0.shape[0]
            eng.set_buoys_dev(p0, c0)
            yx = torch.empty((len(recs), n, 2), dtype=torch.float64, device=dev)
            ll = torch.empty((len(recs), n, 2), dtype=torch.float64, device=dev)
            mk = torch.empty((len(recs), n), dtype=torch.int8, device=dev)
            na = torch.zeros((len(recs),), dtype=torch.int64, device=dev)
            for i, r in enumerate(recs):
                eng.step(r, i, yx[i], ll[i], mk[i], na[i:i + 1])
            torch.cuda.synchronize()
            _, c, a = eng.get_state()
            return yx, ll, mk, na.cpu().numpy(), c, a

        # (1) identity under zero velocity
        yx, ll, mk, na, c, a = run(pos0, cell0, [nrec])
        assert torch.equal(yx[0], pos0) and bool((mk[0] == 1).all()) and na[0] == nP
        assert np.array_equal(c, cell0.cpu().numpy()) and a.all()
        del yx, ll, mk
        # (2) determinism, (4) alive counts
        recs = list(range(nrec))
        yx1, ll1, mk1, na1, c1, a1 = run(pos0, cell0, recs)
        yx2, ll2, mk2, na2, c2, a2 = run(pos0, cell0, recs)
        assert torch.equal(yx1, yx2) and torch.equal(ll1, ll2) and torch.equal(mk1, mk2)
        assert np.array_equal(c1, c2) and np.array_equal(a1, a2) and np.array_equal(na1, na2)
        del yx2, ll2, mk2
        assert (np.diff(na1) <= 0).all() and na1[0] == nP and na1[-1] < nP
        assert np.array_equal(mk1.sum(dim=1).cpu().numpy(), na1)          # row masks = buoys alive at record start
        assert int(a1.sum()) <= na1[-1]
        # (5) projection round trip on the last row (alive buoys)
        sel = mk1[-1] == 1
        llsel = ll1[-1][sel].contiguous()
        yx_back = sit.Geo2CartNPSkm1D(llsel[:200_000].cpu().numpy())
        assert np.abs(yx_back - yx1[-1][sel][:200_000].cpu().numpy()).max() < 1e-6
        # (3) shard invariance: halves tracked separately == whole
        h = (nP // 2 // 256) * 256
        yxa, _, mka, naa, ca, aa = run(pos0[:h].contiguous(), cell0[:h].contiguous(), recs)
        assert torch.equal(yxa, yx1[:, :h]) and torch.equal(mka, mk1[:, :h]) and np.array_equal(ca, c1[:h])
        yxb, _, mkb, nab, cb, ab = run(pos0[h:].contiguous(), cell0[h:].contiguous(), recs)
        assert torch.equal(yxb, yx1[:, h:]) and np.array_equal(cb, c1[h:]) and np.array_equal(naa + nab, na1)
        del yxa, yxb, mka, mkb
        # (6) sub-sample against the oracle
        rng = np.random.default_rng(5)
        sub = np.sort(rng.choice(nP, 20_000, replace=False))
        p_sub, c_sub = pos0.cpu().numpy()[sub], cell0.cpu().numpy()[sub]
        ref = corc.track(g, U, V, IC, p_sub, c_sub.astype(np.int64), history=False)
        sub_t = torch.from_numpy(sub).to(dev)
        assert np.array_equal(yx1[:, sub_t].cpu().numpy(), ref["posC"][1:])
        assert np.array_equal(mk1[:, sub_t].cpu().numpy(), ref["mask"][1:])
        assert np.array_equal(c1[sub], ref["jiT"]) and np.array_equal(a1[sub], ref["alive"])
        got_ll = ll1[:, sub_t].cpu().numpy()
        assert np.abs(got_ll - ref["posG"][1:]).max() < LATLON_TOL_DEG
        # SURVEY 8d parity report: fraction of lat/lon values equal after the f4 rounding of the output file
        assert (got_ll.astype(np.float32) == ref["posG"][1:].astype(np.float32)).mean() > 0.9999


def test_fcc_golden(sit, gold_seed):
    """FCC (locate.py:139-218, geographic variant, not on the tracker's path) against the reference's
    answers: nearest T on the device, local spherical projection, TheCell on the 5x5 box."""
    S, g = gold_seed
    import contextlib, io
    for n, k in enumerate(S["fcc_idx"][:25]):
        with contextlib.redirect_stdout(io.StringIO()):
            ji, vr = sit.FCC((S["SG"][k, 0], S["SG"][k, 1]), g["latT"], g["lonT"], g["latF"], g["lonF"], cellType='T',
                             rd_found_km=2.5, resolkm=g["ResKM"], max_itr=10)
        assert list(ji) == list(S["fcc_ji"][n]) and np.array_equal(np.asarray(vr), S["fcc_vrt"][n])


def test_fast_projection_all_latitudes(sit, corc):
    """The step kernel's own inverse (polynomial latitude for t <= 1/2, table-driven angles, one-Newton
    rcp/rsqrt) against the PROJ-style iteration: Arctic, mid-latitudes, southern hemisphere, the pole,
    the axes and the fill point."""
    from sitrack_b200 import _lib
    rng = np.random.default_rng(44)
    yx = np.concatenate([rng.uniform(-4500, 4500, (200_000, 2)),           # lat >~ 35N
                         rng.uniform(-16000, 16000, (100_000, 2)),         # down to the southern hemisphere
                         rng.uniform(-50, 50, (20_000, 2)),                # around the pole
                         [[0.0, 0.0], [-9999.0, -9999.0], [0.0, 123.0], [123.0, 0.0], [-77.0, 0.0], [0.0, -5.0]]])
    yx = np.ascontiguousarray(yx)
    ll = np.empty_like(yx)
    _lib.check(_lib.lib().st_selftest_xy2latlon_fast(0, yx.shape[0], yx.ctypes.data, ll.ctypes.data, 70.0, -45.0))
    want = corc.inv_stere(yx)
    dlon = np.abs(ll[:, 1] - want[:, 1]); dlon = np.minimum(dlon, 360.0 - dlon)      # +-180 are the same meridian
    assert np.abs(ll[:, 0] - want[:, 0]).max() < LATLON_TOL_DEG
    near_pole = np.hypot(yx[:, 0], yx[:, 1]) < 1e-6
    assert dlon[~near_pole].max() < LATLON_TOL_DEG
    assert ll[-6, 0] == 90.0 and ll[-6, 1] == -45.0                                     # the pole: lam = lon0 like PROJ


def test_config3_end_to_end_vs_oracle(torch, corc):
    """BASELINE config 3 as a config: NANUK4-shaped grid, dense HSS1 seeding (~25 k buoys) plus 1000 SIDFEX-style
    scattered seeds (some on land / outside the domain) -> SeedInit -> 24 hourly records, against the C oracle:
    kept seeds, nearest points and host cells identical, then cells / alive flags / f8 positions bit-exact."""
    import synth
    import sitrack_b200 as sit
    g = synth.make_grid(**synth.GRID_PRESETS["nanuk4"], seed=0)
    nrec = 24
    U, V, IC = synth.make_records(g, nrec, seed=1)
    i1, G1, C1 = synth.hss_seeds(g, IC[0], khss=1)
    i2, G2, C2 = synth.scattered_seeds(g, 1000, seed=2)
    ids, SG, SC = np.concatenate([i1, i2 + i1.size]), np.concatenate([G1, G2]), np.concatenate([C1, C2])
    assert 20_000 < i1.size < 100_000                               # every ocean T-point under ice, lat >= 55
    jiT_o, keep_o, near_o = corc.seed_init(SG, SC, g["latT"], g["lonT"], g["Yf"], g["Xf"], g["ResKM"], g["tmask"], IC[0])
    nPk, SGk, SCk, IDk, jiT, VRT, iKeep = sit.SeedInit(ids, SG, SC, g["latT"], g["lonT"], g["Yf"], g["Xf"], g["ResKM"],
                                                        g["tmask"], xIceConc=IC[0].astype(np.float64))
    ik_o = np.flatnonzero(keep_o)
    assert np.array_equal(iKeep, ik_o) and np.array_equal(jiT, jiT_o[ik_o])
    assert 0 < (keep_o[i1.size:] == 0).sum() < 1000                   # scattered seeds do get dropped, not all of them
    ref = corc.track(g, U, V, IC, SCk, jiT.astype(np.int64))
    with engine_for(g) as eng:
        eng.set_buoys(SCk, jiT)
        r = eng.track((U, V, IC), nrec, pos0=SCk)
        p, c, a = eng.get_state()
    assert np.array_equal(r["posC"], ref["posC"]) and np.array_equal(r["mask"], ref["mask"])
    assert np.array_equal(r["n_alive"], ref["nalive"])
    assert np.array_equal(c, ref["jiT"]) and np.array_equal(a, ref["alive"])
    assert np.abs(r["posG"][1:] - ref["posG"][1:]).max() < 1e-9


def test_seeding_subsample_on_the_twelfth_degree_grid(torch, corc):
    """Seeding at 1/12 degree against the reference's whole-grid search (the oracle scans all 2.5 M T-points per
    seed): 1500 seeds of the cfg4/cfg5 cloud + 500 scattered ones; nearest point, keep mask and host cell equal."""
    import synth
    g = synth.make_grid(**synth.GRID_PRESETS["arctic12"], seed=0)
    U, V, IC = synth.make_records(g, 1, seed=1)
    ids, SG, SC = synth.dense_seeds(g, 200_000, IC[0], seed=3)
    sel = np.sort(np.random.default_rng(4).choice(SC.shape[0], 1500, replace=False))
    i2, G2, C2 = synth.scattered_seeds(g, 500, seed=9)
    SG, SC = np.concatenate([SG[sel], G2]), np.concatenate([SC[sel], C2])
    jiT_o, keep_o, near_o = corc.seed_init(SG, SC, g["latT"], g["lonT"], g["Yf"], g["Xf"], g["ResKM"], g["tmask"], IC[0])
    with engine_for(g) as eng:
        eng.set_locate_grid(g["latT"], g["lonT"], g["ResKM"])
        cell, near, keep = eng.seed_locate(SG, SC, IC[0])
    assert np.array_equal(near, near_o) and np.array_equal(keep, keep_o)
    ik = np.flatnonzero(keep)
    assert np.array_equal(cell[ik], jiT_o[ik]) and ik.size > 1500


def test_seed_locate_near_ties_are_rechecked_on_the_host(torch):
    """Seeds placed (by bisection with numpy's Haversine) where two T-points are equidistant to ~1e-15, and seeds at
    the acceptance radius 0.5 res 1.2^7 to ~1e-15: the device flags them and the host decides them with numpy, the
    reference's own arithmetic (locate.py:253-266, util.py:85-103).  The final nearest points must equal numpy's
    first-minimum argmin / acceptance over the neighbourhood, whichever way the device's last ulp fell."""
    import synth
    from sitrack_b200.locate import _haversine_host
    g = synth.make_grid(**synth.GRID_PRESETS["small"], seed=0)
    U, V, IC = synth.make_records(g, 1, seed=1)
    la, lo, res = g["latT"], g["lonT"], g["ResKM"]
    rng = np.random.default_rng(3)
    SG = []
    for t in range(300):                                         # argmin near-ties between T[j,i] and T[j,i+1]
        j = rng.integers(5, g["Nj"] - 6); i = rng.integers(5, g["Ni"] - 6)
        a, b = np.array([la[j, i], lo[j, i]]), np.array([la[j, i + 1], lo[j, i + 1]])
        if abs(a[1] - b[1]) > 90:
            continue
        x0, x1 = 0.3, 0.7
        f = lambda x: (lambda p: float(_haversine_host(p[0], p[1], a[0], a[1]) - _haversine_host(p[0], p[1], b[0], b[1])))(a + x * (b - a))
        if f(x0) * f(x1) > 0:
            continue
        for _ in range(80):
            xm = 0.5 * (x0 + x1)
            if f(x0) * f(xm) <= 0: x1 = xm
            else: x0 = xm
        SG.append(a + x0 * (b - a))
    n_tie = len(SG)
    for t in range(200):                                         # acceptance near the limit: move away from a T-point of the first row
        i = rng.integers(5, g["Ni"] - 6)
        a = np.array([la[0, i], lo[0, i]]); inward = np.array([la[1, i], lo[1, i]])
        r_lim = 0.5 * res[0, i]
        for _ in range(7):
            r_lim = 1.2 * r_lim
        dirn = a - inward                                        # outward, in (lat,lon) degrees
        x0, x1 = 0.0, 6.0
        f = lambda x: float(_haversine_host(a[0] + x * dirn[0], a[1] + x * dirn[1], a[0], a[1]) - r_lim)
        if f(x1) < 0:
            continue
        for _ in range(80):
            xm = 0.5 * (x0 + x1)
            if f(xm) <= 0: x0 = xm
            else: x1 = xm
        SG.append(a + (x0 if t % 2 else x1) * dirn)
    SG = np.array(SG)
    SG[:, 1] = np.mod(SG[:, 1], 360.0)
    SC = np.zeros_like(SG)                                        # containing cell not under test here
    with engine_for(g) as eng:
        eng.set_locate_grid(la, lo, res)
        st = {}
        cell, near, keep = eng.seed_locate(SG, SC, IC[0], stats=st)
        cell0, near0, keep0 = eng.seed_locate(SG, SC, IC[0], recheck=False)
    assert st["flagged"] >= 0.8 * SG.shape[0], st                 # the constructed seeds are recognised as undecidable
    # numpy's answer: first-minimum argmin over the whole grid, then the reference's ladder
    for p in range(SG.shape[0]):
        d = _haversine_host(SG[p, 0], SG[p, 1], la, lo)
        k = int(np.argmin(d)); jy, jx = divmod(k, la.shape[1])
        rf = 0.5 * res[jy, jx]
        for _ in range(7):
            rf = 1.2 * rf
        want = (jy, jx) if d[jy, jx] < rf else (-1, -1)
        assert (int(near[p, 0]), int(near[p, 1])) == want, (p, near[p], near0[p], want)
    assert n_tie > 100


def test_nearest_point_with_previous_guess_golden(sit, gold_seed):
    """NearestPoint(ji_prv=..., np_box_r=...) (locate.py:241-244,255-256; FCC forwards it) against the reference's
    answers on 648 cases (tests/golden/nearest_box.npz, generated by oracle/make_golden.py from the unmodified
    reference): guesses near and far from the true nearest point, boxes of half-width 3 and 10, seeds that end as
    (-1,-1)."""
    import contextlib, io
    S, g = gold_seed
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "nearest_box.npz"))
    sel = np.arange(0, G["pt"].shape[0], 3)                        # every third case: each one is its own launch
    n_box = 0
    for k in sel:
        with contextlib.redirect_stdout(io.StringIO()):
            got = sit.NearestPoint(tuple(G["pt"][k]), g["latT"], g["lonT"], rd_found_km=2.5, resolkm=g["ResKM"],
                                   ji_prv=tuple(int(v) for v in G["ji_prv"][k]), np_box_r=int(G["np_box_r"][k]), max_itr=10)
        assert tuple(got) == tuple(int(v) for v in G["out"][k]), (k, got, G["out"][k])
    assert (G["out"][sel][:, 0] < 0).sum() > 5 and (G["out"][sel][:, 0] >= 0).sum() > 100
