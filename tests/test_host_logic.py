"""CPU: host-side logic of the drop-in surface (no GPU arithmetic involved), checked against
the reference's own functions when /root/reference is present, and against pinned values."""
import contextlib
import io

import numpy as np
import pytest

import sitrack_b200 as sit
from oracle import ref_loader


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_constants():
    assert sit.rmin_conc == 0.1 and sit.rFoundKM == 2.5 and sit.FillValue == -9999.
    assert sit.tunits_default == 'seconds since 1970-01-01 00:00:00'


def test_epoch_clock():
    assert sit.epoch2clock(850608000) == "1996-12-15_00:00:00"
    assert sit.epoch2clock(850608000, precision='h') == "1996-12-15_00"
    assert sit.clock2epoch("1996-12-15_00:00:00") == 850608000
    assert sit.clock2epoch("1997-04-20", precision='D', cfrmt='guess') == 861494400


def test_updt_ind_in_place():
    for k, (dj, di) in {1: (-1, 0), 2: (0, 1), 3: (1, 0), 4: (0, -1), 5: (-1, -1), 6: (-1, 1), 7: (1, 1), 8: (1, -1)}.items():
        V = np.array([[9, 9, 10, 10], [19, 20, 20, 19]]); ji = np.array([10, 20])
        V2, ji2 = sit.UpdtInd4NewCell(k, V, ji)
        assert V2 is V and ji2 is ji                               # mutated in place, like the reference
        assert list(ji) == [10 + dj, 20 + di]
        assert np.array_equal(V, [[9 + dj, 9 + dj, 10 + dj, 10 + dj], [19 + di, 20 + di, 20 + di, 19 + di]])
    with pytest.raises(SystemExit):
        quiet(sit.UpdtInd4NewCell, 9, np.zeros((2, 4), int), np.zeros(2, int))


def test_get_time_span_pinned():
    t = 850608000 + 1800 + 3600 * np.arange(100)
    Nt, k0, kN, a, b = quiet(sit.GetTimeSpan, 3600., t, 850608000, t[0], t[-1])
    assert (Nt, k0, kN) == (100, 0, 99)
    Nt, k0, kN, a, b = quiet(sit.GetTimeSpan, 3600., t, 850608000 + 7200, t[0], t[-1], iStop=850608000 + 36000)
    assert (k0, kN, Nt) == (2, 9, 8)                               # ties in argmin resolve to the first minimum


@pytest.mark.skipif(not ref_loader.available(), reason="upstream reference not present")
def test_against_reference_host_functions():
    ref = ref_loader.load()
    rng = np.random.default_rng(0)
    t = 850608000 + 1800 + 3600 * np.arange(240)
    for _ in range(50):
        seed = int(850608000 + rng.integers(0, 200) * 3600 + rng.integers(0, 3600))
        stop = int(seed + rng.integers(1, 30) * 3600) if rng.random() < 0.7 else None
        assert quiet(sit.GetTimeSpan, 3600., t, seed, t[0], t[-1], iStop=stop) == \
            quiet(ref.GetTimeSpan, 3600., t, seed, t[0], t[-1], iStop=stop)
    for k in range(1, 9):
        V = rng.integers(5, 50, (2, 4)); ji = rng.integers(5, 50, 2)
        a = sit.UpdtInd4NewCell(k, V.copy(), ji.copy()); b = ref.UpdtInd4NewCell(k, V.copy(), ji.copy())
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(sit.debugSeeding(), ref.debugSeeding())
    x = rng.uniform(0, 360, 20)
    assert np.allclose(sit.degE_to_degWE(x), ref.degE_to_degWE(x)) and sit.degE_to_degWE(270.) == ref.degE_to_degWE(270.)
    for it in (0, 850608000, 861494399):
        for p in "smhD":
            assert sit.epoch2clock(it, precision=p) == ref.epoch2clock(it, precision=p)
    msk = (rng.random((40, 30)) > 0.2).astype('i1'); lat = rng.uniform(50, 90, (40, 30)); lon = rng.uniform(0, 360, (40, 30))
    ic = rng.random((40, 30))
    for kh in (1, 3):
        assert np.array_equal(sit.nemoSeed(msk, lat, lon, ic, khss=kh), ref.nemoSeed(msk, lat, lon, ic, khss=kh))
    a = quiet(sit.nemoSeed, msk, lat, lon, ic, khss=1, platF=lat + 0.1, plonF=lon + 0.1)
    b = quiet(ref.nemoSeed, msk, lat, lon, ic, khss=1, platF=lat + 0.1, plonF=lon + 0.1)
    assert np.array_equal(a, b)
    import os
    dat = os.path.join(os.path.dirname(os.path.dirname(ref.__file__)), "tools", "sidfexloc.dat")
    ll, ids = quiet(sit.SidfexSeeding, dat); ll2, ids2 = quiet(ref.SidfexSeeding, dat)
    assert np.array_equal(ll, ll2) and np.array_equal(ids, ids2)


def test_row_dtype_selection_and_numa_binding_helpers():
    """Host plumbing added with the f4 rows and the multi-GPU paths: the row-buffer dtype picks the C entry
    point (mixing raises before any launch), and the NUMA binding degrades to None without NVML / a GPU."""
    import numpy as np
    from sitrack_b200.engine import _rows_f4
    from sitrack_b200.dist import bind_host_to_gpu, shard_bounds
    a8, a4 = np.zeros((3, 2)), np.zeros((3, 2), np.float32)
    assert _rows_f4(None, None) is False and _rows_f4(a8, None) is False and _rows_f4(a4, a4) is True
    with pytest.raises(TypeError):
        _rows_f4(a4, a8)
    assert bind_host_to_gpu(10 ** 6) is None                     # no such device: no binding, no exception
    b = shard_bounds(636, 2)                                     # the CLI's torchrun test: 512 + 124 buoys
    assert b.tolist() == [0, 512, 636]


def test_mesh_mask_with_levels_and_masked_time_pos(tmp_path, monkeypatch):
    """(a) a real NEMO mesh_mask stores tmask/fmask as (t,z,y,x) with z > 1 -- only the surface plane is
    read (reference ncio.py:28-36); (b) a `time_pos` with fill entries stays masked, so the seed file's time span
    skips them like the reference's np.min / np.max on the masked array (ncio.py:341)."""
    import sitrack_b200.ncio as ncio

    class V:
        def __init__(self, a, units=None):
            self.a, self.units, self.shape = a, units, a.shape

        def __getitem__(self, k):
            return self.a[k]

    class DS:
        def __init__(self, variables, dims):
            self.variables = variables
            self.dimensions = {k: type("D", (), {"size": v})() for k, v in dims.items()}

        def __enter__(self):
            return self

        def __exit__(self, *e):
            pass

    rng = np.random.default_rng(0)
    nz, Nj, Ni = 3, 5, 6
    tm = (rng.random((1, nz, Nj, Ni)) > 0.3).astype('i1')
    ds = DS({"tmask": V(tm), "glamt": V(rng.random((1, Nj, Ni)))}, {})
    assert np.array_equal(ncio._plane(ds.variables["tmask"]), tm[0, 0])
    assert np.array_equal(ncio._plane(ds.variables["glamt"]), ds.variables["glamt"].a[0])
    assert np.array_equal(ncio._plane(V(tm[0, 0])), tm[0, 0])

    t = np.array([1000, 4600], dtype='i4')
    tp = np.ma.masked_equal(np.array([[1000, 1200, -9999], [4600, -9999, 4000]], dtype='i4'), -9999)
    fake = DS({"time": V(t, ncio.tunits_default), "time_pos": V(tp, ncio.tunits_default)}, {"time": 2})
    monkeypatch.setattr(ncio, "open_dataset", lambda fn: fake)
    monkeypatch.setattr(ncio, "chck4f", lambda fn: None)
    Nt, t1d, t2d = ncio.LoadNCtime("x.nc", ltime2d=True)
    assert Nt == 2 and np.ma.is_masked(t2d)
    assert int(np.min(t2d)) == 1000 and int(np.max(t2d)) == 4600
    first, last, name, batch, _ = ncio.SeedFileTimeInfo("a_b_c_d.nc", ltime2d=True)
    assert first == 0 and last == 7200                             # floor / ceil to the hour of 1000 .. 4600, not of -9999


@pytest.mark.skipif(not ref_loader.available(), reason="upstream reference not present")
def test_near_tie_recheck_is_the_reference_nearest_point():
    """The host re-evaluation of flagged seeds (sitrack_b200/locate.py:_recheck_nearest, _haversine_host) against
    the reference's NearestPoint / Haversine: same distances bit for bit, same (jy,jx) or (-1,-1), including seeds
    at the acceptance limit 0.5 res 1.2^7 and beyond it."""
    import synth
    from sitrack_b200.locate import _recheck_nearest, _haversine_host
    ref = ref_loader.load()
    g = synth.make_grid(**synth.GRID_PRESETS["small"], seed=0)
    rng = np.random.default_rng(0)
    n_found = n_lost = 0
    for t in range(200):
        j = rng.integers(2, g["Nj"] - 3); i = rng.integers(2, g["Ni"] - 3)
        sc = rng.choice([0.3, 1.0, 1.75, 1.79, 1.8, 2.5])
        lat = g["latT"][j, i] + sc * 0.05 * rng.normal(); lon = g["lonT"][j, i] + sc * 0.2 * rng.normal()
        if t % 2:                                                   # outside the grid, 1.5 .. 3 cells beyond its first row
            x = rng.choice([1.5, 1.75, 1.79, 1.8, 2.0, 3.0])
            lat = g["latT"][0, i] + x * (g["latT"][0, i] - g["latT"][1, i])
            lon = g["lonT"][0, i] + x * (g["lonT"][0, i] - g["lonT"][1, i])
        want = quiet(ref.NearestPoint, (lat, lon), g["latT"], g["lonT"], rd_found_km=2.5, resolkm=g["ResKM"], max_itr=10)
        d = ref.Haversine(lat, lon, g["latT"], g["lonT"])
        assert np.array_equal(_haversine_host(lat, lon, g["latT"], g["lonT"]), d)
        k = int(np.argmin(d))
        d2 = d.copy().reshape(-1); d2[k] = np.inf
        got = _recheck_nearest((lat, lon), k, int(np.argmin(d2)), g["latT"], g["lonT"], g["ResKM"])
        assert tuple(int(x) for x in want) == got
        n_found += got[0] >= 0; n_lost += got[0] < 0
    assert n_found > 20 and n_lost > 10
