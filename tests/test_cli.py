"""The drop-in command line: CPU test of the file-name / time-axis plumbing and the npz I/O
backend, GPU test of a full `si3_part_tracker.py -F` run against the C oracle."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_seed_name_parsing():
    import si3_part_tracker as cli
    assert quiet(cli.seed_name_info, "sitrack_seeding_nemoTsi3_19961215_00_HSS5.nc") == ("", "_idlSeed")
    assert quiet(cli.seed_name_info, "sitrack_seeding_sidfex_19961215_00_10km.nc") == ("_10km", "_idlSeed")
    assert quiet(cli.seed_name_info, "SELECTION_RGPS_S000_dt72_19970104h00_19970107h00_20km.nc") == ("_20km", "_dt72")


def test_npz_io_roundtrip(tmp_path):
    import sitrack_b200 as sit
    from make_synth_case import write_case
    r = write_case(str(tmp_path), grid="tiny", nrec=6, hss=2)
    Nt, t, d0, dN, conf, exp = quiet(sit.ModelFileTimeInfo, r["si3"])
    assert (Nt, conf, exp) == (6, "SYNTH4", "SYN00") and t.dtype == np.int32
    a, b, name, batch, t2 = quiet(sit.SeedFileTimeInfo, r["seed"])
    assert (a, b, batch) == (850608000, 850608000, "nemoTsi3") and name == "sitrack_seeding_nemoTsi3_19961215_00_HSS2"
    zt, ids, LL, YX = quiet(sit.LoadNCdata, r["seed"], krec=0)
    assert LL.shape == YX.shape == (ids.size, 2) and LL.dtype == np.float64 and (LL[:, 1] >= 0).all()
    # writer: same variables and dtypes as ncSaveCloudBuoys (ncio.py:131-197)
    f = str(tmp_path / "out.npz")
    quiet(sit.ncSaveCloudBuoys, f, np.array([1, 2]), ids, np.tile(YX[:, 0], (2, 1)), np.tile(YX[:, 1], (2, 1)),
          np.tile(LL[:, 0], (2, 1)), np.tile(LL[:, 1], (2, 1)), mask=np.ones((2, ids.size), "i1"))
    z = np.load(f)
    assert z["time"].dtype == np.int32 and z["id_buoy"].dtype == np.int64 and z["y_pos"].dtype == np.float32
    assert z["mask"].dtype == np.int8 and z["latitude"].shape == (2, ids.size)
    zt2, ids2, LL2, YX2, msk = quiet(sit.LoadNCdata, f, krec=1, lmask=True)
    assert np.array_equal(ids2, ids) and np.array_equal(YX2, YX)            # f4 values survive the round trip


@pytest.mark.gpu
def test_cli_full_run_vs_oracle(tmp_path, monkeypatch):
    import si3_part_tracker as cli
    from make_synth_case import write_case
    from oracle import corc
    import sitrack_b200 as sit
    r = write_case(str(tmp_path / "in"), grid="small", nrec=24, hss=3)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["si3_part_tracker.py", "-i", r["si3"], "-m", r["mesh"], "-s", r["seed"],
                                      "-F", "-N", "SYNTH4"])
    quiet(cli.main)
    out = [f for f in os.listdir(tmp_path / "nc") if "_tracking_" in f]
    out12 = [f for f in os.listdir(tmp_path / "nc") if "_tracking12_" in f]
    assert out == ["NEMO-SI3_SYNTH4_SYN00_tracking_nemoTsi3_idlSeed_19961215h00_19961216h00.npz"], out
    assert len(out12) == 1
    z = np.load(tmp_path / "nc" / out[0])
    cache = np.load(tmp_path / "seed" / "Initialized_buoys_sitrack_seeding_nemoTsi3_19961215_00_HSS3_SYNTH4.npz")
    assert sorted(cache.files) == sorted(["nP", "xPosG0", "xPosC0", "IDs", "vJIt", "VRTCS", "idxKeep"])
    # the same run through the oracle, from the grid the CLI derived (device forward projection of lat/lon)
    kmaskt, latT, lonT, Yt, Xt, Yf, Xf, ResKM = quiet(sit.GetModelGrid, r["mesh"])
    Yv, Xv, Yu, Xu = quiet(sit.GetModelUVGrid, r["mesh"])
    g0 = r["grid"]
    assert np.abs(Yf - g0["Yf"]).max() < 1e-6 and np.abs(Xu - g0["Xu"]).max() < 1e-6       # km, projection round trip
    g = dict(Yf=Yf, Xf=Xf, Yu=Yu, Xu=Xu, Yv=Yv, Xv=Xv, tmask=kmaskt)
    U, V, IC = r["records"]
    ref = corc.track(g, U, V, IC, cache["xPosC0"], cache["vJIt"].astype(np.int64))
    assert z["time"].shape == (25,) and z["y_pos"].shape == (25, int(cache["nP"]))
    assert np.array_equal(z["mask"], ref["mask"])
    assert np.array_equal(z["y_pos"], ref["posC"][:, :, 0].astype("f4"))                  # file dtype is f4
    assert np.array_equal(z["x_pos"], ref["posC"][:, :, 1].astype("f4"))
    lat_ref = corc.inv_stere(ref["posC"][1:].reshape(-1, 2)).reshape(24, -1, 2)
    assert np.abs(z["latitude"][1:] - lat_ref[:, :, 0].astype("f4")).max() <= 8e-6        # <= 1 ulp(f4) at 90 deg
    z12 = np.load(tmp_path / "nc" / out12[0])
    assert np.array_equal(z12["y_pos"][1], z["y_pos"][-1]) and z12["time"].tolist() == [int(z["time"][0]), int(z["time"][-1])]
    # second run hits the seeding cache and gives the same file
    quiet(cli.main)
    z2 = np.load(tmp_path / "nc" / out[0])
    assert np.array_equal(z2["y_pos"], z["y_pos"])


@pytest.mark.gpu
def test_readme_workflow_seeding_then_tracking(tmp_path, monkeypatch):
    """The README workflow of the reference, end to end on synthetic files: generate the seeding file
    from the mesh_mask + SI3 file (nemoSeed, HSS5), then track it with the CLI."""
    import si3_part_tracker as cli
    import generate_seeding
    import synth
    from make_synth_case import write_case
    r = write_case(str(tmp_path / "in"), grid="small", nrec=12, hss=5)
    monkeypatch.chdir(tmp_path)
    fseed = quiet(generate_seeding.main, ["-d", "1996-12-15_00:00:00", "-m", r["mesh"], "-i", r["si3"], "-k", "0", "-S", "5"])
    assert fseed == "./nc/sitrack_seeding_nemoTsi3_19961215_00_HSS5.npz"
    z = np.load(fseed)
    ids, SG, SC = synth.hss_seeds(r["grid"], r["records"][2][0], khss=5)
    assert z["id_buoy"].shape == ids.shape and z["latitude"].dtype == np.float32
    assert np.abs(z["latitude"][0] - SG[:, 0]).max() < 1e-5 and np.abs(z["y_pos"][0] - SC[:, 0]).max() < 1e-3
    monkeypatch.setattr(sys, "argv", ["si3_part_tracker.py", "-i", r["si3"], "-m", r["mesh"], "-s", fseed, "-F", "-N", "SYNTH4"])
    quiet(cli.main)
    out = [f for f in os.listdir(tmp_path / "nc") if "_tracking_" in f]
    assert len(out) == 1
    t = np.load(tmp_path / "nc" / out[0])
    assert t["y_pos"].shape[0] == 13 and t["mask"][0].all() and 0 < t["mask"][-1].sum() <= t["mask"].shape[1]


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_cli_under_torchrun_shards_buoys_and_writes_the_same_files(tmp_path):
    """`torchrun --nproc-per-node 2 si3_part_tracker.py ...`: rank 0 seeds, each rank tracks its block of buoys
    (both on cuda:0 here, so the test runs on a one-GPU box), rank 0 writes the files -- identical to the
    single-process run's, variable by variable."""
    import socket
    import subprocess
    from make_synth_case import write_case
    r = write_case(str(tmp_path / "in"), grid="small", nrec=12, hss=2)
    cli = os.path.join(ROOT, "si3_part_tracker.py")
    common = ["-i", r["si3"], "-m", r["mesh"], "-s", r["seed"], "-F", "-N", "SYNTH4", "--device", "0"]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir(); two.mkdir()
    subprocess.run([sys.executable, cli] + common, cwd=one, env=env, check=True, stdout=subprocess.DEVNULL)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), cli] + common,
                   cwd=two, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    files = sorted(os.listdir(one / "nc"))
    assert files == sorted(os.listdir(two / "nc")) and len(files) == 2
    for f in files:
        a, b = np.load(one / "nc" / f), np.load(two / "nc" / f)
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert np.array_equal(a[k], b[k]), (f, k)
        assert a["y_pos"].shape[1] > 256                       # more than one tile: both ranks had work


@pytest.mark.gpu
def test_cli_optional_physics_flags(tmp_path, monkeypatch):
    """--scheme/--interp/--hops route the record loop through st_step_ext: same files and layout, trajectories
    close to the upstream Euler/face-pick ones but not equal (12 records, displacements of a few km)."""
    import si3_part_tracker as cli
    from make_synth_case import write_case
    r = write_case(str(tmp_path / "in"), grid="small", nrec=12, hss=3)
    outs = {}
    for tag, extra in (("euler", []), ("rk4", ["--scheme", "rk4", "--interp", "linear", "--hops", "4"])):
        d = tmp_path / tag
        d.mkdir()
        monkeypatch.chdir(d)
        monkeypatch.setattr(sys, "argv", ["si3_part_tracker.py", "-i", r["si3"], "-m", r["mesh"], "-s", r["seed"],
                                          "-F", "-N", "SYNTH4"] + extra)
        quiet(cli.main)
        f = [f for f in os.listdir(d / "nc") if "_tracking_" in f]
        assert len(f) == 1
        outs[tag] = np.load(d / "nc" / f[0])
    a, b = outs["euler"], outs["rk4"]
    assert a["y_pos"].shape == b["y_pos"].shape and np.array_equal(a["y_pos"][0], b["y_pos"][0])
    both = (a["mask"][-1] == 1) & (b["mask"][-1] == 1)
    assert both.mean() > 0.8
    d = np.hypot(a["y_pos"][-1][both] - b["y_pos"][-1][both], a["x_pos"][-1][both] - b["x_pos"][-1][both])
    assert 0.0 < d.max() < 5.0 and np.median(d) < 1.0            # km


@pytest.mark.gpu
def test_cli_without_F_per_buoy_time_windows(tmp_path, monkeypatch):
    """The default mode of the reference (no -F, `lUse2DTime`): every buoy has its own first and last model record,
    derived from the seed file's 2-D `time_pos` (si3_part_tracker.py:264-312).  The seed file is built so that the
    windows are known by construction; the `tracking12` file (the only one written in this mode, :546-571) must hold
    each buoy's seed at its first record and the oracle's position after its last one."""
    import si3_part_tracker as cli
    from make_synth_case import write_case
    from oracle import corc
    import sitrack_b200 as sit
    from synth.records import T0_EPOCH
    nrec = 24
    r = write_case(str(tmp_path / "in"), grid="small", nrec=nrec, hss=3)
    z = dict(np.load(r["seed"]))
    # the reference stops with an ERROR when SeedInit drops a buoy in this mode (its `zTpos = zTpos[:,idxK]` is
    # commented out, si3_part_tracker.py:250,269-272) and so does the drop-in: seed only buoys that will be kept
    kmaskt, latT, lonT, Yt, Xt, Yf, Xf, ResKM = quiet(sit.GetModelGrid, r["mesh"])
    zt, zid, XG, XC = quiet(sit.LoadNCdata, r["seed"], krec=0)
    ok = quiet(sit.SeedInit, zid, XG, XC, latT, lonT, Yf, Xf, ResKM, kmaskt,
               xIceConc=np.asarray(r["records"][2][2], np.float64))[6]
    assert 0 < ok.size < zid.size
    z["id_buoy"] = z["id_buoy"][ok]
    for k in ("latitude", "longitude", "y_pos", "x_pos"):
        z[k] = z[k][:, ok]
    nP = z["id_buoy"].size
    rng = np.random.default_rng(5)
    tm = T0_EPOCH + 1800 + 3600 * np.arange(nrec)                      # model time axis (record centres)
    first = rng.integers(2, 7, nP); first[0] = 2                       # some buoy starts at the earliest record
    last = rng.integers(14, 21, nP); last[1] = 20
    tpos = np.stack([tm[first] - 1799, tm[last]]).astype("i4")         # see the docstring of record_windows' reference lines
    z["time"] = np.array([tpos[0].min(), tpos[1].max()], "i4")
    for k in ("latitude", "longitude", "y_pos", "x_pos"):
        z[k] = np.concatenate([z[k], z[k]], axis=0)
    z["time_pos"] = tpos
    os.remove(r["seed"])
    np.savez(r["seed"], **z)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["si3_part_tracker.py", "-i", r["si3"], "-m", r["mesh"], "-s", r["seed"],
                                      "-N", "SYNTH4"])
    quiet(cli.main)
    out = os.listdir(tmp_path / "nc")
    assert len(out) == 1 and "_tracking12_" in out[0], out           # no full-series file without -F
    t12 = np.load(tmp_path / "nc" / out[0])
    cache = np.load(tmp_path / "seed" / "Initialized_buoys_sitrack_seeding_nemoTsi3_19961215_00_HSS3_SYNTH4.npz")
    keep = cache["idxKeep"]
    assert keep.size == nP
    # the windows the CLI must have used: kstrt from the time-span logic, per-buoy records by construction
    idA = int(np.floor(tpos[0].min() / 3600.) * 3600); idB = int(np.ceil(tpos[1].max() / 3600.) * 3600)
    Nt, kstrt, kstop, iTmA, iTmB = quiet(sit.GetTimeSpan, 3600, tm, idA, int(tm.min()), int(tm.max()), iStop=idB)
    f_k, l_k = first[keep], last[keep]
    assert f_k.min() >= kstrt and l_k.max() <= kstop
    Yv, Xv, Yu, Xu = quiet(sit.GetModelUVGrid, r["mesh"])
    g = dict(Yf=Yf, Xf=Xf, Yu=Yu, Xu=Xu, Yv=Yv, Xv=Xv, tmask=kmaskt)
    U, V, IC = r["records"]
    ref = corc.track(g, U[kstrt:kstop + 1], V[kstrt:kstop + 1], IC[kstrt:kstop + 1], cache["xPosC0"],
                     cache["vJIt"].astype(np.int64), kstrt=kstrt, rec_first=f_k, rec_last=l_k)
    b = np.arange(f_k.size)
    k0, kN = f_k - kstrt, l_k - kstrt + 1
    assert np.array_equal(t12["y_pos"][0], ref["posC"][k0, b, 0].astype("f4"))
    assert np.array_equal(t12["x_pos"][0], ref["posC"][k0, b, 1].astype("f4"))
    assert np.array_equal(t12["y_pos"][1], ref["posC"][kN, b, 0].astype("f4"))
    assert np.array_equal(t12["x_pos"][1], ref["posC"][kN, b, 1].astype("f4"))
    assert np.array_equal(t12["mask"][0], ref["mask"][k0, b]) and np.array_equal(t12["mask"][1], ref["mask"][kN, b])
    assert t12["mask"][0].all() and 0 < t12["mask"][1].sum()
    # per-buoy times (:334-340, :463): start of the first record, end of the last one (fill when discontinued)
    assert np.array_equal(t12["time_pos"][0], tm[f_k] - 1800)
    alive_end = t12["mask"][1] == 1
    assert np.array_equal(t12["time_pos"][1][alive_end], (tm[l_k] + 1800)[alive_end])
    assert (t12["time_pos"][1][~alive_end] == -9999).all()


class _FakeNC:
    """A minimal in-memory stand-in for the netCDF4 module (absent from this image and from the GPU boxes), with the
    calls the reference's I/O makes: Dataset(fn, 'w'|'r'), createDimension, createVariable(name, dtype, dims,
    fill_value=, zlib=, complevel=), variable slicing / assignment with an unlimited first dimension, attributes."""
    files = {}

    class Var:
        def __init__(self, dtype, shape, fill):
            self.dtype, self.fill = np.dtype(dtype), fill
            self.a = np.zeros([0 if s is None else s for s in shape], self.dtype)
            self.shape0_unlimited = shape and shape[0] is None

        def _grow(self, n):
            if self.shape0_unlimited and n > self.a.shape[0]:
                pad = np.full((n - self.a.shape[0],) + self.a.shape[1:], self.fill if self.fill is not None else 0, self.dtype)
                self.a = np.concatenate([self.a, pad])

        def __setitem__(self, k, v):
            k0 = k[0] if isinstance(k, tuple) else k
            if isinstance(k0, (int, np.integer)):
                self._grow(int(k0) + 1)
            elif isinstance(k0, slice) and k0 == slice(None) and self.shape0_unlimited:
                self._grow(np.shape(v)[0])
            self.a[k] = np.asarray(v).astype(self.dtype)         # like netCDF4: cast to the variable's type

        def __getitem__(self, k):
            return np.ma.masked_equal(self.a[k], self.fill) if self.fill is not None else self.a[k]

        shape = property(lambda self: self.a.shape)

    class Dim:
        def __init__(self, ds, name):
            self.ds, self.name = ds, name

        @property
        def size(self):
            n = self.ds._dims[self.name]
            if n is None:
                return max([v.a.shape[0] for v in self.ds.variables.values() if v.shape0_unlimited] or [0])
            return n

    class Dataset:
        def __init__(self, fn, mode='r', format=None):
            if mode == 'w':
                object.__setattr__(self, "_dims", {}); object.__setattr__(self, "variables", {})
                object.__setattr__(self, "attrs", {}); _FakeNC.files[str(fn)] = self
            else:
                src = _FakeNC.files[str(fn)]
                for k in ("_dims", "variables", "attrs"):
                    object.__setattr__(self, k, getattr(src, k))
            object.__setattr__(self, "dimensions", {n: _FakeNC.Dim(self, n) for n in self._dims})

        def createDimension(self, name, size):
            self._dims[name] = size
            self.dimensions[name] = _FakeNC.Dim(self, name)

        def createVariable(self, name, dtype, dims, fill_value=None, zlib=False, complevel=4):
            assert all(d in self._dims for d in dims) and 0 <= complevel <= 9
            v = _FakeNC.Var(dtype, [self._dims[d] for d in dims], fill_value)
            self.variables[name] = v
            return v

        def __setattr__(self, k, v):
            self.attrs[k] = v

        def close(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *e):
            pass


def test_netcdf_branch_of_the_writer_and_readers(monkeypatch, tmp_path):
    """The netCDF4 code paths of ncio (writer: ncio.py:131-197; readers: :199-326) executed against an in-memory
    stand-in for the module: variables, dtypes, fill values, units and global attributes as the reference writes
    them, rows appended one at a time through the unlimited time dimension, and read back through LoadNCdata."""
    import types
    import sitrack_b200 as sit
    fake = types.ModuleType("netCDF4"); fake.Dataset = _FakeNC.Dataset
    monkeypatch.setitem(sys.modules, "netCDF4", fake)
    monkeypatch.setattr(sit.ncio, "chck4f", lambda fn: None)
    nt, nP = 4, 7
    rng = np.random.default_rng(0)
    t = 850608000 + 3600 * np.arange(nt)
    ids = np.arange(nP) + 11
    Y, X = rng.uniform(-900, 900, (nt, nP)), rng.uniform(-900, 900, (nt, nP))
    La, Lo = rng.uniform(60, 90, (nt, nP)), rng.uniform(-180, 180, (nt, nP))
    M = (rng.random((nt, nP)) > 0.2).astype("i1")
    Y[2, 3] = -9999.
    f = str(tmp_path / "out.nc")
    quiet(sit.ncSaveCloudBuoys, f, t, ids, Y, X, La, Lo, mask=M, corigin="NEMO-SI3_X_Y")
    ds = _FakeNC.files[f]
    assert sorted(ds.variables) == sorted(["time", "buoy", "id_buoy", "latitude", "longitude", "y_pos", "x_pos", "mask"])
    assert ds.variables["y_pos"].dtype == np.float32 and ds.variables["time"].dtype == np.int32
    assert ds.variables["id_buoy"].dtype == np.int64 and ds.variables["mask"].dtype == np.int8
    assert ds.variables["time"].units == sit.tunits_default and ds.attrs["Origin"] == "NEMO-SI3_X_Y"
    assert ds.variables["y_pos"].fill == -9999. and ds.dimensions["time"].size == nt
    zt, ids2, LL, YX, msk = quiet(sit.LoadNCdata, f, krec=2, lmask=True)
    assert zt == t[2] and np.array_equal(ids2, ids) and np.array_equal(msk, M[2])
    assert np.array_equal(np.ma.filled(YX[:, 0], -9999.), Y[2].astype("f4").astype("f8"))
    assert np.allclose(LL[:, 1], np.mod(Lo[2].astype("f4"), 360.))
    # streaming: the same file row by row, with per-buoy times
    f2 = str(tmp_path / "out2.nc")
    tp = np.tile(t[:, None], (1, nP)).astype("i4")
    with contextlib.redirect_stdout(io.StringIO()):
        with sit.CloudBuoyWriter(f2, nt, ids, with_mask=True, with_time_pos=True) as w:
            for jt in range(nt):
                w.write(jt, t[jt], Y[jt], X[jt], La[jt], Lo[jt], mask=M[jt], time_pos=tp[jt])
    d2 = _FakeNC.files[f2]
    for n in ("latitude", "longitude", "y_pos", "x_pos", "mask", "time"):
        assert np.array_equal(d2.variables[n].a, ds.variables[n].a), n
    assert d2.variables["time_pos"].units == sit.tunits_default
    Nt, t1d, t2d = quiet(sit.LoadNCtime, f2, ltime2d=True)
    assert Nt == nt and np.array_equal(np.asarray(t2d), tp)


@pytest.mark.gpu
def test_cli_sort_flag_permutes_buoys_not_trajectories(tmp_path, monkeypatch):
    """--sort stores and writes the buoys in cell-major order: the files list the same buoys (by ID) with the same
    trajectories, bit for bit, in a different order; the engine-level form (set_buoys(sort=True)) hands results
    back in the caller's order."""
    import si3_part_tracker as cli
    from make_synth_case import write_case
    import sitrack_b200 as sit
    r = write_case(str(tmp_path / "in"), grid="small", nrec=12, hss=2)
    # shuffle the seed file so that seed order is far from cell order
    z = dict(np.load(r["seed"]))
    sh = np.random.default_rng(1).permutation(z["id_buoy"].size)
    z["id_buoy"] = z["id_buoy"][sh]
    for k in ("latitude", "longitude", "y_pos", "x_pos"):
        z[k] = z[k][:, sh]
    os.remove(r["seed"]); np.savez(r["seed"], **z)
    outs = {}
    for tag, extra in (("seed", []), ("sorted", ["--sort"])):
        d = tmp_path / tag
        d.mkdir(); monkeypatch.chdir(d)
        monkeypatch.setattr(sys, "argv", ["si3_part_tracker.py", "-i", r["si3"], "-m", r["mesh"], "-s", r["seed"],
                                          "-F", "-N", "SYNTH4"] + extra)
        quiet(cli.main)
        f = [f for f in os.listdir(d / "nc") if "_tracking_" in f]
        outs[tag] = np.load(d / "nc" / f[0])
    a, b = outs["seed"], outs["sorted"]
    assert not np.array_equal(a["id_buoy"], b["id_buoy"]) and sorted(a["id_buoy"]) == sorted(b["id_buoy"])
    ia, ib = np.argsort(a["id_buoy"]), np.argsort(b["id_buoy"])
    for k in ("y_pos", "x_pos", "latitude", "longitude", "mask"):
        assert np.array_equal(a[k][:, ia], b[k][:, ib]), k
    # engine level: sorted storage, caller's order out
    cache = np.load(tmp_path / "seed" / "seed" / os.listdir(tmp_path / "seed" / "seed")[0])
    kmaskt, latT, lonT, Yt, Xt, Yf, Xf, ResKM = quiet(sit.GetModelGrid, r["mesh"])
    Yv, Xv, Yu, Xu = quiet(sit.GetModelUVGrid, r["mesh"])
    U, V, IC = r["records"]
    res = {}
    for srt in (False, True):
        with sit.TrackEngine(Yf, Xf, Yu, Xu, Yv, Xv, tmask=kmaskt) as eng:
            eng.set_buoys(cache["xPosC0"], cache["vJIt"], sort=srt)
            assert (eng.perm is not None) == srt
            res[srt] = (eng.track((U, V, IC), 12, pos0=cache["xPosC0"]), eng.get_state())
    for k in ("posC", "posG", "mask", "n_alive"):
        assert np.array_equal(res[False][0][k], res[True][0][k]), k
    for x, y in zip(res[False][1], res[True][1]):
        assert np.array_equal(x, y)


def test_cloud_buoy_writer_streams_rows_to_npz(tmp_path):
    """CloudBuoyWriter on the .npz backend: rows appended one at a time through on-disk members, the file assembled
    at close with the variables and dtypes of ncSaveCloudBuoys, scratch directory gone, and identical to the
    whole-array writer's file."""
    import sitrack_b200 as sit
    nt, nP = 5, 300
    rng = np.random.default_rng(1)
    t = 850608000 + 3600 * np.arange(nt)
    ids = rng.permutation(nP) + 1
    Y, X = rng.uniform(-900, 900, (nt, nP)), rng.uniform(-900, 900, (nt, nP))
    La, Lo = rng.uniform(60, 90, (nt, nP)), rng.uniform(-180, 180, (nt, nP))
    M = (rng.random((nt, nP)) > 0.2).astype("i1")
    f1, f2 = str(tmp_path / "a.npz"), str(tmp_path / "b.npz")
    with contextlib.redirect_stdout(io.StringIO()):
        with sit.CloudBuoyWriter(f1, nt, ids, with_mask=True) as w:
            for jt in range(nt):
                w.write(jt, t[jt], Y[jt].astype("f4"), X[jt].astype("f4"), La[jt], Lo[jt], mask=M[jt])
        sit.ncSaveCloudBuoys(f2, t, ids, Y, X, La, Lo, mask=M)
    assert sorted(os.listdir(tmp_path)) == ["a.npz", "b.npz"]          # no scratch directory left behind
    a, b = np.load(f1), np.load(f2)
    assert sorted(a.files) == sorted(b.files) == sorted(["time", "buoy", "id_buoy", "latitude", "longitude", "y_pos", "x_pos", "mask"])
    for k in a.files:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
    assert a["y_pos"].dtype == np.float32 and a["time"].dtype == np.int32 and a["id_buoy"].dtype == np.int64


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference on the small config: one JSON line with the keys the driver reads (no GPU involved)."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                        "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=600)
    assert r.returncode == 0
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "buoy-steps/sec" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "buoy-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "lat/lon" in d["cpu_baseline"]["sample"]
    assert d["steps"] == 2 and d["warmup"] == 1 and d["gpu_launches"] == 0
