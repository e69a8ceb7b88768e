"""ctypes binding of the C oracle (oracle/st_oracle.c -> oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(sitrack_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborc.so")
FILL = -9999.0


def build(force=False):
    src = os.path.join(_HERE, "st_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None

_f8 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f4 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i1 = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_i8 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_intersect2seg.argtypes = [_f8, _f8, _f8, _f8]
        L.orc_intersect2seg.restype = C.c_int
        L.orc_inside_quad.argtypes = [C.c_double, C.c_double, _f8]
        L.orc_inside_quad.restype = C.c_int
        L.orc_survive.argtypes = [C.c_long, C.c_long, C.c_int, C.c_int, _i1, _f4, C.c_double]
        L.orc_survive.restype = C.c_int
        L.orc_crossed_edge_c.argtypes = [_f8, _f8, C.c_long, C.c_long, _f8, _f8, C.c_int]
        L.orc_crossed_edge_c.restype = C.c_int
        L.orc_new_host_cell_c.argtypes = [C.c_int, _f8, _f8, C.c_long, C.c_long, _f8, _f8, C.c_int]
        L.orc_new_host_cell_c.restype = C.c_int
        L.orc_inv_stere.argtypes = [C.c_long, _f8, _f8, C.c_double, C.c_double]
        L.orc_inv_stere.restype = None
        L.orc_fwd_stere.argtypes = [C.c_long, _f8, _f8, C.c_double, C.c_double]
        L.orc_fwd_stere.restype = None
        L.orc_inv_stere_ell.argtypes = [C.c_long, _f8, _f8, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_inv_stere_ell.restype = None
        L.orc_fwd_stere_ell.argtypes = [C.c_long, _f8, _f8, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_fwd_stere_ell.restype = None
        L.orc_haversine.argtypes = [C.c_double] * 4
        L.orc_haversine.restype = C.c_double
        L.orc_nearest_point.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, _f8, _f8, C.c_void_p,
                                        C.c_double, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long),
                                        C.POINTER(C.c_double)]
        L.orc_nearest_point.restype = None
        L.orc_find_containing_cell.argtypes = [C.c_double, C.c_double, C.c_long, C.c_long, _f8, _f8, C.c_int,
                                               C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_find_containing_cell.restype = C.c_int
        L.orc_seed_init.argtypes = [C.c_long, _f8, _f8, C.c_int, C.c_int, _f8, _f8, _f8, _f8, _f8, _i1, _f4,
                                    C.c_double, _i8, _i1, C.c_void_p]
        L.orc_seed_init.restype = None
        L.orc_track.argtypes = [C.c_int, C.c_int, _f8, _f8, _f8, _f8, _f8, _f8, _i1,
                                C.c_int, C.c_int, _f4, _f4, _f4,
                                C.c_long, _f8, _f8, _i1, _i8, _i1,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_track.restype = C.c_long
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def intersect2seg(A, B, Cc, D):
    return bool(lib().orc_intersect2seg(_c(A, "f8"), _c(B, "f8"), _c(Cc, "f8"), _c(D, "f8")))


def inside_quad(y, x, quad):
    return bool(lib().orc_inside_quad(float(y), float(x), _c(quad, "f8").reshape(-1)))


def survive(jT, iT, tmask, ic, rmin_conc=0.1):
    Nj, Ni = tmask.shape
    return lib().orc_survive(int(jT), int(iT), Nj, Ni, _c(tmask, "i1"), _c(ic, "f4"), rmin_conc)


def crossed_edge(p1, p2, jT, iT, Yf, Xf):
    return lib().orc_crossed_edge_c(_c(p1, "f8"), _c(p2, "f8"), int(jT), int(iT), _c(Yf, "f8"), _c(Xf, "f8"),
                                    Yf.shape[1])


def new_host_cell(kcross, p1, p2, jT, iT, Yf, Xf):
    return lib().orc_new_host_cell_c(int(kcross), _c(p1, "f8"), _c(p2, "f8"), int(jT), int(iT),
                                     _c(Yf, "f8"), _c(Xf, "f8"), Yf.shape[1])


WGS84 = (6378137.0, 1.0 / 298.257223563)


def inv_stere(yx, lat_ts=70.0, lon0=-45.0, ellipsoid=None):
    """[y,x] km -> [lat,lon] degrees; ellipsoid = (a [m], f), default WGS84 (the reference's case)."""
    yx = _c(yx, "f8").reshape(-1, 2)
    out = np.empty_like(yx)
    if ellipsoid is None:
        lib().orc_inv_stere(yx.shape[0], yx, out, lat_ts, lon0)
    else:
        lib().orc_inv_stere_ell(yx.shape[0], yx, out, lat_ts, lon0, float(ellipsoid[0]), float(ellipsoid[1]))
    return out


def fwd_stere(latlon, lat_ts=70.0, lon0=-45.0, ellipsoid=None):
    ll = _c(latlon, "f8").reshape(-1, 2)
    out = np.empty_like(ll)
    if ellipsoid is None:
        lib().orc_fwd_stere(ll.shape[0], ll, out, lat_ts, lon0)
    else:
        lib().orc_fwd_stere_ell(ll.shape[0], ll, out, lat_ts, lon0, float(ellipsoid[0]), float(ellipsoid[1]))
    return out


def nearest_point(latP, lonP, latT, lonT, reskm, rd_found_km=2.5, max_itr=10):
    Nj, Ni = latT.shape
    jy, jx, d = C.c_long(), C.c_long(), C.c_double()
    r = _c(reskm, "f8") if reskm is not None else None
    lib().orc_nearest_point(float(latP), float(lonP), Nj, Ni, _c(latT, "f8"), _c(lonT, "f8"),
                            r.ctypes.data if r is not None else None, rd_found_km, max_itr,
                            C.byref(jy), C.byref(jx), C.byref(d))
    return jy.value, jx.value, d.value


def find_containing_cell(y, x, kj, ki, Yf, Xf):
    j, i = C.c_long(), C.c_long()
    ok = lib().orc_find_containing_cell(float(y), float(x), int(kj), int(ki), _c(Yf, "f8"), _c(Xf, "f8"),
                                        Yf.shape[1], C.byref(j), C.byref(i))
    return bool(ok), j.value, i.value


def seed_init(SG, SC, latT, lonT, Yf, Xf, reskm, tmask, ic0, rmin_conc=0.1):
    """-> (jiT (nP,2) i8, keep (nP,) i1, jiNearest (nP,2) i8); not compacted."""
    nP = SG.shape[0]
    Nj, Ni = latT.shape
    jiT = np.zeros((nP, 2), np.int64)
    near = np.zeros((nP, 2), np.int64)
    keep = np.zeros(nP, np.int8)
    lib().orc_seed_init(nP, _c(SG, "f8"), _c(SC, "f8"), Nj, Ni, _c(latT, "f8"), _c(lonT, "f8"),
                        _c(Yf, "f8"), _c(Xf, "f8"), _c(reskm, "f8"), _c(tmask, "i1"), _c(ic0, "f4"),
                        rmin_conc, jiT, keep, near.ctypes.data)
    return jiT, keep, near


def track(grid, U, V, IC, pos0, jiT0, kstrt=0, rec_first=None, rec_last=None, uv_strategy=1,
          rdt=3600.0, rmin_conc=0.1, do_latlon=True, history=True, posG0=None, alive0=None):
    """Run the record x buoy loop.  grid: dict with Yf,Xf,Yu,Xu,Yv,Xv (f8) and tmask (i1).
    U,V,IC: (nrec,Nj,Ni) f4.  pos0 (nP,2) [y,x] km, jiT0 (nP,2).  -F semantics unless
    rec_first/rec_last are given (then row k0 = rec_first-kstrt holds the seed)."""
    U, V, IC = _c(U, "f4"), _c(V, "f4"), _c(IC, "f4")
    nrec, Nj, Ni = U.shape
    nP = pos0.shape[0]
    posC = np.full((nrec + 1, nP, 2), FILL)
    posG = np.full((nrec + 1, nP, 2), FILL)
    mask = np.zeros((nrec + 1, nP), np.int8)
    if rec_first is None:
        posC[0] = pos0
        mask[0] = 1
        if posG0 is not None:
            posG[0] = posG0
    else:
        for b in range(nP):
            k0 = int(rec_first[b]) - kstrt
            posC[k0, b] = pos0[b]
            mask[k0, b] = 1
            if posG0 is not None:
                posG[k0, b] = posG0[b]
    jiT = _c(jiT0, "i8").copy()
    alive = np.ones(nP, np.int8) if alive0 is None else _c(alive0, "i1").copy()
    nal = np.zeros(nrec, np.int64)
    jh = np.zeros((nrec + 1, nP, 2), np.int32) if history else None
    ah = np.zeros((nrec + 1, nP), np.int8) if history else None
    rf = _c(rec_first, "i4") if rec_first is not None else None
    rl = _c(rec_last, "i4") if rec_last is not None else None
    g = grid
    ncross = lib().orc_track(Nj, Ni, _c(g["Yf"], "f8"), _c(g["Xf"], "f8"), _c(g["Yu"], "f8"), _c(g["Xu"], "f8"),
                             _c(g["Yv"], "f8"), _c(g["Xv"], "f8"), _c(g["tmask"], "i1"),
                             nrec, kstrt, U, V, IC, nP, posC, posG, mask, jiT, alive,
                             rf.ctypes.data if rf is not None else None,
                             rl.ctypes.data if rl is not None else None,
                             uv_strategy, rdt, rmin_conc, int(do_latlon),
                             nal.ctypes.data, jh.ctypes.data if history else None,
                             ah.ctypes.data if history else None)
    return dict(posC=posC, posG=posG, mask=mask, jiT=jiT, alive=alive, nalive=nal,
                jiT_hist=jh, alive_hist=ah, ncross=ncross)
