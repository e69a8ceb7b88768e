"""Pins for the polar stereographic projection that this repository's author did not compute (the
reference reaches PROJ through cartopy and holds no vector of its own): tests/golden/proj_kat.json quotes NSIDC's
published grid-corner table (Hughes 1980 ellipsoid), the worked example of Snyder (1987) (Clarke 1866 ellipsoid)
and the defining property of the true-scale parallel.  The C oracle is checked here on the CPU; the three CUDA
kernels (inv_stere, inv_stere_fast, fwd_stere) against the same numbers in the gpu-marked tests."""
import json
import os

import numpy as np
import pytest

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "proj_kat.json")))


def _nsidc():
    k = KAT["nsidc_north_grid_corners"]
    f = 1.0 - np.sqrt(1.0 - k["e"] ** 2)                       # e^2 = f (2 - f)
    yx = np.stack([np.array(k["y_km"]), np.array(k["x_km"])], axis=1)
    lon = (np.array(k["lon_deg_east"]) + 180.0) % 360.0 - 180.0
    return k, (k["a_m"], f), yx, np.array(k["lat_deg"]), lon


def _snyder():
    k = KAT["snyder_1987_polar_example"]
    f = 1.0 - np.sqrt(1.0 - k["e2"])
    # south polar aspect -> north-polar formulas with every sign flipped (Snyder p. 161)
    ll = np.array([[-k["phi"], -k["lam"]]])
    yx_pub = np.array([[-k["y_m"], -k["x_m"]]]) / 1000.0
    return k, (k["a_m"], f), ll, -k["phi_c"], -k["lam0"], yx_pub


def check_inverse_nsidc(inv):
    k, ell, yx, lat, lon = _nsidc()
    out = inv(yx, k["lat_ts"], k["lon0"], ell)
    assert np.abs(out[:, 0] - lat).max() <= k["tolerance_deg"] + 1e-9
    dlon = (out[:, 1] - lon + 180.0) % 360.0 - 180.0
    assert np.abs(dlon).max() <= k["tolerance_deg"] + 1e-9


def check_forward_nsidc(fwd, inv):
    """the published corners are rounded to 0.01 degree (~1 km): forward them and compare within that rounding,
    then require the exact round trip of the forward image."""
    k, ell, yx, lat, lon = _nsidc()
    out = fwd(np.stack([lat, lon], axis=1), k["lat_ts"], k["lon0"], ell)
    assert np.abs(out - yx).max() < 1.5                          # km; 0.01 degree of rounding at 31-34 N
    back = inv(out, k["lat_ts"], k["lon0"], ell)
    assert np.abs(back[:, 0] - lat).max() < 1e-9
    assert np.abs((back[:, 1] - lon + 180.0) % 360.0 - 180.0).max() < 1e-9


def check_forward_snyder(fwd):
    k, ell, ll, lat_ts, lon0, yx_pub = _snyder()
    yx = fwd(ll, lat_ts, lon0, ell)
    assert np.abs(yx - yx_pub).max() / np.abs(yx_pub).max() < k["xy_relative_tolerance"]
    # the printed scale factor: k = rho / (a m)
    phi = np.radians(ll[0, 0])
    m = np.cos(phi) / np.sqrt(1.0 - k["e2"] * np.sin(phi) ** 2)
    rho_m = 1000.0 * np.hypot(yx[0, 0], yx[0, 1])
    assert abs(rho_m / (k["a_m"] * m) - k["k"]) < k["k_tolerance"]


def check_inverse_snyder(inv, fwd):
    k, ell, ll, lat_ts, lon0, yx_pub = _snyder()
    out = inv(fwd(ll, lat_ts, lon0, ell), lat_ts, lon0, ell)
    assert abs(out[0, 0] - ll[0, 0]) < 1e-9 and abs((out[0, 1] - ll[0, 1] + 180) % 360 - 180) < 1e-9
    # and from the printed coordinates: within their 1e-5 relative accuracy (16 m at 1639 km is 1.5e-4 degrees)
    out = inv(yx_pub, lat_ts, lon0, ell)
    assert abs(out[0, 0] - ll[0, 0]) < 2e-4 and abs((out[0, 1] - ll[0, 1] + 180) % 360 - 180) < 2e-4


def check_true_scale(fwd):
    """rho(phi_c) = a m(phi_c): the standard parallel is true to scale; k = rho / (a m) is the same along the
    meridian and along the parallel (conformality), by finite differences of the forward map."""
    k = KAT["definitions"]
    a, f = k["a_m"], 1.0 / k["inv_f"]
    e2 = f * (2.0 - f)
    lon = np.array([-45.0, 0.0, 77.0, -160.0])
    yx = fwd(np.stack([np.full(4, k["lat_ts"]), lon], axis=1), k["lat_ts"], k["lon0"], (a, f))
    phi = np.radians(k["lat_ts"])
    m = np.cos(phi) / np.sqrt(1.0 - e2 * np.sin(phi) ** 2)
    assert np.abs(1000.0 * np.hypot(yx[:, 0], yx[:, 1]) / (a * m) - 1.0).max() < 1e-12
    assert abs(yx[0, 1]) < 1e-9 and yx[0, 0] < 0                 # the central meridian runs down the -y axis
    for lat in (55.0, 70.0, 82.5, 89.0):
        phi = np.radians(lat)
        m = np.cos(phi) / np.sqrt(1.0 - e2 * np.sin(phi) ** 2)
        M = a * (1.0 - e2) / (1.0 - e2 * np.sin(phi) ** 2) ** 1.5          # meridional radius of curvature
        h = 1e-4                                                            # degrees
        p = fwd(np.array([[lat, 10.0], [lat + h, 10.0], [lat, 10.0 + h]]), k["lat_ts"], k["lon0"], (a, f)) * 1000.0
        k_mer = np.hypot(*(p[1] - p[0])) / (M * np.radians(h))
        k_par = np.hypot(*(p[2] - p[0])) / (a * m * np.radians(h))
        k_rho = np.hypot(p[0, 0], p[0, 1]) / (a * m)
        assert abs(k_mer / k_rho - 1.0) < 2e-6 and abs(k_par / k_rho - 1.0) < 2e-6
        # meridian and parallel stay perpendicular
        c = np.dot(p[1] - p[0], p[2] - p[0]) / (np.hypot(*(p[1] - p[0])) * np.hypot(*(p[2] - p[0])))
        assert abs(c) < 1e-5


def test_oracle_projection_against_published_values(corc):
    check_inverse_nsidc(corc.inv_stere)
    check_forward_nsidc(corc.fwd_stere, corc.inv_stere)
    check_forward_snyder(corc.fwd_stere)
    check_inverse_snyder(corc.inv_stere, corc.fwd_stere)
    check_true_scale(corc.fwd_stere)


def _gpu_fn(which):
    import ctypes as C
    from sitrack_b200 import _lib
    from sitrack_b200._lib import check, hptr

    def fn(a, lat_ts, lon0, ell):
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 2)
        out = np.empty_like(a)
        check(_lib.lib().st_selftest_proj(0, which, a.shape[0], hptr(a), hptr(out), float(lat_ts), float(lon0),
                                          float(ell[0]), float(ell[1])))
        return out
    return fn


@pytest.mark.gpu
@pytest.mark.parametrize("which", [0, 1], ids=["inv_stere", "inv_stere_fast"])
def test_cuda_inverse_projection_against_published_values(torch, which):
    inv, fwd = _gpu_fn(which), _gpu_fn(2)
    check_inverse_nsidc(inv)
    check_inverse_snyder(inv, fwd)
    check_forward_nsidc(fwd, inv)


@pytest.mark.gpu
def test_cuda_forward_projection_against_published_values(torch):
    fwd = _gpu_fn(2)
    check_forward_snyder(fwd)
    check_true_scale(fwd)
