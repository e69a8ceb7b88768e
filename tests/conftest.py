import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _grid_of(npz):
    return {k[2:]: npz[k] for k in npz.files if k.startswith("g_")}


@pytest.fixture(scope="session")
def gold_pred():
    return np.load(os.path.join(GOLD, "predicates.npz"))


@pytest.fixture(scope="session")
def gold_track():
    z = np.load(os.path.join(GOLD, "track_tiny.npz"))
    return z, _grid_of(z)


@pytest.fixture(scope="session")
def gold_seed():
    z = np.load(os.path.join(GOLD, "seedinit_small.npz"))
    return z, _grid_of(z)


@pytest.fixture(scope="session")
def corc():
    from oracle import corc as m
    m.build()
    return m


@pytest.fixture(scope="session")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="session")
def sit():
    import sitrack_b200
    return sitrack_b200


def engine_for(g, **kw):
    import sitrack_b200 as sit
    return sit.TrackEngine(g["Yf"], g["Xf"], g["Yu"], g["Xu"], g["Yv"], g["Xv"], tmask=g["tmask"], **kw)


TRACK_CASES = {
    "uv1": dict(scale=1.0, uv_strategy=1, kstrt=0, win=False),
    "uv0": dict(scale=1.0, uv_strategy=0, kstrt=0, win=False),
    "fast": dict(scale=4.0, uv_strategy=1, kstrt=0, win=False),
    "win": dict(scale=1.0, uv_strategy=1, kstrt=3, win=True),
}
