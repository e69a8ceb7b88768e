"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/sitrack_b200.h declares; compute entry points fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sitrack_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def so():
    from sitrack_b200 import build
    return build.build()


def test_header_symbols_exported(so):
    L = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export: " + n


def test_binding_covers_header(so):
    from sitrack_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().st_abi_version() == 3


def test_only_sm100a_code(so):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", so], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sitrack_b200 as sit
    with pytest.raises(sit.SitrackCudaError, match="no CUDA device|no CPU fallback|CUDA"):
        sit.IsInsideQuadrangle(2., 2., np.array([[0., 0.], [3., 0.], [4., 4.], [1., 3.5]]))
    z = np.zeros((8, 8))
    with pytest.raises(sit.SitrackCudaError):
        sit.TrackEngine(z, z, z, z, z, z, tmask=np.ones((8, 8), np.int8))


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (or synth/)."""
    pkg = os.path.join(ROOT, "sitrack_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|synth)\b", src, flags=re.M), f
                assert "st_oracle" not in src and "liborc" not in src, f
