"""Import the UNMODIFIED upstream `sitrack` package as `sit` (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing that runs there may call this; it is used by oracle/make_golden.py to
produce the committed fixtures in tests/golden/ and by a few CPU tests that
skip when the reference is absent.

The only shim needed: upstream sitrack/ncio.py:8 imports netCDF4 (absent
here); an empty stub module lets `import sitrack` succeed.  Everything on the
hot path (locate.py, tracking.py, util.Haversine) then runs as shipped.  The
cartopy-backed projections and the netCDF readers/writers stay unusable.
"""
import os
import sys
import types

REF_DIRS = [os.environ.get("SITRACK_REF", ""), "/root/reference"]


def available():
    return any(d and os.path.isdir(os.path.join(d, "sitrack")) for d in REF_DIRS)


def load():
    for d in REF_DIRS:
        if d and os.path.isdir(os.path.join(d, "sitrack")):
            if "netCDF4" not in sys.modules:
                stub = types.ModuleType("netCDF4")
                stub.Dataset = None
                sys.modules["netCDF4"] = stub
            if d not in sys.path:
                sys.path.insert(0, d)
            import sitrack as sit
            assert os.path.realpath(os.path.dirname(sit.__file__)).startswith(os.path.realpath(d)), sit.__file__
            return sit
    raise ImportError("upstream sitrack not found (set SITRACK_REF)")
