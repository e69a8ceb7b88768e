"""Drop-in for `sitrack.ncio` (reference: sitrack/ncio.py).  Host-side I/O stays
Python; netCDF4 is imported lazily (it is not installed in the build image), and
every reader/writer also accepts `.npz` files holding the same variable names so
the whole CLI can run on synthetic data without netCDF.
"""
import numpy as np

tunits_default = 'seconds since 1970-01-01 00:00:00'     # ncio.py:15
FillValue = -9999.                                       # ncio.py:19
