"""Synthetic NANUK4-shaped inputs: grid, hourly records, seeds.

Host-side input GENERATION only -- nothing here is on the tracking path.  All
generators are deterministic in their integer seeds.  The small numpy
projection helpers below exist so that the grid's lat/lon are consistent with
its km coordinates; they are not used by the product (the device kernel has
its own inverse) nor by the oracle (oracle/st_oracle.c has its own).
"""
from .grid import make_grid, GRID_PRESETS          # noqa: F401
from .records import make_records                   # noqa: F401
from .seeds import hss_seeds, scattered_seeds, dense_seeds   # noqa: F401
