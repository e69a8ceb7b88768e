"""Process-wide settings of the drop-in function surface."""
import os

# CUDA device used by the function-style API (sit.SeedInit, sit.IsInsideQuadrangle, ...).
# One process per GPU: under torchrun this follows LOCAL_RANK.
device = int(os.environ.get("SITRACK_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
