"""Multi-GPU plumbing: buoys shard by index (they never interact, reference
si3_part_tracker.py:378-488), one process per GPU, the static grid and the current record
replicated.  The only collectives are the optional per-output-record all-gather of trajectory
rows and the all-reduce of the alive count; both go through torch.distributed (NCCL on GPUs,
gloo in the CPU tests).
"""
import numpy as np

TILE = 256          # shard boundaries fall on kernel tiles


def shard_bounds(n, world, tile=TILE):
    """world+1 offsets of contiguous, tile-aligned, balanced shards of n buoys."""
    ntiles = (n + tile - 1) // tile
    base, extra = divmod(ntiles, world)
    b = [0]
    for r in range(world):
        b.append(min(n, b[-1] + (base + (1 if r < extra else 0)) * tile))
    b[-1] = n
    return np.asarray(b, dtype=np.int64)


def my_shard(n, rank, world, tile=TILE):
    b = shard_bounds(n, world, tile)
    return int(b[rank]), int(b[rank + 1])


class RowGatherer:
    """All-gather of per-rank trajectory rows (nP_local, width) into the global row, in seed order.

    Shards may differ in length, so rows are padded to the longest shard for
    all_gather_into_tensor and the padding is dropped on the way out."""

    def __init__(self, n_global, world, rank, width=2, dtype=None, device="cpu"):
        import torch
        self.torch = torch
        self.b = shard_bounds(n_global, world)
        self.world, self.rank = world, rank
        self.npad = int(np.max(np.diff(self.b))) if world > 0 else 0
        dtype = dtype or torch.float64
        self.send = torch.zeros((self.npad, width), dtype=dtype, device=device)
        self.recv = torch.empty((world * self.npad, width), dtype=dtype, device=device)
        self.nloc = int(self.b[rank + 1] - self.b[rank])

    def gather(self, rows, async_op=False):
        """rows: (nP_local, width) tensor of this rank.  Returns the work handle (async) or None."""
        import torch.distributed as dist
        self.send[: self.nloc].copy_(rows, non_blocking=True)
        return dist.all_gather_into_tensor(self.recv, self.send, async_op=async_op)

    def result(self):
        """(n_global, width) view-free tensor in global buoy order."""
        parts = [self.recv[r * self.npad: r * self.npad + int(self.b[r + 1] - self.b[r])] for r in range(self.world)]
        return self.torch.cat(parts, dim=0)


def allreduce_sum(t):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
