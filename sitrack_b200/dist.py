"""Multi-GPU plumbing: buoys shard by index (they never interact, reference
si3_part_tracker.py:378-488), one process per GPU, the static grid and the current record
replicated.  The only collectives are the optional per-output-record all-gather of trajectory
rows and the all-reduce of the alive count.  Two forms of the all-gather:
  * RowGatherer: torch.distributed all_gather_into_tensor (NCCL on GPUs, gloo in the CPU tests);
  * PeerGather: fused into the step kernel -- every rank's k_advect_warp stores its new positions
    straight into the gathered array of all ranks over NVLink peer memory (st_step_gather; `mode` picks per-thread
    peer stores, copy engines or cp.async.bulk tile stores); here torch.distributed only carries the 64-byte CUDA
    IPC handles once, at set-up.  Shards are tile-aligned (TILE rows): a rank's block of the gathered array must
    start on a multiple of 32 rows for its warps' peer stores to cover whole sectors at the receiver (a quarter of
    the NVLink throughput is lost otherwise).
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


class PeerGather:
    """Fused per-record all-gather of positions (include/sitrack_b200.h: st_gather_*).

        pg = PeerGather(engine, n_global, offset)          # after engine.set_buoys(...)
        for k in range(nrec):
            pg.step(slot, jrec, out_latlon, out_mask, n_alive, stream)   # step + remote row stores
            row = pg.wait(consumer_stream)                 # (n_global,2) tensor, valid on that stream
            ...                                            # consume `row` on consumer_stream
            pg.release(consumer_stream)                    # the buffer may be overwritten nbuf records later
    """

    def __init__(self, engine, n_global, offset, f4=False, nbuf=2, group=None, mode=0):
        import torch.distributed as dist
        self.eng, self.nbuf, self.seq = engine, int(nbuf), 0
        if offset % 32:
            import warnings
            warnings.warn("PeerGather: this rank's block starts at row %d, not a multiple of 32: its peer stores will "
                          "straddle sectors (use dist.shard_bounds)" % offset)
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
        handle = engine.gather_create(rank, world, n_global, offset, f4=f4, nbuf=nbuf)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            engine.gather_connect_ipc(handles)
            dist.barrier(group)                            # every rank has mapped every block
        if mode:
            engine.gather_set_mode(mode)

    def step(self, slot, jrec, out_latlon=None, out_mask=None, n_alive=None, stream=None):
        self.seq += 1
        self.eng.step_gather(slot, jrec, (self.seq - 1) % self.nbuf, self.seq, out_latlon, out_mask, n_alive, stream)
        return self.seq

    def wait(self, stream=None, seq=None):
        seq = self.seq if seq is None else seq
        self.eng.gather_wait(seq, stream)
        return self.eng.gather_buffer((seq - 1) % self.nbuf)

    def release(self, stream=None, seq=None):
        self.eng.gather_ack(self.seq if seq is None else seq, stream)

    def close(self):
        self.eng.gather_destroy()


def bind_host_to_gpu(device):
    """Pin this process to the CPU cores next to `device` (NVML's cpu affinity of the GPU), so that the
    pinned staging buffers it allocates afterwards sit on the GPU's own NUMA node and its PCIe copies do
    not cross the socket interconnect.  Returns the core list, or None when NVML / the call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        return None
    return None
