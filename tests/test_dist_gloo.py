"""CPU, world_size 2 over gloo: the sharding and row-gather plumbing of the multi-GPU path.
The shard a rank tracks, run through the oracle, must reassemble (all-gather in seed order,
all-reduced alive counts) into exactly the single-process result -- buoys never interact."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from sitrack_b200.dist import shard_bounds, my_shard
    for n in (0, 1, 255, 256, 257, 1000, 12_469_235):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all() and len(b) == w + 1
            assert all(x % 256 == 0 for x in b[1:-1] if x < n)           # tile-aligned interior cuts
            if n > 256 * w:
                assert np.diff(b).max() - np.diff(b).min() <= 256 + 255
            assert my_shard(n, w - 1, w) == (int(b[-2]), n)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import corc
    from sitrack_b200.dist import RowGatherer, allreduce_sum, my_shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "track_tiny.npz"))
        g = {k[2:]: z[k] for k in z.files if k.startswith("g_")}
        # a cloud big enough to span several tiles: the golden seeds replicated with tiny offsets
        rng = np.random.default_rng(0)
        rep = 9
        pos0 = np.concatenate([z["pos0"] + rng.uniform(-0.5, 0.5, z["pos0"].shape) for _ in range(rep)])
        cell0 = np.concatenate([z["jiT0"]] * rep)
        n = pos0.shape[0]
        lo, hi = my_shard(n, rank, world)
        nrec = 12
        U, V, IC = 3 * z["U"][:nrec], 3 * z["V"][:nrec], z["IC"][:nrec]
        mine = corc.track(g, U, V, IC, pos0[lo:hi], cell0[lo:hi])
        gat = RowGatherer(n, world, rank, width=2)
        rows = []
        for k in range(nrec):
            gat.gather(torch.from_numpy(mine["posC"][k + 1]))
            rows.append(gat.result().numpy().copy())
        alive = allreduce_sum(torch.from_numpy(mine["nalive"].copy()))
        if rank == 0:
            full = corc.track(g, U, V, IC, pos0, cell0)
            ok = all(np.array_equal(rows[k], full["posC"][k + 1]) for k in range(nrec))
            ok = ok and np.array_equal(alive.numpy(), full["nalive"]) and full["ncross"] > 0
            q.put(("ok" if ok else "mismatch", n, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(100)
    assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
    status, n, lo, hi = q.get(timeout=5)
    assert status == "ok" and n > 512 and (lo, hi) == (0, 512)


def test_shard_bounds_are_tile_aligned_and_balanced():
    """Every rank's block starts on a tile boundary (a multiple of 32 rows is what the peer stores of the fused
    all-gather need to cover whole sectors at the receiver) and the blocks differ by one tile at most."""
    import numpy as np
    from sitrack_b200.dist import shard_bounds, my_shard, TILE
    assert TILE % 32 == 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(0, 5_000_000)); world = int(rng.integers(1, 17)); tile = int(rng.choice([32, 64, 256]))
        b = shard_bounds(n, world, tile)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1 and np.all(np.diff(b) >= 0)
        assert all(int(x) % tile == 0 for x in b[:-1] if x < n)
        sizes = np.diff(b)
        full = sizes[sizes > 0][:-1] if (sizes > 0).any() else sizes
        assert full.size == 0 or full.max() - full.min() <= tile
        r = int(rng.integers(0, world))
        assert my_shard(n, r, world, tile) == (int(b[r]), int(b[r + 1]))
