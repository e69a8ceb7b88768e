"""GPU, world_size 2 over real peer memory (needs two devices; skipped on a one-GPU box): every rank's step
kernel stores its shard's new positions into BOTH ranks' gathered arrays through CUDA-IPC-mapped HBM
(st_step_gather, sitrack_b200.dist.PeerGather; per-thread peer stores, copy engines or cp.async.bulk tile stores).
After each record each rank's gathered array must equal the row of the unsharded run by the C oracle, bit for bit,
and the NCCL all-gather (sitrack_b200.dist.RowGatherer) of every rank's own block of it must reproduce it.  (On a
one-GPU box the same protocol runs with two contexts on one device in tests/test_gpu_parity.py, and every multi-GPU
`bench.py` run checks the exchange against NCCL on all ranks.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, f4, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from conftest import engine_for
    from oracle import corc
    from sitrack_b200.dist import PeerGather, RowGatherer, shard_bounds
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    status = "ok"
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "track_tiny.npz"))
        g = {k[2:]: z[k] for k in z.files if k.startswith("g_")}
        rng = np.random.default_rng(0)
        rep = 40                                            # several tiles per rank, uneven shards
        pos0 = np.concatenate([z["pos0"] + rng.uniform(-0.5, 0.5, z["pos0"].shape) for _ in range(rep)])
        cell0 = np.concatenate([z["jiT0"]] * rep).astype(np.int32)
        n = pos0.shape[0]
        b = shard_bounds(n, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        nrec = 16
        U, V, IC = 3 * z["U"][:nrec], 3 * z["V"][:nrec], z["IC"][:nrec]
        full = corc.track(g, U, V, IC, pos0, cell0.astype(np.int64))
        with engine_for(g, device=rank) as eng:
            eng.set_buoys(pos0[lo:hi], cell0[lo:hi])
            eng.record_slots(2)
            pg = PeerGather(eng, n, lo, f4=f4, nbuf=2, mode=mode)
            rg = RowGatherer(n, world, rank, width=2, dtype=torch.float32 if f4 else torch.float64, device=dev)
            assert np.array_equal(rg.b, b) and TILE_OK(b, rg.b)
            s_cmp, s_con = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            mk = torch.empty((hi - lo,), dtype=torch.int8, device=dev)
            na = torch.zeros((nrec,), dtype=torch.int64, device=dev)
            keep = []
            for k in range(nrec):
                st = eng.staging(k % 2)
                torch.cuda.synchronize()
                st[0], st[1], st[2] = U[k], V[k], IC[k]
                eng.submit_record(k % 2, s_cmp)
                pg.step(k % 2, k, None, mk, na[k:k + 1], s_cmp)
                row = pg.wait(s_con)
                with torch.cuda.stream(s_con):
                    keep.append(row.clone())                # the consumer: a copy on the consumer stream
                pg.release(s_con)
                if k == nrec - 1:                           # the same row through NCCL, from every rank's own block
                    s_con.synchronize()
                    rg.gather(keep[-1][lo:hi])
                    torch.cuda.synchronize()
                    if not torch.equal(rg.result(), keep[-1]):
                        status = "NCCL gather differs on rank %d" % rank
            torch.cuda.synchronize()
            if eng.gather_timed_out():
                status = "timeout"
            for k in range(nrec):
                want = full["posC"][k + 1]
                want = want.astype(np.float32) if f4 else want
                if not np.array_equal(keep[k].cpu().numpy(), want):
                    status = "mismatch at record %d on rank %d" % (k, rank)
                    break
            dist.all_reduce(na)
            if not np.array_equal(na.cpu().numpy(), full["nalive"]):
                status = "alive count mismatch"
            pg.close()
        q.put((rank, status, n, full["ncross"]))
    finally:
        dist.destroy_process_group()


def TILE_OK(b, b2):
    return all(int(x) % 64 == 0 for x in b[:-1])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["thread_stores", "copy_engines", "bulk_stores"])
@pytest.mark.parametrize("f4", [False, True])
def test_two_rank_fused_gather_over_ipc(f4, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, f4, mode, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(240)
    assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [g[1] for g in got] == ["ok", "ok"], got
    assert got[0][2] > 2000 and got[0][3] > 0
