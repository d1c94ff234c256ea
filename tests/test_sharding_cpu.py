"""CPU (-m "not gpu"): the N>1 host logic -- frame sharding and the reporting gather -- with world_size-2 gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance(tsd):
    from tsd_b200 import sharding
    rng = np.random.default_rng(0)
    for F, world in ((1, 2), (7, 2), (64, 4), (100, 8), (3, 8)):
        off = np.concatenate([[0], np.cumsum(rng.integers(0, 400, F))])
        b = sharding.shard_bounds(off, world)
        assert b[0][0] == 0 and b[-1][1] == F and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        if F >= 8 * world:
            loads = [off[hi] - off[lo] for lo, hi in b]
            assert max(loads) - min(loads) <= 2 * 400
    fr = np.arange(10)[:, None]
    bx = np.arange(40).reshape(-1, 1).repeat(4, 1); off = np.arange(11) * 4
    parts = [sharding.local_shard(fr, bx, off, r, 3) for r in range(3)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), fr)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), bx)
    assert all(p[2][0] == 0 and p[2][-1] == len(p[1]) for p in parts)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import tsd_b200
    from tsd_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    det = np.zeros(3 if rank == 0 else 0 if world > 2 and rank == 1 else 5, tsd_b200.DET_DTYPE)
    det["frame"] = np.arange(len(det)); det["x1"] = 100 * rank + np.arange(len(det)); det["id"] = rank + 1
    out = sharding.gather_detections(det, frame_base=10 * rank, dist=dist)
    q.put((rank, out.tobytes(), len(out)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_detections_gloo_world2(tsd):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == 8
    rec = np.frombuffer(res[0][1], tsd.DET_DTYPE)
    assert rec["frame"].tolist() == [0, 1, 2, 10, 11, 12, 13, 14] and rec["id"].tolist() == [1] * 3 + [2] * 5
    assert rec["x1"].tolist() == [0, 1, 2, 100, 101, 102, 103, 104]
