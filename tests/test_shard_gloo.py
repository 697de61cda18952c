"""N > 1 host logic on CPU: two processes over gloo shard 37 pairs, each fills the result records of its own pairs,
one gather brings them to rank 0 in pair order -- the same code path bench.py drives over NCCL."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("semi-direct-visual-odometry_b200")
    shard, capi = pkg.shard, pkg.capi
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_total, rank, world)
    per = -(-n_total // world)
    rec = np.zeros(per, capi.ALIGN_RESULT_DTYPE)          # padded to the common shard size
    for k, i in enumerate(range(lo, hi)):                   # a stand-in for the alignment: pose encodes the pair index
        rec[k]["T_cur"] = (0, 0, 0, 1, i, 2 * i, -i)
        rec[k]["rmse"], rec[k]["status"], rec[k]["evaluations"] = 0.5 * i, i % 3, 20 + i
    local = torch.from_numpy(rec.view(np.uint8).copy())
    bufs = shard.gather_records(local, dist, dst=0)
    if rank == 0:
        out = shard.assemble(bufs, capi.ALIGN_RESULT_DTYPE, n_total, world)
        q.put(out.tobytes())
    else:
        assert bufs is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges(pkg):
    shard = pkg.shard
    for n, w in [(8192, 8), (8192, 4), (1024, 2), (37, 2), (5, 8), (0, 2)]:
        seen = []
        for r in range(w):
            lo, hi = shard.shard_range(n, r, w)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi))
            assert all(shard.owner_of(i, n, w) == r for i in range(lo, hi))
        assert seen == list(range(n))   # contiguous blocks, every pair exactly once, in order
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_shard_ranges_match_the_c_layer(pkg):
    """svo_multi_shard (the C++ multi-GPU layer of libsvo_b200.so) and shard.shard_range (the torch.distributed path) cut a
    batch the same way."""
    for n, w in [(8192, 8), (8192, 3), (1024, 2), (37, 2), (5, 8), (0, 2), (1000, 7)]:
        for r in range(w):
            assert pkg.capi.shard(n, w, r) == pkg.shard.shard_range(n, r, w)


def test_two_rank_gather_gloo(pkg):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    raw = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    out = np.frombuffer(raw, dtype=pkg.capi.ALIGN_RESULT_DTYPE)
    assert len(out) == n_total
    assert np.array_equal(out["T_cur"][:, 4], np.arange(n_total))
    assert np.array_equal(out["evaluations"], 20 + np.arange(n_total))
    assert np.array_equal(out["status"], np.arange(n_total) % 3)
