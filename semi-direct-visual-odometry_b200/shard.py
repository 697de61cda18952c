"""Multi-GPU partitioning of the alignment path: independent frame pairs, contiguous blocks per rank, no
collective during alignment, ONE final gather of the 80-byte pose records.  Pure host logic over torch.distributed:
the same code runs over NCCL on device buffers (bench.py) and over gloo on CPU tensors (tests)."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous block of pair indices [lo, hi) owned by `rank`; block sizes differ by at most one, the first
    n_total % world blocks are the larger ones.  The same arithmetic as svo_multi_shard (csrc/multi.cu), which the
    single-process C++ multi-GPU layer uses; tests/test_shard_gloo.py checks the two against each other."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    q, r = divmod(n_total, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def owner_of(pair, n_total, world):
    q, r = divmod(n_total, world)
    big = r * (q + 1)
    return pair // (q + 1) if pair < big else r + (pair - big) // q


def gather_records(local_bytes, dist, dst=0, out=None):
    """local_bytes: 1-D uint8 torch tensor (this rank's packed result records, device or CPU).  Every rank must pass
    the same length (pad the last shard).  Returns the list of per-rank tensors on `dst`, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    bufs = None
    if rank == dst:
        bufs = out if out is not None else [torch.empty_like(local_bytes) for _ in range(world)]
    dist.gather(local_bytes, bufs, dst=dst)
    return bufs


def assemble(bufs, dtype, n_total, world):
    """Rank-ordered numpy view of the gathered records, trimmed to n_total (padding of the last shard dropped)."""
    parts = []
    for r, b in enumerate(bufs):
        lo, hi = shard_range(n_total, r, world)
        a = np.frombuffer(b.cpu().numpy().tobytes(), dtype=dtype)
        parts.append(a[:hi - lo])
    return np.concatenate(parts)
