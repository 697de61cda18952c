"""Bare host->device bandwidth with all ranks copying at once (the ceiling of bench.py's e2e figure): every rank owns one GPU,
allocates `--mb` MB of page-locked host memory and copies it to its device `--reps` times after a barrier; prints per-rank
and aggregate GB/s.  Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/cuda/h2d_probe.py
(or plain python for one GPU).  Also times the same copy while the GPU runs a memory-bound kernel (a device-to-device copy
loop), since the alignment kernel shares the device with the ingest."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=515)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = a.mb * 1000 * 1000
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()


def run(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        e1.record(stream)
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(2)
bare = run(a.reps)
# the same while the SMs and HBM are busy
busy_a, busy_b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"), torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(400):
        busy_b.copy_(busy_a)
loaded = run(a.reps)
torch.cuda.synchronize()
vals = torch.tensor([bare, loaded], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    b = [float(v[0]) for v in allv]
    l = [float(v[1]) for v in allv]
    print(json.dumps({"probe": "concurrent pinned H2D", "n_gpus": world, "mb_per_copy": a.mb, "reps": a.reps,
                      "bare_gbs_per_gpu": [round(x, 2) for x in b], "bare_gbs_aggregate": round(sum(b), 1),
                      "under_load_gbs_per_gpu": [round(x, 2) for x in l], "under_load_gbs_aggregate": round(sum(l), 1),
                      "cpu_count": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
