"""Pyramid build (svo_frames_rebuild: gradient level 0 + pyrDown of both stacks, 4 levels) of n resident 1241x376 frames:
microseconds per frame and algorithmic GB/s for several batch sizes and chunk sizes (SVO_PYR_CHUNK)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
w, h = 1241, 376
K = (500.0, 500.0, w / 2, h / 2)
dims = [(w, h)]
for _ in range(3):
    dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))
per_frame = 2 * w * h + 2 * sum(a * b for a, b in dims[1:]) + sum(2 * a * b for a, b in dims[1:3])
stream = torch.cuda.Stream()
nmax = 1024
rng = np.random.default_rng(1)
frames = rng.integers(0, 256, (64, h, w), dtype=np.uint8)
with torch.cuda.stream(stream):
    with pkg.Context(w, h, K, levels=4, max_frames=nmax, max_jobs=1, max_features=16, max_fa_items=16, stream=stream.cuda_stream) as ctx:
        for i in range(0, nmax, 64):
            ctx.upload(i, frames)
        ctx.sync()
        for chunk in sys.argv[1:] or ["48"]:
            os.environ["SVO_PYR_CHUNK"] = chunk
            for n in ([int(os.environ["PYR_N"])] if "PYR_N" in os.environ else (1, 3, 16, 64, 256, 1024)):
                for _ in range(3):
                    ctx.rebuild(0, n)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record(stream)
                for _ in range(reps):
                    ctx.rebuild(0, n)
                e1.record(stream)
                torch.cuda.synchronize()
                us = 1e3 * e0.elapsed_time(e1) / reps
                print("chunk %5s n %5d: %8.1f us  %6.3f us/frame  %7.1f GB/s algorithmic" % (chunk, n, us, us / n, per_frame * n / us / 1e3))
