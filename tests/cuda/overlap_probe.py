"""Does the frame ingest (svo_frames_prefetch) overlap a running alignment launch?"""
import importlib, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi, synth = pkg.capi, pkg.synth
n = 1024
batch = synth.make_batch(64, 500)
rep = n // 64
h, w = batch["h"], batch["w"]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = pkg.Context(w, h, batch["K"], levels=4, max_frames=3 * n, max_jobs=n, max_features=512, max_fa_items=16, stream=stream.cuda_stream)
    pin = ctx.pinned(2 * n * h * w)
    frames = pin.array.reshape(2 * n, h, w)
    for r in range(rep):
        frames[r * 64:(r + 1) * 64] = batch["ref"]
        frames[n + r * 64:n + (r + 1) * 64] = batch["cur"]
    jobs = capi.make_jobs(n)
    ident = np.array([0, 0, 0, 1, 0, 0, 0.0])
    jobs["ref_slot"], jobs["kf_slot"], jobs["cur_slot"] = np.arange(n), np.arange(n), np.arange(n) + n
    jobs["n_ref"] = np.tile(batch["n_feat"], rep); jobs["n_kf"] = 0
    feats = np.tile(batch["feats"], rep)
    jobs["feat_offset"] = np.concatenate([[0], np.cumsum(jobs["n_ref"])[:-1]])
    jobs["T_ref"], jobs["T_kf"], jobs["T_cur"] = ident, ident, ident
    kw = dict(patch_size=5, min_level=0, max_level=3, mode=2, max_iter=30)
    ctx.upload(0, frames); ctx.sync()
    ctx.sparse_align_stage(jobs, feats, **kw); ctx.sparse_align_h2d(); ctx.sync()
    def T(f, reps=3):
        ts = []
        for _ in range(reps):
            ctx.sync(); torch.cuda.synchronize(); t0 = time.perf_counter(); f(); ctx.sync(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts)
    print("launch alone: %.2f ms" % T(lambda: ctx.sparse_align_launch()))
    print("prefetch alone (other slots): %.2f ms" % T(lambda: ctx.prefetch(2 * n, frames[n:])))
    print("rebuild alone: %.2f ms" % T(lambda: ctx.rebuild(2 * n, n)))
    def both():
        ctx.sparse_align_launch()
        ctx.prefetch(2 * n, frames[n:])
    print("launch + prefetch (other slots): %.2f ms" % T(both))
    def host_only():
        t0 = time.perf_counter(); ctx.sparse_align_launch(); t1 = time.perf_counter(); ctx.prefetch(2 * n, frames[n:]); t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3
    ctx.sync(); print("host time of the calls: launch %.3f ms, prefetch %.3f ms" % host_only()); ctx.sync()
    # the pipelined loop of bench.py with host timestamps
    jobs_b = jobs.copy(); jobs_b["cur_slot"] = np.arange(n) + 2 * n
    ring = (jobs, jobs_b)
    pj = ctx.pinned(2 * jobs.nbytes + feats.nbytes + 64)
    pjobs = [pj.view(capi.ALIGN_JOB_DTYPE, n, 0), pj.view(capi.ALIGN_JOB_DTYPE, n, jobs.nbytes)]
    pjobs[0][:] = jobs; pjobs[1][:] = jobs_b
    pfeats = pj.view(capi.ALIGN_FEATURE_DTYPE, len(feats), 2 * jobs.nbytes); pfeats[:] = feats
    for variant in ("pinned jobs", "no feats h2d"):
        ctx.sync(); torch.cuda.synchronize()
        marks = []
        t00 = time.perf_counter()
        ctx.prefetch(n, frames[n:])
        steps = 12
        for k in range(steps):
            ta = time.perf_counter()
            ctx.sparse_align_stage(pjobs[k % 2], pfeats, **kw)
            tb = time.perf_counter()
            if variant == "pinned jobs" or k == 0:
                ctx.sparse_align_h2d()
            ctx.sparse_align_launch(); ctx.sparse_align_d2h()
            tc = time.perf_counter()
            if k + 1 < steps:
                ctx.prefetch(n + ((k + 1) % 2) * n, frames[n:])
            td = time.perf_counter()
            ctx.sparse_align_fetch()
            te = time.perf_counter()
            marks.append((ta - t00, tb - ta, tc - tb, td - tc, te - td))
        print(variant, "total %.2f ms/step" % ((time.perf_counter() - t00) * 1e3 / steps))
        for m in marks[:12]:
            print("  t=%.2f ms  stage %.2f  h2d+launch+d2h %.2f  prefetch %.2f  fetch-wait %.2f" % tuple(1e3 * x for x in m))
