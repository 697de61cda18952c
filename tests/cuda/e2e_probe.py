import importlib, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi, synth = pkg.capi, pkg.synth
n = 1024
batch = synth.make_batch(64, 500)
# replicate 64 distinct pairs 16x (probe only cares about copy/launch timing)
rep = n // 64
h, w = batch["h"], batch["w"]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = pkg.Context(w, h, batch["K"], levels=4, max_frames=2 * n, max_jobs=n, max_features=512, max_fa_items=16, stream=stream.cuda_stream)
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
    def T(f, reps=3):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts)
    print("upload 1024 cur frames (pinned 2D copy + pyramid): %.2f ms" % T(lambda: ctx.upload(n, frames[n:])))
    print("pyramid rebuild only 1024 frames: %.2f ms" % T(lambda: ctx.rebuild(n, n)))
    print("stage (host memcpy jobs+feats to pinned): %.2f ms" % T(lambda: ctx.sparse_align_stage(jobs, feats, **kw)))
    print("h2d jobs+feats: %.2f ms" % T(lambda: ctx.sparse_align_h2d()))
    print("launch: %.2f ms" % T(lambda: ctx.sparse_align_launch()))
    print("d2h: %.2f ms" % T(lambda: ctx.sparse_align_d2h()))
    print("fetch: %.2f ms" % T(lambda: ctx.sparse_align_fetch()))
    print("full sparse_align: %.2f ms" % T(lambda: ctx.sparse_align(jobs, feats, want_stats=False, **kw)))
    # raw copies for reference
    a = torch.empty(n * h * w, dtype=torch.uint8, device="cuda")
    src = torch.from_numpy(frames[n:].reshape(-1))
    print("torch 1D pinned->device 478MB: %.2f ms" % T(lambda: a.copy_(src, non_blocking=True)))
    pag = np.array(frames[n:])  # pageable copy
    print("upload from pageable (staged): %.2f ms" % T(lambda: ctx.upload(n, pag)))
    # ingest pipeline pieces
    print("prefetch 1024 cur frames (ingest streams): %.2f ms" % T(lambda: ctx.prefetch(n, frames[n:])))
    half = torch.empty(n * h * w // 2, dtype=torch.uint8, device="cuda")
    for mb in (4, 15, 64, 256):
        nb = mb << 20
        chunks = [(i, min(nb, src.numel() - i)) for i in range(0, src.numel(), nb)]
        def f():
            for o, m in chunks:
                a[o:o + m].copy_(src[o:o + m], non_blocking=True)
        print("torch pinned->device 478MB in %d MB chunks: %.2f ms" % (mb, T(f)))
