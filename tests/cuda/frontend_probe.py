"""Per-frame front end (svo_frontend_run): wall clock per call, with and without the staging copy of the image; run it
under `ncu --metrics gpu__time_duration.sum` for the per-node times of the captured graph."""
import ctypes as C
import importlib, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi, synth = pkg.capi, pkg.synth
pair = synth.make_pair(0, 500)
h, w = pair["h"], pair["w"]
with pkg.Context(w, h, pair["K"], levels=4, max_frames=4, max_jobs=1, max_features=512, max_fa_items=512) as ctx:
    ctx.upload(0, pair["ref"])
    job = capi.make_jobs(1)
    job[0]["n_ref"], job[0]["n_kf"] = pair["n_ref"], 0
    job[0]["T_ref"], job[0]["T_kf"], job[0]["T_cur"] = pair["T_ref"], pair["T_kf"], pair["T_cur_init"]
    kw = dict(cell=30, thr=50, max_features=512, mode=capi.LM_FAITHFUL, fa_patch=7, fa_mode=capi.LM_FAITHFUL)
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for _ in range(3):
        out, sel, ref = ctx.frontend_run(pair["cur"], job, pair["feats"], 0, 0, 1, **kw)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.frontend_run(pair["cur"], job, pair["feats"], 0, 0, 1, **kw)
        ts.append((time.perf_counter() - t0) * 1e6)
    print("frontend_run, image from pageable memory: median %.1f us, min %.1f us" % (np.median(ts), np.min(ts)))
    # zero-copy: the caller writes the image straight into the context's page-locked buffer
    buf = ctx.L.svo_frontend_image_buffer(ctx.h)
    pinned = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(h, w))
    pinned[:] = pair["cur"]
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.frontend_run(pinned, job, pair["feats"], 0, 0, 1, **kw)
        ts.append((time.perf_counter() - t0) * 1e6)
    print("frontend_run, image already in the page-locked buffer: median %.1f us, min %.1f us" % (np.median(ts), np.min(ts)))
    print("selected %d, candidates %d, status %d" % (out["n_selected"], out["n_candidates"], out["align"]["status"]))
