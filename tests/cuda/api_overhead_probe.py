"""Host-side cost of the small synchronous entry points: wall clock of a call with ONE item (kernel time ~0) next to the full
call, so the split between launch/copy/synchronise overhead and kernel time is visible."""
import importlib, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi = pkg.capi
pair = pkg.synth.make_pair(0, 500)
w, h = pair["w"], pair["h"]

def wall(fn, reps=200):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))

with pkg.Context(w, h, pair["K"], levels=4, max_frames=4, max_jobs=1, max_features=512, max_fa_items=4096) as ctx:
    ctx.upload(0, np.stack([pair["ref"], pair["cur"]]))
    pts = pair["feats"]["px"][: pair["n_ref"]].astype(np.float32)
    for n in (1, 100, len(pts)):
        p = np.ascontiguousarray(pts[:n])
        print("klt_track      n=%4d  %.1f us" % (n, wall(lambda: ctx.klt_track(0, 1, p, p, win=11))))
    rng = np.random.default_rng(0)
    items = np.zeros(2000, capi.FA_ITEM_DTYPE)
    items["ref_slot"], items["cur_slot"] = 0, 1
    items["ref_px"] = items["px"] = rng.uniform([20, 20], [w - 20, h - 20], (2000, 2))
    items["A"] = (1, 0, 0, 1)
    for n in (1, 2000):
        it = items[:n]
        print("feature_align  n=%4d  %.1f us" % (n, wall(lambda: ctx.feature_align(it))))
    print("select_grid            %.1f us" % wall(lambda: ctx.select_grid(0, 30, 50)))
