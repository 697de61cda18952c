"""Single-pair latency of the alignment kernel with parts of it disabled (SVO_S5_FORCE=3: constant robust scale, no
selection) -- timing experiments only.  Prints microseconds per launch and per evaluation."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi, synth = pkg.capi, pkg.synth
n = 8
batch = synth.make_batch(n, 500)
F = int(batch["n_feat"].max())
with pkg.Context(batch["w"], batch["h"], batch["K"], levels=4, max_frames=2 * n, max_jobs=n, max_features=F) as ctx:
    ctx.upload(0, np.concatenate([batch["ref"], batch["cur"]]))
    jobs = capi.make_jobs(n)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    jobs["ref_slot"], jobs["kf_slot"], jobs["cur_slot"] = np.arange(n), np.arange(n), np.arange(n) + n
    jobs["n_ref"], jobs["n_kf"], jobs["feat_offset"] = batch["n_feat"], 0, batch["feat_offset"]
    jobs["T_ref"], jobs["T_kf"], jobs["T_cur"] = ident, ident, ident
    for force in sys.argv[1:] or ["0", "3"]:
        os.environ["SVO_S5_FORCE"] = force
        for mode, name in ((2, "GN"), (0, "faithful")):
            tot_us, tot_ev = 0.0, 0
            for p in range(n):
                jj = jobs[p:p + 1].copy()
                jj["feat_offset"] = 0
                ff = batch["feats"][int(batch["feat_offset"][p]):int(batch["feat_offset"][p]) + int(batch["n_feat"][p])]
                ctx.sparse_align_stage(jj, ff, patch_size=5, min_level=0, max_level=3, mode=mode, max_iter=30)
                ctx.sparse_align_h2d()
                for _ in range(3):
                    ctx.sparse_align_launch()
                ctx.sync()
                reps = 20
                t0 = time.perf_counter()
                for _ in range(reps):
                    ctx.sparse_align_launch()
                ctx.sync()
                dt = (time.perf_counter() - t0) / reps
                ctx.sparse_align_d2h()
                res = ctx.sparse_align_fetch()[0]
                tot_us += dt * 1e6
                tot_ev += int(res[0]["evaluations"])
            print("force %s %-8s: %.1f us per pair, %.2f evaluations per pair, %.2f us per evaluation" % (
                force, name, tot_us / n, tot_ev / n, tot_us / tot_ev))
