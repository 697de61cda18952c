"""Finds the evaluations where the selection tiers of the single-CTA kernel disagree (they must not): runs every pair of a
batch alone as job 0 with SVO_S5_FORCE=0 / 1 / 2 and prints the per-evaluation sigma traces of the pairs that differ."""
import importlib, os, sys
import numpy as np
os.environ["SVO_ALIGN_TRACE"] = "1"   # the per-evaluation sigma trace of job 0
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi, synth = pkg.capi, pkg.synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 37
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 24
batch = synth.make_batch(n, nf)


def trace(dbg):
    m = int(dbg[0])
    return [(np.array([dbg[1 + 2 * e]], np.int64).view(np.float64)[0], hex(int(dbg[2 + 2 * e]) & 0xffffffff), (int(dbg[2 + 2 * e]) >> 32) & 0xffffff) for e in range(min(m, 31))]


with pkg.Context(batch["w"], batch["h"], batch["K"], levels=4, max_frames=2 * n, max_jobs=n, max_features=512) as ctx:
    ctx.upload(0, batch["ref"])
    ctx.upload(n, batch["cur"])
    ident = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    bad = 0
    for p in range(n):
        j = capi.make_jobs(1)
        j["ref_slot"], j["kf_slot"], j["cur_slot"] = p, p, p + n
        j["n_ref"], j["n_kf"], j["feat_offset"] = batch["n_feat"][p], 0, 0
        j["T_ref"], j["T_kf"], j["T_cur"] = ident, ident, ident
        ff = batch["feats"][int(batch["feat_offset"][p]):int(batch["feat_offset"][p]) + int(batch["n_feat"][p])]
        out = {}
        for force in ("0", "1", "2"):
            os.environ["SVO_S5_FORCE"] = force
            res, _ = ctx.sparse_align(j, ff, mode=mode, max_iter=12)
            out[force] = (res[0].copy(), trace(ctx.debug_cycles().reshape(-1).copy()))
        r0 = out["0"][0]
        for force in ("1", "2"):
            r = out[force][0]
            if not (np.array_equal(r["T_cur"], r0["T_cur"]) and r["rmse"] == r0["rmse"] and r["evaluations"] == r0["evaluations"]):
                bad += 1
                print("pair %d force %s differs: rmse %.12g vs %.12g, evals %d vs %d" % (p, force, r["rmse"], r0["rmse"], r["evaluations"], r0["evaluations"]))
                for e, (a, b) in enumerate(zip(out["0"][1], out[force][1])):
                    flag = "" if a[0] == b[0] else "   <-- sigma differs"
                    print("   eval %2d  force0 sigma %.12g tiers %s n %d | force%s sigma %.12g tiers %s n %d%s" % (e, a[0], a[1], a[2], force, b[0], b[1], b[2], flag))
                break
    print("pairs with differences:", bad)
