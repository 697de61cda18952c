"""Phase cycle counters of the alignment fast path (library built with `make PROFILE=1`): job 0 of a one-pair launch."""
import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
pair = pkg.synth.make_pair(0, 500)
NAMES = ["push", "pop A", "finish A", "locate A", "pop B", "finish B", "locate B", "priv count", "priv reduce", "priv finish+scan", "other"]
for mode in (2, 1, 0):
    with pkg.Context(pair["w"], pair["h"], pair["K"], levels=4, max_frames=2, max_jobs=1, max_features=512) as ctx:
        ctx.upload(0, np.stack([pair["ref"], pair["cur"]]))
        j = pkg.capi.make_jobs(1)
        j[0]["ref_slot"], j[0]["kf_slot"], j[0]["cur_slot"] = 0, 0, 1
        j[0]["n_ref"], j[0]["n_kf"] = pair["n_ref"], 0
        j[0]["T_ref"], j[0]["T_kf"], j[0]["T_cur"] = pair["T_ref"], pair["T_kf"], pair["T_cur_init"]
        for _ in range(3):
            res, st = ctx.sparse_align(j, pair["feats"], mode=mode, max_iter=30)
        dd = ctx.debug_cycles()
        d = dd[:4]
        print("mode", mode, "evals", res[0]["evaluations"], "tiers", hex(res[0]["reserved"]))
        print("per level [wait for the solve + warp + sample, sigma, -, sums, reduce, -, n_eval]:")
        print(d[:, :7])
        tot = d[:, :6].sum(0); n = d[:, 6].sum()
        print("cycles per evaluation:", (tot / n).round(0), "total", (tot.sum() / n).round(0), "=> us/eval %.2f" % (tot.sum() / n / 1965))
        sel = dd.reshape(-1)[32:43]
        print("selection cycles per evaluation:", ", ".join("%s %.0f" % (nm, v / n) for nm, v in zip(NAMES, sel)))
