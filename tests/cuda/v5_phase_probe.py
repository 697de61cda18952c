"""Phase cycle counters of the single-CTA alignment kernel (library built with `make prof`, selected with
SVO_B200_LIB=.../libsvo_b200_prof.so): job 0 of a one-pair launch."""
import importlib, os, sys, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
NAMES = ["bracket pass", "rank", "count pass", "locate", "wait(bracket)", "wait(rank)", "wait(count)", "-", "#bracket passes", "#count passes", "#ranks", "generic", "-"]
os.environ["SVO_ALIGN_V4"] = "1"   # the default single-CTA kernel (sparse_align_v5.cu)
for idx in (0, 3):
    pair = pkg.synth.make_pair(idx, 500)
    for mode in (2, 0):
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
            print("pair", idx, "mode", mode, "evals", res[0]["evaluations"], "tiers", hex(res[0]["reserved"]))
            print("per level [warp + sample, sigma, -, sums, reduce + solve, level set-up, n_eval]:")
            print(d[:, :7])
            tot = d[:, :6].sum(0); tot[5] = 0; n = d[:, 6].sum()
            print("cycles per evaluation:", (tot / n).round(0), "total", (tot.sum() / n).round(0), "=> us/eval %.2f" % (tot.sum() / n / 1965))
            sel = dd.reshape(-1)[32:45]
            print("solve cycles per level (thread 0):", dd.reshape(-1)[48:52], "per evaluation %.0f" % (dd.reshape(-1)[48:52].sum() / n), "| of which LDLT %.0f, exp map + pose %.0f, quaternion -> R %.0f" % tuple(dd.reshape(-1)[52:55] / n))
            print("selection cycles (total over the pair):", ", ".join("%s %d" % (nm, v) for nm, v in zip(NAMES, sel)))
