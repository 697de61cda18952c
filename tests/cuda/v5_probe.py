"""A/B probe of the single-CTA alignment kernel (sparse_align_v5.cu, select5.cuh) against the cluster kernel
(sparse_align_v3.cu, SVO_ALIGN_V4=0) and the oracle, with every selection tier forced in turn (SVO_S5_FORCE): level stats,
evaluation counts, the sigma trace of job 0, single-pair latency and one-wave / batch throughput.

    python tests/cuda/v5_probe.py [n_pairs_for_timing]
"""
import importlib
import os
import sys
import time

import numpy as np
os.environ["SVO_ALIGN_TRACE"] = "1"   # the per-evaluation sigma trace of job 0

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/oracle")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
capi, synth = pkg.capi, pkg.synth
import oracle as orc


def run(ctx, jobs, feats, v4, mode, force=0, **kw):
    os.environ["SVO_ALIGN_V4"] = "1" if v4 else "0"   # (v4=True: the kernel under test, v5; else the cluster kernel)
    os.environ["SVO_S5_FORCE"] = str(force)
    res, st = ctx.sparse_align(jobs, feats, mode=mode, max_iter=30, **kw)
    dbg = ctx.debug_cycles().reshape(-1).copy()
    return res, st, dbg


def trace(dbg):
    n = int(dbg[0])
    out = []
    for e in range(min(n, 31)):
        sig = np.array([dbg[1 + 2 * e]], np.int64).view(np.float64)[0]
        t = int(dbg[2 + 2 * e])
        out.append((sig, (t & 0xffffff) | ((t >> 56) << 24), (t >> 24) & 0xff, (t >> 32) & 0xffffff))
    return out


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    ok_all = True
    for nfeat, idx in ((500, 0), (500, 3), (120, 5), (65, 7)):
        pair = synth.make_pair(idx, nfeat)
        with pkg.Context(pair["w"], pair["h"], pair["K"], levels=4, max_frames=2, max_jobs=1, max_features=512) as ctx:
            ctx.upload(0, np.stack([pair["ref"], pair["cur"]]))
            j = capi.make_jobs(1)
            j[0]["ref_slot"], j[0]["kf_slot"], j[0]["cur_slot"] = 0, 0, 1
            j[0]["n_ref"], j[0]["n_kf"] = pair["n_ref"], 0
            j[0]["T_ref"], j[0]["T_kf"], j[0]["T_cur"] = pair["T_ref"], pair["T_kf"], pair["T_cur_init"]
            rp, _ = orc.build_pyramid(pair["ref"], 4)
            cp, _ = orc.build_pyramid(pair["cur"], 4)
            for mode in (2, 0, 1):
                r3, s3, _ = run(ctx, j, pair["feats"], False, mode)
                r4, s4, d4 = run(ctx, j, pair["feats"], True, mode)
                for force in (1, 2):
                    rf, sf, _ = run(ctx, j, pair["feats"], True, mode, force=force)
                    same = np.array_equal(sf["sigma"], s4["sigma"]) and np.array_equal(rf["T_cur"], r4["T_cur"]) and \
                        np.array_equal(sf["evaluations"], s4["evaluations"])
                    print("  forced tier %d: %s (tiers %s)" % (force, "bit-identical to the unforced run" if same else "DIFFERS", hex(rf[0]["reserved"])))
                    ok_all = ok_all and same
                rmse, T, status, lv = orc.sparse_align(rp, rp, cp, pair["w"], pair["h"], pair["feats"], pair["n_ref"], 0, pair["T_ref"],
                                                       pair["T_kf"], pair["K"], pair["T_cur_init"], patch_size=5, mode=mode, max_iter=30)
                print("--- features %d pair %d mode %d: evals v3 %d v5 %d oracle %d; tiers v3 %s v5 %s" % (
                    nfeat, idx, mode, r3[0]["evaluations"], r4[0]["evaluations"], sum(l["evaluations"] for l in lv),
                    hex(r3[0]["reserved"]), hex(r4[0]["reserved"])))
                for s in range(4):
                    a, b, o = s3[0, s], s4[0, s], lv[s]
                    hrel = np.abs(b["H"] - o["H"]).max() / np.abs(o["H"]).max()
                    print("  level %d: sigma v3 %.9f v5 %.9f orc %.9f | n_px %d %d %d | evals %d %d %d | status %d %d %d | H rel(v5,orc) %.1e | "
                          "dpose(v5,orc) %.1e" % (s, a["sigma"], b["sigma"], o["sigma"], a["n_px"], b["n_px"], o["n_px"], a["evaluations"],
                                                  b["evaluations"], o["evaluations"], a["status"], b["status"], o["status"], hrel,
                                                  np.abs(b["pose_after"] - o["pose_after"]).max()))
                    if abs(a["sigma"] - b["sigma"]) > 2e-6 * b["sigma"] or a["n_px"] != b["n_px"] or b["evaluations"] != o["evaluations"]:
                        ok_all = False
                dT = np.abs(r4[0]["T_cur"] - T)
                print("  final |dT| v5-orc %.2e, v3-orc %.2e" % (dT.max(), np.abs(r3[0]["T_cur"] - T).max()))
                if mode == 2:
                    print("  sigma trace v5 (sigma, tiers miss|cold<<8|generic<<16|hit<<24, attempts, n):", [("%.6f" % s, hex(t), hex(w), n) for s, t, w, n in trace(d4)])
    print("A/B", "sigma within 2e-6 of v3, n_px / evaluation counts identical, forced tiers bit-identical" if ok_all else "DIFFERENCES (see above)")

    # ---- timing: single pair and a batch ----
    import torch
    batch = synth.make_batch(npairs, 500)
    n = npairs
    F = int(batch["n_feat"].max())
    with pkg.Context(batch["w"], batch["h"], batch["K"], levels=4, max_frames=2 * n, max_jobs=n, max_features=F) as ctx:
        frames = np.concatenate([batch["ref"], batch["cur"]])
        ctx.upload(0, frames)
        jobs = capi.make_jobs(n)
        ident = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
        jobs["ref_slot"], jobs["kf_slot"], jobs["cur_slot"] = np.arange(n), np.arange(n), np.arange(n) + n
        jobs["n_ref"], jobs["n_kf"], jobs["feat_offset"] = batch["n_feat"], 0, batch["feat_offset"]
        jobs["T_ref"], jobs["T_kf"], jobs["T_cur"] = ident, ident, ident
        for v4 in (False, True):
            os.environ["SVO_ALIGN_V4"] = "1" if v4 else "0"
            os.environ["SVO_S5_FORCE"] = "0"
            for label, jj, ff in (("single", jobs[:1], batch["feats"][:int(batch["n_feat"][0])]), ("batch %d" % n, jobs, batch["feats"])):
                ctx.sparse_align_stage(jj, ff, patch_size=5, min_level=0, max_level=3, mode=2, max_iter=30)
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
                tiers = res["reserved"].astype(np.int64)
                print("%s %s: %.1f us per launch, evals/pair %.2f, tiers hit/bracket %d miss/hot %d cold %d generic %d" % (
                    "v5" if v4 else "v3", label, dt * 1e6, res["evaluations"].mean(), ((tiers >> 24) & 0xff).sum(), (tiers & 0xff).sum(),
                    ((tiers >> 8) & 0xff).sum(), ((tiers >> 16) & 0xff).sum()))
            rot = np.array([synth.rotation_angle(res[i]["T_cur"], batch["T_true"][i]) for i in range(n)])
            print("   median rot err %.2e" % np.median(rot))


if __name__ == "__main__":
    main()
