import importlib, os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
import oracle as orc
capi, synth = pkg.capi, pkg.synth
idx, nf, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pair = synth.make_pair(idx, nf)
rp, _ = orc.build_pyramid(pair["ref"], 4); cp, _ = orc.build_pyramid(pair["cur"], 4)
rmse, T, status, lv = orc.sparse_align(rp, rp, cp, pair["w"], pair["h"], pair["feats"], pair["n_ref"], 0, pair["T_ref"], pair["T_kf"], pair["K"], pair["T_cur_init"], patch_size=5, mode=mode, max_iter=30)
with pkg.Context(pair["w"], pair["h"], pair["K"], levels=4, max_frames=2, max_jobs=1, max_features=512) as ctx:
    ctx.upload(0, np.stack([pair["ref"], pair["cur"]]))
    j = capi.make_jobs(1)
    j[0]["ref_slot"], j[0]["kf_slot"], j[0]["cur_slot"] = 0, 0, 1
    j[0]["n_ref"], j[0]["n_kf"] = pair["n_ref"], 0
    j[0]["T_ref"], j[0]["T_kf"], j[0]["T_cur"] = pair["T_ref"], pair["T_kf"], pair["T_cur_init"]
    for v4 in ("0", "1"):
        os.environ["SVO_ALIGN_V4"] = v4
        res, st = ctx.sparse_align(j, pair["feats"], mode=mode, max_iter=30)
        print("v4" if v4 == "1" else "v3", "tiers", hex(int(res[0]["reserved"]) & 0xffffffff), "rot err vs oracle %.3e" % synth.rotation_angle(res[0]["T_cur"], T), "dt %.3e" % np.abs(res[0]["T_cur"][4:] - T[4:]).max())
        for s in range(4):
            b, o = st[0, s], lv[s]
            print("   level %d sigma %.9f orc %.9f n_px %d %d H rel %.2e g rel %.2e chi2 rel %.2e dx rel %.2e" % (s, b["sigma"], o["sigma"], b["n_px"], o["n_px"],
                  np.abs(b["H"] - o["H"]).max() / np.abs(o["H"]).max(), np.abs(b["g"] - o["g"]).max() / np.abs(o["g"]).max(), abs(b["chi2"] - o["chi2"]) / o["chi2"], np.abs(b["dx"] - o["dx"]).max() / np.abs(o["dx"]).max()))
