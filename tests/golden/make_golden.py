"""Generates the committed golden fixtures.  Run from the repo root: python tests/golden/make_golden.py

pyramid_golden.npz  -- a 160x96 crop-sized synthetic frame; image stack by cv2.pyrDown (the live third-party
                       implementation the reference calls, src/image_pyramid.cpp:49-50), gradient level 0 by the
                       numpy statement of Simd::AbsGradientSaturatedSum, lower gradient levels by cv2.pyrDown,
                       grid argmax by numpy.  Nothing here comes from the oracle or the CUDA path.
align_golden.npz    -- per-level H / g / chi2 / sigma / lambda / dx / pose of ImageAlignment::align and results of
                       FeatureAlignment::align on a seeded synthetic pair, as produced by the ORACLE (the reference
                       cannot be built here, so these pin the oracle against accidental drift, not against a
                       reference binary: "parity unpinned").
"""
import importlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
synth = pkg.synth


def main():
    pair = synth.make_pair(index=7, n_features=101)
    img = np.ascontiguousarray(pair["ref"][100:196, 300:460])
    out = {"img": img}
    ci, cg = img, synth.abs_gradient_np(img)
    for l in range(4):
        out["img_l%d" % l], out["grad_l%d" % l] = ci, cg
        ci, cg = cv2.pyrDown(ci), cv2.pyrDown(cg)
    out["select_30_50"] = synth.grid_argmax_np(out["grad_l0"], 30, 50)
    np.savez_compressed(os.path.join(HERE, "pyramid_golden.npz"), **out)

    import oracle as orc
    rp, rg = orc.build_pyramid(pair["ref"], 4)
    cp, cgp = orc.build_pyramid(pair["cur"], 4)
    al = {"index": 7, "n_features": 101}
    for name, mode in (("faithful", orc.LM_FAITHFUL), ("lm", orc.LM_ITERATED), ("gn", orc.GN)):
        rmse, T, st, lv = orc.sparse_align(rp, rp, cp, pair["w"], pair["h"], pair["feats"], pair["n_ref"], 0, pair["T_ref"],
                                           pair["T_kf"], pair["K"], pair["T_cur_init"], mode=mode, max_iter=30)
        al[name + "_rmse"], al[name + "_T"], al[name + "_final_status"] = rmse, T, st
        for k in ("H", "g", "dx", "chi2", "sigma", "lam", "pose_after", "n_px", "status", "iterations"):
            al[name + "_" + k] = np.array([l[k] for l in lv])
    gref = orc.unpack_pyramid(rg, pair["w"], pair["h"], 4)[0]
    gcur = orc.unpack_pyramid(cgp, pair["w"], pair["h"], 4)[0]
    rng = np.random.default_rng(7)
    fa = []
    for i in range(30, 70):  # rows 0..: includes border features (NaN rmse) and interior ones
        px = pair["feats"]["px"][i]
        start = px + rng.uniform(-1.5, 1.5, 2)
        rmse, p, st, it = orc.feature_align(gref, gcur, px, start, patch_size=7, mode=orc.LM_FAITHFUL)
        fa.append([px[0], px[1], start[0], start[1], p[0], p[1], rmse, st, it])
    al["fa"] = np.array(fa)
    np.savez_compressed(os.path.join(HERE, "align_golden.npz"), **al)
    print("wrote goldens")


if __name__ == "__main__":
    main()
