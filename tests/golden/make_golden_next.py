"""Golden vectors for the "next" rows, made by the INDEPENDENT numpy / python restatements in
tests/test_oracle_numerics.py (not by the oracle, not by the CUDA path).  Run from the repo root:
    python tests/golden/make_golden_next.py

next_rows_golden.npz
  ssc_*   FeatureSelection::gradientMagnitudeWithSSC on the gradient of a 160x96 crop (python loop over the stably
          sorted keypoints), two parameter sets
  epi_*   algorithm::matchEpipolarConstraint for 24 seeds of a seeded synthetic pair (numpy restatement), both mean modes
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
synth = pkg.synth
from test_oracle_numerics import _np_epipolar, _py_ssc  # noqa: E402


def main():
    pair = synth.make_pair(index=7, n_features=101)
    img = np.ascontiguousarray(pair["ref"][100:196, 300:460])
    grad = synth.abs_gradient_np(img)
    out = {"img": img, "grad": grad}
    for tag, (thr, k, cell, bucket) in {"a": (60, 60, 16, True), "b": (30, 150, 16, False)}.items():
        feats, info = _py_ssc(grad, thr, k, cell, use_bucketing=bucket)
        out["ssc_%s_params" % tag] = np.array([thr, k, cell, int(bucket)])
        out["ssc_%s_feats" % tag] = feats
        out["ssc_%s_info" % tag] = np.array([info["keypoints"], info["width"], info["iterations"], info["ssc_points"]])
    wide = synth.make_pair(index=8, n_features=120, motion_scale=3.0)
    T_rel = synth.se3_mul(wide["T_cur_true"], synth.se3_inv(wide["T_ref"]))
    rng = np.random.default_rng(99)
    rows = []
    for i in range(0, 120, 5):
        f = wide["feats"][i]
        d = np.linalg.norm(f["point"])
        lo, hi, d0 = d * rng.uniform(0.4, 0.9), d * rng.uniform(1.1, 3.0), d * rng.uniform(0.8, 1.25)
        for mode in (True, False):
            r = _np_epipolar(wide["ref"], wide["cur"], wide["K"], T_rel, f["px"], f["bearing"], d0, lo, hi, eigen_mean=mode)
            rows.append([i, int(mode), d0, lo, hi, float(r["found"]), r["depth"] if r["found"] else 0.0, r["px"][0], r["px"][1],
                         r["steps"], r["score"] if r["score"] is not None else -1.0])
    out["epi_index"], out["epi_motion_scale"] = 8, 3.0
    out["epi_T_rel"] = T_rel
    out["epi_rows"] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "next_rows_golden.npz"), **out)
    print("wrote next_rows_golden.npz")


if __name__ == "__main__":
    main()
