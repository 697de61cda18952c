"""Golden vectors for row f4, produced by the OpenCV binary of this image (cv2.calcOpticalFlowPyrLK called with the reference's
arguments, /root/reference src/algorithm.cpp:60-62).  Run from the repo root: python tests/golden/make_golden_klt.py"""
import importlib, os, sys
import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
synth = importlib.import_module("semi-direct-visual-odometry_b200").synth
out = {}
cases = [(11, 0), (15, 4)]
for k, (win, index) in enumerate(cases):
    pair = synth.make_pair(index=index, n_features=150)
    ref, cur = np.ascontiguousarray(pair["ref"][40:200, 300:620]), np.ascontiguousarray(pair["cur"][40:200, 300:620])   # small crops
    rng = np.random.default_rng(100 + k)
    pts = rng.uniform([0, 0], [319, 159], size=(150, 2)).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-4)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(ref, cur, pts.copy(), pts.copy(), winSize=(win, win), maxLevel=3, criteria=crit,
                                            flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    out.update({"ref%d" % k: ref, "cur%d" % k: cur, "pts%d" % k: pts, "next%d" % k: nxt, "status%d" % k: st.reshape(-1),
                "err%d" % k: err.reshape(-1), "win%d" % k: win})
out["n_cases"] = len(cases)
out["cv2_version"] = cv2.__version__
np.savez_compressed(os.path.join(os.path.dirname(__file__), "klt_golden.npz"), **out)
print("wrote klt_golden.npz with", len(cases), "cases, cv2", cv2.__version__)
