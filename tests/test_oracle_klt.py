"""Row f4: the oracle's restatement of cv::calcOpticalFlowPyrLK (the call of algorithm::computeOpticalFlowSparse,
/root/reference src/algorithm.cpp:60-62) against the OpenCV binary of this image and against committed golden vectors that binary
produced (tests/golden/make_golden_klt.py).  OpenCV sums the window products in float SIMD lanes, the restatement sums them
sequentially: the tolerance below covers that reordering (positions are float32 at coordinates up to 1241, ulp 1.2e-4)."""
import os
import numpy as np
import pytest

TOL_PX = 2e-3        # |dx|, |dy| between the restatement and OpenCV for points both report as tracked
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "klt_golden.npz")


def klt_case(synth, index, n, win, rng):
    pair = synth.make_pair(index=index, n_features=n)
    pts = pair["feats"]["px"][: pair["n_ref"]].astype(np.float32)
    # add points that hug the border / sit in texture-free corners so the status paths are exercised
    h, w = pair["h"], pair["w"]
    extra = np.array([[0.5, 0.5], [w - 1.0, h - 1.0], [w - 2.5, 3.25], [1.0, h - 2.0], [w / 2, 0.0]], np.float32)
    pts = np.concatenate([pts, extra, rng.uniform([0, 0], [w - 1, h - 1], size=(20, 2)).astype(np.float32)])
    return pair, pts


@pytest.mark.parametrize("win,index", [(11, 0), (7, 3), (21, 5), (8, 7)])
def test_klt_against_live_cv2(orc, synth, win, index):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(index)
    pair, pts = klt_case(synth, index, 300, win, rng)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-4)
    want, wst, werr = cv2.calcOpticalFlowPyrLK(pair["ref"], pair["cur"], pts.copy(), pts.copy(), winSize=(win, win), maxLevel=3,
                                               criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    got, gst, gerr, top = orc.klt_track(pair["ref"], pair["cur"], pts, pts, win=win)
    assert top == 3
    wst = wst.reshape(-1)
    assert (gst != wst).mean() <= 0.01              # a point exactly on a threshold may flip with the summation order
    both = (gst == 1) & (wst == 1)
    assert both.sum() > 250
    d = np.abs(got[both] - want[both]).max(axis=1)
    assert np.quantile(d, 0.99) < TOL_PX and d.max() < 0.05, (np.quantile(d, 0.99), d.max())
    assert np.abs(gerr[both] - werr.reshape(-1)[both]).max() < 0.05


def test_klt_without_initial_flow_and_small_image(orc, synth):
    cv2 = pytest.importorskip("cv2")
    pair = synth.make_pair(index=2, n_features=200)
    ref, cur = np.ascontiguousarray(pair["ref"][:90, :150]), np.ascontiguousarray(pair["cur"][:90, :150])
    rng = np.random.default_rng(1)
    pts = rng.uniform([5, 5], [145, 85], size=(120, 2)).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 1e-4)
    want, wst, _ = cv2.calcOpticalFlowPyrLK(ref, cur, pts.copy(), None, winSize=(21, 21), maxLevel=3, criteria=crit)
    got, gst, _, top = orc.klt_track(ref, cur, pts, None, win=21)
    assert top == 2                                  # 150x90 -> 75x45 -> 38x23 -> 19x12 (not larger than the window): 3 levels
    wst = wst.reshape(-1)
    both = (gst == 1) & (wst == 1)
    assert (gst != wst).mean() <= 0.02 and both.sum() > 60
    assert np.quantile(np.abs(got[both] - want[both]).max(axis=1), 0.98) < TOL_PX


def test_klt_golden(orc):
    g = np.load(GOLDEN)
    for k in range(int(g["n_cases"])):
        got, gst, gerr, _ = orc.klt_track(g["ref%d" % k], g["cur%d" % k], g["pts%d" % k], g["pts%d" % k], win=int(g["win%d" % k]))
        wst = g["status%d" % k]
        both = (gst == 1) & (wst == 1)
        assert (gst != wst).mean() <= 0.01 and both.sum() > 50
        assert np.quantile(np.abs(got[both] - g["next%d" % k][both]).max(axis=1), 0.99) < TOL_PX
