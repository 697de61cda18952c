"""Pins the oracle's floating-point building blocks against analytic known answers (SURVEY 8c vi) and its
optimiser behaviour against the quirks of the reference."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_project2d_reference_kat(orc):
    # the reference's own known-answer test, tests/test_camera.cpp:83-96
    uv = orc.project2d((30.3, 40.4, 325.5, 248.8), (17.7, 28.8, 39.9))
    assert abs(uv[0] - 338.9413533834586466165) < 1e-12 and abs(uv[1] - 277.9609022556390977443) < 1e-12


def test_image_jac_matches_central_differences(orc):
    # d pi(exp(xi) p) / d xi at xi = 0, the derivation of python/symbol.py:50-60
    rng = np.random.default_rng(3)
    fx, fy = 721.5377, 700.0
    for _ in range(20):
        p = np.array([rng.uniform(-5, 5), rng.uniform(-3, 3), rng.uniform(4, 30)])
        J = orc.image_jac(p, fx, fy)
        num = np.zeros((2, 6))
        for k in range(6):
            e = np.zeros(6)
            h = 1e-6
            e[k] = h
            pp, pm = orc.se3_act(orc.se3_exp(e), p), orc.se3_act(orc.se3_exp(-e), p)
            num[:, k] = [(fx * pp[0] / pp[2] - fx * pm[0] / pm[2]) / (2 * h), (fy * pp[1] / pp[2] - fy * pm[1] / pm[2]) / (2 * h)]
        assert np.allclose(J, num, rtol=1e-6, atol=1e-5)


def test_se3_exp_against_scipy(orc):
    from scipy.linalg import expm
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(4)
    for scale in (1e-12, 1e-6, 1e-2, 1.0, 3.0):
        xi = rng.normal(size=6) * scale
        T = orc.se3_exp(xi)
        w = xi[3:]
        M = np.zeros((4, 4))
        M[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        M[:3, 3] = xi[:3]
        E = expm(M)
        assert np.allclose(Rotation.from_quat(T[:4]).as_matrix(), E[:3, :3], atol=1e-12)
        assert np.allclose(T[4:], E[:3, 3], atol=1e-12 * max(1, scale))
    # group laws
    a, b = orc.se3_exp(rng.normal(size=6)), orc.se3_exp(rng.normal(size=6))
    p = rng.normal(size=3)
    assert np.allclose(orc.se3_act(orc.se3_mul(a, b), p), orc.se3_act(a, orc.se3_act(b, p)), atol=1e-12)
    assert np.allclose(orc.se3_act(orc.se3_inv(a), orc.se3_act(a, p)), p, atol=1e-12)


def test_bilinear(orc):
    img = np.arange(48, dtype=np.uint8).reshape(6, 8) * 5
    assert orc.bilinear_double(img, 2.0, 3.0) == img[3, 2]
    v = orc.bilinear_double(img, 2.25, 3.5)
    want = 0.5 * (0.75 * img[3, 2] + 0.25 * img[3, 3]) + 0.5 * (0.75 * img[4, 2] + 0.25 * img[4, 3])
    assert abs(v - want) < 1e-12
    # the float variant rounds the two row interpolants to float (src/algorithm.cpp:885-894)
    x, y = 2.1234567891, 3.987654321
    a = np.float32((3 - x) * img[3, 2] + (x - 2) * img[3, 3])
    b = np.float32((3 - x) * img[4, 2] + (x - 2) * img[4, 3])
    assert orc.bilinear_float(img, x, y) == np.float32((4 - y) * float(a) + (y - 3) * float(b))


def test_median_rule(orc):
    # src/algorithm.cpp:834-853: element numValid/2; parity of the TOTAL count selects the branch
    big = np.finfo(np.float64).max
    v = np.array([5.0, 1.0, 4.0, 2.0, 3.0])
    assert orc.median(v, 5) == 3.0
    v6 = np.array([5.0, 1.0, 4.0, 2.0, 3.0, 6.0])
    assert orc.median(v6, 6) == 3.5                      # (vec[2] + vec[3]) / 2 with the exact predecessor
    v7 = np.array([5.0, 1.0, big, 2.0, 3.0, big, 4.0])   # 5 valid of 7 (odd N): vec[2]
    assert orc.median(v7, 5) == 3.0
    v8 = np.array([5.0, 1.0, big, 2.0, 3.0, big, 4.0, big])  # 5 valid of 8 (even N): (vec[1] + vec[2]) / 2
    assert orc.median(v8, 5) == 2.5
    rng = np.random.default_rng(5)
    for n in (49, 64, 12475, 12500):
        r = rng.normal(size=n) * 10
        s = np.sort(r)
        mid = n // 2
        want = s[mid] if n % 2 else 0.5 * (s[mid - 1] + s[mid])
        assert orc.median(r, n) == want
        mad_in = np.abs(r - want)
        sm = np.sort(mad_in)
        wm = sm[mid] if n % 2 else 0.5 * (sm[mid - 1] + sm[mid])
        assert abs(orc.sigma(r, n) - 1.482602218505602 * wm) < 1e-12
        if n % 2:  # both median modes agree with the real reference when N is odd
            assert orc.median(r, n, orc.MEDIAN_LIBSTDCXX) == want


def test_ldlt_solve(orc):
    rng = np.random.default_rng(6)
    for n in (3, 6):
        A = rng.normal(size=(n + 3, n))
        H = A.T @ A + 1e-3 * np.eye(n)
        b = rng.normal(size=n)
        assert np.allclose(orc.ldlt_solve(H, b), np.linalg.solve(H, b), rtol=1e-9, atol=1e-12)
    assert np.array_equal(orc.ldlt_solve(np.zeros((6, 6)), np.ones(6)), np.zeros(6))  # Eigen: zero pivots -> 0


def _align(orc, pair, mode, **kw):
    rp, _ = orc.build_pyramid(pair["ref"], 4)
    cp, _ = orc.build_pyramid(pair["cur"], 4)
    kp, _ = orc.build_pyramid(pair["kf"], 4)
    return orc.sparse_align(rp, kp, cp, pair["w"], pair["h"], pair["feats"], pair["n_ref"], pair["n_kf"], pair["T_ref"],
                            pair["T_kf"], pair["K"], pair["T_cur_init"], mode=mode, **kw)


def test_faithful_mode_is_one_step_per_level(orc, pair_cache):
    # optimizeLM always breaks after one damped step; the error is the PRE-step RMSE of level 0
    pair = pair_cache(0, 200)
    rmse, T, st, lv = _align(orc, pair, orc.LM_FAITHFUL)
    assert [l["iterations"] for l in lv] == [1, 1, 1, 1] and [l["evaluations"] for l in lv] == [1, 1, 1, 1]
    assert st == 0
    assert abs(rmse - np.sqrt(lv[3]["chi2"] / lv[3]["n_px"])) < 1e-12
    for l in lv:
        assert abs(l["lam"] - 1e-2 * np.max(np.diag(l["H"]))) < 1e-9 * l["lam"]  # src/optimizer.cpp:296-299
        dx = np.linalg.solve(l["H"] + l["lam"] * np.eye(6), l["g"])
        assert np.allclose(dx, l["dx"], rtol=1e-8)
        assert np.allclose(l["H"], l["H"].T)


def test_iterated_modes_recover_the_motion(orc, synth, pair_cache):
    pair = pair_cache(0, 500)
    for mode in (orc.LM_ITERATED, orc.GN):
        rmse, T, st, lv = _align(orc, pair, mode, max_iter=30)
        assert synth.rotation_angle(T, pair["T_cur_true"]) < 2e-4
        assert np.abs(T[4:] - pair["T_cur_true"][4:]).max() < 3e-3


def test_world_frame_and_keyframe_features(orc, synth, pair_cache):
    # T_ref != I; features of the last keyframe sample the keyframe image
    T_ref = tuple(synth.se3_from_Rt(synth.rodrigues(np.array([0.02, -0.03, 0.01])), [0.4, -0.2, 1.5]))
    pair = pair_cache(1, 300, n_kf=100, T_ref=T_ref)
    assert pair["n_kf"] == 100
    rmse, T, st, lv = _align(orc, pair, orc.GN, max_iter=30)
    assert synth.rotation_angle(T, pair["T_cur_true"]) < 1e-3
    assert np.abs(T[4:] - pair["T_cur_true"][4:]).max() < 2e-2


def test_no_ref_features_returns_zero(orc, pair_cache):
    pair = dict(pair_cache(0, 200))
    pair["n_ref"], pair["feats"] = 0, pair["feats"][:0]
    rmse, T, st, lv = _align(orc, pair, orc.LM_FAITHFUL)
    assert rmse == 0.0 and np.array_equal(T, pair["T_cur_init"])  # src/image_alignment.cpp:27-28


def test_feature_align_recovers_offset(orc, pair_cache):
    pair = pair_cache(0, 200)
    _, rg = orc.build_pyramid(pair["ref"], 1)
    g = rg.reshape(pair["h"], pair["w"])
    good = 0
    for i in range(60, 90):  # interior cells (the first grid row holds features closer than half+2 to the border)
        px = pair["feats"]["px"][i]
        rmse, p, st, it = orc.feature_align(g, g, px, px + np.array([0.4, -0.3]), mode=orc.GN, max_iter=30)
        good += np.abs(p - px).max() < 0.05
        r1, p1, st1, it1 = orc.feature_align(g, g, px, px + np.array([0.4, -0.3]), mode=orc.LM_FAITHFUL)
        assert it1 == 1
    assert good >= 28
    rmse, p, st, it = orc.feature_align(g, g, (100.0, 100.0), np.array([2.0, 100.0]))
    assert np.isnan(rmse)  # start out of frame: 0 / 0, as the reference


def test_align_golden(orc, pkg):
    g = np.load(os.path.join(GOLD, "align_golden.npz"))
    pair = pkg.synth.make_pair(index=int(g["index"]), n_features=int(g["n_features"]))
    for name, mode in (("faithful", orc.LM_FAITHFUL), ("lm", orc.LM_ITERATED), ("gn", orc.GN)):
        rmse, T, st, lv = _align(orc, pair, mode, max_iter=30)
        assert np.allclose(T, g[name + "_T"], atol=1e-9) and st == int(g[name + "_final_status"])
        assert np.allclose(np.array([l["H"] for l in lv]), g[name + "_H"], rtol=1e-9)
        assert np.array_equal(np.array([l["n_px"] for l in lv]), g[name + "_n_px"])
    _, rg = orc.build_pyramid(pair["ref"], 1)
    _, cg = orc.build_pyramid(pair["cur"], 1)
    gr, gc = rg.reshape(pair["h"], pair["w"]), cg.reshape(pair["h"], pair["w"])
    for row in g["fa"]:
        rmse, p, st, it = orc.feature_align(gr, gc, row[0:2], row[2:4], patch_size=7, mode=orc.LM_FAITHFUL)
        assert np.allclose(p, row[4:6], atol=1e-9) and st == row[7] and it == row[8]
        assert (np.isnan(rmse) and np.isnan(row[6])) or abs(rmse - row[6]) < 1e-9
