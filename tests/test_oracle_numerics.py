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


def test_se3_exp_adversarial_angles(orc):
    """The contracts of Sophus::SE3d::exp the alignment leans on (ImageAlignment::update, src/image_alignment.cpp:372-380),
    pinned against scipy.linalg.expm where closed forms lose digits: theta -> 0 (the series branch and its switch-over),
    theta next to pi (sin theta / theta -> 0, the quaternion's real part -> 0), theta beyond pi (double cover) and a pure
    translation.  The rotation must hold to 1e-12 everywhere.  V(omega) upsilon too, EXCEPT for the cancellation Sophus has
    itself: above its series threshold (theta >= 1e-10) it evaluates (1 - cos theta) / theta^2 directly, which carries an
    absolute error of eps / theta^2, i.e. eps |upsilon| / theta in the translation (5e-9 |upsilon| at theta = 1e-8).  The
    oracle restates that formula, so the bound is part of the contract; the CUDA code uses half-angle forms that do not
    cancel and sits inside the same bound (csrc/math.cuh)."""
    from scipy.linalg import expm
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(40)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    thetas = [0.0, 1e-300, 1e-160, 1e-20, 1e-11, 9.9e-11, 1.1e-10, 1e-9, 1e-8, 3e-8, 1e-6, 1e-4, 1e-2, 0.5, np.pi / 2, 3.0,
              np.pi - 1e-6, np.pi - 1e-12, np.pi, np.pi + 1e-12, np.pi + 1e-6, 4.0, 2 * np.pi - 1e-9, 2 * np.pi, 7.5]
    for th in thetas:
        for ups in (np.zeros(3), rng.normal(size=3), 1e6 * rng.normal(size=3)):
            xi = np.concatenate([ups, th * axis])
            T = np.asarray(orc.se3_exp(xi))
            w = xi[3:]
            M = np.zeros((4, 4))
            M[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            M[:3, 3] = ups
            E = expm(M)
            assert abs(np.linalg.norm(T[:4]) - 1.0) < 1e-14, th
            assert np.abs(Rotation.from_quat(T[:4]).as_matrix() - E[:3, :3]).max() < 1e-12, th
            cancel = 4 * np.finfo(float).eps / th if th >= 1e-10 else 0.0
            assert np.abs(T[4:] - E[:3, 3]).max() <= (1e-12 + cancel) * max(1.0, np.abs(ups).max()), (th, T[4:], E[:3, 3])
    # exp(xi) exp(-xi) = identity, also across the branch switch
    for th in (1e-11, 1e-9, 0.3, np.pi - 1e-9):
        xi = np.concatenate([rng.normal(size=3), th * axis])
        I = np.asarray(orc.se3_mul(orc.se3_exp(xi), orc.se3_exp(-xi)))
        assert np.abs(I[:3]).max() < 1e-15 and abs(abs(I[3]) - 1) < 1e-15 and np.abs(I[4:]).max() < 1e-12 + 8e-16 / th


def test_ldlt_adversarial_systems(orc):
    """Eigen::LDLT::solve as the optimisers call it (src/optimizer.cpp:96,306), restated with its symmetric pivoting on the
    largest |diagonal| and its D^-1 rule.  Against numpy on the systems where pivoting and the zero-pivot rule matter:
    the alignment's own badly scaled normal equations (rotation ~1e6 x translation), near-singular ones, an exactly
    rank-deficient one (Eigen returns the minimum-effort solution with zeros for the null pivots: A x = b still holds on
    the range), an indefinite one (LM with a rejected step can produce it) and permuted diagonals."""
    rng = np.random.default_rng(41)
    # badly scaled SPD, as J^T W J of the alignment: columns differ by 1e3, entries by 1e6
    for trial in range(20):
        S = np.diag(10.0 ** rng.uniform(-3, 3, 6))
        A = rng.normal(size=(40, 6)) @ S
        H = A.T @ A
        b = A.T @ rng.normal(size=40)
        x = np.asarray(orc.ldlt_solve(H, b))
        xr = np.linalg.solve(H, b)
        assert np.abs((x - xr) / np.maximum(np.abs(xr), 1e-300)).max() < 1e-6, trial
        assert np.abs(H @ x - b).max() <= 1e-9 * np.abs(b).max()
    # near-singular: condition 1e14 -- compare residuals, not solutions
    U, _ = np.linalg.qr(rng.normal(size=(6, 6)))
    H = U @ np.diag([1, 1e-2, 1e-5, 1e-8, 1e-11, 1e-14]) @ U.T
    H = 0.5 * (H + H.T)
    b = H @ rng.normal(size=6)
    x = np.asarray(orc.ldlt_solve(H, b))
    assert np.abs(H @ x - b).max() < 1e-12
    # the largest diagonal is pivoted first: a matrix whose leading entry is tiny must not lose the solution
    H = np.diag([1e-18, 3.0, 2.0, 5.0, 1.0, 4.0]) + 1e-3 * np.ones((6, 6))
    b = rng.normal(size=6)
    assert np.allclose(orc.ldlt_solve(H, b), np.linalg.solve(H, b), rtol=1e-9)
    # exactly rank deficient (two features on one line): zeros for the null pivots, consistent on the range
    B = rng.normal(size=(4, 6))
    H = B.T @ B
    b = H @ rng.normal(size=6)
    x = np.asarray(orc.ldlt_solve(H, b))
    assert np.isfinite(x).all() and np.abs(H @ x - b).max() < 1e-9 * max(1.0, np.abs(b).max())
    # symmetric indefinite with a dominant diagonal: LDLT with diagonal pivoting still solves it
    H = np.diag([4.0, -3.0, 5.0, -6.0, 2.0, 7.0]) + 0.1 * (lambda R: R + R.T)(rng.normal(size=(6, 6)))
    b = rng.normal(size=6)
    assert np.allclose(orc.ldlt_solve(H, b), np.linalg.solve(H, b), rtol=1e-8)
    # 3 x 3 (FeatureAlignment's system) with a zero row / column: the reference's pseudo-inverse-like answer 0 there
    H = np.array([[2.0, 0.5, 0.0], [0.5, 3.0, 0.0], [0.0, 0.0, 0.0]])
    b = np.array([1.0, -2.0, 0.0])
    x = np.asarray(orc.ldlt_solve(H, b))
    assert x[2] == 0.0 and np.allclose(H[:2, :2] @ x[:2], b[:2])


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


# ------------------------------------------------------------------------------------------------
# epipolar search of the depth filter: the restatement against independent numpy statements
# ------------------------------------------------------------------------------------------------
def _np_bilinear_float(img, x, y):
    """algorithm::bilinearInterpolation (float), src/algorithm.cpp:885-894, in numpy scalars"""
    x1, y1 = int(x), int(y)
    a = np.float32((x1 + 1 - x) * float(img[y1, x1]) + (x - x1) * float(img[y1, x1 + 1]))
    b = np.float32((x1 + 1 - x) * float(img[y1 + 1, x1]) + (x - x1) * float(img[y1 + 1, x1 + 1]))
    return np.float32((y1 + 1 - y) * float(a) + (y - y1) * float(b))


def _np_epipolar(ref, cur, K, T_rel, px, bearing, depth, dmin, dmax, P=7, eigen_mean=True):
    """independent numpy restatement of src/algorithm.cpp:412-551 (rotation through a matrix, lstsq triangulation)"""
    from scipy.spatial.transform import Rotation
    h, w = ref.shape
    R = Rotation.from_quat(T_rel[:4]).as_matrix()
    t = np.asarray(T_rel[4:])
    half, area = P // 2, P * P

    def bear(x, y):
        b = np.array([(x - K[2]) / K[0], (y - K[3]) / K[1], 1.0])
        return b / np.linalg.norm(b)

    def proj(x, y, d):
        pc = R @ (bear(x, y) * d) + t
        return np.array([K[0] * pc[0] / pc[2] + K[2], K[1] * pc[1] / pc[2] + K[3]])

    def clamp(l):
        return np.array([min(max(l[0], 0.0), w - 1) if l[0] >= w or l[0] < 0 else l[0],
                         min(max(l[1], 0.0), h - 1) if l[1] >= h or l[1] < 0 else l[1]])

    lmin, lmax = clamp(proj(px[0], px[1], dmin)), clamp(proj(px[0], px[1], dmax))
    c = proj(px[0], px[1], depth)
    A = np.stack([(proj(px[0] + half, px[1], depth) - c) / half, (proj(px[0], px[1] + half, depth) - c) / half], 1)

    def warp(img, loc, A, data):
        bnd = A @ np.array([half, half], dtype=np.float64)
        mb = np.ceil(max(abs(bnd[0]), abs(bnd[1]))) + 2
        if not (loc[0] >= mb and loc[1] >= mb and loc[0] < w - mb and loc[1] < h - mb):
            return data
        out = []
        for i in range(-half, half + 1):
            for j in range(-half, half + 1):
                q = loc + A @ np.array([j, i], dtype=np.float64)
                out.append(np.uint8(_np_bilinear_float(img, q[0], q[1])))
        return np.array(out, dtype=np.uint8)

    def score(r, c):
        if eigen_mean:   # Eigen mean() on Matrix<uint8_t>: sum and division in uint8
            rm, cm = float(np.uint8(r.sum(dtype=np.uint8) // np.uint8(area))), float(np.uint8(c.sum(dtype=np.uint8) // np.uint8(area)))
        else:
            rm, cm = r.astype(np.float64).mean(), c.astype(np.float64).mean()
        return float(np.abs((r.astype(np.float64) - rm) - (c.astype(np.float64) - cm)).sum())

    def tri(b_cur):
        Am = np.stack([R @ np.asarray(bearing), -b_cur], 1)
        if np.linalg.det(Am.T @ Am) < 1e-6:
            return None
        return abs(np.linalg.solve(Am.T @ Am, -(Am.T @ t))[0])

    refp = warp(ref, np.asarray(px, dtype=np.float64), np.eye(2), np.zeros(area, np.uint8))
    e = lmax - lmin
    norm = np.linalg.norm(e)
    if norm < 2.0:
        mid = (lmax + lmin) / 2
        d = tri(bear(mid[0], mid[1]))
        return dict(found=d is not None, depth=d, px=mid, steps=0, score=None)
    steps, step = int(np.ceil(norm)), e / norm
    curp = np.zeros(area, np.uint8)
    best, bl = np.inf, None
    for i in range(steps):
        loc = lmin + i * step
        curp = warp(cur, loc, A, curp)
        z = score(refp, curp)
        if z < best:
            best, bl = z, loc
    d = tri(bear(bl[0], bl[1])) if best < area * 128 else None
    return dict(found=d is not None, depth=d, px=bl, steps=steps, score=best)


@pytest.mark.parametrize("eigen_mean", [True, False])
def test_epipolar_match_against_numpy(orc, pkg, eigen_mean):
    synth = pkg.synth
    pair = synth.make_pair(8, 120, motion_scale=3.0)
    T_rel = synth.se3_mul(pair["T_cur_true"], synth.se3_inv(pair["T_ref"]))
    rng = np.random.default_rng(2)
    n_found = 0
    for i in range(0, 120, 3):
        f = pair["feats"][i]
        d = np.linalg.norm(f["point"])
        lo, hi, d0 = d * rng.uniform(0.4, 0.9), d * rng.uniform(1.1, 3.0), d * rng.uniform(0.8, 1.25)
        if i == 3:
            lo, hi = d * 0.9995, d * 1.0005           # short segment: midpoint triangulation
        o = orc.epipolar_match(pair["ref"], pair["cur"], pair["K"], T_rel, f["px"], f["bearing"], d0, lo, hi,
                               mean_mode=orc.MEAN_EIGEN_U8 if eigen_mean else orc.MEAN_EXACT)
        w = _np_epipolar(pair["ref"], pair["cur"], pair["K"], T_rel, f["px"], f["bearing"], d0, lo, hi, eigen_mean=eigen_mean)
        assert o["found"] == w["found"] and o["steps"] == w["steps"], i
        assert np.abs(o["px"] - w["px"]).max() < 1e-9
        if w["score"] is not None:
            assert abs(o["score"] - w["score"]) < 1e-6
        if w["found"]:
            assert abs(o["depth"] - w["depth"]) < 1e-9 * w["depth"]
            n_found += 1
    assert n_found > 30


def test_epipolar_uint8_mean_quirk(orc):
    """computeScore's means are computed in uint8 (Eigen mean() of a Matrix<uint8_t>, src/algorithm.cpp:400-401): two
    flat patches 40 grey levels apart score 49 * |(r - rm) - (c - cm)| with rm, cm the WRAPPED means, not 0."""
    img_r = np.full((40, 40), 100, np.uint8)
    img_c = np.full((40, 40), 140, np.uint8)
    K = (100.0, 100.0, 20.0, 20.0)
    T = np.array([0, 0, 0, 1, 0.5, 0, 0], dtype=np.float64)      # pure x translation: horizontal epipolar line
    b = np.array([0.0, 0.0, 1.0])
    kw = dict(depth=10.0, min_depth=4.0, max_depth=40.0)
    o_q = orc.epipolar_match(img_r, img_c, K, T, (20.0, 20.0), b, mean_mode=orc.MEAN_EIGEN_U8, **kw)
    o_e = orc.epipolar_match(img_r, img_c, K, T, (20.0, 20.0), b, mean_mode=orc.MEAN_EXACT, **kw)
    rm = float(np.uint8((49 * 100) % 256 // 49))     # 4900 mod 256 = 36 -> 0
    cm = float(np.uint8((49 * 140) % 256 // 49))     # 6860 mod 256 = 204 -> 4
    assert o_q["score"] == 49 * abs((100 - rm) - (140 - cm))
    assert o_e["score"] == 0.0 and o_q["steps"] == o_e["steps"] > 2


# ------------------------------------------------------------------------------------------------
# FeatureSelection::gradientMagnitudeWithSSC: the restatement against plain numpy / python
# ------------------------------------------------------------------------------------------------
def _py_ssc(grad, thr, K, cell, occupancy=None, use_bucketing=True):
    """src/feature_selection.cpp:27-89 + :165-248 as a direct python loop (stable sort by response)."""
    import math
    h, w = grad.shape
    ys, xs = np.nonzero(grad > thr)                       # raster order
    resp = grad[ys, xs].astype(np.int32)
    order = np.argsort(-resp, kind="stable")
    ys, xs, resp = ys[order], xs[order], resp[order]
    n = len(ys)
    exp1 = h + w + 2 * K
    exp2 = 4 * w + 4 * K + 4 * h * K + h * h + w * w - 2 * h * w + 4 * h * w * K
    exp3, exp4 = math.sqrt(exp2), 2 * (K - 1)

    def cround(v):   # C round(): half away from zero
        return math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)
    sol1, sol2 = -cround((exp1 + exp3) / exp4), -cround((exp1 - exp3) / exp4)
    high = int(sol1) if sol1 > sol2 else int(sol2)
    low = int(math.sqrt(n / K))
    kmin = int(cround(float(np.float32(K) - np.float32(K) * np.float32(0.1))))
    kmax = int(cround(float(np.float32(K) + np.float32(K) * np.float32(0.1))))
    prev, result, res_vec, iters, wused = -1, [], [], 0, -1
    while True:
        width = low + int((high - low) / 2)               # C integer division truncates toward zero
        if width == prev or low > high or width <= 0:
            res_vec = result
            break
        iters, wused = iters + 1, width
        c = width / 2.0
        ncc, ncr = int(w / c), int(h / c)
        covered = np.zeros((ncr + 1, ncc + 1), bool)
        reach = int(width / c)
        result = []
        rows = (ys.astype(np.float32).astype(np.float64) / c).astype(np.int64)
        cols = (xs.astype(np.float32).astype(np.float64) / c).astype(np.int64)
        for i in range(n):
            r, cc = rows[i], cols[i]
            if not covered[r, cc]:
                result.append(i)
                covered[max(r - reach, 0):min(r + reach, ncr) + 1, max(cc - reach, 0):min(cc + reach, ncc) + 1] = True
        if kmin <= len(result) <= kmax:
            res_vec = result
            break
        if len(result) < kmin:
            high = width - 1
        else:
            low = width + 1
        prev = width
    gcols = w // cell + 1
    grid = np.zeros((h // cell + 1) * gcols, bool) if occupancy is None else np.asarray(occupancy).astype(bool).copy()
    out = []
    for i in res_vec:
        if use_bucketing:
            b = (ys[i] // cell) * gcols + xs[i] // cell
            if grid[b]:
                continue
            grid[b] = True
        out.append((xs[i], ys[i], resp[i]))
    return np.array(out, np.int32).reshape(-1, 3), dict(keypoints=n, width=wused, iterations=iters, ssc_points=len(res_vec))


@pytest.mark.parametrize("thr,k,bucket", [(120, 80, True), (60, 200, True), (60, 200, False), (200, 30, True)])
def test_select_ssc_against_python(orc, thr, k, bucket):
    rng = np.random.default_rng(thr + k)
    h, w = 96, 160
    img = rng.integers(0, 256, (h // 4, w // 4), dtype=np.uint8).repeat(4, 0).repeat(4, 1)   # blocky: many equal responses
    img = (img.astype(np.int32) + rng.integers(0, 12, (h, w))).clip(0, 255).astype(np.uint8)
    grad = orc.abs_gradient(img)
    occ = (rng.random((h // 16 + 1) * (w // 16 + 1)) < 0.25).astype(np.uint8)
    for o in (None, occ):
        got, gi = orc.select_ssc(grad, thr, k, 16, occupancy=o, use_bucketing=bucket)
        want, wi = _py_ssc(grad, thr, k, 16, occupancy=o, use_bucketing=bucket)
        assert gi == wi, (gi, wi)
        assert np.array_equal(got, want)


def _check_next_rows_golden(ssc_fn, epi_fn, synth):
    """shared by the CPU test (oracle) and the GPU test (CUDA path): both must reproduce the golden vectors made by
    the independent python / numpy restatements (tests/golden/make_golden_next.py)"""
    g = np.load(os.path.join(GOLD, "next_rows_golden.npz"))
    for tag in ("a", "b"):
        thr, k, cell, bucket = [int(v) for v in g["ssc_%s_params" % tag]]
        feats, info = ssc_fn(g["img"], g["grad"], thr, k, cell, bool(bucket))
        assert [info["keypoints"], info["width"], info["iterations"], info["ssc_points"]] == list(g["ssc_%s_info" % tag])
        assert np.array_equal(feats, g["ssc_%s_feats" % tag])
    wide = synth.make_pair(index=int(g["epi_index"]), n_features=120, motion_scale=float(g["epi_motion_scale"]))
    rows = g["epi_rows"]
    res = epi_fn(wide, g["epi_T_rel"], rows)
    for row, r in zip(rows, res):
        assert bool(row[5]) == bool(r["found"]) and int(row[9]) == int(r["steps"])
        assert abs(row[7] - r["px"][0]) < 1e-9 and abs(row[8] - r["px"][1]) < 1e-9
        if row[5]:
            assert abs(row[6] - r["depth"]) < 1e-9 * row[6]
        if row[10] >= 0:
            assert abs(row[10] - r["score"]) < 1e-6


def test_next_rows_golden_oracle(orc, pkg):
    def ssc(img, grad, thr, k, cell, bucket):
        assert np.array_equal(orc.abs_gradient(img), grad)
        return orc.select_ssc(grad, thr, k, cell, use_bucketing=bucket)

    def epi(wide, T_rel, rows):
        out = []
        for row in rows:
            f = wide["feats"][int(row[0])]
            out.append(orc.epipolar_match(wide["ref"], wide["cur"], wide["K"], T_rel, f["px"], f["bearing"], row[2], row[3], row[4],
                                          mean_mode=orc.MEAN_EIGEN_U8 if row[1] else orc.MEAN_EXACT))
        return out
    _check_next_rows_golden(ssc, epi, pkg.synth)


# ------------------------------------------------------------------------------------------------
# ImageAlignment::align end to end against an INDEPENDENT numpy restatement: cv2 pyramids, rotation matrices from
# scipy, SE3 exp through scipy.linalg.expm of the 4x4 twist, numpy.linalg.solve, numpy sort for the medians -- no code
# shared with oracle/svo_oracle.cpp.  What the reference executes (one damped LM step per level).
# ------------------------------------------------------------------------------------------------
def _np_bilin(img, x, y):
    """algorithm::bilinearInterpolationDouble, src/algorithm.cpp:896-905"""
    x1, y1 = int(x), int(y)
    a = (x1 + 1 - x) * float(img[y1, x1]) + (x - x1) * float(img[y1, x1 + 1])
    b = (x1 + 1 - x) * float(img[y1 + 1, x1]) + (x - x1) * float(img[y1 + 1, x1 + 1])
    return (y1 + 1 - y) * a + (y - y1) * b


def _np_median_rule(values, num_valid, n_total):
    """algorithm::computeMedian with MEDIAN_EXACT: values = the valid entries; invalid ones sort last"""
    s = np.sort(values)
    mid = num_valid // 2
    return s[mid] if n_total % 2 == 1 else 0.5 * (s[mid - 1] + s[mid])


def _np_align_faithful(ref_pyr, cur_pyr, feats, n_ref, T_ref, T_cur, K, P=5, levels=(3, 2, 1, 0)):
    from scipy.linalg import expm
    from scipy.spatial.transform import Rotation

    def to_mat(T):
        M = np.eye(4)
        M[:3, :3] = Rotation.from_quat(T[:4]).as_matrix()
        M[:3, 3] = T[4:]
        return M
    Mref, Mcur = to_mat(np.asarray(T_ref, float)), to_mat(np.asarray(T_cur, float))
    half, area = P // 2, P * P
    F = len(feats)
    stats = []
    for level in levels:
        ref, cur = ref_pyr[level], cur_pyr[level]
        lh, lw = ref.shape
        scale = 1.0 / (1 << level)
        fx, fy = K[0] / (1 << level), K[1] / (1 << level)
        J = np.zeros((F * area, 6))
        T_patch = np.zeros(F * area)
        vis_ref = np.zeros(F, bool)
        pW = np.zeros((F, 3))
        Cref = -Mref[:3, :3].T @ Mref[:3, 3]            # Frame::cameraInWorld
        for f in range(F):
            ft = feats[f]
            if not ft["has_point"]:
                continue
            u, v = ft["px"][0] * scale, ft["px"][1] * scale
            uI, vI = int(np.floor(u)), int(np.floor(v))
            b = half + 2
            if uI - b < 0 or vI - b < 0 or uI + b >= lw or vI + b >= lh:
                continue
            vis_ref[f] = True
            depth = np.linalg.norm(ft["point"] - Cref)
            pc = ft["bearing"] * depth
            pw = np.linalg.inv(Mref)[:3, :3] @ pc + np.linalg.inv(Mref)[:3, 3]
            pW[f] = pw
            x, y, z = pw
            J0 = np.array([fx / z, 0, -fx * x / z**2, -fx * x * y / z**2, fx * x * x / z**2 + fx, -fx * y / z])
            J1 = np.array([0, fy / z, -fy * y / z**2, -fy * y * y / z**2 - fy, fy * x * y / z**2, fy * x / z])
            k = 0
            for yy in range(-half, half + 1):
                for xx in range(-half, half + 1):
                    T_patch[f * area + k] = _np_bilin(ref, u + xx, v + yy)
                    gx = 0.5 * (_np_bilin(ref, u + xx + 1, v + yy) - _np_bilin(ref, u + xx - 1, v + yy))
                    gy = 0.5 * (_np_bilin(ref, u + xx, v + yy + 1) - _np_bilin(ref, u + xx, v + yy - 1))
                    J[f * area + k] = gx * J0 + gy * J1
                    k += 1
        r = np.zeros(F * area)
        valid = np.zeros(F * area, bool)
        for f in range(F):
            if not vis_ref[f]:
                continue
            pcur = Mcur[:3, :3] @ pW[f] + Mcur[:3, 3]
            u = (K[0] * pcur[0] / pcur[2] + K[2]) * scale
            v = (K[1] * pcur[1] / pcur[2] + K[3]) * scale
            uI, vI = int(np.floor(u)), int(np.floor(v))
            b = half + 2
            if uI - b < 0 or vI - b < 0 or uI + b >= lw or vI + b >= lh:
                continue
            k = 0
            for yy in range(-half, half + 1):
                for xx in range(-half, half + 1):
                    r[f * area + k] = _np_bilin(cur, u + xx, v + yy) - T_patch[f * area + k]
                    valid[f * area + k] = True
                    k += 1
        nv = int(valid.sum())
        med = _np_median_rule(r[valid], nv, F * area)
        mad = _np_median_rule(np.abs(r[valid] - med), nv, F * area)
        sigma = max(1.482602218505602 * mad, np.finfo(float).eps)
        c = 4.6851 * sigma
        wgt = np.where(valid & (np.abs(r) <= c), (1 - r**2 / c**2) ** 2, 0.0)
        chi2 = float((wgt * r * r).sum())
        H = J.T @ (wgt[:, None] * J)
        g = J.T @ (wgt * r)
        lam = 1e-2 * H.diagonal().max()
        dx = np.linalg.solve(H + lam * np.eye(6), g)
        tw = np.zeros((4, 4))                            # pose <- pose * exp(-dx), dx = (upsilon, omega)
        ups, om = -dx[:3], -dx[3:]
        tw[:3, :3] = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        tw[:3, 3] = ups
        Mcur = Mcur @ expm(tw)
        stats.append(dict(H=H, g=g, chi2=chi2, sigma=sigma, lam=lam, dx=dx, n_px=nv, pose=Mcur.copy(), rmse=np.sqrt(chi2 / nv)))
    return stats


@pytest.mark.parametrize("n_features", [101, 60])   # odd N: the reference's median is well defined; even N: MEDIAN_EXACT
def test_align_against_independent_numpy(orc, pkg, n_features):
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial.transform import Rotation
    synth = pkg.synth
    T_ref = synth.se3_from_Rt(synth.rodrigues(np.array([0.02, -0.03, 0.01])), [0.4, -0.2, 1.5])   # world != ref frame
    pair = synth.make_pair(index=11, n_features=n_features, T_ref=tuple(T_ref))
    T0 = synth.se3_mul(synth.se3_from_Rt(synth.rodrigues(np.array([1e-3, -2e-3, 5e-4])), [0.01, -0.02, -0.1]), pair["T_ref"])
    feats = pair["feats"][:pair["n_ref"]]   # the numpy restatement handles the reference frame's features
    def pyr(img):
        out = [img]
        for _ in range(3):
            out.append(cv2.pyrDown(out[-1]))
        return out
    want = _np_align_faithful(pyr(pair["ref"]), pyr(pair["cur"]), feats, len(feats), pair["T_ref"], T0, pair["K"])
    rp, cp = orc.build_pyramid(pair["ref"], 4)[0], orc.build_pyramid(pair["cur"], 4)[0]
    rmse, T, status, lv = orc.sparse_align(rp, rp, cp, pair["w"], pair["h"], feats, len(feats), 0, pair["T_ref"], pair["T_ref"],
                                           pair["K"], T0, mode=orc.LM_FAITHFUL)
    for s, (o, w_) in enumerate(zip(lv, want)):
        assert o["n_px"] == w_["n_px"], s
        assert abs(o["sigma"] - w_["sigma"]) <= 1e-10 * w_["sigma"], (s, o["sigma"], w_["sigma"])
        assert abs(o["chi2"] - w_["chi2"]) <= 1e-9 * w_["chi2"]
        assert np.abs(o["H"] - w_["H"]).max() <= 1e-9 * np.abs(w_["H"]).max()
        assert np.abs(o["g"] - w_["g"]).max() <= 1e-9 * np.abs(w_["H"]).max()
        assert abs(o["lam"] - w_["lam"]) <= 1e-9 * w_["lam"]
        assert np.abs(o["dx"] - w_["dx"]).max() <= 1e-7 * max(1e-3, np.abs(w_["dx"]).max()), (s, o["dx"], w_["dx"])
        Ro = Rotation.from_quat(o["pose_after"][:4]).as_matrix()
        assert np.abs(Ro - w_["pose"][:3, :3]).max() < 1e-9 and np.abs(o["pose_after"][4:] - w_["pose"][:3, 3]).max() < 1e-9, s
    assert abs(rmse - want[-1]["rmse"]) <= 1e-9 * rmse


# ---------------------------------------------------------------------------------------------------------------------
# Row f1: an independent Python restatement of Map::reprojectMap's walk (src/map.cpp:462-489 fill the cells through
# reprojectPoint :491-503; :474-488 visit the cells in m_cellOrders; reprojectCell :505-576 sorts each cell by point type,
# skips DELETED points and accepts the first other candidate after FeatureAlignment::align, whatever it returned).
def _py_reproject_map(orc, grads, cur_grad, K, T_cur, cands, cell, order, max_matches):
    h, w = cur_grad.shape
    cols = -(-w // cell)
    cells = {}
    projected = np.zeros(len(cands), np.uint8)
    pix = []
    for i, c in enumerate(cands):
        px = orc.project2d(K, orc.se3_act(T_cur, c["point"]))                      # Frame::world2image, src/frame.cpp:83-101
        pix.append(px)
        if px[0] >= 3 and px[1] >= 3 and px[0] < w - 3 and px[1] < h - 3:          # PinholeCamera::isInFrame(px, 3), :163-169
            cells.setdefault(int(px[1]) // cell * cols + int(px[0]) // cell, []).append(i)
            projected[i] = 1
    out = []
    for k in order:
        got = False
        for i in sorted(cells.get(int(k), []), key=lambda i: -int(cands["type"][i])):   # sorted() is stable, like the oracle's choice
            if cands["type"][i] == 1:                                              # Point::PointType::DELETED
                continue
            rmse, px, st, _ = orc.feature_align(grads[cands["ref_slot"][i]], cur_grad, cands["ref_px"][i], pix[i])
            out.append((int(k), i, px[0], px[1], rmse, st))
            got = True
            break
        if got and len(out) > max_matches:                                         # `if ( m_matches > 150 ) break;`
            break
    return out, projected


@pytest.mark.parametrize("cell,max_matches", [(30, 150), (48, 25)])
def test_reproject_map_against_python(orc, pkg, pair_cache, cell, max_matches):
    pair = pair_cache(10, 300)
    rng = np.random.default_rng(5 + cell)
    f = pair["feats"][pair["feats"]["has_point"] != 0]
    cands = np.zeros(2 * len(f), pkg.capi.REPROJ_CAND_DTYPE)
    cands["ref_slot"] = np.repeat([0, 1], len(f))
    cands["ref_px"], cands["point"] = np.tile(f["px"], (2, 1)), np.tile(f["point"], (2, 1))
    cands["point"][::7] += 40.0                                                    # some points leave the frame
    cands["type"] = rng.choice([0, 1, 2, 3], size=len(cands), p=[0.4, 0.1, 0.2, 0.3])
    h, w = pair["h"], pair["w"]
    gref, gcur = orc.abs_gradient(pair["ref"]), orc.abs_gradient(pair["cur"])
    order = rng.permutation(-(-w // cell) * -(-h // cell)).astype(np.int32)
    T = pair["T_cur_true"]
    want, wproj = _py_reproject_map(orc, [gref, gref], gcur, pair["K"], T, cands, cell, order, max_matches)
    got, gproj = orc.reproject_map([gref, gref], gcur, pair["K"], T, cands, cell, order, max_matches=max_matches)
    assert np.array_equal(gproj, wproj) and 0 < wproj.sum() < len(cands)
    assert len(got) == len(want) and len(got) > 10
    want = np.array(want, np.float64)
    assert np.array_equal(got[:, :2], want[:, :2]) and np.array_equal(got[:, 5], want[:, 5])
    assert np.array_equal(got[:, 2:5], want[:, 2:5], equal_nan=True)
    if max_matches == 25:
        assert len(got) == 26


# ---------------------------------------------------------------------------------------------------------------------
# FeatureAlignment::align in the reference's mode: a second, independent restatement in numpy (src/feature_alignment.cpp:25-62
# align, :64-110 computeJacobian, :113-168 computeResiduals, :200-205 update; Optimizer::optimizeLM src/optimizer.cpp:161-370 with
# its one damped step; tukeyWeighting :485-507).  float32 bilinear taps as algorithm::bilinearInterpolation (:885-894).
def _np_bilin_f32(img, x, y):
    x1, y1 = int(x), int(y)
    a = np.float32((x1 + 1 - x) * float(img[y1, x1]) + (x - x1) * float(img[y1, x1 + 1]))
    b = np.float32((x1 + 1 - x) * float(img[y1 + 1, x1]) + (x - x1) * float(img[y1 + 1, x1 + 1]))
    return float(np.float32((y1 + 1 - y) * float(a) + (y - y1) * float(b)))


def _np_feature_align_faithful(ref_grad, cur_grad, ref_px, px, P=7):
    h, w = ref_grad.shape
    half, area = P // 2, P * P
    border = half + 2
    inside = lambda p: p[0] >= border and p[1] >= border and p[0] < w - border and p[1] < h - border   # isInFrame(px, border)
    J, T = np.zeros((area, 3)), np.zeros(area)
    if inside(ref_px):                                                      # else J and the template stay zero (:74-77)
        k = 0
        for y in range(-half, half + 1):
            for x in range(-half, half + 1):
                r, c = ref_px[1] + y, ref_px[0] + x
                T[k] = _np_bilin_f32(ref_grad, c, r)
                dx = 0.5 * (_np_bilin_f32(ref_grad, c + 1, r) - _np_bilin_f32(ref_grad, c - 1, r))
                dy = 0.5 * (_np_bilin_f32(ref_grad, c, r + 1) - _np_bilin_f32(ref_grad, c, r - 1))
                J[k] = (dx, dy, 1.0)
                k += 1
    pose = np.array([px[0], px[1], 0.0])
    if not inside(pose[:2]):                                                # computeResiduals returns 0: chi2 / 0 (:126-129)
        return float("nan"), pose[:2]
    res = np.zeros(area)
    k = 0
    for y in range(-half, half + 1):
        for x in range(-half, half + 1):
            res[k] = -(_np_bilin_f32(cur_grad, pose[0] + x, pose[1] + y) - T[k] + pose[2])
            k += 1
    med = np.sort(res)[area // 2]                                           # 49 or 25 residuals: odd, the plain median
    mad = np.sort(np.abs(res - med))[area // 2]
    sigma = max(1.482602218505602 * mad, np.finfo(float).eps)
    c = 4.6851 * sigma
    wgt = np.where(np.abs(res) <= c, (1 - res**2 / c**2) ** 2, 0.0)
    chi2 = float((wgt * res * res).sum())
    H = J.T @ (wgt[:, None] * J)
    g = J.T @ (wgt * res)
    lam = 1e-2 * H.diagonal().max()
    try:
        dxv = np.linalg.solve(H + lam * np.eye(3), g)
    except np.linalg.LinAlgError:                                           # zero Jacobian: Eigen's LDLT::solve treats zero
        dxv = np.linalg.pinv(H) @ g                                         # pivots as a pseudo-inverse does -> dx = 0
    pose = pose + dxv                                                       # FeatureAlignment::update
    return np.sqrt(chi2 / area), pose[:2]                                   # the pre-step RMSE and the moved pixel


@pytest.mark.parametrize("patch", [7, 5])
def test_feature_align_against_independent_numpy(orc, pair_cache, patch):
    pair = pair_cache(3, 200)
    gref, gcur = orc.abs_gradient(pair["ref"]), orc.abs_gradient(pair["cur"])
    rng = np.random.default_rng(12)
    f = pair["feats"][: pair["n_ref"]]
    n = 0
    for i in rng.choice(len(f), 60, replace=False):
        ref_px = f["px"][i].astype(float)
        start = ref_px + rng.uniform(-6, 6, 2)                              # sub-pixel starts around the feature
        want_rmse, want_px = _np_feature_align_faithful(gref, gcur, ref_px, start, patch)
        rmse, px, status, it = orc.feature_align(gref, gcur, ref_px, start, patch_size=patch, mode=orc.LM_FAITHFUL)
        if np.isnan(want_rmse):
            assert np.isnan(rmse)
            continue
        assert abs(rmse - want_rmse) <= 1e-9 * max(1.0, want_rmse), (i, rmse, want_rmse)
        assert np.abs(px - want_px).max() <= 1e-7, (i, px, want_px)
        n += 1
    assert n > 40
    # a reference pixel too close to the border: zero template and Jacobian (:74-77), and a start outside: NaN (0 / 0)
    want_rmse, want_px = _np_feature_align_faithful(gref, gcur, np.array([2.0, 3.0]), np.array([40.0, 40.0]), patch)
    rmse, px, _, _ = orc.feature_align(gref, gcur, [2.0, 3.0], [40.0, 40.0], patch_size=patch, mode=orc.LM_FAITHFUL)
    assert abs(rmse - want_rmse) <= 1e-9 * max(1.0, want_rmse)              # the pre-step RMSE of the raw current patch
    rmse, px, _, _ = orc.feature_align(gref, gcur, f["px"][0], [1.0, 1.0], patch_size=patch, mode=orc.LM_FAITHFUL)
    assert np.isnan(rmse)


# ---------------------------------------------------------------------------------------------------------------------
# ImageAlignment::align with Optimizer::optimizeGN (src/optimizer.cpp:41-159) -- the mode of the headline benchmark (GN, <= 30
# iterations per level) -- as a second, independent numpy restatement: same per-level set-up as _np_align_faithful, then the
# Gauss-Newton loop with its exits (dx.maxCoeff() > 1e3, NaN, chi2 increase -> rollback, dx.dx < 1e-16 or chi2 < 0.1 after
# the update, iteration cap).
def _np_align_gn(ref_pyr, cur_pyr, feats, T_ref, T_cur, K, P=5, levels=(3, 2, 1, 0), max_iter=30):
    from scipy.linalg import expm
    from scipy.spatial.transform import Rotation

    def to_mat(T):
        M = np.eye(4)
        M[:3, :3] = Rotation.from_quat(T[:4]).as_matrix()
        M[:3, 3] = T[4:]
        return M
    Mref, Mcur = to_mat(np.asarray(T_ref, float)), to_mat(np.asarray(T_cur, float))
    half, area = P // 2, P * P
    F = len(feats)
    out = []
    for level in levels:
        ref, cur = ref_pyr[level], cur_pyr[level]
        lh, lw = ref.shape
        scale = 1.0 / (1 << level)
        fx, fy = K[0] / (1 << level), K[1] / (1 << level)
        J = np.zeros((F * area, 6))
        T_patch = np.zeros(F * area)
        vis_ref = np.zeros(F, bool)
        pW = np.zeros((F, 3))
        Minv = np.linalg.inv(Mref)
        Cref = Minv[:3, 3]                                   # camera centre in the world: -R^T t
        b = half + 2
        for f in range(F):
            ft = feats[f]
            if not ft["has_point"]:
                continue
            u, v = ft["px"][0] * scale, ft["px"][1] * scale
            uI, vI = int(np.floor(u)), int(np.floor(v))
            if uI - b < 0 or vI - b < 0 or uI + b >= lw or vI + b >= lh:
                continue
            vis_ref[f] = True
            pw = Minv[:3, :3] @ (ft["bearing"] * np.linalg.norm(ft["point"] - Cref)) + Minv[:3, 3]
            pW[f] = pw
            x, y, z = pw
            J0 = np.array([fx / z, 0, -fx * x / z**2, -fx * x * y / z**2, fx * x * x / z**2 + fx, -fx * y / z])
            J1 = np.array([0, fy / z, -fy * y / z**2, -fy * y * y / z**2 - fy, fy * x * y / z**2, fy * x / z])
            k = 0
            for yy in range(-half, half + 1):
                for xx in range(-half, half + 1):
                    T_patch[f * area + k] = _np_bilin(ref, u + xx, v + yy)
                    gx = 0.5 * (_np_bilin(ref, u + xx + 1, v + yy) - _np_bilin(ref, u + xx - 1, v + yy))
                    gy = 0.5 * (_np_bilin(ref, u + xx, v + yy + 1) - _np_bilin(ref, u + xx, v + yy - 1))
                    J[f * area + k] = gx * J0 + gy * J1
                    k += 1

        def evaluate(M):
            r = np.zeros(F * area)
            valid = np.zeros(F * area, bool)
            for f in range(F):
                if not vis_ref[f]:
                    continue
                pc = M[:3, :3] @ pW[f] + M[:3, 3]
                u = (K[0] * pc[0] / pc[2] + K[2]) * scale
                v = (K[1] * pc[1] / pc[2] + K[3]) * scale
                uI, vI = int(np.floor(u)), int(np.floor(v))
                if uI - b < 0 or vI - b < 0 or uI + b >= lw or vI + b >= lh:
                    continue
                k = 0
                for yy in range(-half, half + 1):
                    for xx in range(-half, half + 1):
                        r[f * area + k] = _np_bilin(cur, u + xx, v + yy) - T_patch[f * area + k]
                        valid[f * area + k] = True
                        k += 1
            nv = int(valid.sum())
            med = _np_median_rule(r[valid], nv, F * area)
            mad = _np_median_rule(np.abs(r[valid] - med), nv, F * area)
            sigma = max(1.482602218505602 * mad, np.finfo(float).eps)
            c = 4.6851 * sigma
            wgt = np.where(valid & (np.abs(r) <= c), (1 - r**2 / c**2) ** 2, 0.0)
            return float((wgt * r * r).sum()), J.T @ (wgt[:, None] * J), J.T @ (wgt * r), nv

        pre_chi2, preM, it, evals, first = np.finfo(float).max, Mcur.copy(), 0, 0, None
        while it < max_iter:
            chi2, H, g, nv = evaluate(Mcur)
            evals += 1
            dx = np.linalg.solve(H, g)
            if first is None:
                first = dict(H=H, g=g, chi2=chi2)
            if dx.max() > 1e3 or np.isnan(dx).any():
                break
            if chi2 > pre_chi2:
                Mcur = preM.copy()
                break
            preM, pre_chi2 = Mcur.copy(), chi2
            tw = np.zeros((4, 4))
            tw[:3, :3] = np.array([[0, dx[5], -dx[4]], [-dx[5], 0, dx[3]], [dx[4], -dx[3], 0]])      # hat(-omega)
            tw[:3, 3] = -dx[:3]
            Mcur = Mcur @ expm(tw)
            if dx @ dx < 1e-16 or chi2 < 1e-1:
                break
            it += 1
        out.append(dict(first=first, pose=Mcur.copy(), rmse=np.sqrt(chi2 / nv), evaluations=evals))
    return out


def test_align_gn_against_independent_numpy(orc, pkg):
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial.transform import Rotation
    synth = pkg.synth
    T_ref = synth.se3_from_Rt(synth.rodrigues(np.array([0.02, -0.03, 0.01])), [0.4, -0.2, 1.5])   # world != ref frame
    pair = synth.make_pair(index=12, n_features=150, T_ref=tuple(T_ref))   # (with ~45 features the coarsest level sees ONE
    T0 = synth.se3_mul(synth.se3_from_Rt(synth.rodrigues(np.array([1e-3, -2e-3, 5e-4])), [0.01, -0.02, -0.1]), pair["T_ref"])
    feats = pair["feats"][:pair["n_ref"]]                                   # feature: singular normal equations, chaotic in any arithmetic)

    def pyr(img):
        out = [img]
        for _ in range(3):
            out.append(cv2.pyrDown(out[-1]))
        return out
    want = _np_align_gn(pyr(pair["ref"]), pyr(pair["cur"]), feats, pair["T_ref"], T0, pair["K"], max_iter=30)
    rp, cp = orc.build_pyramid(pair["ref"], 4)[0], orc.build_pyramid(pair["cur"], 4)[0]
    rmse, T, status, lv = orc.sparse_align(rp, rp, cp, pair["w"], pair["h"], feats, len(feats), 0, pair["T_ref"], pair["T_ref"],
                                           pair["K"], T0, mode=orc.GN, max_iter=30)
    for s, (o, w_) in enumerate(zip(lv, want)):
        # the first iteration of the level: same normal equations
        assert np.abs(o["H"] - w_["first"]["H"]).max() <= 1e-9 * np.abs(w_["first"]["H"]).max(), s
        assert abs(o["chi2"] - w_["first"]["chi2"]) <= 1e-9 * w_["first"]["chi2"], s
        # the level's result: same number of evaluations (same exits taken), same pose
        assert o["evaluations"] == w_["evaluations"], (s, o["evaluations"], w_["evaluations"])
        R = Rotation.from_quat(o["pose_after"][:4]).as_matrix()
        assert np.abs(R - w_["pose"][:3, :3]).max() < 1e-10 and np.abs(o["pose_after"][4:] - w_["pose"][:3, 3]).max() < 1e-9, s
        assert abs(o["rmse"] - w_["rmse"]) <= 1e-9 * w_["rmse"], s
    assert abs(rmse - want[-1]["rmse"]) <= 1e-9 * want[-1]["rmse"]
    assert sum(w_["evaluations"] for w_ in want) >= 10                       # the loop really iterates
