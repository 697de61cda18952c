"""ctypes binding of the CPU ORACLE (oracle/libsvo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsvo_oracle.so")

LM_FAITHFUL, LM_ITERATED, GN = 0, 1, 2
MEDIAN_EXACT, MEDIAN_LIBSTDCXX = 0, 1

u8p = C.POINTER(C.c_uint8)
dp = C.POINTER(C.c_double)


class Feature(C.Structure):
    _fields_ = [("px", C.c_double * 2), ("bearing", C.c_double * 3), ("point", C.c_double * 3),
                ("has_point", C.c_int32), ("reserved", C.c_int32)]


FEATURE_DTYPE = np.dtype([("px", "<f8", 2), ("bearing", "<f8", 3), ("point", "<f8", 3),
                          ("has_point", "<i4"), ("reserved", "<i4")])
assert FEATURE_DTYPE.itemsize == C.sizeof(Feature) == 72


class AlignParams(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("min_level", C.c_int32), ("max_level", C.c_int32),
                ("mode", C.c_int32), ("max_iter", C.c_int32), ("median_mode", C.c_int32)]


class LevelStats(C.Structure):
    _fields_ = [("H", C.c_double * 36), ("g", C.c_double * 6), ("dx", C.c_double * 6), ("chi2", C.c_double),
                ("sigma", C.c_double), ("lam", C.c_double), ("pose_after", C.c_double * 7), ("rmse", C.c_double),
                ("n_px", C.c_int32), ("status", C.c_int32), ("iterations", C.c_int32), ("evaluations", C.c_int32)]


class AlignJob(C.Structure):
    _fields_ = [("ref_pyr", C.c_void_p), ("kf_pyr", C.c_void_p), ("cur_pyr", C.c_void_p), ("feats", C.c_void_p),
                ("n_ref", C.c_int32), ("n_kf", C.c_int32), ("T_ref", C.c_double * 7), ("T_kf", C.c_double * 7),
                ("T_cur", C.c_double * 7), ("rmse", C.c_double), ("status", C.c_int32), ("evaluations", C.c_int32)]


class FaParams(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("mode", C.c_int32), ("max_iter", C.c_int32), ("median_mode", C.c_int32)]


class EpiParams(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("mean_mode", C.c_int32)]


class EpiResult(C.Structure):
    _fields_ = [("depth", C.c_double), ("px", C.c_double * 2), ("score", C.c_double), ("found", C.c_int32),
                ("steps", C.c_int32)]


MEAN_EIGEN_U8, MEAN_EXACT = 0, 1


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "svo_oracle.cpp")
    hdr = os.path.join(_HERE, "svo_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libsvo_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_abs_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orc_pyrdown.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orc_pyramid_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_pyramid_bytes.restype = C.c_int64
    L.orc_build_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_grid_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p,
                                  C.c_void_p, C.c_int]
    L.orc_grid_select.restype = C.c_int
    L.orc_select_ssc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p]
    L.orc_select_ssc.restype = C.c_int
    L.orc_bilinear_double.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    L.orc_bilinear_double.restype = C.c_double
    L.orc_bilinear_float.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    L.orc_bilinear_float.restype = C.c_float
    L.orc_median.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_median.restype = C.c_double
    L.orc_sigma.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_sigma.restype = C.c_double
    L.orc_project2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_image_jac.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    L.orc_se3_exp.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_se3_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_se3_act.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_se3_inv.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_ldlt_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_ldlt_solve.restype = C.c_int
    L.orc_sparse_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(AlignParams), C.c_void_p,
                                   C.c_void_p, C.POINTER(C.c_int32)]
    L.orc_sparse_align.restype = C.c_double
    L.orc_sparse_align_batch.argtypes = [C.POINTER(AlignJob), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.POINTER(AlignParams), C.c_int]
    L.orc_feature_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(FaParams), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.orc_feature_align.restype = C.c_double
    L.orc_epipolar_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_double, C.c_double, C.c_double, C.POINTER(EpiParams), C.POINTER(EpiResult)]
    L.orc_epipolar_match.restype = None
    L.orc_hardware_threads.restype = C.c_int
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _f8(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None:
        assert a.size == n, (a.shape, n)
    return a


def level_dims(w, h, levels):
    out = []
    for _ in range(levels):
        out.append((w, h))
        w, h = (w + 1) // 2, (h + 1) // 2
    return out


def abs_gradient(img):
    img = _c8(img)
    h, w = img.shape
    dst = np.empty_like(img)
    lib().orc_abs_gradient(_p(img), w, h, w, _p(dst), w)
    return dst


def pyrdown(img):
    img = _c8(img)
    h, w = img.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown(_p(img), w, h, w, _p(dst), dst.shape[1])
    return dst


def build_pyramid(img, levels):
    """Returns (img_packed, grad_packed) flat uint8 arrays, levels concatenated."""
    img = _c8(img)
    h, w = img.shape
    n = lib().orc_pyramid_bytes(w, h, levels)
    ip = np.empty(n, np.uint8)
    gp = np.empty(n, np.uint8)
    lib().orc_build_pyramid(_p(img), w, h, w, levels, _p(ip), _p(gp))
    return ip, gp


def unpack_pyramid(packed, w, h, levels):
    out, off = [], 0
    for (lw, lh) in level_dims(w, h, levels):
        out.append(packed[off:off + lw * lh].reshape(lh, lw))
        off += lw * lh
    return out


def grid_select(grad, cell, thr, occupancy=None):
    grad = _c8(grad)
    h, w = grad.shape
    rows, cols = h // cell + 1, w // cell + 1
    out = np.zeros((rows * cols, 3), np.int32)
    occ = None
    if occupancy is not None:
        occ = _c8(np.asarray(occupancy).reshape(-1))
        assert occ.size == rows * cols
    n = lib().orc_grid_select(_p(grad), w, h, w, cell, thr, _p(occ) if occ is not None else None, _p(out), rows * cols)
    return out[:n].copy()


def select_ssc(grad, thr, num_candidates, cell=30, occupancy=None, use_bucketing=True):
    """FeatureSelection::gradientMagnitudeWithSSC.  Returns ((n, 3) int32 array of x, y, magnitude; info dict)."""
    grad = _c8(grad)
    h, w = grad.shape
    cap = h * w
    out = np.zeros((cap, 3), np.int32)
    info = np.zeros(4, np.int32)
    occ = None if occupancy is None else np.ascontiguousarray(np.asarray(occupancy).reshape(-1), dtype=np.uint8)
    n = lib().orc_select_ssc(_p(grad), w, h, w, thr, num_candidates, cell, _p(occ) if occ is not None else None,
                             1 if use_bucketing else 0, _p(out), cap, _p(info))
    return out[:n].copy(), dict(keypoints=int(info[0]), width=int(info[1]), iterations=int(info[2]), ssc_points=int(info[3]))


def bilinear_double(img, x, y):
    img = _c8(img)
    return lib().orc_bilinear_double(_p(img), img.shape[1], x, y)


def bilinear_float(img, x, y):
    img = _c8(img)
    return lib().orc_bilinear_float(_p(img), img.shape[1], x, y)


def median(v, num_valid, mode=MEDIAN_EXACT):
    v = _f8(v)
    return lib().orc_median(_p(v), v.size, num_valid, mode)


def sigma(v, num_valid, mode=MEDIAN_EXACT):
    v = _f8(v)
    return lib().orc_sigma(_p(v), v.size, num_valid, mode)


def project2d(K, p):
    K, p = _f8(K, 4), _f8(p, 3)
    uv = np.zeros(2)
    lib().orc_project2d(_p(K), _p(p), _p(uv))
    return uv


def image_jac(p, fx, fy):
    p = _f8(p, 3)
    J = np.zeros(12)
    lib().orc_image_jac(_p(p), fx, fy, _p(J))
    return J.reshape(2, 6)


def se3_exp(xi):
    xi = _f8(xi, 6)
    out = np.zeros(7)
    lib().orc_se3_exp(_p(xi), _p(out))
    return out


def se3_mul(a, b):
    a, b = _f8(a, 7), _f8(b, 7)
    out = np.zeros(7)
    lib().orc_se3_mul(_p(a), _p(b), _p(out))
    return out


def se3_act(T, p):
    T, p = _f8(T, 7), _f8(p, 3)
    out = np.zeros(3)
    lib().orc_se3_act(_p(T), _p(p), _p(out))
    return out


def se3_inv(T):
    T = _f8(T, 7)
    out = np.zeros(7)
    lib().orc_se3_inv(_p(T), _p(out))
    return out


def ldlt_solve(A, b):
    A = _f8(A)
    b = _f8(b)
    x = np.zeros(b.size)
    lib().orc_ldlt_solve(_p(A), _p(b), b.size, _p(x))
    return x


def stats_to_dict(s):
    return dict(H=np.array(s.H).reshape(6, 6), g=np.array(s.g), dx=np.array(s.dx), chi2=s.chi2, sigma=s.sigma,
                lam=s.lam, pose_after=np.array(s.pose_after), rmse=s.rmse, n_px=s.n_px, status=s.status,
                iterations=s.iterations, evaluations=s.evaluations)


def sparse_align(ref_pyr, kf_pyr, cur_pyr, w, h, feats, n_ref, n_kf, T_ref, T_kf, K, T_cur, patch_size=5, min_level=0,
                 max_level=3, mode=LM_FAITHFUL, max_iter=20, median_mode=MEDIAN_EXACT):
    """Returns (rmse, T_cur_out, status, [per-level dicts])."""
    feats = np.ascontiguousarray(feats, dtype=FEATURE_DTYPE)
    assert feats.size == n_ref + n_kf
    prm = AlignParams(patch_size, min_level, max_level, mode, max_iter, median_mode)
    nl = max_level - min_level + 1
    stats = (LevelStats * nl)()
    T = _f8(T_cur, 7).copy()
    Tr, Tk, Kk = _f8(T_ref, 7), _f8(T_kf, 7), _f8(K, 4)
    status = C.c_int32(0)
    ref_pyr, kf_pyr, cur_pyr = _c8(ref_pyr), _c8(kf_pyr), _c8(cur_pyr)
    rmse = lib().orc_sparse_align(_p(ref_pyr), _p(kf_pyr), _p(cur_pyr), w, h, _p(feats), n_ref, n_kf, _p(Tr), _p(Tk),
                                  _p(Kk), C.byref(prm), _p(T), C.cast(stats, C.c_void_p), C.byref(status))
    return rmse, T, status.value, [stats_to_dict(s) for s in stats]


def sparse_align_batch(jobs, w, h, K, n_threads, patch_size=5, min_level=0, max_level=3, mode=LM_FAITHFUL, max_iter=20,
                       median_mode=MEDIAN_EXACT):
    """jobs: list of dicts(ref_pyr, kf_pyr, cur_pyr, feats, n_ref, n_kf, T_ref, T_kf, T_cur).
    Returns (T_out[n,7], rmse[n], status[n], evaluations[n])."""
    n = len(jobs)
    arr = (AlignJob * n)()
    keep = []
    for i, j in enumerate(jobs):
        rp, kp, cp = _c8(j["ref_pyr"]), _c8(j["kf_pyr"]), _c8(j["cur_pyr"])
        ft = np.ascontiguousarray(j["feats"], dtype=FEATURE_DTYPE)
        keep += [rp, kp, cp, ft]
        arr[i].ref_pyr, arr[i].kf_pyr, arr[i].cur_pyr = rp.ctypes.data, kp.ctypes.data, cp.ctypes.data
        arr[i].feats = ft.ctypes.data
        arr[i].n_ref, arr[i].n_kf = j["n_ref"], j["n_kf"]
        for k in range(7):
            arr[i].T_ref[k] = j["T_ref"][k]
            arr[i].T_kf[k] = j["T_kf"][k]
            arr[i].T_cur[k] = j["T_cur"][k]
    prm = AlignParams(patch_size, min_level, max_level, mode, max_iter, median_mode)
    Kk = _f8(K, 4)
    lib().orc_sparse_align_batch(arr, n, w, h, _p(Kk), C.byref(prm), n_threads)
    T = np.array([[arr[i].T_cur[k] for k in range(7)] for i in range(n)])
    return (T, np.array([arr[i].rmse for i in range(n)]), np.array([arr[i].status for i in range(n)]),
            np.array([arr[i].evaluations for i in range(n)]))


def feature_align(ref_grad, cur_grad, ref_px, px_start, A=None, patch_size=7, mode=LM_FAITHFUL, max_iter=20,
                  median_mode=MEDIAN_EXACT):
    """Returns (rmse, px_out[2], status, iterations)."""
    ref_grad, cur_grad = _c8(ref_grad), _c8(cur_grad)
    h, w = ref_grad.shape
    rp = _f8(ref_px, 2)
    px = _f8(px_start, 2).copy()
    Aa = _f8(A, 4) if A is not None else None
    prm = FaParams(patch_size, mode, max_iter, median_mode)
    st, it = C.c_int32(0), C.c_int32(0)
    rmse = lib().orc_feature_align(_p(ref_grad), _p(cur_grad), w, h, _p(rp), _p(Aa) if Aa is not None else None, _p(px),
                                   C.byref(prm), C.byref(st), C.byref(it))
    return rmse, px, st.value, it.value


def epipolar_match(ref_img, cur_img, K, T_rel, ref_px, ref_bearing, depth, min_depth, max_depth, patch_size=7,
                   mean_mode=MEAN_EIGEN_U8):
    """algorithm::matchEpipolarConstraint for one seed.  Returns dict(found, depth, px, score, steps)."""
    ref_img, cur_img = _c8(ref_img), _c8(cur_img)
    h, w = ref_img.shape
    Kk, Tr, rp, rb = _f8(K, 4), _f8(T_rel, 7), _f8(ref_px, 2), _f8(ref_bearing, 3)
    prm = EpiParams(patch_size, mean_mode)
    out = EpiResult()
    lib().orc_epipolar_match(_p(ref_img), _p(cur_img), w, h, _p(Kk), _p(Tr), _p(rp), _p(rb), float(depth), float(min_depth),
                             float(max_depth), C.byref(prm), C.byref(out))
    return dict(found=bool(out.found), depth=out.depth, px=np.array([out.px[0], out.px[1]]), score=out.score, steps=out.steps)


REPROJ_CAND_DTYPE = np.dtype([("ref_slot", "<i4"), ("type", "<i4"), ("ref_px", "<f8", 2), ("point", "<f8", 3)])


def reproject_map(grads, cur_grad, K, T_cur, cands, cell, cell_order, max_matches=150, patch_size=7, mode=LM_FAITHFUL, max_iter=20):
    """Map::reprojectMap.  grads: list of gradient images indexed by the candidates' ref_slot.
    Returns (matches (m, 6): cell, candidate, px x, px y, rmse, status; projected (n,) uint8)."""
    grads = [_c8(g) for g in grads]
    cur_grad = _c8(cur_grad)
    h, w = cur_grad.shape
    ptrs = (C.c_void_p * len(grads))(*[g.ctypes.data for g in grads])
    cands = np.ascontiguousarray(cands, dtype=REPROJ_CAND_DTYPE).reshape(-1)
    order = np.ascontiguousarray(cell_order, dtype=np.int32)
    Kk, Tc = _f8(K, 4), _f8(T_cur, 7)
    prm = FaParams(patch_size, mode, max_iter, MEDIAN_EXACT)
    out = np.zeros((max_matches + 1, 6), np.float64)
    proj = np.zeros(max(1, cands.size), np.uint8)
    L = lib()
    L.orc_reproject_map.restype = C.c_int
    L.orc_reproject_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_int, C.c_int, C.POINTER(FaParams), C.c_void_p, C.c_void_p]
    m = L.orc_reproject_map(ptrs, _p(cur_grad), w, h, _p(Kk), _p(Tc), _p(cands), cands.size, cell, _p(order), order.size, max_matches,
                            C.byref(prm), _p(out), _p(proj))
    return out[:m].copy(), proj[:cands.size].copy()


class KltParams(C.Structure):
    _fields_ = [("win", C.c_int32), ("max_level", C.c_int32), ("max_count", C.c_int32), ("use_initial_flow", C.c_int32),
                ("epsilon", C.c_double), ("min_eig_threshold", C.c_double)]


def klt_track(ref_img, cur_img, prev_pts, next_pts=None, win=11, max_level=3, max_count=30, epsilon=1e-4, min_eig_threshold=1e-4):
    """cv::calcOpticalFlowPyrLK as algorithm::computeOpticalFlowSparse calls it.  Returns (next_pts (n, 2) float32, status (n,) uint8,
    err (n,) float32, top level used).  next_pts given = OPTFLOW_USE_INITIAL_FLOW."""
    ref_img, cur_img = _c8(ref_img), _c8(cur_img)
    h, w = ref_img.shape
    prev = np.ascontiguousarray(prev_pts, dtype=np.float32).reshape(-1, 2)
    nxt = np.ascontiguousarray(next_pts if next_pts is not None else prev, dtype=np.float32).reshape(-1, 2).copy()
    n = prev.shape[0]
    prm = KltParams(win, max_level, max_count, 1 if next_pts is not None else 0, epsilon, min_eig_threshold)
    status, err = np.zeros(max(1, n), np.uint8), np.zeros(max(1, n), np.float32)
    L = lib()
    L.orc_klt_track.restype = C.c_int
    L.orc_klt_track.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(KltParams),
                                C.c_void_p, C.c_void_p]
    top = L.orc_klt_track(_p(ref_img), _p(cur_img), w, h, _p(prev), _p(nxt), n, C.byref(prm), _p(status), _p(err))
    return nxt, status[:n].copy(), err[:n].copy(), top


def hardware_threads():
    return lib().orc_hardware_threads()
