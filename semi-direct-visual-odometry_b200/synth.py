"""Synthetic KITTI-shaped frame pairs for tests and benchmarks (numpy only).

A textured fronto-parallel plane at depth Z0 in the reference camera is rendered into a
reference frame and, through the exact plane-induced homography of a known motion, into a
current frame.  Camera: KITTI intrinsics (reference tree: resource/kitti.yaml:7-8) on
1241x376 (config/config.json:10-11).  Seeds follow SURVEY.md 8(d): base 20261018 + pair index.
"""
import os

import numpy as np

KITTI_K = (721.5377, 721.5377, 609.5593, 172.8540)
KITTI_W, KITTI_H = 1241, 376
BASE_SEED = 20261018
Z0 = 15.0
MARGIN = 96

FEATURE_DTYPE = np.dtype([("px", "<f8", 2), ("bearing", "<f8", 3), ("point", "<f8", 3),
                          ("has_point", "<i4"), ("reserved", "<i4")])

IDENTITY = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# SE3 helpers; params order qx qy qz qw tx ty tz (Sophus), world -> camera
# ---------------------------------------------------------------------------------------------
def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def se3_from_Rt(R, t):
    return np.concatenate([R_to_quat(R), np.asarray(t, dtype=np.float64)])


def se3_Rt(T):
    return quat_to_R(T[:4]), np.asarray(T[4:7], dtype=np.float64)


def se3_mul(a, b):
    Ra, ta = se3_Rt(a)
    Rb, tb = se3_Rt(b)
    return se3_from_Rt(Ra @ Rb, ta + Ra @ tb)


def se3_inv(T):
    R, t = se3_Rt(T)
    return se3_from_Rt(R.T, -R.T @ t)


def rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def rotation_angle(Ta, Tb):
    """Angle (rad) of the relative rotation between two poses."""
    Ra, _ = se3_Rt(Ta)
    Rb, _ = se3_Rt(Tb)
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    R = Ra.T @ Rb
    s = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arctan2(s, c))


# ---------------------------------------------------------------------------------------------
# texture + rendering
# ---------------------------------------------------------------------------------------------
def make_texture(rng, w=KITTI_W, h=KITTI_H, margin=MARGIN):
    """Smooth random texture (float64, 0..255) of size (h+2m) x (w+2m): noise at 1/4 resolution,
    cubic upsampling, Gaussian blur sigma 1.5, contrast stretch."""
    from scipy import ndimage
    W, H = w + 2 * margin, h + 2 * margin
    lw, lh = (W + 3) // 4 + 2, (H + 3) // 4 + 2
    low = rng.random((lh, lw))
    up = ndimage.zoom(low, 4, order=3)[:H, :W]
    up = ndimage.gaussian_filter(up, 1.5)
    lo, hi = np.percentile(up, 1), np.percentile(up, 99)
    return np.clip((up - lo) / (hi - lo), 0, 1) * 255.0


def sample_bilinear(tex, x, y):
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    x0 = np.clip(x0, 0, tex.shape[1] - 2)
    y0 = np.clip(y0, 0, tex.shape[0] - 2)
    fx = np.clip(x - x0, 0, 1)
    fy = np.clip(y - y0, 0, 1)
    a = tex[y0, x0] * (1 - fx) + tex[y0, x0 + 1] * fx
    b = tex[y0 + 1, x0] * (1 - fx) + tex[y0 + 1, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def render_plane(tex, K, T_cam_ref, w=KITTI_W, h=KITTI_H, margin=MARGIN, z0=Z0):
    """Image seen by a camera whose pose relative to the reference camera is T_cam_ref (ref -> cam)."""
    fx, fy, cx, cy = K
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)  # rays in cam
    R, t = se3_Rt(se3_inv(T_cam_ref))                                       # cam -> ref
    dr = d @ R.T
    s = (z0 - t[2]) / dr[..., 2]
    X = dr * s[..., None] + t
    ur = fx * X[..., 0] / X[..., 2] + cx
    vr = fy * X[..., 1] / X[..., 2] + cy
    img = sample_bilinear(tex, ur + margin, vr + margin)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def abs_gradient_np(img):
    """numpy statement of Simd::AbsGradientSaturatedSum (independent of oracle and product)."""
    s = img.astype(np.int32)
    g = np.zeros_like(s)
    g[1:-1, 1:-1] = np.abs(s[1:-1, 2:] - s[1:-1, :-2]) + np.abs(s[2:, 1:-1] - s[:-2, 1:-1])
    return np.minimum(g, 255).astype(np.uint8)


def grid_argmax_np(grad, cell, thr):
    h, w = grad.shape
    out = []
    for r in range(h // cell + 1):
        for c in range(w // cell + 1):
            roi = grad[r * cell:min((r + 1) * cell, h), c * cell:min((c + 1) * cell, w)]
            if roi.size == 0:
                continue
            k = int(np.argmax(roi))  # first maximum in raster order
            i, j = divmod(k, roi.shape[1])
            if int(roi[i, j]) > thr:
                out.append((c * cell + j, r * cell + i, int(roi[i, j])))
    return np.array(out, dtype=np.int32).reshape(-1, 3)


def random_motion(rng, trans_xy=0.05, trans_z=(0.2, 0.8), rot_deg=0.5):
    """KITTI-like 10 Hz inter-frame motion T_cur_ref."""
    t = np.array([rng.uniform(-trans_xy, trans_xy), rng.uniform(-trans_xy, trans_xy), -rng.uniform(*trans_z)])
    wv = np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3))
    return se3_from_Rt(rodrigues(wv), t)


def make_features(px_xy, K, T_ref, z0=Z0, has_point=None):
    """Feature records for pixels of the reference frame whose 3D points lie on the plane."""
    fx, fy, cx, cy = K
    n = len(px_xy)
    feats = np.zeros(n, FEATURE_DTYPE)
    Tinv = se3_inv(T_ref)
    R, t = se3_Rt(Tinv)
    for i, (x, y) in enumerate(px_xy):
        ray = np.array([(x - cx) / fx, (y - cy) / fy, 1.0])
        feats["px"][i] = (x, y)
        feats["bearing"][i] = ray / np.linalg.norm(ray)   # PinholeCamera::inverseProject2d returns a unit vector
        feats["point"][i] = R @ (ray * z0) + t            # world point on the plane
        feats["has_point"][i] = 1 if has_point is None else int(has_point[i])
    return feats


def make_pair(index=0, n_features=500, cell=30, thr=50, T_ref=None, n_kf=0, base_seed=BASE_SEED, motion_scale=1.0,
              w=KITTI_W, h=KITTI_H, K=KITTI_K):
    """One synthetic frame pair.  Returns a dict with ref/cur/kf images, feature records (n_ref then
    n_kf), poses (T_ref, T_kf, T_cur_true, T_cur_init) and K."""
    rng = np.random.default_rng(base_seed + index)
    tex = make_texture(rng, w, h)
    T_ref = IDENTITY.copy() if T_ref is None else np.asarray(T_ref, dtype=np.float64)
    T_cur_ref = random_motion(rng, 0.05 * motion_scale, (0.2 * motion_scale, 0.8 * motion_scale), 0.5 * motion_scale)
    ref = render_plane(tex, K, IDENTITY, w, h)
    cur = render_plane(tex, K, T_cur_ref, w, h)
    sel = grid_argmax_np(abs_gradient_np(ref), cell, thr)
    n_ref = min(n_features - n_kf, len(sel))
    feats_ref = make_features(sel[:n_ref, :2].astype(np.float64), K, T_ref)
    out = dict(ref=ref, cur=cur, K=np.array(K), T_ref=T_ref, w=w, h=h, T_cur_true=se3_mul(T_cur_ref, T_ref),
               T_cur_init=T_ref.copy(), n_ref=n_ref, n_kf=0, feats=feats_ref, kf=ref, T_kf=T_ref.copy())
    if n_kf > 0:
        # last keyframe: a second view of the same plane, slightly behind the reference frame
        T_kf_ref = random_motion(rng, 0.05, (-0.6, -0.3), 0.3)
        kf = render_plane(tex, K, T_kf_ref, w, h)
        T_kf = se3_mul(T_kf_ref, T_ref)
        selk = grid_argmax_np(abs_gradient_np(kf), cell, thr)
        # features of the keyframe: plane points seen from the keyframe
        Rk, tk = se3_Rt(se3_inv(T_kf_ref))  # kf -> ref
        fxk, fyk, cxk, cyk = K
        fk = np.zeros(min(n_kf, len(selk)), FEATURE_DTYPE)
        Rw, tw = se3_Rt(se3_inv(T_ref))
        for i in range(len(fk)):
            x, y = float(selk[i, 0]), float(selk[i, 1])
            ray = np.array([(x - cxk) / fxk, (y - cyk) / fyk, 1.0])
            dr = Rk @ ray
            s = (Z0 - tk[2]) / dr[2]
            Xref = dr * s + tk
            fk["px"][i] = (x, y)
            fk["bearing"][i] = ray / np.linalg.norm(ray)
            fk["point"][i] = Rw @ Xref + tw
            fk["has_point"][i] = 1
        out.update(kf=kf, T_kf=T_kf, n_kf=len(fk), feats=np.concatenate([feats_ref, fk]))
    return out


# ---------------------------------------------------------------------------------------------
# batched workloads (configs 4 and 5): many independent pairs, generated in parallel on the host
# ---------------------------------------------------------------------------------------------
GROUP = 32          # pairs that share one (larger) texture, each at its own crop
GROUP_SPREAD = 256  # the shared texture is this much larger than a frame in both directions


def _make_group(args):
    """Worker: pairs [first, first + count) of a batch.  Every pair sees its own crop of the group texture,
    its own motion and the grid-argmax features of its own reference frame (cell raster order)."""
    first, count, n_features, cell, thr, base_seed, w, h, K = args
    rng = np.random.default_rng(base_seed + first)
    tex = make_texture(rng, w + GROUP_SPREAD, h + GROUP_SPREAD)
    out = []
    for i in range(count):
        prng = np.random.default_rng(base_seed + first + i)
        ox, oy = (int(v) for v in prng.integers(0, GROUP_SPREAD, 2))
        sub = tex[oy:oy + h + 2 * MARGIN, ox:ox + w + 2 * MARGIN]
        T_cur_ref = random_motion(prng)
        ref = np.clip(np.rint(sub[MARGIN:MARGIN + h, MARGIN:MARGIN + w]), 0, 255).astype(np.uint8)  # identity view
        cur = render_plane(sub, K, T_cur_ref, w, h)
        sel = grid_argmax_np(abs_gradient_np(ref), cell, thr)
        n = min(n_features, len(sel))
        feats = make_features(sel[:n, :2].astype(np.float64), K, IDENTITY)
        out.append((ref, cur, feats, T_cur_ref))
    return first, out


def make_batch(n_pairs, n_features=500, cell=None, thr=50, base_seed=BASE_SEED, first_index=0, workers=None,
               w=KITTI_W, h=KITTI_H, K=KITTI_K):
    """n_pairs independent synthetic pairs (T_ref = I, prior = T_ref).  Returns dict of arrays:
    ref, cur (n, h, w) u8; feats (concatenated records); n_feat (n,), feat_offset (n,); T_true (n, 7)."""
    import multiprocessing as mp
    if cell is None:
        cell = 30 if n_features <= 500 else 20  # cell 30 -> <= 546 features, cell 20 -> <= 1197
    tasks = [(first_index + g, min(GROUP, n_pairs - g), n_features, cell, thr, base_seed, w, h, tuple(K))
             for g in range(0, n_pairs, GROUP)]
    workers = workers or min(len(tasks), os.cpu_count() or 1)
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            groups = pool.map(_make_group, tasks)
    else:
        groups = [_make_group(t) for t in tasks]
    ref = np.empty((n_pairs, h, w), np.uint8)
    cur = np.empty((n_pairs, h, w), np.uint8)
    T_true = np.empty((n_pairs, 7))
    feats, n_feat = [], np.zeros(n_pairs, np.int32)
    for first, items in sorted(groups, key=lambda g: g[0]):
        for i, (r, c, f, T) in enumerate(items):
            k = first - first_index + i
            ref[k], cur[k], T_true[k], n_feat[k] = r, c, T, len(f)
            feats.append(f)
    offs = np.concatenate([[0], np.cumsum(n_feat)[:-1]]).astype(np.int32)
    return dict(ref=ref, cur=cur, feats=np.concatenate(feats), n_feat=n_feat, feat_offset=offs, T_true=T_true,
                K=np.array(K), w=w, h=h, cell=cell)
