"""ctypes binding of libsvo_b200.so, the C ABI declared in include/svo_b200.h.

This is the Python-side harness over the product library (tests, bench.py, smoke); the
reference-shaped C++ host classes live in host/.  There is NO CPU fallback: loading fails
loudly if the library has not been built, and Context() fails if there is no CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVO_B200_LIB", os.path.join(_HERE, "libsvo_b200.so"))  # override: A/B builds of the same ABI

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_NO_DEVICE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
LM_FAITHFUL, LM_ITERATED, GN = 0, 1, 2
ST_SUCCESS, ST_FAILED = 0, 9  # Optimizer::Status (include/optimizer.hpp:21-33), the two the harness tests against
MAX_LEVELS = 8

# every symbol include/svo_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "svo_create", "svo_destroy", "svo_last_error", "svo_version", "svo_sync", "svo_launch_count", "svo_stream",
    "svo_host_alloc", "svo_host_free", "svo_level_dims", "svo_frames_upload", "svo_frames_prefetch", "svo_frames_upload_device",
    "svo_frames_rebuild", "svo_frame_download", "svo_select_grid", "svo_sparse_align", "svo_sparse_align_stage",
    "svo_sparse_align_h2d", "svo_sparse_align_launch", "svo_sparse_align_d2h", "svo_sparse_align_fetch",
    "svo_sparse_align_results_device", "svo_debug_cycles",
    "svo_feature_align", "svo_feature_align_stage", "svo_feature_align_h2d", "svo_feature_align_launch",
    "svo_feature_align_d2h", "svo_feature_align_fetch", "svo_frontend_run", "svo_frontend_image_buffer",
    "svo_epipolar_match", "svo_select_ssc", "svo_reproject_map", "svo_klt_track",
    "svo_multi_create", "svo_multi_destroy", "svo_multi_devices", "svo_multi_ctx", "svo_multi_last_error", "svo_multi_shard",
    "svo_multi_frames_upload", "svo_multi_sparse_align", "svo_multi_sparse_align_stage", "svo_multi_sparse_align_launch",
    "svo_multi_sparse_align_fetch", "svo_multi_sync", "svo_multi_time_launches",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("levels", C.c_int32),
                ("max_frames", C.c_int32), ("max_jobs", C.c_int32), ("max_features", C.c_int32),
                ("max_fa_items", C.c_int32), ("reserved", C.c_int32), ("stream", C.c_void_p), ("K", C.c_double * 4)]


class AlignParams(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("min_level", C.c_int32), ("max_level", C.c_int32), ("mode", C.c_int32),
                ("max_iter", C.c_int32), ("reserved", C.c_int32)]


class FaParams(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("mode", C.c_int32), ("max_iter", C.c_int32), ("reserved", C.c_int32)]


class KltParams(C.Structure):
    _fields_ = [("win", C.c_int32), ("max_level", C.c_int32), ("max_count", C.c_int32), ("use_initial_flow", C.c_int32),
                ("epsilon", C.c_double), ("min_eig_threshold", C.c_double)]


class FrontendParams(C.Structure):
    _fields_ = [("ref_slot", C.c_int32), ("kf_slot", C.c_int32), ("cur_slot", C.c_int32), ("cell", C.c_int32),
                ("thr", C.c_uint32), ("max_features", C.c_int32), ("align", AlignParams), ("fa", FaParams)]


FEATURE_PX_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("magnitude", "<i4")])
ALIGN_FEATURE_DTYPE = np.dtype([("px", "<f8", 2), ("bearing", "<f8", 3), ("point", "<f8", 3), ("has_point", "<i4"),
                                ("reserved", "<i4")])
ALIGN_JOB_DTYPE = np.dtype([("ref_slot", "<i4"), ("kf_slot", "<i4"), ("cur_slot", "<i4"), ("n_ref", "<i4"),
                            ("n_kf", "<i4"), ("feat_offset", "<i4"), ("T_ref", "<f8", 7), ("T_kf", "<f8", 7),
                            ("T_cur", "<f8", 7)])
ALIGN_RESULT_DTYPE = np.dtype([("T_cur", "<f8", 7), ("rmse", "<f8"), ("status", "<i4"), ("evaluations", "<i4"),
                               ("iterations", "<i4"), ("reserved", "<i4")])
FRONTEND_RESULT_DTYPE = np.dtype([("align", ALIGN_RESULT_DTYPE), ("n_selected", "<i4"), ("n_candidates", "<i4")])
ALIGN_STATS_DTYPE = np.dtype([("H", "<f8", (6, 6)), ("g", "<f8", 6), ("dx", "<f8", 6), ("chi2", "<f8"), ("sigma", "<f8"),
                              ("lam", "<f8"), ("pose_after", "<f8", 7), ("rmse", "<f8"), ("n_px", "<i4"),
                              ("status", "<i4"), ("iterations", "<i4"), ("evaluations", "<i4")])
FA_ITEM_DTYPE = np.dtype([("ref_slot", "<i4"), ("cur_slot", "<i4"), ("ref_px", "<f8", 2), ("px", "<f8", 2),
                          ("A", "<f8", 4), ("use_affine", "<i4"), ("reserved", "<i4")])
REPROJ_CAND_DTYPE = np.dtype([("ref_slot", "<i4"), ("type", "<i4"), ("ref_px", "<f8", 2), ("point", "<f8", 3)])
REPROJ_MATCH_DTYPE = np.dtype([("cell", "<i4"), ("candidate", "<i4"), ("px", "<f8", 2), ("rmse", "<f8"), ("status", "<i4"),
                               ("reserved", "<i4")])
EPI_ITEM_DTYPE = np.dtype([("ref_slot", "<i4"), ("cur_slot", "<i4"), ("T_rel", "<f8", 7), ("px", "<f8", 2), ("bearing", "<f8", 3),
                           ("depth", "<f8"), ("min_depth", "<f8"), ("max_depth", "<f8")])
EPI_RESULT_DTYPE = np.dtype([("depth", "<f8"), ("px", "<f8", 2), ("score", "<f8"), ("found", "<i4"), ("steps", "<i4")])
MEAN_EIGEN_U8, MEAN_EXACT = 0, 1


class EpiParams(C.Structure):
    _fields_ = [("patch_size", C.c_int32), ("mean_mode", C.c_int32)]


FA_RESULT_DTYPE = np.dtype([("px", "<f8", 2), ("rmse", "<f8"), ("status", "<i4"), ("iterations", "<i4")])
assert ALIGN_FEATURE_DTYPE.itemsize == 72 and ALIGN_JOB_DTYPE.itemsize == 192 and ALIGN_RESULT_DTYPE.itemsize == 80
assert ALIGN_STATS_DTYPE.itemsize == 488 and FA_ITEM_DTYPE.itemsize == 80 and FA_RESULT_DTYPE.itemsize == 32

_lib = None


class SvoError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("svo_b200 error %d: %s" % (code, text))
        self.code = code


def load():
    """dlopen libsvo_b200.so; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SVO_B200_LIB", LIB_PATH)  # diagnostics: the instrumented twin built by `make prof`
    if not os.path.exists(path):
        raise ImportError("libsvo_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    L = C.CDLL(path)
    vp, i, i64 = C.c_void_p, C.c_int, C.c_int64
    L.svo_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.svo_destroy.argtypes = [vp]
    L.svo_destroy.restype = None
    L.svo_last_error.argtypes = [vp]
    L.svo_last_error.restype = C.c_char_p
    L.svo_version.restype = C.c_char_p
    L.svo_sync.argtypes = [vp]
    L.svo_launch_count.argtypes = [vp]
    L.svo_launch_count.restype = i64
    L.svo_stream.argtypes = [vp]
    L.svo_stream.restype = vp
    L.svo_host_alloc.argtypes = [vp, i64, C.POINTER(vp)]
    L.svo_host_free.argtypes = [vp, vp]
    L.svo_level_dims.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.svo_frames_upload.argtypes = [vp, i, i, vp, i, i64]
    L.svo_frames_prefetch.argtypes = [vp, i, i, vp, i, i64]
    L.svo_frames_upload_device.argtypes = [vp, i, i, vp, i, i64]
    L.svo_frames_rebuild.argtypes = [vp, i, i]
    L.svo_frame_download.argtypes = [vp, i, i, i, vp, i]
    L.svo_select_grid.argtypes = [vp, i, i, C.c_uint32, vp, vp, i, C.POINTER(i)]
    L.svo_sparse_align.argtypes = [vp, vp, i, vp, i, C.POINTER(AlignParams), vp, vp]
    L.svo_sparse_align_stage.argtypes = [vp, vp, i, vp, i, C.POINTER(AlignParams), i]
    for n in ("h2d", "launch", "d2h"):
        getattr(L, "svo_sparse_align_" + n).argtypes = [vp]
        getattr(L, "svo_feature_align_" + n).argtypes = [vp]
    L.svo_sparse_align_fetch.argtypes = [vp, vp, vp]
    L.svo_sparse_align_results_device.argtypes = [vp]
    L.svo_debug_cycles.argtypes = [vp, vp]
    L.svo_multi_create.argtypes = [C.POINTER(Config), vp, i, C.POINTER(vp)]
    L.svo_multi_destroy.argtypes = [vp]
    L.svo_multi_destroy.restype = None
    L.svo_multi_devices.argtypes = [vp]
    L.svo_multi_ctx.argtypes = [vp, i]
    L.svo_multi_ctx.restype = vp
    L.svo_multi_last_error.argtypes = [vp]
    L.svo_multi_last_error.restype = C.c_char_p
    L.svo_multi_shard.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    L.svo_multi_shard.restype = None
    L.svo_multi_frames_upload.argtypes = [vp, i, i, vp, i, i64, i]
    L.svo_multi_sparse_align.argtypes = [vp, vp, i, vp, i, C.POINTER(AlignParams), vp, vp]
    L.svo_multi_sparse_align_stage.argtypes = [vp, vp, i, vp, i, C.POINTER(AlignParams), i]
    L.svo_multi_sparse_align_launch.argtypes = [vp]
    L.svo_multi_sparse_align_fetch.argtypes = [vp, vp, vp]
    L.svo_multi_sync.argtypes = [vp]
    L.svo_multi_time_launches.argtypes = [vp, i, i, C.POINTER(C.c_double)]
    L.svo_sparse_align_results_device.restype = vp
    L.svo_reproject_map.argtypes = [vp, i, vp, vp, i, i, vp, i, i, C.POINTER(FaParams), vp, C.POINTER(i), vp]
    L.svo_klt_track.argtypes = [vp, i, i, vp, vp, i, C.POINTER(KltParams), vp, vp]
    L.svo_select_ssc.argtypes = [vp, i, C.c_uint32, i, i, vp, i, vp, i, C.POINTER(i), vp]
    L.svo_epipolar_match.argtypes = [vp, vp, i, C.POINTER(EpiParams), vp]
    L.svo_frontend_run.argtypes = [vp, C.POINTER(FrontendParams), vp, i, vp, vp, i, vp, vp, vp, i, vp]
    L.svo_frontend_image_buffer.argtypes = [vp]
    L.svo_frontend_image_buffer.restype = vp
    L.svo_feature_align.argtypes = [vp, vp, i, C.POINTER(FaParams), vp]
    L.svo_feature_align_stage.argtypes = [vp, vp, i, C.POINTER(FaParams)]
    L.svo_feature_align_fetch.argtypes = [vp, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PinnedBuffer:
    """Page-locked host memory from svo_host_alloc, viewed as a numpy uint8 array."""

    def __init__(self, ctx, nbytes):
        self._ctx = ctx
        p = C.c_void_p()
        ctx._check(ctx.L.svo_host_alloc(ctx.h, nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def view(self, dtype, count, offset=0):
        """A structured-array view (no copy) of `count` records at byte `offset`."""
        dt = np.dtype(dtype)
        return self.array[offset:offset + count * dt.itemsize].view(dt)

    def free(self):
        if self.ptr:
            self._ctx.L.svo_host_free(self._ctx.h, self.ptr)
            self.ptr = None
            self.array = None


def shard(n_items, n_parts, part):
    """svo_multi_shard: block [lo, hi) of `part` when n_items are cut into n_parts contiguous blocks"""
    lo, hi = C.c_int(), C.c_int()
    load().svo_multi_shard(n_items, n_parts, part, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


class MultiContext:
    """svo_multi: one context + host worker thread per device of this process, batches cut into contiguous shards."""

    def __init__(self, width, height, K, n_devices, devices=None, levels=4, max_frames=4, max_jobs=1, max_features=1024,
                 max_fa_items=16):
        self.L = load()
        cfg = Config(0, width, height, levels, max_frames, max_jobs, max_features, max_fa_items, 0, None,
                     (C.c_double * 4)(*[float(k) for k in K]))
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        st = self.L.svo_multi_create(C.byref(cfg), _ptr(dv), n_devices, C.byref(h))
        if st != OK:
            raise SvoError(st, self.L.svo_last_error(None).decode())
        self.h, self.n, self.width, self.height = h, n_devices, width, height
        self._staged = None

    def _check(self, st):
        if st != OK:
            raise SvoError(st, self.L.svo_multi_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.svo_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, first_slot, imgs, prefetch=False):
        a = np.ascontiguousarray(imgs, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[1:] == (self.height, self.width)
        self._check(self.L.svo_multi_frames_upload(self.h, first_slot, a.shape[0], a.ctypes.data, a.strides[1], a.strides[0],
                                                   1 if prefetch else 0))
        return a

    def sparse_align(self, jobs, feats, patch_size=5, min_level=0, max_level=3, mode=LM_FAITHFUL, max_iter=20, want_stats=True):
        jobs = np.ascontiguousarray(jobs, dtype=ALIGN_JOB_DTYPE).reshape(-1)
        feats = np.ascontiguousarray(feats, dtype=ALIGN_FEATURE_DTYPE).reshape(-1)
        prm = AlignParams(patch_size, min_level, max_level, mode, max_iter, 0)
        res = np.zeros(jobs.size, ALIGN_RESULT_DTYPE)
        stats = np.zeros((jobs.size, max_level - min_level + 1), ALIGN_STATS_DTYPE) if want_stats else None
        self._check(self.L.svo_multi_sparse_align(self.h, _ptr(jobs), jobs.size, _ptr(feats), feats.size, C.byref(prm), _ptr(res),
                                                  _ptr(stats)))
        return res, stats

    def stage(self, jobs, feats, patch_size=5, min_level=0, max_level=3, mode=LM_FAITHFUL, max_iter=20):
        jobs = np.ascontiguousarray(jobs, dtype=ALIGN_JOB_DTYPE).reshape(-1)
        feats = np.ascontiguousarray(feats, dtype=ALIGN_FEATURE_DTYPE).reshape(-1)
        prm = AlignParams(patch_size, min_level, max_level, mode, max_iter, 0)
        self._check(self.L.svo_multi_sparse_align_stage(self.h, _ptr(jobs), jobs.size, _ptr(feats), feats.size, C.byref(prm), 0))
        self._staged = jobs.size

    def launch(self):
        self._check(self.L.svo_multi_sparse_align_launch(self.h))

    def fetch(self):
        res = np.zeros(self._staged, ALIGN_RESULT_DTYPE)
        self._check(self.L.svo_multi_sparse_align_fetch(self.h, _ptr(res), None))
        return res

    def sync(self):
        self._check(self.L.svo_multi_sync(self.h))

    def time_launches(self, warmup, steps):
        ms = C.c_double()
        self._check(self.L.svo_multi_time_launches(self.h, warmup, steps, C.byref(ms)))
        return ms.value


class Context:
    """One svo_ctx: a CUDA stream, a device arena of frame slots (pyramids) and pinned staging."""

    def __init__(self, width, height, K, levels=4, max_frames=4, max_jobs=1, max_features=1024, max_fa_items=2048,
                 device=0, stream=None):
        self.L = load()
        cfg = Config(device, width, height, levels, max_frames, max_jobs, max_features, max_fa_items, 0, stream,
                     (C.c_double * 4)(*[float(k) for k in K]))
        h = C.c_void_p()
        st = self.L.svo_create(C.byref(cfg), C.byref(h))
        if st != OK:
            raise SvoError(st, self.L.svo_last_error(None).decode())
        self.h = h
        self.width, self.height, self.levels, self.max_frames = width, height, levels, max_frames
        self.dims = []
        for l in range(levels):
            w, hh, p = C.c_int(), C.c_int(), C.c_int()
            self._check(self.L.svo_level_dims(self.h, l, C.byref(w), C.byref(hh), C.byref(p)))
            self.dims.append((w.value, hh.value, p.value))

    def _check(self, st):
        if st != OK:
            raise SvoError(st, self.L.svo_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.svo_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def sync(self):
        self._check(self.L.svo_sync(self.h))

    @property
    def stream(self):
        return self.L.svo_stream(self.h)

    @property
    def launches(self):
        return self.L.svo_launch_count(self.h)

    def pinned(self, nbytes):
        return PinnedBuffer(self, nbytes)

    # ---- ImagePyramid ----
    def upload(self, first_slot, imgs):
        """imgs: (h, w) or (n, h, w) uint8; builds image + gradient stacks of the slots."""
        a = np.asarray(imgs)
        assert a.dtype == np.uint8
        if a.ndim == 2:
            a = a[None]
        assert a.shape[1:] == (self.height, self.width), a.shape
        if not (a.strides[2] == 1 and a.strides[1] >= self.width):
            a = np.ascontiguousarray(a)
        self._check(self.L.svo_frames_upload(self.h, first_slot, a.shape[0], a.ctypes.data, a.strides[1], a.strides[0]))
        return a  # keep alive until the stream has consumed it (pinned path is asynchronous)

    def prefetch(self, first_slot, imgs):
        """upload() on the ingest streams: overlaps work already enqueued (svo_frames_prefetch); the slots must not be
        referenced by a batch that is still in flight."""
        a = np.asarray(imgs)
        assert a.dtype == np.uint8
        if a.ndim == 2:
            a = a[None]
        assert a.shape[1:] == (self.height, self.width), a.shape
        if not (a.strides[2] == 1 and a.strides[1] >= self.width):
            a = np.ascontiguousarray(a)
        self._check(self.L.svo_frames_prefetch(self.h, first_slot, a.shape[0], a.ctypes.data, a.strides[1], a.strides[0]))
        return a

    def upload_device(self, first_slot, n, dptr, pitch, frame_stride):
        self._check(self.L.svo_frames_upload_device(self.h, first_slot, n, dptr, pitch, frame_stride))

    def rebuild(self, first_slot, n):
        self._check(self.L.svo_frames_rebuild(self.h, first_slot, n))

    def download(self, slot, level, which=0):
        w, h, _ = self.dims[level]
        out = np.empty((h, w), np.uint8)
        self._check(self.L.svo_frame_download(self.h, slot, level, which, out.ctypes.data, w))
        return out

    # ---- FeatureSelection::gradientMagnitudeByValue ----
    def select_grid(self, slot, cell=30, thr=50, occupancy=None):
        rows, cols = self.height // cell + 1, self.width // cell + 1
        out = np.zeros(rows * cols, FEATURE_PX_DTYPE)
        occ = None
        if occupancy is not None:
            occ = np.ascontiguousarray(np.asarray(occupancy).reshape(-1), dtype=np.uint8)
            assert occ.size == rows * cols
        n = C.c_int()
        self._check(self.L.svo_select_grid(self.h, slot, cell, thr, _ptr(occ), _ptr(out), out.size, C.byref(n)))
        return out[:n.value].copy()

    # ---- FeatureSelection::gradientMagnitudeWithSSC ----
    def select_ssc(self, slot, thr=50, num_candidates=250, cell=30, occupancy=None, use_bucketing=True):
        rows, cols = self.height // cell + 1, self.width // cell + 1
        out = np.zeros(4096, FEATURE_PX_DTYPE)
        occ = None
        if occupancy is not None:
            occ = np.ascontiguousarray(np.asarray(occupancy).reshape(-1), dtype=np.uint8)
            assert occ.size == rows * cols
        n = C.c_int()
        info = np.zeros(4, np.int32)
        self._check(self.L.svo_select_ssc(self.h, slot, thr, num_candidates, cell, _ptr(occ), 1 if use_bucketing else 0, _ptr(out),
                                          out.size, C.byref(n), _ptr(info)))
        return out[:n.value].copy(), dict(keypoints=int(info[0]), width=int(info[1]), iterations=int(info[2]), ssc_points=int(info[3]))

    # ---- ImageAlignment::align ----
    def sparse_align(self, jobs, feats, patch_size=5, min_level=0, max_level=3, mode=LM_FAITHFUL, max_iter=20,
                     want_stats=True):
        jobs = np.ascontiguousarray(jobs, dtype=ALIGN_JOB_DTYPE).reshape(-1)
        feats = np.ascontiguousarray(feats, dtype=ALIGN_FEATURE_DTYPE).reshape(-1)
        prm = AlignParams(patch_size, min_level, max_level, mode, max_iter, 0)
        res = np.zeros(jobs.size, ALIGN_RESULT_DTYPE)
        nl = max_level - min_level + 1
        stats = np.zeros((jobs.size, nl), ALIGN_STATS_DTYPE) if want_stats else None
        self._check(self.L.svo_sparse_align(self.h, _ptr(jobs), jobs.size, _ptr(feats), feats.size, C.byref(prm),
                                            _ptr(res), _ptr(stats)))
        return res, stats

    def sparse_align_stage(self, jobs, feats, patch_size=5, min_level=0, max_level=3, mode=LM_FAITHFUL, max_iter=20,
                           want_stats=False):
        jobs = np.ascontiguousarray(jobs, dtype=ALIGN_JOB_DTYPE).reshape(-1)
        feats = np.ascontiguousarray(feats, dtype=ALIGN_FEATURE_DTYPE).reshape(-1)
        prm = AlignParams(patch_size, min_level, max_level, mode, max_iter, 0)
        self._check(self.L.svo_sparse_align_stage(self.h, _ptr(jobs), jobs.size, _ptr(feats), feats.size, C.byref(prm),
                                                  1 if want_stats else 0))
        self._staged = (jobs.size, max_level - min_level + 1, want_stats)

    def sparse_align_h2d(self):
        self._check(self.L.svo_sparse_align_h2d(self.h))

    def sparse_align_launch(self):
        self._check(self.L.svo_sparse_align_launch(self.h))

    def sparse_align_d2h(self):
        self._check(self.L.svo_sparse_align_d2h(self.h))

    def sparse_align_fetch(self):
        n, nl, ws = self._staged
        res = np.zeros(n, ALIGN_RESULT_DTYPE)
        stats = np.zeros((n, nl), ALIGN_STATS_DTYPE) if ws else None
        self._check(self.L.svo_sparse_align_fetch(self.h, _ptr(res), _ptr(stats)))
        return res, stats

    def debug_cycles(self):
        out = np.zeros((8, 8), np.int64)
        self._check(self.L.svo_debug_cycles(self.h, _ptr(out)))
        return out

    @property
    def results_device_ptr(self):
        return self.L.svo_sparse_align_results_device(self.h)

    # ---- Map::reprojectMap: projection, per-cell choice, one FeatureAlignment launch ----
    def reproject_map(self, cur_slot, T_cur, cands, cell, cell_order, max_matches=150, patch_size=7, mode=LM_FAITHFUL, max_iter=20):
        cands = np.ascontiguousarray(cands, dtype=REPROJ_CAND_DTYPE).reshape(-1)
        order = np.ascontiguousarray(cell_order, dtype=np.int32)
        T = np.ascontiguousarray(T_cur, dtype=np.float64)
        prm = FaParams(patch_size, mode, max_iter, 0)
        out = np.zeros(max_matches + 1, REPROJ_MATCH_DTYPE)
        proj = np.zeros(max(1, cands.size), np.uint8)
        n = C.c_int()
        self._check(self.L.svo_reproject_map(self.h, cur_slot, _ptr(T), _ptr(cands), cands.size, cell, _ptr(order), order.size,
                                             max_matches, C.byref(prm), _ptr(out), C.byref(n), _ptr(proj)))
        return out[:n.value].copy(), proj[:cands.size].copy()

    # ---- cv::calcOpticalFlowPyrLK as algorithm::computeOpticalFlowSparse calls it ----
    def klt_track(self, ref_slot, cur_slot, prev_pts, next_pts=None, win=11, max_level=3, max_count=30, epsilon=1e-4,
                  min_eig_threshold=1e-4):
        """returns (next_pts (n, 2) float32, status (n,) uint8, err (n,) float32); next_pts given = OPTFLOW_USE_INITIAL_FLOW"""
        prev = np.ascontiguousarray(prev_pts, dtype=np.float32).reshape(-1, 2)
        nxt = np.ascontiguousarray(next_pts if next_pts is not None else prev, dtype=np.float32).reshape(-1, 2).copy()
        n = prev.shape[0]
        assert nxt.shape[0] == n
        prm = KltParams(win, max_level, max_count, 1 if next_pts is not None else 0, epsilon, min_eig_threshold)
        status, err = np.zeros(max(1, n), np.uint8), np.zeros(max(1, n), np.float32)
        self._check(self.L.svo_klt_track(self.h, ref_slot, cur_slot, _ptr(prev), _ptr(nxt), n, C.byref(prm), _ptr(status), _ptr(err)))
        return nxt, status[:n].copy(), err[:n].copy()

    # ---- algorithm::matchEpipolarConstraint, batched over depth-filter seeds ----
    def epipolar_match(self, items, patch_size=7, mean_mode=MEAN_EIGEN_U8):
        items = np.ascontiguousarray(items, dtype=EPI_ITEM_DTYPE).reshape(-1)
        prm = EpiParams(patch_size, mean_mode)
        res = np.zeros(items.size, EPI_RESULT_DTYPE)
        self._check(self.L.svo_epipolar_match(self.h, _ptr(items), items.size, C.byref(prm), _ptr(res)))
        return res

    # ---- the whole per-frame front end, one CUDA graph launch ----
    def frontend_run(self, img, job, feats, ref_slot, kf_slot, cur_slot, cell=30, thr=50, max_features=512, patch_size=5,
                     min_level=0, max_level=3, mode=LM_FAITHFUL, max_iter=20, fa_patch=7, fa_mode=LM_FAITHFUL, fa_max_iter=20,
                     occupancy=None):
        """returns (align result record, selected features, refined per-feature records)"""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        assert img.shape == (self.height, self.width)
        job = np.ascontiguousarray(job, dtype=ALIGN_JOB_DTYPE).reshape(-1)
        feats = np.ascontiguousarray(feats, dtype=ALIGN_FEATURE_DTYPE).reshape(-1)
        prm = FrontendParams(ref_slot, kf_slot, cur_slot, cell, thr, max_features,
                             AlignParams(patch_size, min_level, max_level, mode, max_iter, 0),
                             FaParams(fa_patch, fa_mode, fa_max_iter, 0))
        rows, cols = self.height // cell + 1, self.width // cell + 1
        occ = None
        if occupancy is not None:
            occ = np.ascontiguousarray(np.asarray(occupancy).reshape(-1), dtype=np.uint8)
            assert occ.size == rows * cols
        res = np.zeros(1, FRONTEND_RESULT_DTYPE)
        sel = np.zeros(rows * cols, FEATURE_PX_DTYPE)
        ref = np.zeros(max(1, feats.size), FA_RESULT_DTYPE)
        self._check(self.L.svo_frontend_run(self.h, C.byref(prm), img.ctypes.data, img.strides[0], _ptr(job), _ptr(feats),
                                            feats.size, _ptr(occ), _ptr(res), _ptr(sel), sel.size, _ptr(ref)))
        return res[0], sel[:int(res[0]["n_selected"])].copy(), ref[:feats.size]

    # ---- FeatureAlignment::align ----
    def feature_align(self, items, patch_size=7, mode=LM_FAITHFUL, max_iter=20):
        items = np.ascontiguousarray(items, dtype=FA_ITEM_DTYPE).reshape(-1)
        prm = FaParams(patch_size, mode, max_iter, 0)
        res = np.zeros(items.size, FA_RESULT_DTYPE)
        self._check(self.L.svo_feature_align(self.h, _ptr(items), items.size, C.byref(prm), _ptr(res)))
        return res

    def feature_align_stage(self, items, patch_size=7, mode=LM_FAITHFUL, max_iter=20):
        items = np.ascontiguousarray(items, dtype=FA_ITEM_DTYPE).reshape(-1)
        prm = FaParams(patch_size, mode, max_iter, 0)
        self._check(self.L.svo_feature_align_stage(self.h, _ptr(items), items.size, C.byref(prm)))
        self._staged_fa = items.size

    def feature_align_h2d(self):
        self._check(self.L.svo_feature_align_h2d(self.h))

    def feature_align_launch(self):
        self._check(self.L.svo_feature_align_launch(self.h))

    def feature_align_d2h(self):
        self._check(self.L.svo_feature_align_d2h(self.h))

    def feature_align_fetch(self):
        res = np.zeros(self._staged_fa, FA_RESULT_DTYPE)
        self._check(self.L.svo_feature_align_fetch(self.h, _ptr(res)))
        return res


def make_jobs(n):
    return np.zeros(n, ALIGN_JOB_DTYPE)
