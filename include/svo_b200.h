/*
 * svo_b200.h -- C ABI of the B200 (sm_100a) photometric-alignment library, libsvo_b200.so.
 *
 * Drop-in boundary for the hot path of amin-abouee/semi-direct-visual-odometry.  The reference
 * has no FFI: the path sits behind plain C++ classes.  Every entry point below is what the body
 * of one of those classes binds (citations relative to the reference tree); the host-side C++
 * mirror of the classes lives in semi-direct-visual-odometry_b200/host/ and INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an svo_status (0 = OK, < 0 = error)
 *     and never throws; svo_last_error() gives the text of the last failure of a context.
 *   - poses: Sophus::SE3d::params() order  qx qy qz qw tx ty tz, world -> camera
 *     (include/frame.hpp:198).  K: fx fy cx cy (src/pinhole_camera.cpp:123-139).
 *   - images: 8-bit gray, row-major.  Pyramids live on the device in "frame slots" of a
 *     per-context arena; a slot is what a reference Frame's m_imagePyramid is.
 *   - there is NO CPU fallback: without a CUDA device svo_create fails with SVO_ERR_NO_DEVICE.
 *   - compute is enqueued on one CUDA stream (svo_config.stream, or a private one); host->device copies and the
 *     kernels of svo_frames_prefetch run on private copy / ingest streams of the context, ordered with the main
 *     stream by events.  Every entry point holds the context's lock for the duration of the call: calls from
 *     several threads (the reference's depth-filter thread next to the tracker) are serialised, never
 *     interleaved.  The phase-split forms (..._stage / _h2d / _launch / _d2h / _fetch) keep state in the context
 *     between calls: one thread drives such a sequence at a time.
 */
#ifndef SVO_B200_H
#define SVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_MAX_LEVELS 8

typedef int32_t svo_status;
enum {
    SVO_OK              = 0,
    SVO_ERR_INVALID     = -1, /* bad argument */
    SVO_ERR_CUDA        = -2, /* a CUDA call failed; see svo_last_error */
    SVO_ERR_CAPACITY    = -3, /* batch / feature count above what the context was created for */
    SVO_ERR_NO_DEVICE   = -4, /* no usable CUDA device: this library has no CPU path */
    SVO_ERR_UNSUPPORTED = -5  /* e.g. gradientMagnitudeByValue(useBucketing=false), SURVEY 9.7 */
};

/* optimisation modes: what Optimizer::optimizeLM really does (one damped step),
 * the same loop with the always-true exit clause removed, and Optimizer::optimizeGN */
enum { SVO_LM_FAITHFUL = 0, SVO_LM_ITERATED = 1, SVO_GN = 2 };

/* Optimizer::Status, include/optimizer.hpp:21-33 */
enum {
    SVO_ST_SUCCESS = 0, SVO_ST_MAX_COFF_DX = 1, SVO_ST_NAN_IN_DX = 2, SVO_ST_SMALL_STEP = 3,
    SVO_ST_LAMBDA = 4, SVO_ST_NORM_INF_DIFF = 5, SVO_ST_NON_SUFF_POINTS = 6,
    SVO_ST_INCREASE_CHI2 = 7, SVO_ST_SMALL_CHI2 = 8, SVO_ST_FAILED = 9
};

typedef struct svo_ctx svo_ctx;

typedef struct {
    int32_t device;        /* CUDA device ordinal */
    int32_t width, height; /* camera / level-0 image size (config/config.json:10-11) */
    int32_t levels;        /* pyramid levels per frame = max_level + 1 (src/system.cpp:36) */
    int32_t max_frames;    /* frame slots in the device pyramid arena */
    int32_t max_jobs;      /* sparse-alignment jobs (frame pairs) per batch */
    int32_t max_features;  /* features per job, n_ref + n_kf */
    int32_t max_fa_items;  /* FeatureAlignment items per batch */
    int32_t reserved;
    void* stream;          /* cudaStream_t to enqueue on; NULL = the context creates one */
    double K[4];           /* fx fy cx cy */
} svo_config;

svo_status svo_create(const svo_config* cfg, svo_ctx** out);
void svo_destroy(svo_ctx* ctx);
const char* svo_last_error(const svo_ctx* ctx);
const char* svo_version(void);
svo_status svo_sync(svo_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t svo_launch_count(const svo_ctx* ctx);
/* the CUDA stream the context enqueues on (cudaStream_t), so a caller can bracket it with events */
void* svo_stream(const svo_ctx* ctx);
/* page-locked host memory: images handed to svo_frames_upload from such a buffer are DMA'd without staging */
svo_status svo_host_alloc(svo_ctx* ctx, int64_t bytes, void** out);
svo_status svo_host_free(svo_ctx* ctx, void* p);
/* level geometry: w_l = (w_{l-1}+1)/2 as cv::pyrDown; pitch is the device row pitch */
svo_status svo_level_dims(const svo_ctx* ctx, int level, int* w, int* h, int* pitch);

/* ---------------------------------------------------------------------------------------------
 * ImagePyramid::createImagePyramid (src/image_pyramid.cpp:36-52), called from Frame::Frame
 * (src/frame.cpp:26).  Builds the image stack AND the gradient stack
 * (Simd::AbsGradientSaturatedSum at level 0, cv::pyrDown below) for n consecutive slots.
 * imgs: host pointer, n images of `pitch` bytes per row, `frame_stride` bytes apart.
 * Asynchronous on the context stream (the host buffer is staged through pinned memory first,
 * so it may be reused as soon as the call returns).
 * ------------------------------------------------------------------------------------------- */
svo_status svo_frames_upload(svo_ctx* ctx, int first_slot, int n, const uint8_t* imgs, int pitch,
                             int64_t frame_stride);
/* Same work, enqueued on the context's ingest streams so that it runs CONCURRENTLY with what is already enqueued on
 * the main stream (the alignment of the previous batch of frames): what a pipelined caller of Frame::Frame
 * (src/frame.cpp:26) does with the next camera frame while the tracker still works on the current one.  Every later
 * call that touches frame slots is ordered after it.  The caller guarantees that no job enqueued but not yet fetched
 * references these slots (double-buffer the slots of the incoming frames). */
svo_status svo_frames_prefetch(svo_ctx* ctx, int first_slot, int n, const uint8_t* imgs, int pitch,
                               int64_t frame_stride);
/* same, level-0 images already on the device (dptr: device pointer) */
svo_status svo_frames_upload_device(svo_ctx* ctx, int first_slot, int n, const void* dptr, int pitch,
                                    int64_t frame_stride);
/* rebuild the stacks of slots whose level-0 image is already in the arena (asynchronous) */
svo_status svo_frames_rebuild(svo_ctx* ctx, int first_slot, int n);
/* ImagePyramid::getImageAtLevel / getGradientAtLevel (src/image_pyramid.cpp:64-100): copies one
 * level to a continuous host buffer (which: 0 image, 1 gradient).  Synchronous. */
svo_status svo_frame_download(svo_ctx* ctx, int slot, int level, int which, uint8_t* dst, int dst_pitch);

/* ---------------------------------------------------------------------------------------------
 * FeatureSelection::gradientMagnitudeByValue(frame, thr, useBucketing = true)
 * (src/feature_selection.cpp:91-146).  occupancy: m_occupancyGrid as (h/cell+1)*(w/cell+1) bytes,
 * row-major, nullable.  out: features in cell raster order.  Synchronous.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t x, y;      /* Feature::m_pixelPosition */
    int32_t magnitude; /* Feature::m_gradientMagnitude */
} svo_feature_px;
svo_status svo_select_grid(svo_ctx* ctx, int slot, int cell, uint32_t thr, const uint8_t* occupancy,
                           svo_feature_px* out, int max_out, int* n_out);

/* FeatureSelection::gradientMagnitudeWithSSC(frame, thr, numberCandidate, useBucketing) (src/feature_selection.cpp:27-89,
 * SSC :165-248) -- what System calls on every keyframe (src/system.cpp:81,253,429); SURVEY 8(f) row f2.  Keypoints of
 * equal response are ordered by raster position (the reference's std::sort leaves their order unspecified).  out: features
 * in the order the reference appends them.  info (nullable, 4 ints): keypoints above thr, last SSC width, SSC iterations,
 * points kept before bucketing.  num_candidates <= ~3,700 (at most 4,096 points may survive the suppression).
 * Synchronous. */
svo_status svo_select_ssc(svo_ctx* ctx, int slot, uint32_t thr, int num_candidates, int cell, const uint8_t* occupancy,
                          int use_bucketing, svo_feature_px* out, int max_out, int* n_out, int32_t* info);

/* ---------------------------------------------------------------------------------------------
 * ImageAlignment::align(refFrame, curFrame) (src/image_alignment.cpp:25-67), batched over
 * independent frame pairs ("jobs").
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    double px[2];      /* Feature::m_pixelPosition (level 0) */
    double bearing[3]; /* Feature::m_bearingVec (unit norm, src/pinhole_camera.cpp:100) */
    double point[3];   /* Feature::m_point->m_position (world) */
    int32_t has_point; /* Feature::m_point != nullptr */
    int32_t reserved;
} svo_align_feature;

typedef struct {
    int32_t ref_slot;    /* refFrame->m_imagePyramid */
    int32_t kf_slot;     /* refFrame->m_lastKeyframe->m_imagePyramid (must exist) */
    int32_t cur_slot;    /* curFrame->m_imagePyramid */
    int32_t n_ref;       /* refFrame->numberObservation() */
    int32_t n_kf;        /* lastKF->numberObservation() */
    int32_t feat_offset; /* first feature of this job in the feats array: n_ref of ref, then n_kf of lastKF */
    double T_ref[7];     /* refFrame->m_absPose */
    double T_kf[7];      /* lastKF->m_absPose */
    double T_cur[7];     /* curFrame->m_absPose on entry (the prior, src/system.cpp:309) */
} svo_align_job;

typedef struct {
    int32_t patch_size; /* m_patchSize, 5 (config/config.json:30); even sizes: SURVEY 9.2 extension */
    int32_t min_level;  /* 0 */
    int32_t max_level;  /* 3 */
    int32_t mode;       /* SVO_LM_FAITHFUL (reference behaviour) | SVO_LM_ITERATED | SVO_GN */
    int32_t max_iter;   /* Optimizer::m_maxIteration, 20 (src/optimizer.cpp:18) */
    int32_t reserved;
} svo_align_params;

typedef struct {
    double T_cur[7]; /* curFrame->m_absPose after align */
    double rmse;     /* value align() returns: RMSE of the last level (0 when n_ref == 0) */
    int32_t status;  /* Optimizer::Status of the last level */
    int32_t evaluations; /* residual evaluations over all levels */
    int32_t iterations;  /* solves over all levels */
    int32_t reserved;    /* diagnostics of the fast path: robust-scale selections served by the hot | cold << 8 |
                            generic << 16 tier (0 from the generic kernel) */
} svo_align_result;

/* per-level diagnostics; same meaning and layout as orc_level_stats of the oracle */
typedef struct {
    double H[36];         /* undamped J^T W J of the first iteration, row-major */
    double g[6];          /* J^T W r */
    double dx[6];         /* first solved step */
    double chi2;          /* sum w r^2 before the first step */
    double sigma;         /* 1.4826 MAD before the first step */
    double lambda;        /* damping of the first step (0 for GN) */
    double pose_after[7]; /* pose at the end of the level */
    double rmse;
    int32_t n_px;
    int32_t status;
    int32_t iterations;
    int32_t evaluations;
} svo_align_level_stats;

/* One call = the whole align: stage -> H2D -> kernel -> D2H -> fetch (synchronous).
 * results: n_jobs records.  stats: n_jobs * (max_level-min_level+1) records (coarse to fine), nullable. */
svo_status svo_sparse_align(svo_ctx* ctx, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats,
                            int n_feats, const svo_align_params* params, svo_align_result* results,
                            svo_align_level_stats* stats);
/* The same, split so a caller can overlap or time the phases; all asynchronous except fetch. */
svo_status svo_sparse_align_stage(svo_ctx* ctx, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats,
                                  int n_feats, const svo_align_params* params, int want_stats);
svo_status svo_sparse_align_h2d(svo_ctx* ctx);
svo_status svo_sparse_align_launch(svo_ctx* ctx);
svo_status svo_sparse_align_d2h(svo_ctx* ctx);
svo_status svo_sparse_align_fetch(svo_ctx* ctx, svo_align_result* results, svo_align_level_stats* stats);
/* device copy of the last batch's svo_align_result records (valid once the launch has completed on the stream):
 * what a multi-GPU caller hands to its final pose gather (ncclGather) without a host round trip */
const void* svo_sparse_align_results_device(const svo_ctx* ctx);
/* diagnostics: per level (8 slots each, coarse to fine) SM-clock cycles job 0 of the last fast-path launch spent in
 * [0] warp+sample [1] median select [2] MAD select [3] weights+sums [4] reductions [5] solve [6] #evaluations */
svo_status svo_debug_cycles(svo_ctx* ctx, int64_t* out64);

/* ---------------------------------------------------------------------------------------------
 * FeatureAlignment::align(refFeature, curFrame, pixelPos) (src/feature_alignment.cpp:25-62),
 * batched: one item per (feature, frame) call the reference makes serially from
 * Map::reprojectCell / addCandidateToFrame (src/map.cpp:538,608).  Runs on gradient level 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t ref_slot;   /* refFeature->m_frame->m_imagePyramid */
    int32_t cur_slot;   /* curFrame->m_imagePyramid */
    double ref_px[2];   /* refFeature->m_pixelPosition */
    double px[2];       /* pixelPos on entry */
    double A[4];        /* optional 2x2 row-major affine warp of the template */
    int32_t use_affine; /* 0 = identity = the reference */
    int32_t reserved;
} svo_fa_item;

typedef struct {
    int32_t patch_size; /* 7 (src/map.cpp:18) */
    int32_t mode;       /* SVO_LM_FAITHFUL | SVO_LM_ITERATED | SVO_GN */
    int32_t max_iter;   /* 20 */
    int32_t reserved;
} svo_fa_params;

typedef struct {
    double px[2];  /* pixelPos on return */
    double rmse;   /* value align() returns (NaN when the start is out of frame, as the reference) */
    int32_t status;
    int32_t iterations;
} svo_fa_result;

svo_status svo_feature_align(svo_ctx* ctx, const svo_fa_item* items, int n, const svo_fa_params* params,
                             svo_fa_result* results);
svo_status svo_feature_align_stage(svo_ctx* ctx, const svo_fa_item* items, int n, const svo_fa_params* params);
svo_status svo_feature_align_h2d(svo_ctx* ctx);
svo_status svo_feature_align_launch(svo_ctx* ctx);
svo_status svo_feature_align_d2h(svo_ctx* ctx);
svo_status svo_feature_align_fetch(svo_ctx* ctx, svo_fa_result* results);

/* ---------------------------------------------------------------------------------------------
 * Map::reprojectMap(refFrame, curFrame, overlapKeyFrames) (src/map.cpp:260-489), SURVEY 8(f) row f1: reprojectPoint
 * (:492-504) for every candidate, the per-cell choice of reprojectCell (:506-579) and ONE FeatureAlignment launch.
 * The caller lists the candidates in the reference's insertion order (features with a point of refFrame, then of its
 * last keyframe, :462-476) and keeps the Point bookkeeping (m_lastProjectedKFId, m_succeededProjection, types).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t ref_slot;  /* candidate.m_feature->m_frame->m_imagePyramid */
    int32_t type;      /* Point::PointType: 0 GOOD, 1 DELETED, 2 CANDIDATE, 3 UNKNOWN (include/point.hpp:18-24) */
    double ref_px[2];  /* candidate.m_feature->m_pixelPosition */
    double point[3];   /* candidate.m_point->m_position */
} svo_reproj_candidate;
typedef struct {
    int32_t cell;      /* grid cell (row-major over ceil(h/cell) x ceil(w/cell)) */
    int32_t candidate; /* index into the candidate array: the cell's first candidate after the sort by type, descending */
    double px[2];      /* pixel position of the new Feature (:564): the projected pixel refined by FeatureAlignment::align */
    double rmse;       /* align's return value (the reference ignores it: the match test is commented out, :528-549) */
    int32_t status;
    int32_t reserved;
} svo_reproj_match;
/* cell_order: Map::m_grid.m_cellOrders (a permutation of the cells, shuffled once at start-up, :237-247).  matches:
 * capacity max_matches + 1 (the walk stops once m_matches exceeds max_matches = 150, :484-487), in cell_order order.
 * projected (nullable, n bytes): reprojectPoint's return value per candidate (for overlapKeyFrames' counters).
 * Synchronous. */
svo_status svo_reproject_map(svo_ctx* ctx, int cur_slot, const double T_cur[7], const svo_reproj_candidate* cands, int n,
                             int cell_size, const int32_t* cell_order, int n_cells, int max_matches, const svo_fa_params* fa,
                             svo_reproj_match* matches, int* n_matches, uint8_t* projected);

/* ---------------------------------------------------------------------------------------------
 * algorithm::matchEpipolarConstraint(refFrame, curFrame, refFeature, patchSize, initialDepth, minDepth, maxDepth,
 * estimatedDepth) (src/algorithm.cpp:412-551), batched: one item per depth-filter seed and frame, as
 * DepthEstimator::updateFilters calls it serially (src/depth_estimator.cpp:245).  Runs on IMAGE level 0.
 * the first component next to the alignment path, same parity bar.
 * ------------------------------------------------------------------------------------------- */
enum { SVO_MEAN_EIGEN_U8 = 0, /* what computeScore executes: Eigen's mean() of a Matrix<uint8_t> sums and divides in
                                 uint8 arithmetic, (sum mod 256) / area (src/algorithm.cpp:400-401) */
       SVO_MEAN_EXACT = 1 };  /* the arithmetic mean the code evidently meant */
typedef struct {
    int32_t ref_slot;   /* depthFilter.m_feature->m_frame->m_imagePyramid */
    int32_t cur_slot;   /* frame->m_imagePyramid */
    double T_rel[7];    /* algorithm::computeRelativePose(ref, cur) = T_cur * T_ref^-1 (src/algorithm.cpp:705-709) */
    double px[2];       /* refFeature->m_pixelPosition */
    double bearing[3];  /* refFeature->m_bearingVec */
    double depth;       /* initialDepth = 1 / mu */
    double min_depth;   /* 1 / (mu + var) */
    double max_depth;   /* 1 / max(mu - var, 1e-7) */
} svo_epi_item;
typedef struct {
    int32_t patch_size; /* 7; odd */
    int32_t mean_mode;  /* SVO_MEAN_EIGEN_U8 (reference behaviour) | SVO_MEAN_EXACT */
} svo_epi_params;
typedef struct {
    double depth;   /* estimatedDepth, valid when found */
    double px[2];   /* best location on the epipolar segment (its midpoint when the segment is shorter than 2 px) */
    double score;   /* minimum score over the steps (DBL_MAX when none was scored) */
    int32_t found;  /* the function's return value */
    int32_t steps;  /* 1-pixel steps walked (0: short-segment branch) */
} svo_epi_result;
svo_status svo_epipolar_match(svo_ctx* ctx, const svo_epi_item* items, int n, const svo_epi_params* params,
                              svo_epi_result* results);

/* ---------------------------------------------------------------------------------------------
 * Next row f4: the tracker of algorithm::computeOpticalFlowSparse (src/algorithm.cpp:29-107, System's initialisation,
 * src/system.cpp:129; the same call at src/map.cpp:322,403), i.e.
 *   cv::calcOpticalFlowPyrLK(refImg, curImg, refPoints, curPoints, status, errors, cv::Size(win, win), 3,
 *                            cv::TermCriteria(COUNT + EPS, 30, 1e-4), cv::OPTFLOW_USE_INITIAL_FLOW)
 * on the image pyramids of two frame slots (OpenCV rebuilds the same cv::pyrDown pyramids on every call).
 * prev_pts / next_pts: n (x, y) float pairs (cv::Point2f); next_pts is the initial guess on entry when use_initial_flow
 * and the tracked position on return (also for points whose status is 0, as OpenCV leaves them).  status: 1 = tracked.
 * err (may be NULL): mean absolute window difference at level 0.  The top level is min(max_level, the last level
 * larger than the window) and must exist in the context (levels > top).  Capacity: n <= max_fa_items.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t win;              /* window side, 3..21 (m_patchSizeOpticalFlow; 11 in src/map.cpp) */
    int32_t max_level;        /* 3 */
    int32_t max_count;        /* 30 */
    int32_t use_initial_flow; /* 1 */
    double epsilon;           /* 1e-4 */
    double min_eig_threshold; /* 1e-4, OpenCV's default */
} svo_klt_params;
svo_status svo_klt_track(svo_ctx* ctx, int ref_slot, int cur_slot, const float* prev_pts, float* next_pts, int n,
                         const svo_klt_params* params, uint8_t* status, float* err);

/* ---------------------------------------------------------------------------------------------
 * The per-frame front end as ONE CUDA graph launch: what System::processNewFrame runs for a new camera image between
 * Frame::Frame (src/system.cpp:36, src/frame.cpp:26) and the candidate matching of Map::reprojectMap /
 * addCandidateToFrame (src/system.cpp:313-330, src/map.cpp:595-627):
 *   pyramids of the new image -> gradientMagnitudeByValue on it -> ImageAlignment::align(ref, new) ->
 *   Frame::world2image of every tracked point with the ALIGNED pose + isInFrame(px, 3) (src/map.cpp:601-602) ->
 *   FeatureAlignment::align of those candidates against their owner frame's gradient image.
 * The graph is captured once per (slots, parameters) configuration and cached in the context.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t ref_slot, kf_slot, cur_slot; /* frame slots; the new image is written to cur_slot */
    int32_t cell;                        /* FeatureSelection cell size (config/config.json: 30) */
    uint32_t thr;                        /* gradient threshold (50) */
    int32_t max_features;                /* capacity of the captured graph: n_ref + n_kf <= max_features */
    svo_align_params align;              /* patch 4 or 5 */
    svo_fa_params fa;
} svo_frontend_params;

typedef struct {
    svo_align_result align; /* curFrame->m_absPose after ImageAlignment::align, its return value and status */
    int32_t n_selected;     /* features gradientMagnitudeByValue found on the new frame */
    int32_t n_candidates;   /* tracked features that reprojected into the new frame and were aligned */
} svo_frontend_result;

/* job: n_ref, n_kf and the three poses are used (slots come from prm, feat_offset is 0).  feats: n_feats records, n_ref
 * of the reference frame then n_kf of the last keyframe.  occupancy: as svo_select_grid, nullable.  selected: the new
 * features, cell raster order.  refined: n_feats records, one per feature: px = the refined pixel position in the new
 * frame; features without a point or not reprojected into the frame have status SVO_ST_FAILED, iterations 0 and a NaN
 * rmse.  Synchronous.  img may be svo_frontend_image_buffer(ctx) (page-locked, w*h bytes dense): no staging copy. */
svo_status svo_frontend_run(svo_ctx* ctx, const svo_frontend_params* prm, const uint8_t* img, int pitch,
                            const svo_align_job* job, const svo_align_feature* feats, int n_feats,
                            const uint8_t* occupancy, svo_frontend_result* result, svo_feature_px* selected,
                            int max_selected, svo_fa_result* refined);
uint8_t* svo_frontend_image_buffer(svo_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * Several GPUs from one process (BASELINE config 5: 8,192 independent frame pairs over 2 / 4 / 8
 * B200s).  The reference's System runs one sequence on one thread (src/system.cpp:36-60); a host
 * that serves many sequences hands this layer a batch of independent ImageAlignment::align calls
 * (src/image_alignment.cpp:25-67).  One context and one host worker thread per device; item j of a
 * batch of n belongs to the device whose contiguous block svo_multi_shard() holds j; there is no
 * exchange between devices, the final gather is every device writing its block of the caller's
 * result array.  cfg: as svo_create, capacities PER DEVICE, cfg->device and cfg->stream ignored.
 * devices: CUDA ordinals (NULL: 0 .. n_devices-1).
 * ------------------------------------------------------------------------------------------- */
typedef struct svo_multi svo_multi;
svo_status svo_multi_create(const svo_config* cfg, const int32_t* devices, int n_devices, svo_multi** out);
void svo_multi_destroy(svo_multi* m);
int svo_multi_devices(const svo_multi* m);
svo_ctx* svo_multi_ctx(svo_multi* m, int i); /* the i-th device's context, for every per-context call above */
const char* svo_multi_last_error(const svo_multi* m);
/* block [lo, hi) of part `part` when n_items are cut into n_parts contiguous blocks (sizes differ by at most one) */
void svo_multi_shard(int n_items, int n_parts, int part, int* lo, int* hi);
/* svo_frames_upload (prefetch = 0) / svo_frames_prefetch (1) of n frames: frame j goes to the device of its block, into
 * that device's slot first_slot + (j - lo) */
svo_status svo_multi_frames_upload(svo_multi* m, int first_slot, int n, const uint8_t* imgs, int pitch,
                                   int64_t frame_stride, int prefetch);
/* svo_sparse_align over the devices.  jobs[j] runs on the device of its block; its slots are slots of THAT device;
 * feat_offset indexes the global feats array.  results / stats: n_jobs (x levels) records, written block by block. */
svo_status svo_multi_sparse_align(svo_multi* m, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats,
                                  int n_feats, const svo_align_params* params, svo_align_result* results,
                                  svo_align_level_stats* stats);
/* the same in phases: stage (+ H2D), launch (asynchronous on every device), fetch (D2H + copy out) */
svo_status svo_multi_sparse_align_stage(svo_multi* m, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats,
                                        int n_feats, const svo_align_params* params, int want_stats);
svo_status svo_multi_sparse_align_launch(svo_multi* m);
svo_status svo_multi_sparse_align_fetch(svo_multi* m, svo_align_result* results, svo_align_level_stats* stats);
svo_status svo_multi_sync(svo_multi* m);
/* `steps` launches of the staged batch on every device after `warmup` untimed ones, timed per device with CUDA events on
 * its stream; *ms_max = the slowest device's time for all steps */
svo_status svo_multi_time_launches(svo_multi* m, int warmup, int steps, double* ms_max);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H */
