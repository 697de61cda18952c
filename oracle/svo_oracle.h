/*
 * svo_oracle.h -- C ABI of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a dependency-free, double-precision C++17 restatement of the
 * photometric-alignment hot path of amin-abouee/semi-direct-visual-odometry
 * (image pyramid -> grid feature selection -> sparse SE3 image alignment ->
 * per-feature 2D alignment).  Each function cites the reference file:line it
 * follows.  Only tests/, __graft_entry__.smoke(), bench.py's CPU-baseline /
 * `--impl reference` legs and bench_configs.py's CPU baselines may load it.  The product
 * (libsvo_b200.so) never does.
 *
 * PARITY PINNING: the reference's own tests hold no numeric golden vector for
 * this path and the reference cannot be compiled here
 * (Eigen, Sophus, OpenCV C++, g2o and libSIMD are absent).  What IS pinned:
 *   - orc_pyrdown       bit-exact against cv2.pyrDown 4.13 (tests/test_oracle_pyramid.py)
 *   - orc_project2d     the reference's own KAT, tests/test_camera.cpp:83-96
 *   - orc_image_jac     against central differences of the projection (python/symbol.py:50-60)
 *   - orc_se3_exp       against the closed form / scipy Rotation
 *   - orc_sparse_align  (reference mode, four levels) against a second, independent numpy restatement to 1e-9 per
 *                       level (tests/test_oracle_numerics.py::test_align_against_independent_numpy)
 *   - orc_sparse_align  (GN mode: Optimizer::optimizeGN, the headline benchmark's) against an independent numpy restatement
 *                       of the loop and its exits: same evaluation counts, pose to 1e-10 per level
 *                       (tests/test_oracle_numerics.py::test_align_gn_against_independent_numpy)
 *   - orc_feature_align (reference mode) against a second, independent numpy restatement: RMSE to 1e-9, moved pixel to
 *                       1e-7 (tests/test_oracle_numerics.py::test_feature_align_against_independent_numpy)
 *   - orc_select_ssc, orc_epipolar_match, orc_reproject_map  against direct python / numpy restatements and the golden
 *                       vectors they produced (tests/golden/next_rows_golden.npz)
 *   - orc_klt_track     against the RUNNING cv2 4.13 binary (cv2.calcOpticalFlowPyrLK with the reference's arguments) and
 *                       golden vectors that binary wrote (tests/test_oracle_klt.py, tests/golden/klt_golden.npz)
 * The alignment numerics (H, g, pose) are therefore "parity unpinned" against a
 * running reference binary; they are pinned to this line-by-line restatement and to the
 * independent restatements above.
 */
#ifndef SVO_ORACLE_H
#define SVO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* optimisation modes */
enum { ORC_LM_FAITHFUL = 0, ORC_LM_ITERATED = 1, ORC_GN = 2 };
/* median rule */
enum { ORC_MEDIAN_EXACT = 0, ORC_MEDIAN_LIBSTDCXX = 1 };
/* Optimizer::Status, include/optimizer.hpp:21-33 */
enum {
    ORC_ST_SUCCESS = 0, ORC_ST_MAX_COFF_DX = 1, ORC_ST_NAN_IN_DX = 2, ORC_ST_SMALL_STEP = 3,
    ORC_ST_LAMBDA = 4, ORC_ST_NORM_INF_DIFF = 5, ORC_ST_NON_SUFF_POINTS = 6,
    ORC_ST_INCREASE_CHI2 = 7, ORC_ST_SMALL_CHI2 = 8, ORC_ST_FAILED = 9
};

/* One feature of the reference frame or of its last keyframe (include/feature.hpp:31-38,
 * include/point.hpp:28).  Layout is identical to svo_align_feature in include/svo_b200.h. */
typedef struct {
    double px[2];      /* Feature::m_pixelPosition, level 0 */
    double bearing[3]; /* Feature::m_bearingVec (unit norm) */
    double point[3];   /* Feature::m_point->m_position (world) */
    int32_t has_point; /* Feature::m_point != nullptr */
    int32_t reserved;
} orc_feature;

typedef struct {
    int32_t patch_size;  /* 5 (config/config.json:30) */
    int32_t min_level;   /* 0 */
    int32_t max_level;   /* 3 */
    int32_t mode;        /* ORC_LM_FAITHFUL | ORC_LM_ITERATED | ORC_GN */
    int32_t max_iter;    /* Optimizer::m_maxIteration, 20 in the reference */
    int32_t median_mode; /* ORC_MEDIAN_EXACT | ORC_MEDIAN_LIBSTDCXX */
} orc_align_params;

/* Per-level record.  H, g, chi2, sigma, n_px, lambda, dx are those of the FIRST
 * iteration of the level (the only one in LM_FAITHFUL); pose_after, status,
 * iterations, evaluations describe the end of the level. */
typedef struct {
    double H[36];         /* undamped J^T W J, row-major */
    double g[6];          /* J^T W r */
    double dx[6];         /* first solved step */
    double chi2;          /* sum w r^2 before the first step */
    double sigma;         /* 1.4826 * MAD before the first step */
    double lambda;        /* damping used in the first step (0 for GN) */
    double pose_after[7]; /* qx qy qz qw tx ty tz after the level */
    double rmse;          /* value optimizeLM/GN returns for the level */
    int32_t n_px;         /* residual rows written in the first evaluation */
    int32_t status;       /* Optimizer::Status at the end of the level */
    int32_t iterations;   /* solves performed */
    int32_t evaluations;  /* residual evaluations performed */
} orc_level_stats;

/* ---- image pyramid (src/image_pyramid.cpp:36-52) ---- */
/* Simd::AbsGradientSaturatedSum semantics, 3rd_party/simd/include/Simd/SimdLib.h:856-884 */
void orc_abs_gradient(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
/* cv::pyrDown (5x5 [1 4 6 4 1]^2/256, round half up, BORDER_REFLECT_101), dst = (w+1)/2 x (h+1)/2 */
void orc_pyrdown(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
/* packed pyramid: level l is (w_l x h_l) continuous, levels concatenated; returns bytes needed */
int64_t orc_pyramid_bytes(int w, int h, int levels);
/* builds image AND gradient stacks exactly as createImagePyramid does */
void orc_build_pyramid(const uint8_t* img, int w, int h, int pitch, int levels, uint8_t* img_pyr, uint8_t* grad_pyr);

/* ---- grid feature selection (src/feature_selection.cpp:19-25,91-146) ---- */
/* occupancy: rows*cols bytes (nullable); out_xym: 3 ints per feature (x, y, magnitude) in cell raster order */
int orc_grid_select(const uint8_t* grad, int w, int h, int pitch, int cell, uint32_t thr, const uint8_t* occupancy,
                    int32_t* out_xym, int max_out);

/* FeatureSelection::gradientMagnitudeWithSSC (src/feature_selection.cpp:27-89) with FeatureSelection::SSC (:165-248):
 * every pixel with gradient > thr, sorted by response (std::sort there: the order of EQUAL responses is unspecified in
 * the reference; here: stable, i.e. raster order -- the documented choice), suppression via square covering with a
 * binary search over the square width, then (use_bucketing) first-come bucketing on the occupancy grid.
 * out: (x, y, magnitude) triples in emission order.  info (nullable): [0] keypoints above thr, [1] final width,
 * [2] SSC iterations, [3] points returned by SSC (before bucketing).  Returns the number of features. */
int orc_select_ssc(const uint8_t* grad, int w, int h, int pitch, uint32_t thr, int num_candidates, int cell,
                   const uint8_t* occupancy, int use_bucketing, int32_t* out, int max_out, int32_t* info);

/* ---- numerics (src/algorithm.cpp:834-905, src/pinhole_camera.cpp:50-57) ---- */
double orc_bilinear_double(const uint8_t* img, int pitch, double x, double y);
float orc_bilinear_float(const uint8_t* img, int pitch, double x, double y);
double orc_median(const double* v, int n, int num_valid, int median_mode);
double orc_sigma(const double* v, int n, int num_valid, int median_mode);
void orc_project2d(const double K[4], const double p[3], double uv[2]);
void orc_image_jac(const double p[3], double fx, double fy, double J[12]);
void orc_se3_exp(const double xi[6], double qt[7]);
void orc_se3_mul(const double a[7], const double b[7], double out[7]);
void orc_se3_act(const double T[7], const double p[3], double out[3]);
void orc_se3_inv(const double T[7], double out[7]);
int orc_ldlt_solve(const double* A, const double* b, int n, double* x);

/* ---- sparse image alignment (src/image_alignment.cpp:25-67) ---- */
/* *_pyr: packed image pyramids built by orc_build_pyramid (levels 0..max_level at least).
 * feats: n_ref features of the reference frame followed by n_kf features of its last keyframe.
 * T_*: Sophus params order qx qy qz qw tx ty tz, world -> camera.  K: fx fy cx cy.
 * stats: (max_level - min_level + 1) records, coarse to fine, nullable.
 * Returns the value ImageAlignment::align returns (RMSE of the last level). */
double orc_sparse_align(const uint8_t* ref_pyr, const uint8_t* kf_pyr, const uint8_t* cur_pyr, int w, int h,
                        const orc_feature* feats, int n_ref, int n_kf, const double T_ref[7], const double T_kf[7],
                        const double K[4], const orc_align_params* params, double T_cur[7], orc_level_stats* stats,
                        int32_t* status_out);

typedef struct {
    const uint8_t* ref_pyr;
    const uint8_t* kf_pyr;
    const uint8_t* cur_pyr;
    const orc_feature* feats;
    int32_t n_ref, n_kf;
    double T_ref[7], T_kf[7], T_cur[7]; /* T_cur in/out */
    double rmse;                        /* out */
    int32_t status;                     /* out */
    int32_t evaluations;                /* out: residual evaluations over all levels */
} orc_align_job;
/* one job per task over n_threads std::threads (CPU baseline for the batched configs) */
void orc_sparse_align_batch(orc_align_job* jobs, int n_jobs, int w, int h, const double K[4],
                            const orc_align_params* params, int n_threads);

/* ---- per-feature 2D alignment (src/feature_alignment.cpp:25-62) ---- */
typedef struct {
    int32_t patch_size;  /* 7 (src/map.cpp:18) */
    int32_t mode;        /* ORC_LM_FAITHFUL | ORC_LM_ITERATED | ORC_GN */
    int32_t max_iter;    /* 20 */
    int32_t median_mode;
} orc_fa_params;
/* ref_grad / cur_grad: gradient level 0 (w x h, pitch w).  A: optional 2x2 row-major affine warp of the
 * template, NULL = identity = the reference.  px_inout: start / result pixel in cur.
 * Returns the RMSE FeatureAlignment::align returns. */
double orc_feature_align(const uint8_t* ref_grad, const uint8_t* cur_grad, int w, int h, const double ref_px[2],
                         const double* A, double px_inout[2], const orc_fa_params* params, int32_t* status_out,
                         int32_t* iterations_out);

/* ---- epipolar search of the depth filter: algorithm::matchEpipolarConstraint (src/algorithm.cpp:412-551), with
 * getAffineWarp (:335-367), applyAffineWarp (:369-394), computeScore (:396-410) and depthFromTriangulation
 * (:682-703).  Called per depth-filter seed from DepthEstimator::updateFilters (src/depth_estimator.cpp:245). ---- */
enum { ORC_MEAN_EIGEN_U8 = 0, ORC_MEAN_EXACT = 1 };
typedef struct {
    int32_t patch_size; /* 7 (src/depth_estimator.cpp:245) */
    int32_t mean_mode;  /* ORC_MEAN_EIGEN_U8: what computeScore executes -- Eigen's mean() on a Matrix<uint8_t> sums and
                           divides in uint8 arithmetic ((sum mod 256) / area, an integer); ORC_MEAN_EXACT: the real mean */
} orc_epi_params;
typedef struct {
    double depth;    /* estimatedDepth (valid when found) */
    double px[2];    /* bestLocation on the epipolar line (the segment midpoint when the segment is shorter than 2 px) */
    double score;    /* minimum score over the steps (DBL_MAX when no step was scored) */
    int32_t found;   /* the function's return value */
    int32_t steps;   /* pixelStep (0 for the short-segment branch) */
} orc_epi_result;
/* ref_img / cur_img: image level 0 (w x h, pitch w).  T_rel = computeRelativePose(ref, cur) = T_cur T_ref^-1.
 * ref_px, ref_bearing: refFeature->m_pixelPosition / m_bearingVec. */
void orc_epipolar_match(const uint8_t* ref_img, const uint8_t* cur_img, int w, int h, const double K[4], const double T_rel[7],
                        const double ref_px[2], const double ref_bearing[3], double depth, double min_depth, double max_depth,
                        const orc_epi_params* params, orc_epi_result* out);

/* ---- Map::reprojectMap (src/map.cpp:260-489): reprojectPoint (:492-504) per candidate, reprojectCell (:506-579) per
 * cell in cell_order; equal Point types keep their insertion order (the reference's std::sort is unstable).
 * grads: gradient level 0 of the frame in each slot (w x h, pitch w), indexed by the candidates' ref_slot; cur_grad: the
 * new frame's.  matches: (cell, candidate, px x, px y, rmse, status) per match as doubles, capacity max_matches + 1.
 * Returns the number of matches. ---- */
typedef struct {
    int32_t ref_slot, type;
    double ref_px[2];
    double point[3];
} orc_reproj_candidate;
int orc_reproject_map(const uint8_t* const* grads, const uint8_t* cur_grad, int w, int h, const double K[4], const double T_cur[7],
                      const orc_reproj_candidate* cands, int n, int cell, const int32_t* cell_order, int n_cells, int max_matches,
                      const orc_fa_params* fa, double* matches, uint8_t* projected);

/* ---- algorithm::computeOpticalFlowSparse's tracker (src/algorithm.cpp:29-107, the initialisation; the same call sits at
 * src/map.cpp:322,403): cv::calcOpticalFlowPyrLK(refImg, curImg, refPoints, curPoints, status, errors, Size(win, win), 3,
 * TermCriteria(COUNT + EPS, 30, 1e-4), OPTFLOW_USE_INITIAL_FLOW).  OpenCV is a dependency of the reference that is not
 * vendored (CMake find_package(OpenCV)); this restates the published algorithm of OpenCV 4.x modules/video/src/lkpyramid.cpp
 * (buildOpticalFlowPyramid: pyrDown levels padded BORDER_REFLECT_101; calcSharrDeriv: int16 Scharr 3-10-3 with reflected
 * borders, zero outside the image; LKTrackerInvoker: 14-bit fixed-point bilinear weights, window values scaled by 32, float
 * sums, the eps / oscillation exits) and is pinned against the cv2 4.13 binary of this image (tests/test_oracle_klt.py).
 * prev_pts / next_pts: n (x, y) float pairs; next_pts holds the initial guess when use_initial_flow.  status: 1 tracked.
 * err: mean absolute window difference at level 0 (OpenCV's default error measure).  Returns the top level used. ---- */
typedef struct {
    int32_t win;              /* window side (cv::Size(win, win)) */
    int32_t max_level;        /* 3 in the reference */
    int32_t max_count;        /* 30 */
    int32_t use_initial_flow; /* 1 */
    double epsilon;           /* 1e-4 (squared internally, as OpenCV does) */
    double min_eig_threshold; /* 1e-4, OpenCV's default */
} orc_klt_params;
int orc_klt_track(const uint8_t* ref_img, const uint8_t* cur_img, int w, int h, const float* prev_pts, float* next_pts, int n,
                  const orc_klt_params* params, uint8_t* status, float* err);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* SVO_ORACLE_H */
