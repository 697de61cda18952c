// epipolar.cu -- the depth filter's epipolar search, algorithm::matchEpipolarConstraint (src/algorithm.cpp:412-551),
// batched over seeds: one warp per seed (DepthEstimator::updateFilters calls it once per depth filter and frame,
// src/depth_estimator.cpp:245).  Per seed: project the seed's ray at min / max depth into the current frame and clamp
// (:430-451), affine warp at the initial depth (getAffineWarp, :335-367), reference patch (applyAffineWarp with the
// identity, :464-467), then walk the epipolar segment in 1-pixel steps, warp the current patch (float bilinear taps
// truncated to uint8, :369-394), score it (computeScore, :396-410) and triangulate the best location
// (depthFromTriangulation, :682-703).  Lanes own patch pixels (lane, lane + 32); the per-seed geometry is computed
// redundantly by every lane in FP64.  Quirks mirrored: an out-of-frame step scores the PREVIOUS patch again; the
// patch means are Eigen's uint8 mean ((sum mod 256) / area) unless mean_mode asks for the exact one.
#include <float.h>

#include "ctx.h"
#include "math.cuh"

namespace {

struct EpiArgs {
    ArenaView view;
    const svo_epi_item* items;
    svo_epi_result* results;
    int n;
    svo_epi_params prm;
    double K[4];
};

// algorithm::bilinearInterpolation (float), src/algorithm.cpp:885-894
__device__ __forceinline__ float epi_bilinear_float(const uint8_t* __restrict__ img, int pitch, double x, double y)
{
    const int x1 = (int)x, y1 = (int)y;
    const int x2 = x1 + 1, y2 = y1 + 1;
    const uint8_t* p = img + (long long)y1 * pitch + x1;
    const float a = (float)((x2 - x) * (double)__ldg(p) + (x - x1) * (double)__ldg(p + 1));
    const float b = (float)((x2 - x) * (double)__ldg(p + pitch) + (x - x1) * (double)__ldg(p + pitch + 1));
    return (float)((y2 - y) * (double)a + (y - y1) * (double)b);
}

struct EpiCam {
    double fx, fy, cx, cy;
    int w, h;
    __device__ __forceinline__ void bearing(double x, double y, double b[3]) const
    {
        b[0] = (x - cx) / fx;
        b[1] = (y - cy) / fy;
        b[2] = 1.0;
        const double n = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        b[0] /= n;
        b[1] /= n;
        b[2] /= n;
    }
    __device__ __forceinline__ bool inFrame(double x, double y, double bd) const { return x >= bd && y >= bd && x < w - bd && y < h - bd; }
};

__device__ __forceinline__ void epi_project_at_depth(const EpiCam& cam, const svo::Pose& T, double x, double y, double depth, double uv[2])
{
    double b[3], pr[3], pc[3];
    cam.bearing(x, y, b);
    pr[0] = b[0] * depth, pr[1] = b[1] * depth, pr[2] = b[2] * depth;
    svo::quat_rotate(T.q, pr, pc);
    pc[0] += T.t[0], pc[1] += T.t[1], pc[2] += T.t[2];
    uv[0] = cam.fx * (pc[0] / pc[2]) + cam.cx;
    uv[1] = cam.fy * (pc[1] / pc[2]) + cam.cy;
}

__device__ __forceinline__ bool epi_triangulate(const svo::Pose& T, const double bref[3], const double bcur[3], double* depth)
{
    double a0[3];
    svo::quat_rotate(T.q, bref, a0);
    const double a1[3] = {-bcur[0], -bcur[1], -bcur[2]};
    const double m00 = a0[0] * a0[0] + a0[1] * a0[1] + a0[2] * a0[2];
    const double m01 = a0[0] * a1[0] + a0[1] * a1[1] + a0[2] * a1[2];
    const double m11 = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
    const double det = m00 * m11 - m01 * m01;
    if (det < 0.000001) return false;
    const double r0 = a0[0] * T.t[0] + a0[1] * T.t[1] + a0[2] * T.t[2];
    const double r1 = a1[0] * T.t[0] + a1[1] * T.t[1] + a1[2] * T.t[2];
    *depth          = fabs(-((m11 * r0 - m01 * r1) / det));
    return true;
}

__global__ void __launch_bounds__(128) k_epipolar_match(const EpiArgs a)
{
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= a.n) return;
    const svo_epi_item it = a.items[item];
    const int P = a.prm.patch_size, half = P / 2, area = P * P;
    const EpiCam cam{a.K[0], a.K[1], a.K[2], a.K[3], a.view.w[0], a.view.h[0]};
    const int pitch     = a.view.pitch[0];
    const uint8_t* refI = a.view.img[0] + (long long)it.ref_slot * a.view.plane_stride[0];
    const uint8_t* curI = a.view.img[0] + (long long)it.cur_slot * a.view.plane_stride[0];
    svo::Pose T;
    for (int i = 0; i < 4; i++) T.q[i] = it.T_rel[i];
    for (int i = 0; i < 3; i++) T.t[i] = it.T_rel[4 + i];

    svo_epi_result res;
    res.depth = 0.0, res.px[0] = res.px[1] = 0.0, res.score = DBL_MAX, res.found = 0, res.steps = 0;

    double locMin[2], locMax[2];
    epi_project_at_depth(cam, T, it.px[0], it.px[1], it.min_depth, locMin);
    epi_project_at_depth(cam, T, it.px[0], it.px[1], it.max_depth, locMax);
    auto clampLoc = [&](double* l) {  // :435-451
        l[0] = l[0] >= 0 ? l[0] : 0.0;
        l[0] = l[0] < cam.w ? l[0] : cam.w - 1;
        l[1] = l[1] >= 0 ? l[1] : 0.0;
        l[1] = l[1] < cam.h ? l[1] : cam.h - 1;
    };
    clampLoc(locMin);
    clampLoc(locMax);
    const double ex = locMax[0] - locMin[0], ey = locMax[1] - locMin[1];
    double A[4];
    {
        double c[2], du[2], dv[2];
        epi_project_at_depth(cam, T, it.px[0], it.px[1], it.depth, c);
        epi_project_at_depth(cam, T, it.px[0] + half, it.px[1], it.depth, du);
        epi_project_at_depth(cam, T, it.px[0], it.px[1] + half, it.depth, dv);
        A[0] = (du[0] - c[0]) / half;
        A[2] = (du[1] - c[1]) / half;
        A[1] = (dv[0] - c[0]) / half;
        A[3] = (dv[1] - c[1]) / half;
    }
    const double norm = sqrt(ex * ex + ey * ey);

    // pixel slots of this lane, raster order (i outer, j inner) as applyAffineWarp fills `data`
    const int p0 = lane, p1 = lane + 32;
    const bool has0 = p0 < area, has1 = p1 < area;
    const int i0 = p0 / P - half, j0 = p0 % P - half;
    const int i1 = p1 / P - half, j1 = p1 % P - half;

    // reference patch, identity warp: boundary ceil(max(|half|, |half|)) + 2
    uint32_t r0 = 0, r1 = 0;
    if (cam.inFrame(it.px[0], it.px[1], (double)half + 2.0)) {
        if (has0) r0 = __float2uint_rz(epi_bilinear_float(refI, pitch, it.px[0] + j0, it.px[1] + i0));
        if (has1) r1 = __float2uint_rz(epi_bilinear_float(refI, pitch, it.px[0] + j1, it.px[1] + i1));
    }
    const bool eigenMean = a.prm.mean_mode == SVO_MEAN_EIGEN_U8;
    const uint32_t rsum  = __reduce_add_sync(0xffffffffu, r0 + r1);
    const double refMean = eigenMean ? (double)((rsum & 255u) / (uint32_t)(area & 255)) : (double)rsum / area;
    const int refMeanI   = (int)((rsum & 255u) / (uint32_t)(area & 255));

    if (norm < 2.0) {  // :469-483
        const double cx = (locMax[0] + locMin[0]) / 2.0, cy = (locMax[1] + locMin[1]) / 2.0;
        double bc[3];
        cam.bearing(cx, cy, bc);
        res.px[0] = cx, res.px[1] = cy;
        res.found = epi_triangulate(T, it.bearing, bc, &res.depth) ? 1 : 0;
        if (lane == 0) a.results[item] = res;
        return;
    }
    const uint32_t pixelStep = (uint32_t)ceil(norm);
    const double sx = ex / norm, sy = ey / norm;
    const double bx = A[0] * half + A[1] * half, by = A[2] * half + A[3] * half;
    const double maxBoundary = ceil(fmax(fabs(bx), fabs(by))) + 2.0;
    double minScore = DBL_MAX, bestX = 0.0, bestY = 0.0;
    uint32_t c0 = 0, c1 = 0;  // the current patch persists across out-of-frame steps (curPatchIntensities is never reset)
    for (uint32_t i = 0; i < pixelStep; i++) {
        const double lx = locMin[0] + i * sx, ly = locMin[1] + i * sy;
        if (cam.inFrame(lx, ly, maxBoundary)) {
            if (has0) c0 = __float2uint_rz(epi_bilinear_float(curI, pitch, lx + (A[0] * j0 + A[1] * i0), ly + (A[2] * j0 + A[3] * i0)));
            if (has1) c1 = __float2uint_rz(epi_bilinear_float(curI, pitch, lx + (A[0] * j1 + A[1] * i1), ly + (A[2] * j1 + A[3] * i1)));
        }
        const uint32_t csum = __reduce_add_sync(0xffffffffu, c0 + c1);
        double z;
        if (eigenMean) {
            // the reference's means are integers here ((sum mod 256) / area in uint8 arithmetic): every term of the score is
            // an integer, the double sum is exact in any order -- one hardware reduction instead of a double butterfly
            const int cm = (int)((csum & 255u) / (uint32_t)(area & 255));
            int zi       = 0;
            if (has0) zi += abs(((int)r0 - refMeanI) - ((int)c0 - cm));
            if (has1) zi += abs(((int)r1 - refMeanI) - ((int)c1 - cm));
            z = (double)__reduce_add_sync(0xffffffffu, zi);
        } else {
            const double curMean = (double)csum / area;
            z                    = 0.0;
            if (has0) z += fabs(((double)r0 - refMean) - ((double)c0 - curMean));
            if (has1) z += fabs(((double)r1 - refMean) - ((double)c1 - curMean));
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        }
        if (z < minScore) {
            minScore = z;
            bestX = lx, bestY = ly;
        }
    }
    res.steps = (int32_t)pixelStep;
    res.score = minScore;
    res.px[0] = bestX, res.px[1] = bestY;
    if (minScore < (double)((uint32_t)area * 128u)) {
        double bc[3];
        cam.bearing(bestX, bestY, bc);
        res.found = epi_triangulate(T, it.bearing, bc, &res.depth) ? 1 : 0;
    }
    if (lane == 0) a.results[item] = res;
}

}  // namespace

svo_status launch_epipolar_match(svo_ctx* ctx, int n, const svo_epi_params& prm)
{
    if (n == 0) return SVO_OK;
    EpiArgs args;
    args.view    = make_view(ctx->arena);
    args.items   = ctx->d_epi_items;
    args.results = ctx->d_epi_results;
    args.n       = n;
    args.prm     = prm;
    for (int i = 0; i < 4; i++) args.K[i] = ctx->cfg.K[i];
    k_epipolar_match<<<(n * 32 + 127) / 128, 128, 0, ctx->stream>>>(args);
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
