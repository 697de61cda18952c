// feature_align.cu -- FeatureAlignment::align (src/feature_alignment.cpp:25-62), batched: one warp per
// (feature, frame) item, the whole 3-parameter optimisation (x, y, intensity offset) stays in registers.
// Everything is FP64 except the bilinear taps, which reproduce algorithm::bilinearInterpolation's FLOAT
// rounding (src/algorithm.cpp:885-894) bit for bit.  Runs on GRADIENT level 0 (:69,118).
// Patch pixels are spread over the lanes (up to 64 pixels = 2 per lane); the 3x3 normal equations are a
// shuffle reduction; the median / MAD of the <= 64 residuals is an exact rank count over shuffles.
#include <float.h>

#include "ctx.h"
#include "math.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct FaArgs {
    ArenaView view;
    const svo_fa_item* items;
    svo_fa_result* results;
    int n;
    svo_fa_params prm;
};

// algorithm::bilinearInterpolation (float), src/algorithm.cpp:885-894
__device__ __forceinline__ float bilinear_float(const uint8_t* __restrict__ img, int pitch, double x, double y)
{
    const int x1 = (int)x, y1 = (int)y;
    const int x2 = x1 + 1, y2 = y1 + 1;
    const uint8_t* p = img + (long long)y1 * pitch + x1;
    const float a = (float)((x2 - x) * (double)__ldg(p) + (x - x1) * (double)__ldg(p + 1));
    const float b = (float)((x2 - x) * (double)__ldg(p + pitch) + (x - x1) * (double)__ldg(p + pitch + 1));
    return (float)((y2 - y) * (double)a + (y - y1) * (double)b);
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// k-th smallest (0-based) of the 64 slots {v0 (lanes), v1 (lanes + 32)}; exact, ties broken by slot index
__device__ double warp_kth(double v0, double v1, int k)
{
    const int lane = threadIdx.x & 31;
    int rank0 = 0, rank1 = 0;
    for (int j = 0; j < 32; j++) {
        const double a = shfl_d(v0, j);
        const double b = shfl_d(v1, j);
        rank0 += (a < v0 || (a == v0 && j < lane)) + (b < v0);
        rank1 += (a < v1 || a == v1) + (b < v1 || (b == v1 && j < lane));
    }
    double out = 0.0;
    const unsigned m0 = __ballot_sync(FULL, rank0 == k);
    const unsigned m1 = __ballot_sync(FULL, rank1 == k);
    if (m0)
        out = shfl_d(v0, __ffs(m0) - 1);
    else if (m1)
        out = shfl_d(v1, __ffs(m1) - 1);
    return out;
}

// median with the reference's rule over `area` entries padded to 64 slots with +inf
__device__ double warp_median(double v0, double v1, int area, int numValid)
{
    const int mid = numValid / 2;
    const double hi = warp_kth(v0, v1, mid);
    if ((area & 1) || mid == 0) return hi;
    return (warp_kth(v0, v1, mid - 1) + hi) / 2.0;
}

__global__ void __launch_bounds__(128) k_feature_align(const FaArgs a)
{
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= a.n) return;
    const svo_fa_item it = a.items[item];
    if (it.ref_slot < 0) {  // slot without a candidate (svo_frontend_run: no 3D point / not reprojected into the frame)
        if (lane == 0) {
            svo_fa_result res;
            res.px[0]       = it.px[0];
            res.px[1]       = it.px[1];
            res.rmse        = __longlong_as_double(0x7ff8000000000000LL);
            res.status      = SVO_ST_FAILED;
            res.iterations  = 0;
            a.results[item] = res;
        }
        return;
    }
    const int P = a.prm.patch_size, area = P * P, half = P / 2, pb = -half;
    const int w = a.view.w[0], h = a.view.h[0], pitch = a.view.pitch[0];
    const uint8_t* refG = a.view.grad[0] + (long long)it.ref_slot * a.view.plane_stride[0];
    const uint8_t* curG = a.view.grad[0] + (long long)it.cur_slot * a.view.plane_stride[0];
    const double INF    = __longlong_as_double(0x7ff0000000000000LL);

    // pixel slots of this lane
    const int p0 = lane, p1 = lane + 32;
    const bool has0 = p0 < area, has1 = p1 < area;
    const int oy0 = pb + p0 / P, ox0 = pb + p0 % P;
    const int oy1 = pb + p1 / P, ox1 = pb + p1 % P;

    double A0 = 1, A1 = 0, A2 = 0, A3 = 1;
    double refBorder = half + 2;  // :67
    if (it.use_affine) {
        A0 = it.A[0];
        A1 = it.A[1];
        A2 = it.A[2];
        A3 = it.A[3];
        double m = 0;
        for (int sx = -1; sx <= 1; sx += 2)
            for (int sy = -1; sy <= 1; sy += 2) {
                const double cx = sx * (half + 1), cy = sy * (half + 1);
                m = fmax(m, fabs(A0 * cx + A1 * cy));
                m = fmax(m, fabs(A2 * cx + A3 * cy));
            }
        refBorder = ceil(m) + 1;
    }
    auto inFrame = [w, h](double x, double y, double b) { return x >= b && y >= b && x < w - b && y < h - b; };

    // ---- computeJacobian, :64-110: template + (gx, gy, 1) rows from the REFERENCE gradient image ----
    double T0 = 0, T1 = 0, gx0 = 0, gy0 = 0, gx1 = 0, gy1 = 0, j20 = 0, j21 = 0;
    if (inFrame(it.ref_px[0], it.ref_px[1], refBorder)) {
        auto sample = [&](double ox, double oy) -> float {
            return bilinear_float(refG, pitch, it.ref_px[0] + A0 * ox + A1 * oy, it.ref_px[1] + A2 * ox + A3 * oy);
        };
        if (has0) {
            T0  = sample(ox0, oy0);
            gx0 = 0.5 * (double)(sample(ox0 + 1, oy0) - sample(ox0 - 1, oy0));  // float subtraction, as the reference
            gy0 = 0.5 * (double)(sample(ox0, oy0 + 1) - sample(ox0, oy0 - 1));
            j20 = 1.0;
        }
        if (has1) {
            T1  = sample(ox1, oy1);
            gx1 = 0.5 * (double)(sample(ox1 + 1, oy1) - sample(ox1 - 1, oy1));
            gy1 = 0.5 * (double)(sample(ox1, oy1 + 1) - sample(ox1, oy1 - 1));
            j21 = 1.0;
        }
    }

    double px = it.px[0], py = it.px[1], pz = 0.0;  // flow = (x, y, 0), :41-43
    const double curBorder = half + 2;
    const int mode    = a.prm.mode;
    const int maxIter = a.prm.max_iter > 0 ? a.prm.max_iter : 20;

    double r0 = INF, r1 = INF, w0 = 0, w1 = 0, chi2 = 0;
    int cnt = 0;
    // computeResiduals :113-168 + tukeyWeighting + chi2; all lanes hold identical scalars afterwards
    auto evaluate = [&](double x, double y, double z) {
        r0 = INF;
        r1 = INF;
        w0 = 0;
        w1 = 0;
        cnt = 0;
        const bool vis = inFrame(x, y, curBorder);
        if (vis) {
            if (has0) r0 = -((double)bilinear_float(curG, pitch, x + ox0, y + oy0) - T0 + z);  // :152
            if (has1) r1 = -((double)bilinear_float(curG, pitch, x + ox1, y + oy1) - T1 + z);
            cnt = area;
        }
        double sigma;
        if (cnt == 0) {
            sigma = DBL_EPSILON;
        } else {
            const double med = warp_median(r0, r1, area, cnt);
            const double d0 = has0 ? fabs(r0 - med) : INF, d1 = has1 ? fabs(r1 - med) : INF;
            const double mad = warp_median(d0, d1, area, cnt);
            sigma = 1.482602218505602 * mad;
            if (sigma <= DBL_EPSILON) sigma = DBL_EPSILON;
        }
        const double c = 4.6851 * sigma, c2 = c * c;
        if (vis) {
            if (has0 && fabs(r0) <= c) {
                const double t = 1.0 - (r0 * r0) / c2;
                w0 = t * t;
            }
            if (has1 && fabs(r1) <= c) {
                const double t = 1.0 - (r1 * r1) / c2;
                w1 = t * t;
            }
        }
        double s = 0;
        if (vis && has0) s += r0 * r0 * w0;
        if (vis && has1) s += r1 * r1 * w1;
        chi2 = warp_sum(s);
    };
    // H (3x3) and g from the current r, w
    auto normal = [&](double* H, double* g) {
        const double rr0 = w0 != 0.0 ? r0 : 0.0, rr1 = w1 != 0.0 ? r1 : 0.0;  // sentinel rows have weight 0
        double s[9];
        s[0] = w0 * gx0 * gx0 + w1 * gx1 * gx1;
        s[1] = w0 * gx0 * gy0 + w1 * gx1 * gy1;
        s[2] = w0 * gx0 * j20 + w1 * gx1 * j21;
        s[3] = w0 * gy0 * gy0 + w1 * gy1 * gy1;
        s[4] = w0 * gy0 * j20 + w1 * gy1 * j21;
        s[5] = w0 * j20 * j20 + w1 * j21 * j21;
        s[6] = w0 * gx0 * rr0 + w1 * gx1 * rr1;
        s[7] = w0 * gy0 * rr0 + w1 * gy1 * rr1;
        s[8] = w0 * j20 * rr0 + w1 * j21 * rr1;
#pragma unroll
        for (int i = 0; i < 9; i++) s[i] = warp_sum(s[i]);
        H[0] = s[0];
        H[1] = s[1];
        H[2] = s[2];
        H[3] = s[1];
        H[4] = s[3];
        H[5] = s[4];
        H[6] = s[2];
        H[7] = s[4];
        H[8] = s[5];
        g[0] = s[6];
        g[1] = s[7];
        g[2] = s[8];
    };

    int status = SVO_ST_FAILED, iters = 0;
    double rmse = 0.0;
    if (mode == SVO_GN) {  // Optimizer::optimizeGN, src/optimizer.cpp:41-159
        double preChi2 = DBL_MAX, qx = px, qy = py, qz = pz;
        int itc = 0;
        while (itc < maxIter) {
            evaluate(px, py, pz);
            double H[9], g[3], dx[3];
            normal(H, g);
            svo::ldlt_solve<3>(H, g, dx);
            iters++;
            const double mx = fmax(dx[0], fmax(dx[1], dx[2]));
            if (mx > 1e3) {
                status = SVO_ST_MAX_COFF_DX;
                break;
            }
            if (isnan(dx[0]) || isnan(dx[1]) || isnan(dx[2])) {
                status = SVO_ST_NAN_IN_DX;
                break;
            }
            if (chi2 > preChi2) {
                status = SVO_ST_INCREASE_CHI2;
                px = qx;
                py = qy;
                pz = qz;
                break;
            }
            qx = px;
            qy = py;
            qz = pz;
            preChi2 = chi2;
            const double step = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            px += dx[0];
            py += dx[1];
            pz += dx[2];
            if (step < 1e-16 || chi2 < 1e-1) {
                status = step < 1e-16 ? SVO_ST_SMALL_STEP : status;
                status = chi2 < 1e-1 ? SVO_ST_SMALL_CHI2 : status;
                break;
            }
            status = SVO_ST_SUCCESS;
            ++itc;
        }
        rmse = sqrt(chi2 / (double)cnt);
    } else {  // Optimizer::optimizeLM, src/optimizer.cpp:161-370
        const bool faithful = mode == SVO_LM_FAITHFUL;
        double lambda = 1e-2, nu = 2.0;
        evaluate(px, py, pz);
        // accepted state
        double a_r0 = r0, a_r1 = r1, a_w0 = w0, a_w1 = w1, a_chi2 = chi2;
        int a_cnt = cnt;
        double qx = px, qy = py, qz = pz, preChi2 = 0;
        bool success = true;
        int itc = 0;
        while (itc < maxIter) {
            if (success) {
                qx = px;
                qy = py;
                qz = pz;
                preChi2 = a_chi2;
                status  = SVO_ST_SUCCESS;
            }
            r0 = a_r0;
            r1 = a_r1;
            w0 = a_w0;
            w1 = a_w1;
            double H[9], g[3], dx[3];
            normal(H, g);
            if (itc == 0) lambda *= fmax(H[0], fmax(H[4], H[8]));
            H[0] += lambda;
            H[4] += lambda;
            H[8] += lambda;
            svo::ldlt_solve<3>(H, g, dx);
            px += dx[0];
            py += dx[1];
            pz += dx[2];
            iters++;
            const double mx = fmax(dx[0], fmax(dx[1], dx[2]));
            if (mx > 1e3) {
                status = SVO_ST_MAX_COFF_DX;
                break;
            }
            if (isnan(dx[0]) || isnan(dx[1]) || isnan(dx[2])) {
                status = SVO_ST_NAN_IN_DX;
                break;
            }
            const double step = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            if (step < 1e-16 || lambda >= 1e14 || lambda <= 1e-14 || faithful) {
                status = step < 1e-16 ? SVO_ST_SMALL_STEP : status;
                status = fabs(lambda) >= 1e14 ? SVO_ST_LAMBDA : status;
                break;
            }
            evaluate(px, py, pz);
            success = svo::nielsen_update(preChi2, chi2, lambda, nu);
            if (success) {
                a_r0 = r0;
                a_r1 = r1;
                a_w0 = w0;
                a_w1 = w1;
                a_chi2 = chi2;
                a_cnt  = cnt;
            } else {
                px = qx;
                py = qy;
                pz = qz;
            }
            ++itc;
        }
        rmse = sqrt(a_chi2 / (double)a_cnt);
    }
    if (lane == 0) {
        svo_fa_result res;
        res.px[0]       = px;
        res.px[1]       = py;
        res.rmse        = rmse;
        res.status      = status;
        res.iterations  = iters;
        a.results[item] = res;
    }
}

}  // namespace

svo_status launch_feature_align(svo_ctx* ctx, svo_fa_result* results)
{
    if (ctx->staged_fa == 0) return SVO_OK;
    FaArgs args;
    args.view    = make_view(ctx->arena);
    args.items   = ctx->d_fa_items;
    args.results = results ? results : ctx->d_fa_results;  // override: mapped host memory (front-end graph)
    args.n       = ctx->staged_fa;
    args.prm     = ctx->staged_fa_params;
    const int blocks = (ctx->staged_fa * 32 + 127) / 128;
    k_feature_align<<<blocks, 128, 0, ctx->stream>>>(args);
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
