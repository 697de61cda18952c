// sparse_align.cu -- ImageAlignment::align (src/image_alignment.cpp:25-67) as ONE persistent kernel:
// one CTA per frame pair ("job") runs all pyramid levels and all optimizer iterations without
// returning to the host.  Per level:
//   precompute  (computeJacobian, :69-192)   template patch T, gradients gx gy per patch pixel (L2-resident
//               scratch) and the 2x6 image Jacobian rows A, B per feature (computeImageJac, :194-248)
//   evaluate    (computeResiduals, :251-370) warp the features with the current pose (FP64, thread per
//               feature), bilinear-sample the current image (FP32, warp per feature, lane per patch pixel),
//               r = I - T into shared memory
//   sigma       (tukeyWeighting/computeSigma, src/optimizer.cpp:485-507, src/algorithm.cpp:834-872) two exact
//               order statistics (median, MAD) by a block-wide 8-bit MSD radix select on the float keys
//   reduce      J^T W J, J^T W r, chi2.  Because every Jacobian row of a feature is gx*A + gy*B, the 6x6
//               sum factorises per feature into  sxx AA^T + sxy (AB^T + BA^T) + syy BB^T  with the three
//               patch sums sxx = sum w gx^2, sxy = sum w gx gy, syy = sum w gy^2 (and bx, by for J^T W r),
//               so the per-pixel work is 6 FMAs instead of 27; features are accumulated in FP64.
//   solve       (optimizeLM :161-370 / optimizeGN :41-159) damping, pivoted LDLT 6x6, pose <- pose exp(-dx)
//               in FP64 on thread 0; the accept/reject logic of the reference stays on the device.
// Tensor cores are not used: the path is a gather plus a 28-scalar reduction.
#include <float.h>
#include <stdlib.h>

#include "ctx.h"
#include "math.cuh"
#include "align_common.cuh"

namespace {

constexpr int GEO_INVALID = -32768;
constexpr unsigned FULL   = 0xffffffffu;

struct AlignArgs {
    ArenaView view;
    const svo_align_job* jobs;
    const svo_align_feature* feats;
    svo_align_result* results;
    svo_align_level_stats* stats;  // nullable
    float* scratch_tpl;            // [job][3][tpl_stride]
    float* scratch_jac;            // [job][max_features][12]
    int tpl_stride;                // max_features * area
    int max_features;
    svo_align_params prm;
    double K[4];
};

struct __align__(16) Geo {  // per feature, per evaluation
    int uI, vI;
    float fu, fv;
};

__device__ __forceinline__ uint32_t f2key(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct KeySigned {
    __device__ __forceinline__ uint32_t operator()(float r) const { return f2key(r); }
};
struct KeyAbsDev {
    float med;
    __device__ __forceinline__ uint32_t operator()(float r) const { return __float_as_uint(fabsf(r - med)); }
};

// Block-wide exact k-th smallest (0-based) of key(r[0..N)) by MSD radix select, 8 bits per pass.
// hist: 64 * NT words, zero on entry and on exit.  Each thread owns one word column (bank = tid % 32,
// conflict-free); a word packs the 8-bit counters of 4 adjacent digits.  Requires N <= 255 * NT and N < 65536.
// Returns the key; *less_out = number of elements strictly smaller than it.
template <int NT, class KeyFn>
__device__ uint32_t block_select(const float* r, int N, int k, KeyFn key, uint32_t* hist, uint32_t* tot, int* less_out)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    uint32_t prefix = 0, mask = 0;
    const int k0 = k;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < N; i += NT) {
            const uint32_t kk = key(r[i]);
            if ((kk & mask) == prefix) {
                const uint32_t d = (kk >> shift) & 255u;
                hist[(d >> 2) * NT + tid] += 1u << ((d & 3u) * 8u);
            }
        }
        __syncthreads();
        for (int row = warp; row < 64; row += NW) {
            uint32_t lo = 0, hi = 0;
            for (int j = lane; j < NT; j += 32) {
                const uint32_t wv   = hist[row * NT + j];
                hist[row * NT + j] = 0;
                lo += wv & 0x00ff00ffu;
                hi += (wv >> 8) & 0x00ff00ffu;
            }
            lo = __reduce_add_sync(FULL, lo);
            hi = __reduce_add_sync(FULL, hi);
            if (lane == 0) {
                tot[row * 4 + 0] = lo & 0xffffu;
                tot[row * 4 + 1] = hi & 0xffffu;
                tot[row * 4 + 2] = lo >> 16;
                tot[row * 4 + 3] = hi >> 16;
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t c[8];
            uint32_t s = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c[i] = tot[lane * 8 + i];
                s += c[i];
            }
            uint32_t incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            uint32_t cum = incl - s;
            if ((uint32_t)k >= cum && (uint32_t)k < incl) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if ((uint32_t)k >= cum && (uint32_t)k < cum + c[i]) {
                        tot[256] = lane * 8 + i;
                        tot[257] = (uint32_t)k - cum;
                    }
                    cum += c[i];
                }
            }
        }
        __syncthreads();
        const uint32_t digit = tot[256];
        k                    = (int)tot[257];
        prefix |= digit << shift;
        mask |= 255u << shift;
    }
    *less_out = k0 - k;
    return prefix;
}

// largest key strictly below `bound` (0 if none)
template <int NT, class KeyFn>
__device__ uint32_t block_max_below(const float* r, int N, uint32_t bound, KeyFn key, uint32_t* tot)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    uint32_t m = 0;
    for (int i = tid; i < N; i += NT) {
        const uint32_t kk = key(r[i]);
        if (kk < bound) m = max(m, kk);
    }
    m = __reduce_max_sync(FULL, m);
    __syncthreads();  // tot[256..] may still be read by slower threads of the previous select
    if (lane == 0) tot[260 + warp] = m;
    __syncthreads();
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) out = max(out, tot[260 + i]);
    return out;
}

// median with the reference's rule (src/algorithm.cpp:834-853): element numValid/2; when the TOTAL count N is
// even the mean of that element and its exact predecessor (MEDIAN_EXACT)
template <int NT, class KeyFn, class Inv>
__device__ double block_median(const float* r, int N, int numValid, KeyFn key, Inv inv, uint32_t* hist, uint32_t* tot)
{
    const int mid = numValid / 2;
    int less;
    const uint32_t hiKey = block_select<NT>(r, N, mid, key, hist, tot, &less);
    const double hi      = (double)inv(hiKey);
    if ((N & 1) || mid == 0) return hi;
    uint32_t loKey = hiKey;
    if (mid - 1 < less) loKey = block_max_below<NT>(r, N, hiKey, key, tot);
    return ((double)inv(loKey) + hi) / 2.0;
}

struct InvSigned {
    __device__ __forceinline__ float operator()(uint32_t k) const { return key2f(k); }
};
struct InvAbs {
    __device__ __forceinline__ float operator()(uint32_t k) const { return __uint_as_float(k); }
};

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_sparse_align(const AlignArgs a)
{
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int job = blockIdx.x;
    const svo_align_job* J = a.jobs + job;
    const int nRef = J->n_ref, nKf = J->n_kf;
    const int F    = nRef + nKf;
    const int P    = a.prm.patch_size;
    const int area = P * P;
    const int half = P / 2;
    const int pb   = -half;
    const int border = half + 2;
    const int N    = F * area;
    const int nLevels = a.prm.max_level - a.prm.min_level + 1;

    // ---- shared memory carve-up (sized for max_features) ----
    Ctrl* ctrl    = reinterpret_cast<Ctrl*>(smem);
    size_t off    = (sizeof(Ctrl) + 15) & ~size_t(15);
    double* red   = reinterpret_cast<double*>(smem + off);  // [NW][28]
    off += sizeof(double) * NW * 28;
    double* pW = reinterpret_cast<double*>(smem + off);  // [F][3]
    off += sizeof(double) * 3 * a.max_features;
    off = (off + 15) & ~size_t(15);
    Geo* geo = reinterpret_cast<Geo*>(smem + off);  // [F]
    off += sizeof(Geo) * a.max_features;
    float* r = reinterpret_cast<float*>(smem + off);  // [F * area]; doubles as FP64 scratch during precompute
    off += sizeof(float) * (size_t)a.max_features * area;
    off = (off + 15) & ~size_t(15);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + off);  // [64][NT]
    off += sizeof(uint32_t) * 64 * NT;
    uint32_t* tot = reinterpret_cast<uint32_t*>(smem + off);  // [256 + 4 + 4 + NW]
    off += sizeof(uint32_t) * (264 + NW);
    uint8_t* fflag = smem + off;  // [F] bit0 has_point, bit1 visible in the reference image at this level

    float* tplT  = a.scratch_tpl + (size_t)job * 3 * a.tpl_stride;
    float* tplGx = tplT + a.tpl_stride;
    float* tplGy = tplGx + a.tpl_stride;
    float* jac   = a.scratch_jac + (size_t)job * a.max_features * 12;

    if (nRef == 0) {  // src/image_alignment.cpp:27-28
        if (tid == 0) {
            svo_align_result res;
            for (int i = 0; i < 7; i++) res.T_cur[i] = J->T_cur[i];
            res.rmse        = 0.0;
            res.status      = SVO_ST_SUCCESS;
            res.evaluations = 0;
            res.iterations  = 0;
            res.reserved    = 0;
            a.results[job]  = res;
        }
        return;
    }

    for (int i = tid; i < 64 * NT; i += NT) hist[i] = 0;

    // ---- prologue: world points of all features (level independent) ----
    // p_W = T_frame^-1 (bearing * |P - C_frame|)   src/image_alignment.cpp:153-155
    for (int f = tid; f < F; f += NT) {
        const svo_align_feature* ft = a.feats + J->feat_offset + f;
        const double* Tp            = f < nRef ? J->T_ref : J->T_kf;
        Pose Tf;
#pragma unroll
        for (int i = 0; i < 4; i++) Tf.q[i] = Tp[i];
#pragma unroll
        for (int i = 0; i < 3; i++) Tf.t[i] = Tp[4 + i];
        double C[3];
        svo::pose_camera_in_world(Tf, C);
        const double d0 = ft->point[0] - C[0], d1 = ft->point[1] - C[1], d2 = ft->point[2] - C[2];
        const double depthNorm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double pC[3]     = {ft->bearing[0] * depthNorm, ft->bearing[1] * depthNorm, ft->bearing[2] * depthNorm};
        double w3[3];
        svo::pose_inv_act(Tf, pC, w3);
        pW[3 * f + 0] = w3[0];
        pW[3 * f + 1] = w3[1];
        pW[3 * f + 2] = w3[2];
        fflag[f]      = ft->has_point ? 1 : 0;
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) ctrl->pose.q[i] = J->T_cur[i];
#pragma unroll
        for (int i = 0; i < 3; i++) ctrl->pose.t[i] = J->T_cur[4 + i];
        set_Rt(ctrl);
        ctrl->evals_total = 0;
        ctrl->iters_total = 0;
        ctrl->status      = SVO_ST_FAILED;
        ctrl->rmse        = 0.0;
    }
    __syncthreads();

    const double fx0 = a.K[0], fy0 = a.K[1], cx0 = a.K[2], cy0 = a.K[3];

#pragma unroll 1
    for (int level = a.prm.max_level, si = 0; level >= a.prm.min_level; level--, si++) {
        const int lw = a.view.w[level], lh = a.view.h[level], lpitch = a.view.pitch[level];
        const double denom = (double)(1 << level);
        const double scale = 1.0 / denom;
        const uint8_t* refImg = a.view.img[level] + (long long)J->ref_slot * a.view.plane_stride[level];
        const uint8_t* kfImg  = a.view.img[level] + (long long)J->kf_slot * a.view.plane_stride[level];
        const uint8_t* curImg = a.view.img[level] + (long long)J->cur_slot * a.view.plane_stride[level];

        // ================= precompute (computeJacobian) =================
        double* fracs = reinterpret_cast<double*>(r);  // [F][2] FP64 fractions, r is idle here
        for (int f = tid; f < F; f += NT) {
            Geo g;
            g.uI = GEO_INVALID;
            g.vI = 0;
            g.fu = g.fv = 0.f;
            uint8_t fl = fflag[f] & 1;
            if (fl) {
                const svo_align_feature* ft = a.feats + J->feat_offset + f;
                const double u = ft->px[0] * scale, v = ft->px[1] * scale;
                const double uf = floor(u), vf = floor(v);
                const int uI = (int)uf, vI = (int)vf;
                if (!((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh)) {
                    fl |= 2;
                    g.uI             = uI;
                    g.vI             = vI;
                    fracs[2 * f + 0] = u - uf;
                    fracs[2 * f + 1] = v - vf;
                    // computeImageJac at the WORLD point, scaled focal lengths
                    const double fx = fx0 / denom, fy = fy0 / denom;
                    const double x = pW[3 * f], y = pW[3 * f + 1], z = pW[3 * f + 2];
                    const double x2 = x * x, y2 = y * y, z2 = z * z;
                    float* jf = jac + f * 12;
                    jf[0]     = (float)(fx / z);
                    jf[1]     = 0.f;
                    jf[2]     = (float)(-(fx * x) / z2);
                    jf[3]     = (float)(-(fx * x * y) / z2);
                    jf[4]     = (float)((fx * x2) / z2 + fx);
                    jf[5]     = (float)(-(fx * y) / z);
                    jf[6]     = 0.f;
                    jf[7]     = (float)(fy / z);
                    jf[8]     = (float)(-(fy * y) / z2);
                    jf[9]     = (float)(-(fy * y2) / z2 - fy);
                    jf[10]    = (float)((fy * x * y) / z2);
                    jf[11]    = (float)((fy * x) / z);
                }
            }
            fflag[f] = fl;
            geo[f]   = g;
        }
        __syncthreads();
        for (int f = warp; f < F; f += NW) {
            if (!(fflag[f] & 2)) continue;
            const Geo g        = geo[f];
            const double fu = fracs[2 * f], fv = fracs[2 * f + 1];
            const uint8_t* img = f < nRef ? refImg : kfImg;
            for (int p = lane; p < area; p += 32) {
                const int py = p / P, px = p - py * P;
                const uint8_t* c = img + (long long)(g.vI + pb + py) * lpitch + (g.uI + pb + px);
                // 12-byte cross footprint around the bilinear cell
                const double i00 = __ldg(c), i01 = __ldg(c + 1), i10 = __ldg(c + lpitch), i11 = __ldg(c + lpitch + 1);
                const double l0 = __ldg(c - 1), l1 = __ldg(c + lpitch - 1);
                const double r0 = __ldg(c + 2), r1 = __ldg(c + lpitch + 2);
                const double t0 = __ldg(c - lpitch), t1 = __ldg(c - lpitch + 1);
                const double b0 = __ldg(c + 2 * lpitch), b1 = __ldg(c + 2 * lpitch + 1);
                const double wu0 = 1.0 - fu, wv0 = 1.0 - fv;
                auto bil = [&](double a00, double a01, double a10, double a11) {
                    const double ta = wu0 * a00 + fu * a01;
                    const double tb = wu0 * a10 + fu * a11;
                    return wv0 * ta + fv * tb;
                };
                const double T  = bil(i00, i01, i10, i11);
                const double gx = 0.5 * (bil(i01, r0, i11, r1) - bil(l0, i00, l1, i10));
                const double gy = 0.5 * (bil(i10, i11, b0, b1) - bil(t0, t1, i00, i01));
                const int idx   = f * area + p;
                tplT[idx]       = (float)T;
                tplGx[idx]      = (float)gx;
                tplGy[idx]      = (float)gy;
            }
        }
        if (tid == 0) {
            ctrl->lambda      = 1e-2;
            ctrl->nu          = 2.0;
            ctrl->it          = 0;
            ctrl->done        = 0;
            ctrl->success     = 1;
            ctrl->status      = SVO_ST_FAILED;
            ctrl->first       = 1;
            ctrl->evals_level = 0;
            ctrl->iters_level = 0;
            ctrl->preChi2     = DBL_MAX;
            ctrl->pre_pose    = ctrl->pose;
            ctrl->nvis        = 0;
        }
        __syncthreads();

        // ================= evaluate (computeResiduals + tukeyWeighting + normal equations) =================
        auto evaluate = [&]() {
            // --- A: warp features (FP64), thread per feature ---
            const double R0 = ctrl->R[0], R1 = ctrl->R[1], R2 = ctrl->R[2], R3 = ctrl->R[3], R4 = ctrl->R[4],
                         R5 = ctrl->R[5], R6 = ctrl->R[6], R7 = ctrl->R[7], R8 = ctrl->R[8];
            const double t0 = ctrl->t[0], t1 = ctrl->t[1], t2 = ctrl->t[2];
            int nv = 0;
            for (int f0 = 0; f0 < F; f0 += NT) {
                const int f = f0 + tid;
                bool vis    = false;
                Geo g;
                g.uI = GEO_INVALID;
                g.vI = 0;
                g.fu = g.fv = 0.f;
                if (f < F && (fflag[f] & 2)) {
                    const double x = pW[3 * f], y = pW[3 * f + 1], z = pW[3 * f + 2];
                    const double cxp = R0 * x + R1 * y + R2 * z + t0;
                    const double cyp = R3 * x + R4 * y + R5 * z + t1;
                    const double czp = R6 * x + R7 * y + R8 * z + t2;
                    // PinholeCamera::project2d, src/pinhole_camera.cpp:55-56 (no z > 0 test)
                    const double u = (fx0 * (cxp / czp) + cx0) * scale;
                    const double v = (fy0 * (cyp / czp) + cy0) * scale;
                    if (isfinite(u) && isfinite(v) && fabs(u) < 1e6 && fabs(v) < 1e6) {
                        const double uf = floor(u), vf = floor(v);
                        const int uI = (int)uf, vI = (int)vf;
                        if (!((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh)) {
                            vis  = true;
                            g.uI = uI;
                            g.vI = vI;
                            g.fu = (float)(u - uf);
                            g.fv = (float)(v - vf);
                        }
                    }
                }
                if (f < F) geo[f] = g;
                nv += __popc(__ballot_sync(FULL, vis));
            }
            if (lane == 0 && nv) atomicAdd(&ctrl->nvis, nv);
            __syncthreads();
            // --- B: residuals, warp per feature, lane per patch pixel ---
            for (int f = warp; f < F; f += NW) {
                const Geo g = geo[f];
                if (g.uI == GEO_INVALID) {
                    for (int p = lane; p < area; p += 32) r[f * area + p] = __int_as_float(0x7f800000);
                    continue;
                }
                const float fu = g.fu, fv = g.fv, wu0 = 1.f - fu, wv0 = 1.f - fv;
                const uint8_t* base = curImg + (long long)(g.vI + pb) * lpitch + (g.uI + pb);
                for (int p = lane; p < area; p += 32) {
                    const int py = p / P, px = p - py * P;
                    const uint8_t* c = base + py * lpitch + px;
                    const float i00 = __ldg(c), i01 = __ldg(c + 1), i10 = __ldg(c + lpitch), i11 = __ldg(c + lpitch + 1);
                    const float ta = wu0 * i00 + fu * i01;
                    const float tb = wu0 * i10 + fu * i11;
                    const int idx  = f * area + p;
                    r[idx]         = (wv0 * ta + fv * tb) - tplT[idx];
                }
            }
            __syncthreads();
            // --- C: sigma = 1.4826 MAD (two exact order statistics) ---
            const int nvis     = ctrl->nvis;
            const int numValid = nvis * area;
            double sigma;
            if (nvis == 0) {
                sigma = DBL_EPSILON;  // the reference gets MAD = 0 from all-sentinel input
            } else {
                const double med = block_median<NT>(r, N, numValid, KeySigned{}, InvSigned{}, hist, tot);
                KeyAbsDev kd{(float)med};
                const double mad = block_median<NT>(r, N, numValid, kd, InvAbs{}, hist, tot);
                sigma            = 1.482602218505602 * mad;
                if (sigma <= DBL_EPSILON) sigma = DBL_EPSILON;
            }
            const double cD  = 4.6851 * sigma;
            const float cF   = (float)cD;
            const float ic2  = (float)(1.0 / (cD * cD));
            // --- D: per-feature patch sums (FP32) -> 6x6 contribution, accumulated over features in FP64 ---
            double acc = 0.0;
            int ea = 0, eb = 0;
            if (lane < 21) {
                ea = c_pairA[lane];
                eb = c_pairB[lane];
            } else if (lane < 27) {
                ea = lane - 21;
            }
            for (int f = warp; f < F; f += NW) {
                if (geo[f].uI == GEO_INVALID) continue;
                float sxx = 0.f, sxy = 0.f, syy = 0.f, bx = 0.f, by = 0.f, ch = 0.f;
                for (int p = lane; p < area; p += 32) {
                    const int idx  = f * area + p;
                    const float rr = r[idx];
                    const float gx = tplGx[idx], gy = tplGy[idx];
                    float w        = 0.f;
                    if (fabsf(rr) <= cF) {
                        const float t = 1.f - rr * rr * ic2;
                        w             = t * t;
                    }
                    const float wgx = w * gx, wgy = w * gy, wr = w * rr;
                    sxx += wgx * gx;
                    sxy += wgx * gy;
                    syy += wgy * gy;
                    bx += wr * gx;
                    by += wr * gy;
                    ch += wr * rr;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sxx += __shfl_xor_sync(FULL, sxx, o);
                    sxy += __shfl_xor_sync(FULL, sxy, o);
                    syy += __shfl_xor_sync(FULL, syy, o);
                    bx += __shfl_xor_sync(FULL, bx, o);
                    by += __shfl_xor_sync(FULL, by, o);
                    ch += __shfl_xor_sync(FULL, ch, o);
                }
                const float* jf = jac + f * 12;
                float val       = 0.f;
                if (lane < 21) {
                    const float Aa = jf[ea], Ab = jf[eb], Ba = jf[6 + ea], Bb = jf[6 + eb];
                    val = sxx * (Aa * Ab) + sxy * (Aa * Bb + Ba * Ab) + syy * (Ba * Bb);
                } else if (lane < 27) {
                    val = bx * jf[ea] + by * jf[6 + ea];
                } else if (lane == 27) {
                    val = ch;
                }
                acc += (double)val;
            }
            if (lane < 28) red[warp * 28 + lane] = acc;
            __syncthreads();
            if (tid < 28) {
                double s = 0.0;
#pragma unroll
                for (int wI = 0; wI < NW; wI++) s += red[wI * 28 + tid];
                ctrl->E[tid] = s;
            }
            if (tid == 32) {
                ctrl->sigma  = sigma;
                ctrl->n_eval = numValid;
                ctrl->nvis   = 0;
                ctrl->evals_level++;
            }
            __syncthreads();
        };

        // record the first iteration of the level for the stats record
        auto record_first = [&](const double* H, const double* g, double chi2, double lambda, int n) {
            if (!ctrl->first) return;
            ctrl->first = 0;
            if (a.stats) {
                svo_align_level_stats* s = a.stats + (size_t)job * nLevels + si;
                for (int i = 0; i < 36; i++) s->H[i] = H[i];
                for (int i = 0; i < 6; i++) s->g[i] = g[i];
                s->chi2   = chi2;
                s->sigma  = ctrl->first_sigma;
                s->lambda = lambda;
                s->n_px   = n;
            }
        };

        if (a.prm.mode == SVO_GN) {
            // ---------------- Optimizer::optimizeGN, src/optimizer.cpp:41-159 ----------------
            const int maxIter = a.prm.max_iter > 0 ? a.prm.max_iter : 20;
#pragma unroll 1
            while (true) {
                evaluate();
                if (tid == 0) {
                    double H[36], g[6], dx[6];
                    expand_H(ctrl->E, H, g);
                    const double chi2 = ctrl->E[27];
                    if (ctrl->first) ctrl->first_sigma = ctrl->sigma;
                    const bool wasFirst = ctrl->first;
                    record_first(H, g, chi2, 0.0, ctrl->n_eval);
                    svo::ldlt_solve<6>(H, g, dx);
                    if (wasFirst && a.stats)
                        for (int i = 0; i < 6; i++) a.stats[(size_t)job * nLevels + si].dx[i] = dx[i];
                    ctrl->iters_level++;
                    double mx = dx[0];
                    bool nan  = false;
                    for (int i = 0; i < 6; i++) {
                        mx = fmax(mx, dx[i]);  // fmax ignores NaN, the NaN test below catches it
                        nan |= isnan(dx[i]);
                    }
                    if (mx > 1e3) {
                        ctrl->status = SVO_ST_MAX_COFF_DX;
                        ctrl->done   = 1;
                    } else if (nan) {
                        ctrl->status = SVO_ST_NAN_IN_DX;
                        ctrl->done   = 1;
                    } else if (chi2 > ctrl->preChi2) {
                        ctrl->status = SVO_ST_INCREASE_CHI2;
                        ctrl->pose   = ctrl->pre_pose;  // rollback, :113-118
                        ctrl->done   = 1;
                    } else {
                        ctrl->pre_pose = ctrl->pose;
                        ctrl->preChi2  = chi2;
                        double step    = 0;
                        for (int i = 0; i < 6; i++) step += dx[i] * dx[i];
                        svo::pose_update_right_exp_neg(ctrl->pose, dx);
                        if (step < 1e-16 || chi2 < 1e-1) {
                            int st = ctrl->status;
                            st     = step < 1e-16 ? SVO_ST_SMALL_STEP : st;
                            st     = chi2 < 1e-1 ? SVO_ST_SMALL_CHI2 : st;
                            ctrl->status = st;
                            ctrl->done   = 1;
                        } else {
                            ctrl->status = SVO_ST_SUCCESS;
                            ctrl->it++;
                            if (ctrl->it >= maxIter) ctrl->done = 1;
                        }
                    }
                    set_Rt(ctrl);
                    ctrl->rmse = sqrt(chi2 / (double)ctrl->n_eval);
                }
                __syncthreads();
                if (ctrl->done) break;
            }
        } else {
            // ---------------- Optimizer::optimizeLM, src/optimizer.cpp:161-370 ----------------
            const bool faithful = a.prm.mode == SVO_LM_FAITHFUL;
            const int maxIter   = a.prm.max_iter > 0 ? a.prm.max_iter : 20;
            evaluate();
            if (tid == 0) {
                for (int i = 0; i < 28; i++) ctrl->curE[i] = ctrl->E[i];
                ctrl->cur_n       = ctrl->n_eval;
                ctrl->first_sigma = ctrl->sigma;
            }
            __syncthreads();
#pragma unroll 1
            while (true) {
                if (tid == 0) {
                    if (ctrl->success) {  // :224-233 snapshot
                        ctrl->pre_pose = ctrl->pose;
                        ctrl->preChi2  = ctrl->curE[27];
                        ctrl->status   = SVO_ST_SUCCESS;
                    }
                    double H[36], g[6], dx[6];
                    expand_H(ctrl->curE, H, g);
                    if (ctrl->it == 0) {
                        double mx = H[0];
                        for (int i = 1; i < 6; i++) mx = fmax(mx, H[i * 6 + i]);
                        ctrl->lambda *= mx;  // :296-299
                    }
                    const double lambda = ctrl->lambda;
                    const bool wasFirst = ctrl->first;
                    record_first(H, g, ctrl->curE[27], lambda, ctrl->cur_n);
                    for (int i = 0; i < 6; i++) H[i * 6 + i] += lambda;
                    svo::ldlt_solve<6>(H, g, dx);
                    if (wasFirst && a.stats)
                        for (int i = 0; i < 6; i++) a.stats[(size_t)job * nLevels + si].dx[i] = dx[i];
                    svo::pose_update_right_exp_neg(ctrl->pose, dx);  // :310 applied before any check
                    ctrl->iters_level++;
                    double mx = dx[0], step = 0;
                    bool nan = false;
                    for (int i = 0; i < 6; i++) {
                        mx = fmax(mx, dx[i]);
                        nan |= isnan(dx[i]);
                        step += dx[i] * dx[i];
                    }
                    if (mx > 1e3) {
                        ctrl->status = SVO_ST_MAX_COFF_DX;
                        ctrl->done   = 1;
                    } else if (nan) {
                        ctrl->status = SVO_ST_NAN_IN_DX;
                        ctrl->done   = 1;
                    } else if (step < 1e-16 || lambda >= 1e14 || lambda <= 1e-14 || faithful) {
                        // :328 -- in the reference the clause `normDiffPose < m_normInfDiff` is always true
                        int st = ctrl->status;
                        st     = step < 1e-16 ? SVO_ST_SMALL_STEP : st;
                        st     = fabs(lambda) >= 1e14 ? SVO_ST_LAMBDA : st;
                        ctrl->status = st;
                        ctrl->done   = 1;
                    }
                    set_Rt(ctrl);
                }
                __syncthreads();
                if (ctrl->done) break;
                evaluate();
                if (tid == 0) {
                    const bool ok = svo::nielsen_update(ctrl->preChi2, ctrl->E[27], ctrl->lambda, ctrl->nu);
                    ctrl->success = ok;
                    if (ok) {
                        for (int i = 0; i < 28; i++) ctrl->curE[i] = ctrl->E[i];
                        ctrl->cur_n = ctrl->n_eval;
                    } else {
                        ctrl->pose = ctrl->pre_pose;  // :346-360 rollback
                        set_Rt(ctrl);
                    }
                    ctrl->it++;
                    if (ctrl->it >= maxIter) ctrl->done = 1;
                }
                __syncthreads();
                if (ctrl->done) break;
            }
            if (tid == 0) ctrl->rmse = sqrt(ctrl->curE[27] / (double)ctrl->cur_n);
        }
        if (tid == 0) {
            ctrl->evals_total += ctrl->evals_level;
            ctrl->iters_total += ctrl->iters_level;
            if (a.stats) {
                svo_align_level_stats* s = a.stats + (size_t)job * nLevels + si;
                for (int i = 0; i < 4; i++) s->pose_after[i] = ctrl->pose.q[i];
                for (int i = 0; i < 3; i++) s->pose_after[4 + i] = ctrl->pose.t[i];
                s->rmse        = ctrl->rmse;
                s->status      = ctrl->status;
                s->iterations  = ctrl->iters_level;
                s->evaluations = ctrl->evals_level;
            }
        }
        __syncthreads();
    }

    if (tid == 0) {
        svo_align_result res;
        for (int i = 0; i < 4; i++) res.T_cur[i] = ctrl->pose.q[i];
        for (int i = 0; i < 3; i++) res.T_cur[4 + i] = ctrl->pose.t[i];
        res.rmse        = ctrl->rmse;
        res.status      = ctrl->status;
        res.evaluations = ctrl->evals_total;
        res.iterations  = ctrl->iters_total;
        res.reserved    = 0;
        a.results[job]  = res;
    }
}

}  // namespace

size_t sparse_align_smem_bytes(int nthreads, int max_features, int area)
{
    const int NW = nthreads / 32;
    size_t off   = (sizeof(Ctrl) + 15) & ~size_t(15);
    off += sizeof(double) * NW * 28;
    off += sizeof(double) * 3 * max_features;
    off = (off + 15) & ~size_t(15);
    off += sizeof(Geo) * max_features;
    off += sizeof(float) * (size_t)max_features * area;
    off = (off + 15) & ~size_t(15);
    off += sizeof(uint32_t) * 64 * nthreads;
    off += sizeof(uint32_t) * (264 + NW);
    off += max_features;
    return (off + 15) & ~size_t(15);
}

svo_status launch_sparse_align(svo_ctx* ctx)
{
    const svo_align_params& prm = ctx->staged_params;
    const int area              = prm.patch_size * prm.patch_size;
    const int nJobs             = ctx->staged_jobs;
    if (nJobs == 0) return SVO_OK;
    // feature capacity of this launch: the largest job, rounded up (keeps shared memory small for small jobs)
    int maxF = 1;
    for (int j = 0; j < nJobs; j++) maxF = std::max(maxF, ctx->src_jobs[j].n_ref + ctx->src_jobs[j].n_kf);
    // fast path (sparse_align_v3.cu): patch 4 / 5, a thread-block cluster per pair, <= 2,048 features per pair.
    // SVO_ALIGN_GENERIC=1 forces the generic kernel below (A/B measurements); both are CUDA paths, there is no CPU
    // fallback anywhere.
    {
        const char* e = getenv("SVO_ALIGN_GENERIC");
        const char* v = getenv("SVO_ALIGN_V4");  // "0": keep the cluster kernel for <= 512 features too (A/B measurements)
        if (!(e && e[0] == '1')) {
            if (!(v && v[0] == '0') && sparse_align_v5_supported(ctx, maxF)) return launch_sparse_align_v5(ctx, maxF);
            if (sparse_align_v3_supported(ctx, maxF)) return launch_sparse_align_v3(ctx, maxF);
        }
    }
    maxF = (maxF + 15) & ~15;
    if ((int64_t)maxF * area >= 65536) SVO_FAIL(SVO_ERR_CAPACITY, "features * patch area must stay below 65536");

    AlignArgs args;
    args.view         = make_view(ctx->arena);
    args.jobs         = ctx->d_jobs;
    args.feats        = ctx->d_feats;
    args.results      = ctx->d_results;
    args.stats        = ctx->staged_want_stats ? ctx->d_stats : nullptr;
    args.scratch_tpl  = ctx->d_scratch_tpl;
    args.scratch_jac  = ctx->d_scratch_jac;
    args.tpl_stride   = ctx->cfg.max_features * area;
    args.max_features = maxF;
    args.prm          = prm;
    for (int i = 0; i < 4; i++) args.K[i] = ctx->cfg.K[i];

    const size_t s512 = sparse_align_smem_bytes(512, maxF, area);
    const size_t s256 = sparse_align_smem_bytes(256, maxF, area);
    if (s512 <= (size_t)ctx->max_smem_optin && (int64_t)maxF * area <= 255 * 512) {
        SVO_CUDA(cudaFuncSetAttribute(k_sparse_align<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s512));
        k_sparse_align<512><<<nJobs, 512, s512, ctx->stream>>>(args);
    } else if (s256 <= (size_t)ctx->max_smem_optin && (int64_t)maxF * area <= 255 * 256) {
        SVO_CUDA(cudaFuncSetAttribute(k_sparse_align<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s256));
        k_sparse_align<256><<<nJobs, 256, s256, ctx->stream>>>(args);
    } else {
        SVO_FAIL(SVO_ERR_CAPACITY, "sparse alignment job does not fit in shared memory");
    }
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
