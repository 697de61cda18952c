// sparse_align_v5.cu -- ImageAlignment::align (src/image_alignment.cpp:25-67) for patch sizes 4 / 5 and at most 512
// features per pair: ONE kernel, one CTA per frame pair, persistent over all pyramid levels and all optimiser
// iterations (optimizeGN / optimizeLM, src/optimizer.cpp:41-370).  Thread f owns feature f.
//
//   level set-up   computeJacobian (:69-192) fused into the kernel: every thread reads the (P+3)^2 reference bytes of its
//                  feature, builds the (P+2)^2 bilinear template grid in FP64 and keeps, feature-minor in shared memory,
//                  the template T and twice the central differences (gx, gy) of every patch pixel, the 2x6 image
//                  Jacobian (computeImageJac, :194-248) and the world point.  Nothing is staged through global memory.
//   evaluation     computeResiduals (:251-370) + tukeyWeighting + normal equations:
//       warp       p_cur = R p_W + t, project, scale (FP64)
//       sample     the (P+1)^2 footprint of the current image lives in REGISTERS (8-byte row windows, one spare row and
//                  column on either side), re-fetched from L2 only when the feature leaves them; bilinear in FP32
//       sigma      1.4826 MAD: exact median and median absolute deviation, no shared-memory atomics per key (select5.cuh)
//       reduce     per-feature patch sums sxx sxy syy bx by chi2 -> 28 entries of J^T W J, J^T W r, chi2 through the
//                  factorisation J_row = gx A + gy B; transposed warp-shuffle reduction, FP64 across warps
//       solve      damping, LDLT 6x6 in registers, pose <- pose exp(-dx) in FP64 by one thread; accept / reject on device
#include <float.h>
#include <stdlib.h>

#include "align_common.cuh"
#include "select5.cuh"

namespace {

constexpr unsigned FULL5 = 0xffffffffu;

template <int P>
struct PatchGeo5 {
    static constexpr int PB   = -(P / 2);  // first offset; odd P: -half..half, even P: -P/2..P/2-1
    static constexpr int AREA = P * P;
    static constexpr int FW   = P + 1;     // footprint of the bilinear taps of a patch
    static constexpr int GW   = P + 2;     // template grid: patch plus one ring for the central differences
};

struct V5Args {
    ArenaView view;
    const svo_align_job* jobs;
    const svo_align_feature* feats;
    svo_align_result* results;
    svo_align_level_stats* stats;  // nullable
    svo_align_params prm;
    double K[4];
    int force;       // selection tier forced by SVO_S5_FORCE (tests): 0 none, 1 no predictions, 2 generic only
    long long* dbg;  // nullable: per-evaluation trace of job 0 (svo_debug_cycles): [0] evaluations, then sigma bits, tier | nvis << 8
};

// bytes p[0..7] of an arbitrarily aligned address as two words; never touches memory beyond the aligned 8 bytes that
// hold p[7]
__device__ __forceinline__ void load8(const uint8_t* p, uint32_t& lo, uint32_t& hi)
{
    const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
    const uint2* p8    = reinterpret_cast<const uint2*>(ad & ~uintptr_t(7));
    const uint32_t sh  = (uint32_t)(ad & 7u);
    const uint2 w0     = __ldg(p8);
    uint2 w1           = make_uint2(0u, 0u);
    if (sh) w1 = __ldg(p8 + 1);
    const uint32_t s4 = sh & 3u;
    uint32_t a0 = w0.x, a1 = w0.y, a2 = w1.x, a3 = w1.y;
    if (sh & 4u) {
        a0 = a1;
        a1 = a2;
        a2 = a3;
    }
    lo = __funnelshift_r(a0, a1, s4 * 8u);
    hi = __funnelshift_r(a1, a2, s4 * 8u);
}

// dx = (H + diag_add I)^-1 g, E = (21 H upper triangle, 6 g) in shared memory: copied to registers, register-resident LDLT,
// pivoted Eigen-rule fallback (out of line) for semi-definite systems
__device__ __forceinline__ void solve6_inline(const double* E, double diag_add, double* dx)
{
    double e[27];
#pragma unroll
    for (int i = 0; i < 27; i++) e[i] = E[i];
    if (svo::ldlt6_nopivot(e, diag_add, e + 21, dx)) return;
    solve6(E, diag_add, dx);
}

template <int P, int NT>
__host__ __device__ constexpr size_t v5_smem_bytes()
{
    using G = PatchGeo5<P>;
    size_t b = 0;
    b += (size_t)G::AREA * NT * 4;      // T
    b += (size_t)G::AREA * NT * 8;      // (gx, gy)
    b += (size_t)12 * NT * 4;           // image Jacobian rows A, B
    b += (size_t)3 * NT * 8;            // world points
    b += s5_smem_words<NT>() * 4;       // selection
    b = (b + 15) & ~size_t(15);
    b += (size_t)(NT / 32) * 32 * 8;    // red
    b += sizeof(Ctrl) + 64;
    return b + 256;
}

template <int P, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_align_v5(const V5Args a)
{
    using G          = PatchGeo5<P>;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int job = blockIdx.x;
    const svo_align_job* J = a.jobs + job;
    const int F       = J->n_ref + J->n_kf;
    const int nLevels = a.prm.max_level - a.prm.min_level + 1;
    svo_align_level_stats* statsOut = a.stats;

    // ---- shared memory carve-up ----
    unsigned char* sp = smem_raw;
    float* Tm         = reinterpret_cast<float*>(sp);  // [AREA][NT]
    sp += (size_t)G::AREA * NT * 4;
    float2* Gm = reinterpret_cast<float2*>(sp);        // [AREA][NT]
    sp += (size_t)G::AREA * NT * 8;
    float* ABm = reinterpret_cast<float*>(sp);         // [12][NT] rows of the 2x6 image Jacobian (read once per evaluation)
    sp += (size_t)12 * NT * 4;
    double* PWm = reinterpret_cast<double*>(sp);       // [3][NT] world points
    sp += (size_t)3 * NT * 8;
    S5Smem<NT, G::AREA * NT * 12 + 12 * NT * 4 + 3 * NT * 8> sel;  // (views derived from the shared-memory symbol: LDS / STS / ATOMS, never generic)
    sp += (s5_smem_words<NT>() * 4 + 15) & ~size_t(15);
    double* red = reinterpret_cast<double*>(sp);       // [NW][32]
    sp += (size_t)NW * 32 * 8;
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(sp);

    if (J->n_ref == 0) {  // src/image_alignment.cpp:27-28
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
    sel.clear();
    sel.rp = 0;
#ifdef SVO_PROFILE
    if (a.dbg && job == 0 && tid == 0)
        for (int i = 48; i < 56; i++) a.dbg[i] = 0;
#endif
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
#ifndef SVO_PROFILE
        if (a.dbg && job == 0) a.dbg[0] = 0;
#endif
    }

    // ---- world point of this thread's feature: p_W = T_frame^-1 (bearing |P - C_frame|), src/image_alignment.cpp:153-155 ----
    const int f = tid;
    const svo_align_feature* ft = a.feats + J->feat_offset + f;
    bool hasPoint = false;
    double pwx = 0, pwy = 0, pwz = 1, ftu = 0, ftv = 0;
    if (f < F && ft->has_point) {
        hasPoint         = true;
        const double* Tp = f < J->n_ref ? J->T_ref : J->T_kf;
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
        double pW[3];
        svo::pose_inv_act(Tf, pC, pW);
        pwx = pW[0], pwy = pW[1], pwz = pW[2];
        ftu = ft->px[0], ftv = ft->px[1];
    }
    PWm[f] = pwx, PWm[NT + f] = pwy, PWm[2 * NT + f] = pwz;  // (own column: no barrier needed)
    const double fx0 = a.K[0], fy0 = a.K[1], cx0 = a.K[2], cy0 = a.K[3];
    const int border = P / 2 + 2;
    const int refSlot = f < J->n_ref ? J->ref_slot : J->kf_slot;
    uint32_t tierCount = 0;  // diagnostics, evaluations by selection route: prediction + count passes | no prediction << 8 | bisection << 16 | predicted brackets alone << 24
    int parity         = 0;
    S5Pred pred;
    s5_pred_init(pred);
    __syncthreads();  // selection buffers are zero, ctrl is initialised

#pragma unroll 1
    for (int level = a.prm.max_level, si = 0; level >= a.prm.min_level; level--, si++) {
        const int lw = a.view.w[level], lh = a.view.h[level], lpitch = a.view.pitch[level];
        const double scale    = 1.0 / (double)(1 << level);
        const uint8_t* curImg = a.view.img[level] + (long long)J->cur_slot * a.view.plane_stride[level];

#ifdef SVO_PROFILE
        const long long cs0 = clock64();
#endif
        // ================= level set-up: template, gradients, image Jacobian (computeJacobian) =================
        bool refvis = false;
        if (hasPoint) {
            const double u = ftu * scale, v = ftv * scale;
            const double uf = floor(u), vf = floor(v);
            const int uI = (int)uf, vI = (int)vf;
            if (!((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh)) {
                refvis = true;
                // computeImageJac at the WORLD point, scaled focal lengths
                const double denom = (double)(1 << level);
                const double fx = a.K[0] / denom, fy = a.K[1] / denom;
                const double x = pwx, y = pwy, z = pwz;
                // (ONE division: the rows are rounded to FP32, a reciprocal-multiply differs from the reference's divisions by
                // 1e-16 relative; ten FP64 divisions per thread were a fifth of the level set-up)
                const double iz = 1.0 / z, iz2 = iz * iz;
                const double x2 = x * x, y2 = y * y;
                float A[6], B[6];
                A[0] = (float)(fx * iz);
                A[1] = 0.f;
                A[2] = (float)(-(fx * x) * iz2);
                A[3] = (float)(-(fx * x * y) * iz2);
                A[4] = (float)((fx * x2) * iz2 + fx);
                A[5] = (float)(-(fx * y) * iz);
                B[0] = 0.f;
                B[1] = (float)(fy * iz);
                B[2] = (float)(-(fy * y) * iz2);
                B[3] = (float)(-(fy * y2) * iz2 - fy);
                B[4] = (float)((fy * x * y) * iz2);
                B[5] = (float)((fy * x) * iz);
#pragma unroll
                for (int i = 0; i < 6; i++) ABm[i * NT + f] = A[i], ABm[(6 + i) * NT + f] = B[i];
                // template grid: bilinear samples at (u + gx - 1 + PB, v + gy - 1 + PB), gx, gy in [0, GW); three grid
                // rows roll through registers, a patch row is emitted as soon as the row below it exists
                const uint8_t* img = a.view.img[level] + (long long)refSlot * a.view.plane_stride[level] +
                                     (long long)(vI + G::PB - 1) * lpitch + (uI + G::PB - 1);
                const double fu = u - uf, fv = v - vf, wu0 = 1.0 - fu, wv0 = 1.0 - fv;
                double prev[G::GW];
                float g0[G::GW], g1[G::GW], g2[G::GW];
#pragma unroll
                for (int cx = 0; cx < G::GW; cx++) prev[cx] = 0.0, g0[cx] = g1[cx] = g2[cx] = 0.f;
                // all source rows first (independent loads in flight together), then the arithmetic
                uint32_t rlo[G::GW + 1], rhi[G::GW + 1];
#pragma unroll
                for (int ry = 0; ry <= G::GW; ry++) load8(img + (long long)ry * lpitch, rlo[ry], rhi[ry]);  // GW + 1 <= 8 bytes of the source row
#pragma unroll
                for (int ry = 0; ry <= G::GW; ry++) {
                    const uint32_t lo = rlo[ry], hi = rhi[ry];
                    double cur[G::GW];
                    // byte -> double without the conversion unit: 2^52 + b has b in its low mantissa bits
                    double left = __hiloint2double(0x43300000, (int)(lo & 0xffu)) - 4503599627370496.0;
#pragma unroll
                    for (int cx = 0; cx < G::GW; cx++) {
                        const int c1        = cx + 1;
                        const uint32_t word = c1 < 4 ? lo : hi;
                        const double right  = __hiloint2double(0x43300000, (int)((word >> (8 * (c1 & 3))) & 0xffu)) - 4503599627370496.0;
                        cur[cx]             = wu0 * left + fu * right;  // src/algorithm.cpp:901-902
                        left                = right;
                    }
                    if (ry > 0) {
#pragma unroll
                        for (int cx = 0; cx < G::GW; cx++) {
                            g0[cx] = g1[cx];
                            g1[cx] = g2[cx];
                            g2[cx] = (float)(wv0 * prev[cx] + fv * cur[cx]);  // :903
                        }
                        const int gy = ry - 1;  // grid row just completed
                        if (gy >= 2) {
                            const int y = gy - 2;  // patch row: grid rows y, y + 1, y + 2 are g0, g1, g2
#pragma unroll
                            for (int xx = 0; xx < P; xx++) {
                                Tm[(size_t)(y * P + xx) * NT + f] = g1[xx + 1] * 65536.f;  // exact: the residuals are kept x 2^16
                                Gm[(size_t)(y * P + xx) * NT + f] = make_float2(g1[xx + 2] - g1[xx], g2[xx + 1] - g0[xx + 1]);
                            }
                        }
                    }
#pragma unroll
                    for (int cx = 0; cx < G::GW; cx++) prev[cx] = cur[cx];
                }
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
        }
        __syncthreads();  // ctrl init visible
        pred.haveMove = false;  // the deviation changes with the level: no predicted bracket at its first evaluation

        // current-image window of this feature: FW + 1 rows x 8 bytes starting at column wx, row wy -- one column of slack on
        // either side and one row: a warp re-fetches (an L2 round trip for all its lanes) only when one of its features moves
        // by more than about half a pixel from where its window was loaded, not at every integer crossing
        uint32_t winLo[G::FW + 1], winHi[G::FW + 1];
        int wx = 0, wy = 0;
        bool winValid = false;

        // ================= evaluate (computeResiduals + tukeyWeighting + normal equations) =================
        const bool gn       = a.prm.mode == SVO_GN;
        const bool faithful = a.prm.mode == SVO_LM_FAITHFUL;
        const int maxIter   = a.prm.max_iter > 0 ? a.prm.max_iter : 20;
        bool lmFirst        = true;  // thread 0 only
        int evalInLevel     = 0;
#ifdef SVO_PROFILE
        long long* const tph = reinterpret_cast<long long*>(sel.prof()) + 16;  // (thread 33 only; zero at the start of a level)
        if (tid == 33) tph[5] += clock64() - cs0;  // level set-up including its barrier
#define V5_TICK(var) const long long var = clock64()
#else
#define V5_TICK(var)
#endif
#pragma unroll 1
        while (true) {
            V5_TICK(c0);
            // --- warp the feature (FP64) ---
            bool vis = false;
            int uI = 0, vI = 0;
            float fu = 0.f, fv = 0.f;
            if (refvis) {
                const double* R = ctrl->R;
                const double pwx = PWm[f], pwy = PWm[NT + f], pwz = PWm[2 * NT + f];
                const double cxp = R[0] * pwx + R[1] * pwy + R[2] * pwz + ctrl->t[0];
                const double cyp = R[3] * pwx + R[4] * pwy + R[5] * pwz + ctrl->t[1];
                const double czp = R[6] * pwx + R[7] * pwy + R[8] * pwz + ctrl->t[2];
                // PinholeCamera::project2d, src/pinhole_camera.cpp:55-56 (no z > 0 test)
                // (two true divisions, as the reference: with an identity prior the features project onto INTEGER pixels, and
                // a reciprocal-multiply that is 1 ulp off flips floor() and with it the border test)
                const double u = (fx0 * (cxp / czp) + cx0) * scale;
                const double v = (fy0 * (cyp / czp) + cy0) * scale;
                if (isfinite(u) && isfinite(v) && fabs(u) < 1e6 && fabs(v) < 1e6) {
                    const double uf = floor(u), vf = floor(v);
                    uI = (int)uf;
                    vI = (int)vf;
                    if (!((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh)) {
                        vis = true;
                        fu  = (float)(u - uf);
                        fv  = (float)(v - vf);
                    }
                }
            }
            // --- sample: refresh the register window if the footprint left it, bilinear taps, residuals x 2^16 ---
            float rs[G::AREA];
            if (vis) {
                const int x0 = uI + G::PB, y0 = vI + G::PB;  // footprint origin
                if (!winValid || (y0 != wy && y0 != wy + 1) || x0 < wx || x0 + G::FW > wx + 8) {
                    wx       = x0 - 1;
                    wy       = fv >= 0.5f ? y0 : y0 - 1;  // (the spare row on the side the feature is closer to; both are inside the image)
                    winValid = true;
#pragma unroll
                    for (int r = 0; r <= G::FW; r++) load8(curImg + (long long)(wy + r) * lpitch + wx, winLo[r], winHi[r]);
                }
                const bool down = y0 != wy;  // the footprint starts at window row 1
                const uint32_t off = (uint32_t)(x0 - wx);  // 0 .. 8 - FW
                const float wu0 = 1.f - fu;
                const float wv0s = (1.f - fv) * 65536.f, fvs = fv * 65536.f;  // vertical weights x 2^16 (exact scaling)
                float prevRow[P];
#pragma unroll
                for (int r = 0; r < G::FW; r++) {
                    const uint32_t rl = down ? winLo[r + 1] : winLo[r], rh = down ? winHi[r + 1] : winHi[r];
                    const uint32_t lo = __funnelshift_r(rl, rh, off * 8u);
                    const uint32_t hi = rh >> (off * 8u);
                    float px[G::FW];
#pragma unroll
                    for (int c = 0; c < G::FW; c++)  // byte -> float without the conversion pipe: 0x4B0000bb is 2^23 + bb
                        px[c] = __uint_as_float(__byte_perm(c < 4 ? lo : hi, 0x4B000000u, 0x7650u | (uint32_t)(c & 3))) - 8388608.f;
                    float curRow[P];
#pragma unroll
                    for (int c = 0; c < P; c++) curRow[c] = wu0 * px[c] + fu * px[c + 1];
                    if (r > 0) {
#pragma unroll
                        for (int c = 0; c < P; c++) {
                            // 2^16 (I - T) in two fused multiply-adds
                            const float Ts       = Tm[(size_t)((r - 1) * P + c) * NT + f];
                            rs[(r - 1) * P + c] = fmaf(fvs, curRow[c], fmaf(wv0s, prevRow[c], -Ts));
                        }
                    }
#pragma unroll
                    for (int c = 0; c < P; c++) prevRow[c] = curRow[c];
                }
            } else {
#pragma unroll
                for (int i = 0; i < G::AREA; i++) rs[i] = S5_SKIP;  // (above every range of the selection; never read by the sums)
            }

            V5_TICK(c1);
            // --- sigma = 1.4826 MAD, src/optimizer.cpp:485-507, src/algorithm.cpp:834-872 (MEDIAN_EXACT) ---
            double sigma = DBL_EPSILON;  // no visible pixel: the reference gets MAD = 0 from all-sentinel input
            int numValid = 0, lastWhy = 0;
            {
                double mad;
                uint32_t nvis;
                int tier;
                if (s5_sigma<G::AREA, NT>(rs, vis, F * G::AREA, pred, sel, parity, a.force, &mad, &nvis, &tier)) {
                    numValid = (int)nvis * G::AREA;
                    sigma    = 1.482602218505602 * mad * (1.0 / 65536.0);
                    if (sigma <= DBL_EPSILON) sigma = DBL_EPSILON;
                    tierCount += tier == 0 ? 0x1000000 : (tier == 1 ? 1 : (tier == 2 ? 0x100 : 0x10000));  // hit << 24 | miss | cold << 8 | generic << 16
                    lastWhy = sel.why & 0xff;
                    // Below the coarsest level the pose is already close: from the first to the second evaluation of a level
                    // the median and the deviation move by a few hundredths of an intensity unit (measured: < 0.05), while the
                    // jump ACROSS the level change, which `moved` holds now, says nothing about it.
                    if (evalInLevel == 0) {
                        pred.haveMove = si > 0;
                        pred.moved[0] = pred.moved[1] = 2622.f;  // (x 1.25 = 0.05 intensity units)
                    }
                }
                evalInLevel++;
            }
            V5_TICK(c2);
            const double cD  = 4.6851 * sigma;
            const float nic2 = (float)(-1.0 / (cD * cD * 4294967296.0));  // -1 / c^2 for residuals x 2^16

            // --- per-feature patch sums (FP32) and the feature's 28 contributions ---
            float val[32];
#pragma unroll
            for (int i = 0; i < 32; i++) val[i] = 0.f;
            if (vis) {
                float sxx = 0.f, sxy = 0.f, syy = 0.f, bx = 0.f, by = 0.f, ch = 0.f;
#pragma unroll
                for (int i = 0; i < G::AREA; i++) {
                    const float rr = rs[i];                   // 2^16 r: rescaled once per feature
                    const float2 g = Gm[(size_t)i * NT + f];  // twice the central differences: rescaled once per feature
                    const float t  = fmaxf(fmaf(rr * rr, nic2, 1.f), 0.f);  // Tukey: (1 - r^2 / c^2)^2 for |r| <= c, else 0
                    const float w  = t * t;
                    const float wgx = w * g.x, wgy = w * g.y, wr = w * rr;
                    sxx += wgx * g.x;
                    sxy += wgx * g.y;
                    syy += wgy * g.y;
                    bx += wr * g.x;
                    by += wr * g.y;
                    ch += wr * rr;
                }
                sxx *= 0.25f, sxy *= 0.25f, syy *= 0.25f, bx *= 0.5f / 65536.f, by *= 0.5f / 65536.f, ch *= 1.f / 4294967296.f;
                // J_row = gx A + gy B:  H += sxx A A^T + sxy (A B^T + B A^T) + syy B B^T = A u^T + B v^T
                float uu[6], vv[6], A[6], B[6];
#pragma unroll
                for (int i = 0; i < 6; i++) A[i] = ABm[i * NT + f], B[i] = ABm[(6 + i) * NT + f];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    uu[i] = sxx * A[i] + sxy * B[i];
                    vv[i] = sxy * A[i] + syy * B[i];
                }
                int kx = 0;
#pragma unroll
                for (int i = 0; i < 6; i++)
#pragma unroll
                    for (int j = i; j < 6; j++, kx++) val[kx] = A[i] * uu[j] + B[i] * vv[j];
#pragma unroll
                for (int i = 0; i < 6; i++) val[21 + i] = bx * A[i] + by * B[i];
                val[27] = ch;
            }
            V5_TICK(c3);
            // --- transposed warp reduction: afterwards lane i holds the warp total of val[i] ---
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const bool up = (lane & s) != 0;
#pragma unroll
                for (int j = 0; j < s; j++) {
                    const float send = up ? val[j] : val[j + s];
                    const float keep = up ? val[j + s] : val[j];
                    val[j]           = keep + __shfl_xor_sync(FULL5, send, s);
                }
            }
            red[warp * 32 + lane] = (double)val[0];
            __syncthreads();
            if (warp == 0) {
                if (lane < 28) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; w++) s += red[w * 32 + lane];
                    ctrl->E[lane] = s;
                }
                __syncwarp();
                if (lane == 0) {
#ifdef SVO_PROFILE
                    const long long ts0 = clock64();
#endif
                    ctrl->sigma  = sigma;
                    ctrl->n_eval = numValid;
                    ctrl->evals_level++;
#ifndef SVO_PROFILE
                    if (a.dbg && job == 0) {
                        const long long n = a.dbg[0];
                        if (n < 31) {
                            a.dbg[1 + 2 * n] = __double_as_longlong(sigma);
                            a.dbg[2 + 2 * n] = (long long)(tierCount & 0xffffffu) | ((long long)(lastWhy & 0xff) << 24) | ((long long)numValid << 32) | ((long long)(tierCount >> 24) << 56);
                            a.dbg[0]         = n + 1;
                        }
                    }
#endif
                    // the control flow of optimizeGN (src/optimizer.cpp:41-159) and optimizeLM (:161-370) as a state
                    // machine after every evaluation
                    auto record_first = [&](const double* E, double chi2, double lambda, int n) {
                        if (!ctrl->first) return;
                        ctrl->first = 0;
                        if (statsOut) {
                            svo_align_level_stats* s = statsOut + (size_t)job * nLevels + si;
                            double H[36], g[6];
                            expand_H(E, H, g);
                            for (int i = 0; i < 36; i++) s->H[i] = H[i];
                            for (int i = 0; i < 6; i++) s->g[i] = g[i];
                            s->chi2   = chi2;
                            s->sigma  = ctrl->first_sigma;
                            s->lambda = lambda;
                            s->n_px   = n;
                        }
                    };
                    double dx[6];
                    if (gn) {
                        const double chi2 = ctrl->E[27];
                        if (ctrl->first) ctrl->first_sigma = ctrl->sigma;
                        const bool wasFirst = ctrl->first;
                        record_first(ctrl->E, chi2, 0.0, ctrl->n_eval);
#ifdef SVO_PROFILE
                        const long long tq0 = clock64();
#endif
                        solve6_inline(ctrl->E, 0.0, dx);
#ifdef SVO_PROFILE
                        if (a.dbg && job == 0) a.dbg[52] += clock64() - tq0;
#endif
                        if (wasFirst && statsOut)
                            for (int i = 0; i < 6; i++) statsOut[(size_t)job * nLevels + si].dx[i] = dx[i];
                        ctrl->iters_level++;
                        double mx = dx[0];
                        bool nan  = false;
                        for (int i = 0; i < 6; i++) {
                            mx = fmax(mx, dx[i]);
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
#ifdef SVO_PROFILE
                            const long long tq1 = clock64();
#endif
                            svo::pose_update_right_exp_neg_fast(ctrl->pose, dx);
#ifdef SVO_PROFILE
                            if (a.dbg && job == 0) a.dbg[53] += clock64() - tq1;
#endif
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
                        ctrl->rmse = sqrt(chi2 / (double)ctrl->n_eval);
                    } else {
                        if (lmFirst) {  // the evaluation before the loop, :199-206
                            lmFirst = false;
                            for (int i = 0; i < 28; i++) ctrl->curE[i] = ctrl->E[i];
                            ctrl->cur_n       = ctrl->n_eval;
                            ctrl->first_sigma = ctrl->sigma;
                        } else {  // the re-evaluation after a step, :338-360
                            const bool ok = svo::nielsen_update(ctrl->preChi2, ctrl->E[27], ctrl->lambda, ctrl->nu);
                            ctrl->success = ok;
                            if (ok) {
                                for (int i = 0; i < 28; i++) ctrl->curE[i] = ctrl->E[i];
                                ctrl->cur_n = ctrl->n_eval;
                            } else {
                                ctrl->pose = ctrl->pre_pose;  // rollback
                            }
                            ctrl->it++;
                            if (ctrl->it >= maxIter) ctrl->done = 1;
                        }
                        if (!ctrl->done) {
                            if (ctrl->success) {  // :224-233 snapshot
                                ctrl->pre_pose = ctrl->pose;
                                ctrl->preChi2  = ctrl->curE[27];
                                ctrl->status   = SVO_ST_SUCCESS;
                            }
                            if (ctrl->it == 0) {
                                // diagonal entries of the packed upper triangle: 0, 6, 11, 15, 18, 20
                                const double* Ec = ctrl->curE;
                                const double mx  = fmax(fmax(fmax(Ec[0], Ec[6]), fmax(Ec[11], Ec[15])), fmax(Ec[18], Ec[20]));
                                ctrl->lambda *= mx;  // :296-299
                            }
                            const double lambda = ctrl->lambda;
                            const bool wasFirst = ctrl->first;
                            record_first(ctrl->curE, ctrl->curE[27], lambda, ctrl->cur_n);
                            solve6_inline(ctrl->curE, lambda, dx);
                            if (wasFirst && statsOut)
                                for (int i = 0; i < 6; i++) statsOut[(size_t)job * nLevels + si].dx[i] = dx[i];
                            svo::pose_update_right_exp_neg_fast(ctrl->pose, dx);  // :310 applied before any check
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
                        }
                        if (ctrl->done) ctrl->rmse = sqrt(ctrl->curE[27] / (double)ctrl->cur_n);
                    }
#ifdef SVO_PROFILE
                    const long long tq2 = clock64();
#endif
                    set_Rt(ctrl);
#ifdef SVO_PROFILE
                    if (a.dbg && job == 0) a.dbg[54] += clock64() - tq2;
                    if (a.dbg && job == 0) a.dbg[48 + si] += clock64() - ts0;
#endif
                }
            }
            __syncthreads();
#ifdef SVO_PROFILE
            {
                const long long c4 = clock64();
                if (tid == 33) {
                tph[0] += c1 - c0;  // warp + window + bilinear + residual (+ the wait for the previous solve)
                tph[1] += c2 - c1;  // sigma
                tph[3] += c3 - c2;  // weights + patch sums + expansion
                tph[4] += c4 - c3;  // reductions (the solve shows up in the next [0])
                tph[6] += 1;
                }
            }
#endif
            if (ctrl->done) break;
        }
#ifdef SVO_PROFILE
        if (a.dbg && job == 0 && tid == 33) {
            for (int i = 0; i < 8; i++) a.dbg[si * 8 + i] = tph[i], tph[i] = 0;
            for (int i = 0; i < 16; i++) a.dbg[32 + i] = reinterpret_cast<long long*>(sel.prof())[i];  // selection phases, summed over the levels so far
        }
#endif
        if (tid == 0) {
            ctrl->evals_total += ctrl->evals_level;
            ctrl->iters_total += ctrl->iters_level;
            if (statsOut) {
                svo_align_level_stats* s = statsOut + (size_t)job * nLevels + si;
                for (int i = 0; i < 4; i++) s->pose_after[i] = ctrl->pose.q[i];
                for (int i = 0; i < 3; i++) s->pose_after[4 + i] = ctrl->pose.t[i];
                s->rmse        = ctrl->rmse;
                s->status      = ctrl->status;
                s->iterations  = ctrl->iters_level;
                s->evaluations = ctrl->evals_level;
            }
        }
        __syncthreads();  // the level's bookkeeping is done before thread 0 re-initialises ctrl
    }
    if (tid == 0) {
        svo_align_result res;
        for (int i = 0; i < 4; i++) res.T_cur[i] = ctrl->pose.q[i];
        for (int i = 0; i < 3; i++) res.T_cur[4 + i] = ctrl->pose.t[i];
        res.rmse        = ctrl->rmse;
        res.status      = ctrl->status;
        res.evaluations = ctrl->evals_total;
        res.iterations  = ctrl->iters_total;
        res.reserved    = (int32_t)tierCount;
        a.results[job]  = res;
    }
}

template <int P, int NT>
svo_status launch_v5_nt(svo_ctx* ctx, const V5Args& args, int nJobs)
{
    const size_t smem = v5_smem_bytes<P, NT>();
    SVO_CUDA(cudaFuncSetAttribute(k_align_v5<P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_align_v5<P, NT><<<nJobs, NT, smem, ctx->stream>>>(args);
    return SVO_OK;
}

template <int P>
svo_status launch_v5(svo_ctx* ctx, int maxF)
{
    V5Args args;
    args.view    = make_view(ctx->arena);
    args.jobs    = ctx->d_jobs;
    args.feats   = ctx->d_feats;
    args.results = ctx->d_results;
    args.stats   = ctx->staged_want_stats ? ctx->d_stats : nullptr;
    args.prm     = ctx->staged_params;
    // the per-evaluation trace of job 0 (svo_debug_cycles) costs that job a global-memory round trip per evaluation on the serial
    // path of its solver: only on request (probes set SVO_ALIGN_TRACE=1; the instrumented twin of the library always traces)
#ifdef SVO_PROFILE
    args.dbg = ctx->d_dbg;
#else
    {
        const char* te = getenv("SVO_ALIGN_TRACE");
        args.dbg       = (te && te[0] == '1') ? ctx->d_dbg : nullptr;
    }
#endif
    {
        const char* fe = getenv("SVO_S5_FORCE");
        args.force     = fe ? atoi(fe) : 0;
    }
    for (int i = 0; i < 4; i++) args.K[i] = ctx->cfg.K[i];
    int nt = 64;
    while (nt < 512 && nt < maxF) nt *= 2;
    const int nJobs = ctx->staged_jobs;
    svo_status st;
    switch (nt) {
        case 64: st = launch_v5_nt<P, 64>(ctx, args, nJobs); break;
        case 128: st = launch_v5_nt<P, 128>(ctx, args, nJobs); break;
        case 256: st = launch_v5_nt<P, 256>(ctx, args, nJobs); break;
        default: st = launch_v5_nt<P, 512>(ctx, args, nJobs); break;
    }
    if (st != SVO_OK) return st;
    ctx->launches += 1;
    ctx->last_align_nt = nt;
    ctx->last_align_c  = 1;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}

}  // namespace

// the single-CTA fast path handles patch 4 / 5 and at most 512 features per pair
bool sparse_align_v5_supported(const svo_ctx* ctx, int maxF)
{
    const svo_align_params& prm = ctx->staged_params;
    return (prm.patch_size == 4 || prm.patch_size == 5) && maxF <= 512;
}

svo_status launch_sparse_align_v5(svo_ctx* ctx, int maxF)
{
    return ctx->staged_params.patch_size == 5 ? launch_v5<5>(ctx, maxF) : launch_v5<4>(ctx, maxF);
}
