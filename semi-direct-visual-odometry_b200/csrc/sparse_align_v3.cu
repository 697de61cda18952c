// sparse_align_v3.cu -- fast path of ImageAlignment::align (src/image_alignment.cpp:25-67) for patch sizes 4 / 5: one
// THREAD-BLOCK CLUSTER per frame pair.  Two kernels:
//
// k_align_precompute<P>  (computeJacobian, :69-192, for ALL levels at once; one thread per job x level x feature)
//     FP64: world point p_W = T_frame^-1 (bearing |P - C|), visibility at the level, the 2x6 image Jacobian rows A, B
//     (computeImageJac, :194-248) and the (P+2)^2 grid of bilinear template samples around the feature, from which
//     T (centre P x P), gx and gy (central differences) follow.  Written feature-minor, one contiguous block per
//     (CTA of the cluster, level), so that the iterate kernel stages a level with ONE bulk async copy.
//
// k_align_cluster<P, NT> (optimizeLM / optimizeGN, src/optimizer.cpp:41-370, over all levels; a cluster of C CTAs of
//                         NT threads per pair, persistent: never returns to the host between iterations)
//     CTA r of the cluster owns features [r NT, (r+1) NT), thread f owns one feature.  C x NT >= features: 500
//     features run as 2 x 256 (two CTAs of DIFFERENT pairs share an SM, so the serial phases of one pair overlap the
//     parallel phases of another), 1,000 features as 4 x 256, a single latency-critical pair as 8 x 64 on 8 SMs.
//     Per evaluation (computeResiduals, :251-370 + tukeyWeighting + normal equations):
//       warp      p_cur = R p_W + t, project, scale (FP64)
//       sample    the (P+1)^2 footprint of the current image lives in REGISTERS (8-byte row windows) and is re-fetched
//                 from L2 only when the feature's integer position leaves the window; bilinear in FP32
//       residual  r = I - T, kept in registers as fixed point q = rint(r 2^16) (|r| <= 255 is exact in 25 bits)
//       sigma     1.4826 MAD by two exact order statistics over the keys of the whole cluster (cluster_select.cuh:
//                 per-CTA shared-memory histograms combined through distributed shared memory, one cluster barrier
//                 per sweep)
//       reduce    per-feature patch sums sxx sxy syy bx by chi2 -> 28 entries of J^T W J, J^T W r, chi2 through the
//                 factorisation J_row = gx A + gy B; transposed warp-shuffle reduction, FP64 across warps, then across
//                 the CTAs of the cluster in a fixed order (every CTA gets bit-identical sums)
//       solve     damping, LDLT 6x6, pose <- pose exp(-dx) in FP64, redundantly in every CTA (no broadcast); accept /
//                 reject on device
#include <float.h>
#include <stdlib.h>

#include "align_common.cuh"
#include "cluster_select.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int P>
struct PatchGeo {
    static constexpr int PB   = -(P / 2);  // first offset; odd P: -half..half, even P: -P/2..P/2-1
    static constexpr int AREA = P * P;
    static constexpr int FW   = P + 1;     // footprint of the bilinear taps of a patch
    static constexpr int GW   = P + 2;     // template grid: patch plus one ring for the central differences
    static constexpr int GA   = GW * GW;
    static constexpr int ROWW = GA + 12 + 1;  // scratch words per feature and level: grid, A, B, flag
};

struct V3Args {
    ArenaView view;
    const svo_align_job* jobs;
    const svo_align_feature* feats;
    svo_align_result* results;
    svo_align_level_stats* stats;  // nullable
    float* scratch;                // [job][chunk = CTA of the cluster][ pW: 3*NT doubles | level blocks: ROWW*NT floats each ]
    long long job_stride;          // floats per job in scratch
    long long chunk_stride;        // floats per chunk
    int NT;                        // threads (= features) per CTA
    int C;                         // CTAs per cluster (= chunks per job)
    svo_align_params prm;
    double K[4];
    long long* dbg;                // nullable: per-phase cycle counters of job 0 (svo_debug_cycles)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------------------
// precompute
// ------------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(128) k_align_precompute(const V3Args a)
{
    using G = PatchGeo<P>;
    const int f   = blockIdx.x * blockDim.x + threadIdx.x;  // feature slot of the job
    const int si  = blockIdx.y;                             // 0 = coarsest
    const int job = blockIdx.z;
    const int NT  = a.NT;
    if (f >= a.C * NT) return;
    const int level        = a.prm.max_level - si;
    const svo_align_job* J = a.jobs + job;
    const int nRef = J->n_ref, F = J->n_ref + J->n_kf;
    const int fl   = f % NT;  // column inside the chunk
    float* base    = a.scratch + (long long)job * a.job_stride + (long long)(f / NT) * a.chunk_stride;
    double* pWout  = reinterpret_cast<double*>(base);
    float* blk     = base + 6 * NT + (long long)si * G::ROWW * NT;
    uint32_t flag  = 0;
    if (f < F) {
        const svo_align_feature* ft = a.feats + J->feat_offset + f;
        if (ft->has_point) {
            // p_W = T_frame^-1 (bearing * |P - C_frame|)   src/image_alignment.cpp:153-155
            const double* Tp = f < nRef ? J->T_ref : J->T_kf;
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
            if (si == 0) {
                pWout[0 * NT + fl] = pW[0];
                pWout[1 * NT + fl] = pW[1];
                pWout[2 * NT + fl] = pW[2];
            }
            const int lw = a.view.w[level], lh = a.view.h[level], lpitch = a.view.pitch[level];
            const double denom = (double)(1 << level);
            const double scale = 1.0 / denom;
            const int border   = P / 2 + 2;
            const double u = ft->px[0] * scale, v = ft->px[1] * scale;
            const double uf = floor(u), vf = floor(v);
            const int uI = (int)uf, vI = (int)vf;
            if (!((uI - border) < 0 || (vI - border) < 0 || (uI + border) >= lw || (vI + border) >= lh)) {
                flag = 1;
                // computeImageJac at the WORLD point, scaled focal lengths
                const double fx = a.K[0] / denom, fy = a.K[1] / denom;
                const double x = pW[0], y = pW[1], z = pW[2];
                const double x2 = x * x, y2 = y * y, z2 = z * z;
                float* jf = blk + G::GA * NT + fl;
                jf[0 * NT]  = (float)(fx / z);
                jf[1 * NT]  = 0.f;
                jf[2 * NT]  = (float)(-(fx * x) / z2);
                jf[3 * NT]  = (float)(-(fx * x * y) / z2);
                jf[4 * NT]  = (float)((fx * x2) / z2 + fx);
                jf[5 * NT]  = (float)(-(fx * y) / z);
                jf[6 * NT]  = 0.f;
                jf[7 * NT]  = (float)(fy / z);
                jf[8 * NT]  = (float)(-(fy * y) / z2);
                jf[9 * NT]  = (float)(-(fy * y2) / z2 - fy);
                jf[10 * NT] = (float)((fy * x * y) / z2);
                jf[11 * NT] = (float)((fy * x) / z);
                // template grid: bilinear samples at (u + gx - 1 + PB, v + gy - 1 + PB), gx, gy in [0, GW)
                const uint8_t* img = a.view.img[level] +
                                     (long long)(f < nRef ? J->ref_slot : J->kf_slot) * a.view.plane_stride[level] +
                                     (long long)(vI + G::PB - 1) * lpitch + (uI + G::PB - 1);
                const double fu = u - uf, fv = v - vf, wu0 = 1.0 - fu, wv0 = 1.0 - fv;
                double prev[G::GW];  // horizontal interpolants of the previous source row
#pragma unroll
                for (int ry = 0; ry <= G::GW; ry++) {
                    double cur[G::GW];
                    double left = (double)__ldg(img + (long long)ry * lpitch);
#pragma unroll
                    for (int cx = 0; cx < G::GW; cx++) {
                        const double right = (double)__ldg(img + (long long)ry * lpitch + cx + 1);
                        cur[cx]            = wu0 * left + fu * right;  // src/algorithm.cpp:901-902
                        left               = right;
                    }
                    if (ry > 0) {
#pragma unroll
                        for (int cx = 0; cx < G::GW; cx++)
                            blk[((ry - 1) * G::GW + cx) * NT + fl] = (float)(wv0 * prev[cx] + fv * cur[cx]);  // :903
                    }
#pragma unroll
                    for (int cx = 0; cx < G::GW; cx++) prev[cx] = cur[cx];
                }
            }
        }
    }
    reinterpret_cast<uint32_t*>(blk)[(G::GA + 12) * NT + fl] = flag;
}

// ------------------------------------------------------------------------------------------------------------
// iterate
// ------------------------------------------------------------------------------------------------------------
template <int P, int NT>
__host__ __device__ constexpr size_t v3_smem_bytes()
{
    using G = PatchGeo<P>;
    size_t b = 0;
    b += (size_t)G::ROWW * NT * 4;    // level block (grid, jac, flags): one bulk copy
    b += (size_t)3 * NT * 8;          // pW
    b += cs_smem_bytes<NT>();         // selection rounds (cluster_select.cuh)
    b += (size_t)(NT / 32) * 32 * 8;  // red
    b += 2 * 32 * 8;                  // xE: this CTA's 28 sums, ping-pong, read by the whole cluster
    b += sizeof(Ctrl) + 64;
    return b + 1024;                  // alignment slack
}

template <int NT>
struct V3Occ {  // CTAs per SM the register file allows at 128 registers per thread
    static constexpr int MINB = NT >= 512 ? 1 : (NT == 256 ? 2 : 4);
};

template <int P, int NT>
__global__ void __launch_bounds__(NT, V3Occ<NT>::MINB) k_align_cluster(const V3Args a)
{
    using G          = PatchGeo<P>;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C        = a.C;
    const int job      = blockIdx.x / C;
    const int crank    = blockIdx.x % C;  // == %cluster_ctarank: clusters are C consecutive CTAs along x
    const svo_align_job* J = a.jobs + job;
    const int F       = J->n_ref + J->n_kf;
    const int nLevels = a.prm.max_level - a.prm.min_level + 1;
    svo_align_level_stats* statsOut = crank == 0 ? a.stats : nullptr;  // CTA 0 of the cluster reports

    // ---- shared memory carve-up ----
    unsigned char* sp = smem_raw;
    float* blk        = reinterpret_cast<float*>(sp);  // [ROWW][NT]
    sp += (size_t)G::ROWW * NT * 4;
    double* pWs = reinterpret_cast<double*>(sp);  // [3][NT]
    sp += (size_t)3 * NT * 8;
    SelCtx<NT> sc;
    sc.init(sp, C);
#ifdef SVO_PROFILE
    for (int i = 0; i < 12; i++) sc.prof[i] = 0;
    for (int i = 0; i < 20; i++) sc.stat[i] = 0;
    sc.tlast = clock64();
#endif
    sp += cs_smem_bytes<NT>();
    static_assert(((size_t)G::ROWW * NT * 4 + (size_t)3 * NT * 8 + cs_smem_bytes<NT>()) % 16 == 0, "carve-up keeps 16-byte alignment");
    // (no integer round trip of the pointer: the compiler keeps the shared address space and emits LDS / STS for
    // every access below instead of generic loads)
    double* red = reinterpret_cast<double*>(sp);  // [NW][32]
    sp += (size_t)NW * 32 * 8;
    double* xE = reinterpret_cast<double*>(sp);   // [2][32]
    sp += 2 * 32 * 8;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sp);
    sp += 16;
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(sp);
    int pe     = 0;  // which half of xE the next reduction uses

    const float* gridS    = blk;
    const float* jacS     = blk + (size_t)G::GA * NT;
    const uint32_t* flagS = reinterpret_cast<const uint32_t*>(blk + (size_t)(G::GA + 12) * NT);

    if (J->n_ref == 0) {  // src/image_alignment.cpp:27-28 (uniform over the cluster)
        if (tid == 0 && crank == 0) {
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

    const float* scr = a.scratch + (long long)job * a.job_stride + (long long)crank * a.chunk_stride;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    // the selection buffers of every CTA are zero before any CTA of the cluster reads them
    if (C > 1)
        cs_cluster_sync();
    else
        __syncthreads();
    uint32_t mphase = 0;
    // world points (level independent): bulk copy HBM -> shared
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(3 * NT * 8);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(pWs)),
                     "l"(scr), "r"(bytes), "r"(smem_u32(mbar))
                     : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(mbar)), "r"(mphase)
                         : "memory");
        mphase ^= 1;
    }
    const int f  = tid;               // column of this thread's feature inside the CTA's chunk
    const int fg = crank * NT + tid;  // feature slot of the job
    double pwx = 0, pwy = 0, pwz = 1;
    if (fg < F) {
        pwx = pWs[f];
        pwy = pWs[NT + f];
        pwz = pWs[2 * NT + f];
    }
    const double fx0 = a.K[0], fy0 = a.K[1], cx0 = a.K[2], cy0 = a.K[3];
    const int border = P / 2 + 2;
    uint32_t tierCount = 0;  // diagnostics: hot | cold << 8 | generic << 16 selections
    Bracket brMed{0u, 7, false, false}, brMad{0u, 7, false, false};

#pragma unroll 1
    for (int level = a.prm.max_level, si = 0; level >= a.prm.min_level; level--, si++) {
        const int lw = a.view.w[level], lh = a.view.h[level], lpitch = a.view.pitch[level];
        const double scale    = 1.0 / (double)(1 << level);
        const uint8_t* curImg = a.view.img[level] + (long long)J->cur_slot * a.view.plane_stride[level];

        // ---- the level's template block: one bulk async copy HBM -> shared ----
        __syncthreads();  // everyone is done with the previous level's block
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(G::ROWW * NT * 4);
            const float* src     = scr + 6 * NT + (long long)si * G::ROWW * NT;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(blk)),
                         "l"(src), "r"(bytes), "r"(smem_u32(mbar))
                         : "memory");
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
        {
            uint32_t done = 0;
            while (!done)
                asm volatile(
                    "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                    : "=r"(done)
                    : "r"(smem_u32(mbar)), "r"(mphase)
                    : "memory");
            mphase ^= 1;
        }
        __syncthreads();  // ctrl init visible
        const bool refvis = fg < F && flagS[f] != 0;

        // current-image window of this feature: FW rows x 8 bytes starting at column wx, row wy
        uint32_t winLo[G::FW], winHi[G::FW];
        int wx = 0, wy = 0;
        bool winValid = false;
        // selection brackets carried between evaluations.  Across a level change only the MEDIAN's is worth trying, and
        // only when the level above was iterated: measured on 12 pairs (tests/cuda/bracket_probe.py), the widest bracket
        // around the previous level's median holds the new one in 36 of 36 cases with GN, in 14 of 36 with the reference's
        // one step per level, and the MAD's in 2 of 36 / 0 of 36 (the residual scale changes with the level) -- a miss
        // costs a whole sweep before the cold tier starts.
        if (si == 0) {
            brMed = Bracket{0u, 7, false, false};
            brMad = Bracket{0u, 7, false, false};
        } else {
            brMed.valid = brMed.have && a.prm.mode != SVO_LM_FAITHFUL;
            brMad.valid = false;
            brMed.shift = brMad.shift = 7;
            brMed.have = brMad.have = false;
        }

        // ================= evaluate (computeResiduals + tukeyWeighting + normal equations) =================
#ifdef SVO_PROFILE
        long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define SVO_TICK(var) const long long var = clock64()
#else
#define SVO_TICK(var)
#endif
        auto evaluate = [&]() {
            SVO_TICK(c0);
            // --- warp the feature (FP64) ---
            bool vis = false;
            int uI = 0, vI = 0;
            float fu = 0.f, fv = 0.f;
            if (refvis) {
                const double* R = ctrl->R;
                const double cxp = R[0] * pwx + R[1] * pwy + R[2] * pwz + ctrl->t[0];
                const double cyp = R[3] * pwx + R[4] * pwy + R[5] * pwz + ctrl->t[1];
                const double czp = R[6] * pwx + R[7] * pwy + R[8] * pwz + ctrl->t[2];
                // PinholeCamera::project2d, src/pinhole_camera.cpp:55-56 (no z > 0 test)
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
            // --- refresh the register window if the footprint left it ---
            uint32_t key[G::AREA];  // residuals as order-preserving fixed-point keys: rint(r 2^16) + 2^25
            if (vis) {
                const int x0 = uI + G::PB, y0 = vI + G::PB;  // footprint origin
                if (!winValid || y0 != wy || x0 < wx || x0 + G::FW > wx + 8) {
                    wx       = x0 - 1;
                    wy       = y0;
                    winValid = true;
#pragma unroll
                    for (int r = 0; r < G::FW; r++) {
                        const uint8_t* p   = curImg + (long long)(wy + r) * lpitch + wx;
                        const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
                        const uint2* p8    = reinterpret_cast<const uint2*>(ad & ~uintptr_t(7));
                        const uint32_t sh  = (uint32_t)(ad & 7u);
                        const uint2 w0     = __ldg(p8);
                        uint2 w1           = make_uint2(0u, 0u);
                        if (sh) w1 = __ldg(p8 + 1);  // never touched when aligned: may lie past the plane
                        // bytes sh .. sh+7 of the 16-byte pair
                        const uint32_t s4 = sh & 3u;
                        uint32_t a0 = w0.x, a1 = w0.y, a2 = w1.x, a3 = w1.y;
                        if (sh & 4u) {
                            a0 = a1;
                            a1 = a2;
                            a2 = a3;
                        }
                        winLo[r] = __funnelshift_r(a0, a1, s4 * 8u);
                        winHi[r] = __funnelshift_r(a1, a2, s4 * 8u);
                    }
                }
                // --- bilinear taps (FP32) and residuals ---
                const uint32_t off = (uint32_t)(x0 - wx);  // 0 .. 8 - FW
                const float wu0 = 1.f - fu, wv0 = 1.f - fv;
                float prevRow[P];
#pragma unroll
                for (int r = 0; r < G::FW; r++) {
                    // FW bytes of the row starting at byte `off`
                    const uint32_t lo = __funnelshift_r(winLo[r], winHi[r], off * 8u);
                    const uint32_t hi = winHi[r] >> (off * 8u);
                    float px[G::FW];
#pragma unroll
                    for (int c = 0; c < G::FW; c++) {
                        // byte -> float without the conversion pipe: 0x4B0000bb is 2^23 + bb exactly
                        px[c] = __uint_as_float(__byte_perm(c < 4 ? lo : hi, 0x4B000000u, 0x7650u | (uint32_t)(c & 3))) - 8388608.f;
                    }
                    float curRow[P];
#pragma unroll
                    for (int c = 0; c < P; c++) curRow[c] = wu0 * px[c] + fu * px[c + 1];
                    if (r > 0) {
#pragma unroll
                        for (int c = 0; c < P; c++) {
                            const float val = wv0 * prevRow[c] + fv * curRow[c];
                            const float T   = gridS[(size_t)((r - 1 + 1) * G::GW + (c + 1)) * NT + f];
                            key[(r - 1) * P + c] = (uint32_t)(__float2int_rn((val - T) * 65536.f) + (1 << 25));
                        }
                    }
#pragma unroll
                    for (int c = 0; c < P; c++) prevRow[c] = curRow[c];
                }
            } else {
#pragma unroll
                for (int i = 0; i < G::AREA; i++) key[i] = 0;
            }
            // visible features of the cluster: summed at the first barrier of the median selection
            sc.add_aux((uint32_t)__popc(__ballot_sync(FULL, vis)));
            SVO_TICK(c1);
            const int N = F * G::AREA;

            // --- sigma = 1.4826 MAD, src/optimizer.cpp:485-507, src/algorithm.cpp:834-872 (MEDIAN_EXACT) ---
            double sigma     = DBL_EPSILON;  // no visible pixel: the reference gets MAD = 0 from all-sentinel input
            int med2         = 0;            // twice the median, fixed point
            uint32_t negMask = 0;            // sign of (2 q - med2) per patch pixel, to undo the MAD key transform below
            int mid          = 0;            // numValid / 2: element `mid` of the sorted N-vector (invalid rows sort last)
            int numValid     = 0;
            // ONE call site for both order statistics (median of r, then median of |r - med|): the selection code
            // is large and the kernel is bound by instruction fetch as much as by issue
#pragma unroll 1
            for (int s = 0; s < 2; s++) {
                Bracket b = s ? brMad : brMed;
#ifdef SVO_PROFILE
                const bool brMadTried = brMad.valid, brMedTried = brMed.valid;
#endif
                uint32_t kHi, kLo;
                int tier;
                // mean of elements mid-1 and mid when N is even: the selection returns both
                const bool ok = cs_tiered_select<G::AREA, NT>(key, vis, s ? 0 : 512 - 32, mid, s == 0, N, b, sc, &kHi, &kLo, &tier);
                if (s)
                    brMad = b;
                else
                    brMed = b;
#ifdef SVO_PROFILE
                if (si > 0 && ctrl->evals_level == 0 && (s ? brMadTried : brMedTried)) {  // first evaluation of a level: carried bracket
                    sc.stat[16 + 2 * s]++;
                    sc.stat[17 + 2 * s] += tier == 1;
                }
#endif
                if (!ok) break;  // no visible pixel (s == 0 only)
                tierCount += tier == 1 ? 1 : (tier == 2 ? 0x100 : 0x10000);
                if (s == 0) {
                    numValid = (int)sc.auxTotal * G::AREA;
                    med2     = ((int)kHi - (1 << 25)) + ((int)kLo - (1 << 25));
                    // keys in place: |2 q - med2|, deviations in half fixed-point units (2^-17)
#pragma unroll
                    for (int i = 0; i < G::AREA; i++) {
                        const int d = 2 * ((int)key[i] - (1 << 25)) - med2;
                        negMask |= (d < 0 ? 1u : 0u) << i;
                        key[i] = (uint32_t)abs(d);
                    }
                } else {
                    const double mad = 0.5 * ((double)kHi + (double)kLo);
                    sigma            = 1.482602218505602 * mad * (1.0 / 131072.0);
                    if (sigma <= DBL_EPSILON) sigma = DBL_EPSILON;
                }
            }
            SVO_TICK(c2);
            const double cD = 4.6851 * sigma;
            const float cF  = (float)cD;
            const float ic2 = (float)(1.0 / (cD * cD));

            // --- per-feature patch sums (FP32) and the feature's 28 contributions ---
            float val[32];
#pragma unroll
            for (int i = 0; i < 32; i++) val[i] = 0.f;
            if (vis) {
                float sxx = 0.f, sxy = 0.f, syy = 0.f, bx = 0.f, by = 0.f, ch = 0.f;
#pragma unroll
                for (int y = 0; y < P; y++) {
#pragma unroll
                    for (int x = 0; x < P; x++) {
                        // 2 q = med2 +/- |2 q - med2| (exact), r = q 2^-16
                        const int i2   = y * P + x;
                        const int q2   = med2 + (((negMask >> i2) & 1u) ? -(int)key[i2] : (int)key[i2]);
                        const float rr = (float)q2 * (1.f / 131072.f);
                        const float gl = gridS[(size_t)((y + 1) * G::GW + x) * NT + f];
                        const float gr = gridS[(size_t)((y + 1) * G::GW + x + 2) * NT + f];
                        const float gu = gridS[(size_t)(y * G::GW + x + 1) * NT + f];
                        const float gd = gridS[(size_t)((y + 2) * G::GW + x + 1) * NT + f];
                        const float gx = gr - gl, gy = gd - gu;  // twice the central differences: rescaled once per feature
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
                }
                sxx *= 0.25f, sxy *= 0.25f, syy *= 0.25f, bx *= 0.5f, by *= 0.5f;
                float A[6], B[6];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    A[i] = jacS[(size_t)i * NT + f];
                    B[i] = jacS[(size_t)(6 + i) * NT + f];
                }
                int k = 0;
#pragma unroll
                for (int i = 0; i < 6; i++)
#pragma unroll
                    for (int j = i; j < 6; j++, k++) val[k] = sxx * (A[i] * A[j]) + sxy * (A[i] * B[j] + B[i] * A[j]) + syy * (B[i] * B[j]);
#pragma unroll
                for (int i = 0; i < 6; i++) val[21 + i] = bx * A[i] + by * B[i];
                val[27] = ch;
            }
            SVO_TICK(c3);
            // --- transposed warp reduction: afterwards lane i holds the warp total of val[i] ---
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const bool up = (lane & s) != 0;
#pragma unroll
                for (int j = 0; j < s; j++) {
                    const float send = up ? val[j] : val[j + s];
                    const float keep = up ? val[j + s] : val[j];
                    val[j]           = keep + __shfl_xor_sync(FULL, send, s);
                }
            }
            red[warp * 32 + lane] = (double)val[0];
            __syncthreads();
            if (C == 1) {
                if (tid < 28) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; w++) s += red[w * 32 + tid];
                    ctrl->E[tid] = s;
                }
            } else {
                // this CTA's sums -> xE, then every CTA adds the C copies in the same order: bit-identical E everywhere
                if (tid < 28) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; w++) s += red[w * 32 + tid];
                    xE[pe * 32 + tid] = s;
                }
                cs_cluster_sync();
                if (tid < 28) {
                    ctrl->E[tid] = cs_gather_sum_f64(smem_u32(xE + pe * 32 + tid), C);
                }
                pe ^= 1;
            }
            if (tid == 32) {
                ctrl->sigma  = sigma;
                ctrl->n_eval = numValid;
                ctrl->evals_level++;
            }
            __syncthreads();
#ifdef SVO_PROFILE
            const long long c4 = clock64();
            tph[0] += c1 - c0;  // warp + window + bilinear + residual (+ the wait for the previous solve)
            tph[1] += c2 - c1;  // sigma: median + MAD selections
            tph[3] += c3 - c2;  // weights + patch sums + expansion
            tph[4] += c4 - c3;  // reductions
            tph[6] += 1;
#endif
        };

        // record the first iteration of the level for the stats record
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

        // One evaluate() call site for all three modes (keeps the per-thread register state in registers): the
        // control flow of optimizeGN (src/optimizer.cpp:41-159) and optimizeLM (:161-370) runs as a state machine
        // on thread 0 after every evaluation.
        const bool gn       = a.prm.mode == SVO_GN;
        const bool faithful = a.prm.mode == SVO_LM_FAITHFUL;
        const int maxIter   = a.prm.max_iter > 0 ? a.prm.max_iter : 20;
        bool lmFirst        = true;  // thread 0 only
#pragma unroll 1
        while (true) {
            evaluate();
            if (tid == 0) {
                double dx[6];
                if (gn) {
                    const double chi2 = ctrl->E[27];
                    if (ctrl->first) ctrl->first_sigma = ctrl->sigma;
                    const bool wasFirst = ctrl->first;
                    record_first(ctrl->E, chi2, 0.0, ctrl->n_eval);
                    solve6(ctrl->E, 0.0, dx);
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
                        solve6(ctrl->curE, lambda, dx);
                        if (wasFirst && statsOut)
                            for (int i = 0; i < 6; i++) statsOut[(size_t)job * nLevels + si].dx[i] = dx[i];
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
                    }
                    if (ctrl->done) ctrl->rmse = sqrt(ctrl->curE[27] / (double)ctrl->cur_n);
                }
                set_Rt(ctrl);
            }
            __syncthreads();
            // (bar.sync does not block at issue: a clock read here shows the barrier's ISSUE time; the wait for the
            // single-thread solve is charged to the first dependent read of the next evaluation, phase [0])
            if (ctrl->done) break;
        }
#ifdef SVO_PROFILE
        if (a.dbg && job == 0 && crank == 0 && tid == 33) {
            for (int i = 0; i < 8; i++) a.dbg[si * 8 + i] = tph[i];
            for (int i = 0; i < 12; i++) a.dbg[32 + i] = sc.prof[i];  // selection phases, summed over the levels so far
            for (int i = 0; i < 20; i++) a.dbg[44 + i] = sc.stat[i];  // bracket statistics
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
    }
    __syncthreads();
    if (tid == 0 && crank == 0) {
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
    if (C > 1) cs_cluster_sync();  // no CTA leaves while a peer may still read its shared memory
}

template <int P, int NT>
svo_status launch_v3_nt(svo_ctx* ctx, const V3Args& args, int nJobs)
{
    const size_t smem = v3_smem_bytes<P, NT>();
    SVO_CUDA(cudaFuncSetAttribute(k_align_cluster<P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = dim3((unsigned)(nJobs * args.C), 1, 1);
    cfg.blockDim         = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id               = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)args.C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs              = at;
    cfg.numAttrs           = 1;
    SVO_CUDA(cudaLaunchKernelEx(&cfg, k_align_cluster<P, NT>, args));
    return SVO_OK;
}

// Shape of the launch: NT threads (= features) per CTA, C CTAs per pair.  Measured on B200 (profiles/): a cluster
// barrier with release / acquire semantics costs more than the work it spreads as long as one CTA can hold the pair,
// so C = 1 up to 512 features (the smallest CTA that covers them: more co-resident CTAs per SM), and clusters of 2 / 4
// CTAs of 512 threads beyond (1,000 features = 2 x 512).  SVO_ALIGN_NT / SVO_ALIGN_C override (experiments).
void v3_shape(const svo_ctx* ctx, int maxF, int nJobs, int* NT, int* C)
{
    (void)ctx;
    (void)nJobs;
    int c = (maxF + 511) / 512;
    if (c < 1) c = 1;
    if (c == 3) c = 4;
    if (c > 4) c = 8;
    int nt = 64;
    while (nt < 512 && nt * c < maxF) nt *= 2;
    if (const char* e = getenv("SVO_ALIGN_NT")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256 || v == 512) {
            nt = v;
            c  = (maxF + nt - 1) / nt;
            if (c < 1) c = 1;
            if (c == 3) c = 4;
            if (c > 4 && c < 8) c = 8;
        }
    }
    if (const char* e = getenv("SVO_ALIGN_C")) {
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4 || v == 8) && v * nt >= maxF) c = v;
    }
    *NT = nt;
    *C  = c;
}

template <int P>
svo_status launch_v3(svo_ctx* ctx, int maxF)
{
    using G                     = PatchGeo<P>;
    const svo_align_params& prm = ctx->staged_params;
    const int nJobs             = ctx->staged_jobs;
    const int nLevels           = prm.max_level - prm.min_level + 1;
    int NT, C;
    v3_shape(ctx, maxF, nJobs, &NT, &C);
    const long long chunk_stride = 6LL * NT + (long long)nLevels * G::ROWW * NT;  // floats
    const long long job_stride   = chunk_stride * C;
    const size_t need            = (size_t)job_stride * 4 * nJobs;
    if ((long long)NT * C < maxF) SVO_FAIL(SVO_ERR_CAPACITY, "sparse alignment: more than 4,096 features per pair");
    if (need > ctx->scratch2_bytes) {
        // the front-end graphs captured this pointer as a kernel argument: they are stale once it is freed
        frontend_invalidate_graphs(ctx);
        SVO_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_scratch2) SVO_CUDA(cudaFree(ctx->d_scratch2));
        ctx->d_scratch2     = nullptr;
        ctx->scratch2_bytes = 0;
        SVO_CUDA(cudaMalloc(&ctx->d_scratch2, need));
        ctx->scratch2_bytes = need;
    }
    V3Args args;
    args.view         = make_view(ctx->arena);
    args.jobs         = ctx->d_jobs;
    args.feats        = ctx->d_feats;
    args.results      = ctx->d_results;
    args.stats        = ctx->staged_want_stats ? ctx->d_stats : nullptr;
    args.scratch      = ctx->d_scratch2;
    args.job_stride   = job_stride;
    args.chunk_stride = chunk_stride;
    args.NT           = NT;
    args.C            = C;
    args.prm          = prm;
    args.dbg          = ctx->d_dbg;
    for (int i = 0; i < 4; i++) args.K[i] = ctx->cfg.K[i];

    dim3 pgrid((C * NT + 127) / 128, nLevels, nJobs);
    k_align_precompute<P><<<pgrid, 128, 0, ctx->stream>>>(args);
    svo_status st;
    switch (NT) {
        case 64: st = launch_v3_nt<P, 64>(ctx, args, nJobs); break;
        case 128: st = launch_v3_nt<P, 128>(ctx, args, nJobs); break;
        case 256: st = launch_v3_nt<P, 256>(ctx, args, nJobs); break;
        default: st = launch_v3_nt<P, 512>(ctx, args, nJobs); break;
    }
    if (st != SVO_OK) return st;
    ctx->launches += 2;
    ctx->last_align_nt = NT;
    ctx->last_align_c  = C;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}

}  // namespace

// returns true when the cluster fast path handles this batch: patch 4 / 5, at most 8 CTAs (4,096 features) per pair
bool sparse_align_v3_supported(const svo_ctx* ctx, int maxF)
{
    const svo_align_params& prm = ctx->staged_params;
    if (prm.patch_size != 4 && prm.patch_size != 5) return false;
    int NT, C;
    v3_shape(ctx, maxF, ctx->staged_jobs, &NT, &C);
    return C <= 8 && (long long)NT * C >= maxF;  // 8 CTAs of 512 threads hold 4,096 features
}

svo_status launch_sparse_align_v3(svo_ctx* ctx, int maxF)
{
    return ctx->staged_params.patch_size == 5 ? launch_v3<5>(ctx, maxF) : launch_v3<4>(ctx, maxF);
}
