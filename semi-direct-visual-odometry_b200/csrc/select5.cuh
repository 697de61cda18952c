// select5.cuh -- robust scale of the alignment (tukeyWeighting / computeSigma, src/optimizer.cpp:485-507, median rule
// src/algorithm.cpp:834-872, MEDIAN_EXACT of SURVEY 9.3) for ONE CTA of NT threads, each holding AREA residuals (FP32,
// scaled by 2^16) in registers.  Exact order statistics of those floats: the median first, then the median of
// |x - median|.  Invisible features hold S5_SKIP in every residual (above every range, so no pass needs a visibility test).
//
// Shared-memory atomics cost about two cycles per LANE on this SM (a histogram of 12,500 keys is 25,000 cycles), so no
// pass issues one per key.  Both selections are the same two passes over the registers, run by ONE copy of the code
// (a loop over the two phases), each followed by one block barrier:
//
//   bracket pass   for a range [lo, lo + W) believed to hold the target and few keys: count the keys below lo (sign bits,
//                  one warp reduction), append the keys inside to a list -- a warp ballot per key, the keys of a warp
//                  compacted into the warp's staging row, ONE atomic per warp reserves its part of the list.  The counts
//                  decide whether the target (and, for the even rule, its predecessor) is in the list; if so the list is
//                  ranked by counting with all threads.
//   count pass     16 bins over a range (14 equal inner bins, bin 0 / 15 hold everything below / above), counted per thread in
//                  4-bit fields of two 64-bit registers (keys 0..12 and 13..24: a field never exceeds 13), widened to
//                  16-bit pairs, one REDUX per pair and one 8-lane atomic per warp.  The bin(s) of the target become the
//                  next range: a bracket pass if they hold few keys, another count pass otherwise.
//
// Every pass recounts the keys below its own range and verifies by rank arithmetic that the target is inside, so a range
// derived in floating point from the previous pass needs no exact agreement with it -- only its margin.  The first range
// is the previous evaluation's value +- three times its last movement (bracket pass if that is expected to hold < 96 keys,
// from the density the last bracket pass measured), else a range around the last value scaled by the last deviation.
//
// Safety net ("generic"): bitwise bisection over the ordered-integer image of the floats, 33 counting passes, always right
// (heavy ties, ranges that keep missing).  tests/test_gpu_parity.py forces every tier through the C ABI (SVO_S5_FORCE) and
// compares with the oracle.
#pragma once
#include <stdint.h>

namespace {

#ifdef SVO_PROFILE
// phase cycle counters of the instrumented twin of the library: ONE thread (33) accumulates into shared memory, so the
// other threads carry no profiling state (an accumulator per thread in registers distorts what it measures)
#define S5_T(i)                                                               \
    do {                                                                      \
        if (threadIdx.x == 33) {                                              \
            const long long t__ = clock64();                                  \
            reinterpret_cast<long long*>(sm.prof())[i] += t__ - sm.tlast;     \
            sm.tlast = t__;                                                   \
        }                                                                     \
    } while (0)
#define S5_N(i)                                                               \
    do {                                                                      \
        if (threadIdx.x == 33) reinterpret_cast<long long*>(sm.prof())[i] += 1; \
    } while (0)
#else
#define S5_T(i)
#define S5_N(i)
#endif

constexpr unsigned S5_FULL = 0xffffffffu;
constexpr float S5_SKIP    = 3.0e38f;     // residual value of an invisible feature
constexpr int S5_DEPTH     = 8;           // key stack of a thread (a thread with DEPTH - 1 or more keys inside a bracket: overflow)
constexpr float S5_MAGIC   = 12582912.f;  // 1.5 * 2^23: x + MAGIC has rint(x) in its low mantissa bits
constexpr int S5_PASSES    = 12;          // passes a phase may take before the generic tier

template <int NT>
__host__ __device__ constexpr int s5_cap()  // list entries a bracket pass may hold
{
    return NT >= 256 ? 256 : NT;
}

template <int NT>
__host__ __device__ constexpr size_t s5_smem_words()
{
    return (size_t)3 * 16 + 8 + s5_cap<NT>() + (size_t)NT * S5_DEPTH + 64;
}

// Views derived from the `extern __shared__` symbol with compile-time offsets (LDS / STS / ATOMS, never generic).
template <int NT, int OFF>
struct S5Smem {
    int why;  // diagnostics of the last evaluation: passes of phase 0 | passes of phase 1 << 4
    int rp;   // parity of the result slots (they alternate from rank to rank; starts at 0)
#ifdef SVO_PROFILE
    long long tlast;  // (thread 33) prof(): cycles 0 bracket pass 1 rank 2 count pass 3 locate, 4 5 6 the barrier waits after 0 1 2, 11 generic,
                      // 16..23 the phases of the kernel's evaluation loop
#endif
    __device__ __forceinline__ static uint32_t* base()
    {
        extern __shared__ __align__(128) unsigned char s5_dynamic_smem[];
        return reinterpret_cast<uint32_t*>(s5_dynamic_smem + OFF);
    }
    // [3][16] counters of a pass, three rotating sets: 0 keys below 1 list fill 2 visible features 3 max below (generic)
    // 4 a key stack overflowed 5..7 (fused pass) pushed keys below the median bracket / inside it / deviation candidates;
    // 8..15 the 16 bin totals of a count pass as 16-bit pairs
    __device__ __forceinline__ static uint32_t* cnt() { return base(); }
    __device__ __forceinline__ static uint32_t* res() { return cnt() + 48; }               // [2][2] target, predecessor (ordered-integer image; atomicMax), two alternating sets
    __device__ __forceinline__ static uint32_t* list() { return res() + 8; }               // [cap]
    __device__ __forceinline__ static uint32_t* stack() { return list() + s5_cap<NT>(); }  // [DEPTH][NT]
    __device__ __forceinline__ static uint32_t* prof() { return stack() + NT * S5_DEPTH; }  // [32] 64-bit cycle counters (instrumented build)
    __device__ __forceinline__ static void clear()
    {
        uint32_t* b = base();
        for (int i = threadIdx.x; i < (int)s5_smem_words<NT>(); i += NT) b[i] = 0;
    }
};

struct S5Pred {      // carried from evaluation to evaluation (uniform over the CTA); key units = 2^-16 intensity
    float v[2];      // last median, last deviation
    float moved[2];  // how far each moved at the last evaluation
    float rho[2];    // keys per key unit around each target (measured by the last bracket pass)
    bool have;       // v[] holds values (of this level or the one above)
    bool haveMove;   // moved[] is meaningful for the next evaluation
};

__device__ __forceinline__ void s5_pred_init(S5Pred& p)
{
    p.v[0] = p.v[1] = 0.f, p.moved[0] = p.moved[1] = 0.f, p.rho[0] = p.rho[1] = 0.f;
    p.have = p.haveMove = false;
}

// order-preserving image of a float in the unsigned integers
__device__ __forceinline__ uint32_t s5_ord(float f)
{
    const uint32_t b = __float_as_uint(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float s5_unord(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// Element idx (and idx - 1 when needPred) of the n <= CAP floats of the list, by counting with all threads: G threads per
// entry count the entries strictly below it (16-byte loads); element idx is the LARGEST value with at most idx entries below
// it, found with one warp reduction and one atomicMax per warp.  One barrier.  The result slots alternate between calls.
template <int NT, class SM>
__device__ __forceinline__ void s5_rank(SM& sm, uint32_t n, uint32_t idx, bool needPred, float* hiOut, float* lwOut)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int lgN    = n <= 4u ? 2 : 32 - __clz(n - 1u);  // entries rounded up to a power of two (at least one load)
    const int lgG    = min(31 - __clz(NT) - lgN, 5);      // threads per entry: NT / entries, one warp at most
    const uint32_t G = 1u << lgG;
    const uint32_t e = (uint32_t)tid >> lgG, sub = (uint32_t)tid & (G - 1u);
    const float me   = __uint_as_float(sm.list()[min(e, n - 1u)]);
    float less       = 0.f;
    const uint4* l4  = reinterpret_cast<const uint4*>(sm.list());
    uint32_t j = sub;
    for (; 4u * j + 3u < n; j += G) {
        const uint4 x = l4[j];
        less += (__uint_as_float(x.x) < me ? 1.f : 0.f) + (__uint_as_float(x.y) < me ? 1.f : 0.f);
        less += (__uint_as_float(x.z) < me ? 1.f : 0.f) + (__uint_as_float(x.w) < me ? 1.f : 0.f);
    }
    for (uint32_t i = 4u * j; i < n; i++) less += __uint_as_float(sm.list()[i]) < me ? 1.f : 0.f;  // (the last, partial group)
    for (uint32_t o = G >> 1; o >= 1u; o >>= 1) less += __shfl_xor_sync(S5_FULL, less, o);
    const uint32_t below = (uint32_t)less;
    const bool live      = e < n && sub == 0;
    const uint32_t oHi   = __reduce_max_sync(S5_FULL, live && below <= idx ? s5_ord(me) : 0u);
    const uint32_t oLo   = __reduce_max_sync(S5_FULL, live && needPred && below + 1u <= idx ? s5_ord(me) : 0u);
    uint32_t* r          = sm.res() + sm.rp * 2;
    if (tid < 2) sm.res()[(sm.rp ^ 1) * 2 + tid] = 0;  // (last read before the previous barrier)
    if (lane == 0) {
        if (oHi) atomicMax(&r[0], oHi);
        if (oLo) atomicMax(&r[1], oLo);
    }
    S5_T(1);
    __syncthreads();
    S5_T(5);
    S5_N(10);
    *hiOut = s5_unord(r[0]);
    *lwOut = needPred ? s5_unord(r[1]) : *hiOut;
    sm.rp ^= 1;
}

// Returns false when no feature is visible.  On success *mad = median of |x - median| (for the even rule the mean of the two
// middle elements, as the median itself).  force: 0 normal, 1 never use the prediction, 2 generic tier only.
// *tier: 0 both phases found their target in the predicted bracket, 1 a prediction needed count passes, 2 no prediction,
// 4 generic.  nTotal = rows of the reference's residual vector (its parity picks the median rule); vis = this thread's
// feature is visible (its residuals are real), else rs[] must hold S5_SKIP.
template <int AREA, int NT, class SM>
__device__ __forceinline__ bool s5_sigma(const float (&rs)[AREA], bool vis, int nTotal, S5Pred& pr, SM& sm, int& rot, int force, double* madOut,
                                         uint32_t* nvisOut, int* tier)
{
    constexpr int CAP   = s5_cap<NT>();
    constexpr int HALF0 = (AREA + 1) / 2;  // keys counted in the first accumulator of a count pass
    static_assert(HALF0 <= 15 && AREA - HALF0 <= 15, "4-bit fields");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t nvis = 0, k = 0, kLow = 0;
    bool needPred = false, known = false;
    float result0 = 0.f, result1 = 0.f;  // median, deviation (already averaged for the even rule)
    float center     = 0.f;
    uint32_t absmask = 0xffffffffu;
    int worst = 0;
    sm.why    = 0;
#ifdef SVO_PROFILE
    sm.tlast = clock64();
#endif
    if (force == 3) {  // timing experiments only: a constant scale, no selection at all (results are NOT the reference's)
        uint32_t* c       = sm.cnt() + rot * 16;
        const uint32_t nv = __popc(__ballot_sync(S5_FULL, vis));
        if (lane == 0 && nv) atomicAdd(&c[2], nv);
        __syncthreads();
        nvis = c[2];
        if (tid < 16) sm.cnt()[((rot + 2) % 3) * 16 + tid] = 0;
        rot      = (rot + 1) % 3;
        *nvisOut = nvis;
        *madOut  = 3.0 * 65536.0;
        *tier    = 4;
        return nvis != 0;
    }
    // ---- fused pass: both targets at once, when both are predicted into brackets expected to hold few keys ----
    //   median bracket [mA, mA + wM) with centre M = mA + hm; deviation bracket [dA, dA + wD0).  With e = |x - M|:
    //   e < dA - hm            "inner": |x - med| < dA for EVERY median in the bracket      -> counted
    //   dA - hm <= e < dB + hm  candidates of the deviation                                   -> pushed
    //   the rest               |x - med| >= dB                                              -> nothing
    // The median is element k - cM + Lb of the list (cM keys below mA, Lb of them pushed as candidates); the deviation is
    // element k - cInner of the candidates' |x - med| PROVIDED it lies inside [dA, dB) by a margin: then every inner key is
    // below it and every other key above, and the rank arithmetic is a proof.
    int phaseStart = 0;
    bool fusedTried = false;
    if (pr.have && pr.haveMove && force == 0 && pr.rho[0] > 0.f && pr.rho[1] > 0.f) {
        const float hm = fminf(fmaxf(fmaxf(1.25f * pr.moved[0], __fdividef(8.f, pr.rho[0])), 16.f), 1.0e7f);
        const float hd = fminf(fmaxf(fmaxf(1.25f * pr.moved[1], __fdividef(8.f, pr.rho[1])), 16.f), 1.0e7f);
        if (2.f * hm * pr.rho[0] + 2.f * (hd + hm) * pr.rho[1] <= 200.f) {
            fusedTried     = true;
            const float mA = pr.v[0] - hm, wM = 2.f * hm, dA = pr.v[1] - hd, wD0 = 2.f * hd;
            const float gLo = dA - hm, wD = wD0 + 2.f * hm;
            const uint32_t wMb = __float_as_uint(wM), wDb = __float_as_uint(wD);
            uint32_t* c = sm.cnt() + rot * 16;
            float nbM = 0.f, nbI = 0.f;
            uint32_t top = (uint32_t)tid;
            const uint32_t lim = (uint32_t)tid + (uint32_t)((S5_DEPTH - 1) * NT);
#pragma unroll
            for (int i = 0; i < AREA; i++) {
                const float a = rs[i] - mA;
                const float g = fabsf(a - hm) - gLo;
                nbM += __saturatef(fmaf(a, 1.2676506e30f, 1.f));
                nbI += __saturatef(fmaf(g, 1.2676506e30f, 1.f));
                if (__float_as_uint(a) < wMb || __float_as_uint(g) < wDb) {
                    sm.stack()[top] = __float_as_uint(rs[i]);
                    top += NT;
                }
                top = min(top, lim);
            }
            uint32_t cM = __reduce_add_sync(S5_FULL, (uint32_t)AREA - (uint32_t)nbM);
            uint32_t cI = __reduce_add_sync(S5_FULL, (uint32_t)AREA - (uint32_t)nbI);
            const uint32_t mine = (top - (uint32_t)tid) / NT;
            uint32_t incl       = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(S5_FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t wcount = __shfl_sync(S5_FULL, incl, 31);
            const bool full       = __any_sync(S5_FULL, mine >= (uint32_t)(S5_DEPTH - 1));
            uint32_t wbase = 0;
            if (lane == 0 && wcount) wbase = atomicAdd(&c[1], wcount);
            wbase = __shfl_sync(S5_FULL, wbase, 0) + incl - mine;
            uint32_t lb = 0, nm = 0, nd = 0;  // my pushed keys: below mA, inside the median bracket, deviation candidates
            for (uint32_t j = 0; j < mine; j++) {
                const uint32_t xb = sm.stack()[j * NT + tid];
                const float a     = __uint_as_float(xb) - mA;
                const float g     = fabsf(a - hm) - gLo;
                lb += __float_as_uint(a) >> 31;
                nm += __float_as_uint(a) < wMb ? 1u : 0u;
                nd += __float_as_uint(g) < wDb ? 1u : 0u;
                if (wbase + j < (uint32_t)CAP) sm.list()[wbase + j] = xb;
            }
            const uint32_t packed = __reduce_add_sync(S5_FULL, lb | (nm << 10) | (nd << 20));  // (each below 256 per warp)
            const uint32_t nv     = __popc(__ballot_sync(S5_FULL, vis));
            if (lane == 0) {
                if (cM) atomicAdd(&c[0], cM);
                if (cI) atomicAdd(&c[3], cI);
                if (nv) atomicAdd(&c[2], nv);
                if (packed) atomicAdd(&c[5], packed & 1023u), atomicAdd(&c[6], (packed >> 10) & 1023u), atomicAdd(&c[7], packed >> 20);
                if (full) c[4] = 1;
            }
            S5_T(0);
            __syncthreads();
            S5_T(4);
            S5_N(8);
            const uint32_t CM = c[0], n = c[1], CI = c[3], Lb = c[5], nM = c[6], nD = c[7];
            const bool ovf = c[4] != 0 || n > (uint32_t)CAP;
            nvis     = c[2];
            known    = true;
            k        = nvis * (uint32_t)AREA / 2u;  // numValid / 2
            needPred = !(nTotal & 1) && k > 0;
            kLow     = needPred ? k - 1u : k;
            if (tid < 16) sm.cnt()[((rot + 2) % 3) * 16 + tid] = 0;
            rot = (rot + 1) % 3;
            if (nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            if (!ovf && CM <= kLow && k < CM + nM) {
                float hi, lw;
                s5_rank<NT>(sm, n, k - CM + Lb, needPred, &hi, &lw);
                result0    = needPred ? fmaf(0.5f, hi, 0.5f * lw) : hi;
                pr.rho[0]  = __fdividef((float)max(nM, 1u), wM);
                phaseStart = 1;
                center     = result0;
                absmask    = 0x7fffffffu;
                // the list becomes the candidates' deviations (+inf for the other entries)
                if ((uint32_t)tid < n) {
                    const float x = __uint_as_float(sm.list()[tid]);
                    const float g = fabsf((x - mA) - hm) - gLo;
                    sm.list()[tid] = __float_as_uint(g) < wDb ? __float_as_uint(fabsf(x - center)) : 0x7f800000u;
                }
                __syncthreads();
                if (CI <= kLow && k < CI + nD) {
                    s5_rank<NT>(sm, n, k - CI, needPred, &hi, &lw);
                    const float mgn = 2.f + 1.0e-6f * (dA + wD0);
                    if (lw >= dA + mgn && hi < dA + wD0 - mgn) {  // the proof (see above)
                        result1    = needPred ? fmaf(0.5f, hi, 0.5f * lw) : hi;
                        pr.rho[1]  = __fdividef((float)max(nD, 1u), wD);
                        phaseStart = 2;
                    }
                }
            }
            if (phaseStart < 2) worst = 1;
        }
    }
#pragma unroll 1
    for (int phase = phaseStart; phase < 2; phase++) {
        // ---- first range ----
        const float pv = phase == 0 ? pr.v[0] : pr.v[1], pd = pr.v[1];
        const bool predicted = pr.have && pr.haveMove && force == 0;
        bool bracket = false;
        float lo, W;
        if (predicted) {
            const float rho = phase == 0 ? pr.rho[0] : pr.rho[1];
            float h         = 1.25f * (phase == 0 ? pr.moved[0] : pr.moved[1]);  // (Gauss-Newton steps shrink several-fold from one to the next)
            if (rho > 0.f) h = fmaxf(h, __fdividef(8.f, rho));
            h       = fminf(fmaxf(h, 16.f), 1.0e7f);
            lo      = pv - h;
            W       = 2.f * h;
            bracket = rho > 0.f && W * rho <= 200.f && !fusedTried;  // (a failed fused pass: the same bracket would fail again)
            if (!bracket) worst = max(worst, 1);
        } else if (pr.have && force != 1) {  // first evaluation of a level: the values of the level above
            worst         = max(worst, 2);
            const float s = fmaxf(pd, 1024.f);
            lo            = phase == 0 ? pv - s : 0.5f * s;
            W             = phase == 0 ? 2.f * s : 1.5f * s;
        } else {  // nothing known: +- 16 / [0, 32) intensity units
            worst = max(worst, 2);
            lo    = phase == 0 ? -1048576.f : 0.f;
            W     = 2097152.f;
        }
        bool done   = false;
        int attempt = 0;
        float found = 0.f, rhoNew = 0.f;
        if (force == 2) attempt = S5_PASSES;
#pragma unroll 1
        for (; attempt < S5_PASSES && !done; attempt++) {
            uint32_t* c = sm.cnt() + rot * 16;
            uint32_t C0 = 0, n = 0;
            if (bracket) {
                // ---- bracket pass: count below, push the keys inside on this thread's stack; branch-free (the keys of a pass
                // are independent instruction streams), the count on the FP32 pipe: sat(u 2^100 + 1) is 1 for u >= 0, else 0 ----
                const uint32_t Wbits = __float_as_uint(W);
                float notBelow = 0.f;
                uint32_t top = (uint32_t)tid;                                      // word offset into the stacks [DEPTH][NT], column tid
                const uint32_t lim = (uint32_t)tid + (uint32_t)((S5_DEPTH - 1) * NT);  // (a full stack keeps overwriting its last slot)
#pragma unroll
                for (int i = 0; i < AREA; i++) {
                    const float v = __uint_as_float(__float_as_uint(rs[i] - center) & absmask);
                    const float u = v - lo;
                    notBelow += __saturatef(fmaf(u, 1.2676506e30f, 1.f));
                    if (__float_as_uint(u) < Wbits) {  // 0 <= u < W (a negative u has the sign bit set)
                        sm.stack()[top] = __float_as_uint(v);
                        top += NT;
                    }
                    top = min(top, lim);
                }
                uint32_t c0 = (uint32_t)AREA - (uint32_t)notBelow;
                c0          = __reduce_add_sync(S5_FULL, c0);
                // compaction: the stacks of a warp go to one contiguous part of the list, ONE atomic per warp reserves it
                const uint32_t mine = (top - (uint32_t)tid) / NT;        // (DEPTH - 1 = full or overflowed: treated as overflow)
                uint32_t incl       = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(S5_FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const uint32_t wcount = __shfl_sync(S5_FULL, incl, 31);
                const bool full       = __any_sync(S5_FULL, mine >= (uint32_t)(S5_DEPTH - 1));
                uint32_t wbase = 0;
                if (lane == 0) {
                    if (c0) atomicAdd(&c[0], c0);
                    if (wcount) wbase = atomicAdd(&c[1], wcount);
                    if (full) c[4] = 1;
                }
                if (!known) {
                    const uint32_t nv = __popc(__ballot_sync(S5_FULL, vis));
                    if (lane == 0 && nv) atomicAdd(&c[2], nv);
                }
                wbase = __shfl_sync(S5_FULL, wbase, 0) + incl - mine;
                for (uint32_t j = 0; j < mine; j++)
                    if (wbase + j < (uint32_t)CAP) sm.list()[wbase + j] = sm.stack()[j * NT + tid];
                S5_T(0);
                __syncthreads();
                S5_T(4);
                S5_N(8);
                C0 = c[0], n = c[1];
            } else {
                // ---- count pass: 16 bins, 4-bit fields ----
                // bin f = rint(60 t) >> 2 with t = sat((v - lo') / W'): inner bins 1..14 tile [lo, lo + W)
                const float sc = __fdividef(14.f, 15.f * W), of = -(lo - W * (1.f / 16.f)) * sc;
                uint64_t acc0 = 0, acc1 = 0;
#pragma unroll
                for (int i = 0; i < AREA; i++) {
                    const float v     = __uint_as_float(__float_as_uint(rs[i] - center) & absmask);
                    const float t     = __saturatef(fmaf(v, sc, of));
                    const uint32_t sh = __float_as_uint(fmaf(t, 60.f, S5_MAGIC)) & 0x3cu;
                    if (i < HALF0)
                        acc0 += 1ull << sh;
                    else
                        acc1 += 1ull << sh;
                }
                // (invisible features sit in bin 15 with all their keys: taken out below through nvis)
                const uint64_t m4 = 0x0f0f0f0f0f0f0f0full;
                const uint64_t ev = (acc0 & m4) + (acc1 & m4), od = ((acc0 >> 4) & m4) + ((acc1 >> 4) & m4);  // bytes: bins 0 2 4 .. / 1 3 5 ..
                const uint32_t w4[4] = {(uint32_t)ev, (uint32_t)(ev >> 32), (uint32_t)od, (uint32_t)(od >> 32)};
                uint32_t mine = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t a = __reduce_add_sync(S5_FULL, w4[q] & 0x00ff00ffu);         // bytes 0, 2 of the word
                    const uint32_t b = __reduce_add_sync(S5_FULL, (w4[q] >> 8) & 0x00ff00ffu);  // bytes 1, 3
                    if (lane == 2 * q) mine = a;
                    if (lane == 2 * q + 1) mine = b;
                }
                if (lane < 8 && mine) atomicAdd(&c[8 + lane], mine);
                if (!known) {
                    const uint32_t nv = __popc(__ballot_sync(S5_FULL, vis));
                    if (lane == 0 && nv) atomicAdd(&c[2], nv);
                }
                S5_T(2);
                __syncthreads();
                S5_T(6);
                S5_N(9);
            }
            if (!known) {
                nvis     = c[2];
                known    = true;
                k        = nvis * (uint32_t)AREA / 2u;  // numValid / 2
                needPred = !(nTotal & 1) && k > 0;
                kLow     = needPred ? k - 1u : k;
            }
            const bool ovf = c[4] != 0;
            if (tid < 16) sm.cnt()[((rot + 2) % 3) * 16 + tid] = 0;  // last read before the previous barrier
            rot = (rot + 1) % 3;
            if (nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            if (bracket) {
                const bool inside = C0 <= kLow && k < C0 + n;
                if (inside && n <= (uint32_t)CAP && !ovf) {
                    float hi, lw;
                    s5_rank<NT>(sm, n, k - C0, needPred, &hi, &lw);
                    found  = needPred ? fmaf(0.5f, hi, 0.5f * lw) : hi;
                    rhoNew = __fdividef((float)max(n, 1u), W);
                    done   = true;
                    S5_T(1);
                } else {
                    if (attempt == 0) worst = max(worst, 1);
                    bracket = false;
                    if (!inside) {  // a miss: the bracket becomes the middle of a range fourteen times as wide
                        lo -= 6.5f * W;
                        W *= 14.f;
                    }               // (else: the bracket holds the target and too many keys: count inside it)
                }
            } else {
                // ---- locate: lane f of every warp owns bin f (f < 16); word q of the totals holds two bins in its halves:
                //   q 0: bins 0, 4   q 1: bins 2, 6   q 2: bins 8, 12   q 3: bins 10, 14
                //   q 4: bins 1, 5   q 5: bins 3, 7   q 6: bins 9, 13   q 7: bins 11, 15
                const int f  = lane & 15;
                const int q  = ((f & 1) << 2) | ((f >> 3) << 1) | ((f >> 1) & 1);
                uint32_t cum = (c[8 + q] >> ((f & 4) << 2)) & 0xffffu;  // this bin's total
                if (f == 15) cum = 0;                                    // (bin 15 holds the invisible features' keys too: never needed)
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(S5_FULL, cum, o, 16);
                    if (f >= o) cum += t;
                }  // keys in bins <= f
                const uint32_t reachHi = __ballot_sync(S5_FULL, k < cum) & 0x7fffu, reachLo = __ballot_sync(S5_FULL, kLow < cum) & 0x7fffu;
                const int bHi = reachHi ? __ffs(reachHi) - 1 : 15, bLo = reachLo ? __ffs(reachLo) - 1 : 15;
                const uint32_t cLo    = __shfl_sync(S5_FULL, cum, max(bLo - 1, 0)) * (bLo > 0 ? 1u : 0u);  // keys below bin bLo
                const uint32_t cHiEnd = bHi < 15 ? __shfl_sync(S5_FULL, cum, bHi) : nvis * (uint32_t)AREA;
                const uint32_t m      = cHiEnd - cLo;  // keys in bins bLo .. bHi
                const float bw   = W * (1.f / 14.f);
                float nlo = lo + (float)(bLo - 1) * bw, nhi = lo + (float)bHi * bw;
                if (bLo == 0) nlo = nhi - fmaxf(16.f * W, 4.f * fabsf(nhi));   // below the range: open it downwards
                if (bHi == 15) nhi = nlo + fmaxf(16.f * W, 4.f * fabsf(nlo));  // above: upwards
                const float mg = 0.02f * bw + 1.0e-6f * fmaxf(fabsf(nlo), fabsf(nhi));
                lo      = nlo - mg;
                W       = (nhi - nlo) + 2.f * mg;
                bracket = bLo > 0 && bHi < 15 && m <= 192u;
                if (!bracket && bLo > 0 && bHi < 15 && !(bw > 1.0e-2f)) attempt = S5_PASSES;  // ties: the bins cannot split them
                S5_T(3);
            }
        }
        sm.why |= min(attempt, 15) << (4 * phase);
        if (!done) {
            // ---- generic: bisection over the ordered-integer image of the keys, always right ----
            worst = 4;
            if (!known) {  // (forced: nobody has counted the visible features yet)
                uint32_t* c       = sm.cnt() + rot * 16;
                const uint32_t nv = __popc(__ballot_sync(S5_FULL, vis));
                if (lane == 0 && nv) atomicAdd(&c[2], nv);
                __syncthreads();
                nvis = c[2];
                if (tid < 16) sm.cnt()[((rot + 2) % 3) * 16 + tid] = 0;
                rot      = (rot + 1) % 3;
                known    = true;
                k        = nvis * (uint32_t)AREA / 2u;
                needPred = !(nTotal & 1) && k > 0;
                kLow     = needPred ? k - 1u : k;
                if (nvis == 0) {
                    *nvisOut = 0;
                    return false;
                }
            }
            uint32_t T = 0, lessT = 0, maxBelow = 0;
#pragma unroll 1
            for (int b = 31; b >= -1; b--) {  // b = -1: the final pass counts the keys below T and finds the largest of them
                const uint32_t cand = b >= 0 ? (T | (1u << b)) : T;
                uint32_t c0 = 0, mb = 0;
#pragma unroll
                for (int i = 0; i < AREA; i++) {
                    const uint32_t o = s5_ord(__uint_as_float(__float_as_uint(rs[i] - center) & absmask));
                    if (o < cand) c0++, mb = max(mb, o);
                }
                c0          = __reduce_add_sync(S5_FULL, c0);
                mb          = __reduce_max_sync(S5_FULL, mb);
                uint32_t* c = sm.cnt() + rot * 16;
                if (lane == 0 && c0) atomicAdd(&c[0], c0), atomicMax(&c[3], mb);
                __syncthreads();
                const uint32_t C0 = c[0];
                maxBelow          = c[3];
                if (tid < 16) sm.cnt()[((rot + 2) % 3) * 16 + tid] = 0;
                rot = (rot + 1) % 3;
                if (b >= 0) {
                    if (C0 <= k) T = cand;
                } else
                    lessT = C0;
            }
            const float hi = s5_unord(T);
            const float lw = (needPred && lessT > kLow) ? s5_unord(maxBelow) : hi;  // (lessT <= k - 1: element k - 1 equals element k)
            found  = needPred ? fmaf(0.5f, hi, 0.5f * lw) : hi;
            rhoNew = 0.f;
            S5_T(11);
        }
        if (phase == 0)
            result0 = found, pr.rho[0] = rhoNew;
        else
            result1 = found, pr.rho[1] = rhoNew;
        // ---- the deviations are taken from the median just found ----
        center  = result0;
        absmask = 0x7fffffffu;
    }
    // ---- prediction for the next evaluation ----
    if (pr.have) {
        pr.moved[0] = fabsf(result0 - pr.v[0]);
        pr.moved[1] = fabsf(result1 - pr.v[1]);
        pr.haveMove = true;
    }
    pr.v[0] = result0, pr.v[1] = result1, pr.have = true;
    *tier    = worst;
    *madOut  = (double)result1;
    *nvisOut = nvis;
    return true;
}

}  // namespace
