// select4.cuh -- robust scale of the alignment (tukeyWeighting / computeSigma, src/optimizer.cpp:485-507, median rule
// src/algorithm.cpp:834-872, MEDIAN_EXACT of SURVEY 9.3) for ONE CTA of NT threads, each holding AREA residuals in
// registers: the median and the median absolute deviation are found TOGETHER.
//
// A ROUND histograms the keys (fixed point rint(r 2^16), biased) over three windows of one global grid (bin = key >> s):
// M around the predicted median, L and R around median -/+ deviation.  From the three histograms and the number of keys
// below each window, `s4_locate` derives the bin(s) [bMin, bMax] of the median (and of its predecessor when the even
// rule needs it), bounds j0 2^s < d* <= j1 2^s on the k-th smallest deviation by counting the keys that MUST / CAN lie
// within j bins of any median in those bins, the candidate bins L' = [bMin - j1, bMax - j0], R' = [bMin + j0, bMax + j1]
// and `base`, the number of keys closer to the median than every candidate.  `s4_lists` then ranks the few keys of the
// candidate bins exactly.  tests/model_select4.py is the executable statement of this arithmetic;
// tests/test_model_select4.py pins it against sorted arrays on the CPU.
//
// Every step after a fill is spread over ALL threads of the CTA (thread t owns bin t of each window, list entry t / 8,
// ...) with a block barrier between dependent steps: the steps are short, and work that every warp repeats costs
// sixteen times its instruction count in issue slots -- far more than a barrier.
//
// Rounds:   atomic   NT bins per window, shared-memory atomics by the FEW keys inside the windows (they are pushed on a
//                    thread-private stack in shared memory by a branch-free pass over the registers); windows come
//                    from the previous evaluation ("hot") or from a coarse round.
//           private  one contiguous span of 64 bins counted with thread-private packed byte counters (no atomics: every
//                    key takes part), summed over the columns by one warp per counter row ("cold": no usable prediction).
//           generic  exact radix select over all bits with atomics by every key: the safety net (degenerate
//                    distributions, overflowing stacks or lists), never the common path.
//
// Keys: the residuals arrive scaled by 2^16 in FP32.  key = bits(rs + 1.5 2^23) is rint(rs) + 0x4B400000 for |rs| < 2^22
// (|r| < 64 intensity units) and monotone everywhere: one FADD per key and pass, no conversion.  Windows always lie
// inside the linear range; the generic tier converts properly.
#pragma once
#include <stdint.h>

namespace {

#ifdef SVO_PROFILE
#define S4_T(i)                          \
    do {                                 \
        const long long t__ = clock64(); \
        sm.prof[i] += t__ - sm.tlast;    \
        sm.tlast = t__;                  \
    } while (0)
#else
#define S4_T(i)
#endif

constexpr unsigned S4_FULL    = 0xffffffffu;
constexpr uint32_t S4_BIAS    = 0x4B400000u;  // bits of 1.5 * 2^23
constexpr float S4_MAGIC      = 12582912.f;
constexpr uint32_t S4_DMAX    = 30u << 16;    // deviations up to 30 intensity units are handled by the window tiers: the widest span
                                              // (64 bins of one unit around a median within 16 units of zero) ends inside the linear range
constexpr int S4_CAPM         = 64;  // candidate list of the median
constexpr int S4_CAPD         = 128; // candidate list of the deviation
constexpr int S4_ATOMIC_LIMIT = 3072;  // a round with more keys inside its windows than this is counted privately instead
constexpr int S4_MARGIN       = 4;
constexpr int S4_NONE         = 0x7fffffff;

// shared-memory words of the selection for a CTA of NT threads with AREA >= 16 keys each (every thread's key stack holds
// all its keys: the residuals of a patch are strongly correlated, a whole patch inside one window is common)
template <int NT, int AREA>
__host__ __device__ constexpr size_t s4_smem_words()
{
    return (size_t)AREA * NT + 3 * NT + 3 * (NT + 8) + 48 + 16 + 32 + 16 + S4_CAPM + S4_CAPD;
}

// The selection's shared memory lives at byte offset OFF of the kernel's dynamic shared memory.  Every view is derived
// from the `extern __shared__` symbol with compile-time offsets, never from a stored pointer: the compiler keeps the
// shared address space and emits LDS / STS / ATOMS (a pointer that has been through a struct or a call degrades to
// generic loads and generic atomics, which are several times slower).
template <int NT, int OFF, int AREA>
struct S4Smem {
    static_assert(AREA >= 16, "the private counters of a coarse round need 16 rows");
    int why;  // diagnostics: why the last evaluation left the hot / cold tier (0: it did not)
    int shiftUsed;  // diagnostics: bin width (log2) of the last atomic round
#ifdef SVO_PROFILE
    long long prof[16], tlast;  // cycles: 0 hot fill 1 hot locate 2 - 3 lists 4 private fill 5 private reduce 6 private locate 7 -
                                //         8 refine + atomic fill 9 atomic locate 10 - 11 generic
#endif
    __device__ __forceinline__ static uint32_t* base()
    {
        extern __shared__ __align__(128) unsigned char s4_dynamic_smem[];
        return reinterpret_cast<uint32_t*>(s4_dynamic_smem + OFF);
    }
    // [AREA][NT] key stacks; the private counters [16][NT] of a coarse round live here too
    __device__ __forceinline__ static uint32_t* stack() { return base(); }
    // [3][NT] windows M, L, R; zero between rounds (the scan clears what it reads)
    __device__ __forceinline__ static uint32_t* hist() { return base() + AREA * NT; }
    // [3][NT + 8] keys below every bin edge of the three windows (absolute counts): cab[X][i] = keys with grid bin < gX + i
    __device__ __forceinline__ static uint32_t* cab() { return hist() + 3 * NT; }
    // [3][16] warp totals
    __device__ __forceinline__ static uint32_t* wtot() { return cab() + 3 * (NT + 8); }
    // [16] non-empty bins of the M window, one mask per warp
    __device__ __forceinline__ static uint32_t* nzm() { return wtot() + 48; }
    // [2][16] per round parity: 0..2 keys below M / L / R, 3 visible features, 4 flags; [5], [6] list fills
    __device__ __forceinline__ static uint32_t* cnt() { return nzm() + 16; }
    // [16] 0 max-below (generic) 1 visible (generic) 2..5 median: bin, rank, count, predecessor bin  6, 7 first j  8..11 ranks
    //      12, 13 bin and rank of a generic pass
    __device__ __forceinline__ static uint32_t* misc() { return cnt() + 32; }
    __device__ __forceinline__ static uint32_t* listM() { return misc() + 16; }
    __device__ __forceinline__ static uint32_t* listD() { return listM() + S4_CAPM; }
    __device__ __forceinline__ static void clear()
    {
        uint32_t* b = base();
        for (int i = threadIdx.x; i < (int)s4_smem_words<NT, AREA>(); i += NT) b[i] = 0;
    }
};

struct S4Win {  // three windows on the grid of bins of width 2^s (starts in bins), ordered L < M < R, at most NT bins each
    int s, gM, gL, gR, nbM, nbL, nbR;
    bool contig;
};

__device__ __forceinline__ S4Win s4_make_win(int s, int gM, int nbM, int gL, int nbL, int gR, int nbR)
{
    S4Win w;
    w.s = s, w.gM = gM, w.gL = gL, w.gR = gR, w.nbM = nbM, w.nbL = nbL, w.nbR = nbR;
    w.contig = (gL + nbL == gM) && (gM + nbM == gR);
    return w;
}

// M centred on m0, R centred on m0 + d0, L on m0 - d0 (contiguous triple when they would overlap)
__device__ __forceinline__ S4Win s4_predicted(uint32_t m0, uint32_t d0, int s, int nb)
{
    const int gM = (int)(m0 >> s) - nb / 2;
    int gR       = (int)((m0 + d0) >> s) - nb / 2;
    int gL       = (int)((m0 - d0) >> s) - (nb - 1) / 2;
    if (gR < gM + nb || gL + nb > gM) {
        gR = gM + nb;
        gL = gM - nb;
    }
    return s4_make_win(s, gM, nb, gL, nb, gR, nb);
}

struct S4Loc {  // result of s4_locate (uniform over the CTA)
    int bMin, bMax;      // grid bins of the median's predecessor / the median
    uint32_t hMb, rM;    // keys in bin bMax, rank of the median inside it
    int j0, j1;
    uint32_t base;
    int Ll, Lh, Rl, Rh;  // candidate bins of the deviation
    uint32_t nCandM, nCandD;  // keys in the candidate bins (exact)
    bool ok;
};

struct S4Counters {
    uint32_t below[3];  // keys below the M, L, R windows
    uint32_t nvis;      // visible features
    uint32_t flags;     // 1: a key stack was full
};

__device__ __forceinline__ uint32_t s4_warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(S4_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// keys with grid bin < b; b must be an edge of the window the rule picks (guaranteed by the ranges of s4_locate)
template <int NT, class SM>
__device__ __forceinline__ uint32_t s4_cf(const SM& sm, const S4Win& w, int b)
{
    const int X = b >= w.gR ? 2 : (b >= w.gM ? 0 : 1);
    const int g = X == 2 ? w.gR : (X == 0 ? w.gM : w.gL);
    const int i = min(max(b - g, 0), NT);  // (only out of range for probes whose result is discarded)
    return sm.cab()[X * (NT + 8) + i];
}

// ---------------------------------------------------------------------------------------------------------------
// After the fill barrier.  Four short steps over all threads, three barriers (model_select4.locate):
//   1  thread t scans bin t of the three windows within its warp, publishes warp totals and M's non-empty mask
//   2  absolute counts below every bin edge -> shared memory; the thread whose M bin holds rank k publishes the median's
//      bin, its rank inside and, when the even rule needs it, the last non-empty bin before it
//   3  thread t takes the bin edges it owns as the upper argument of Glo(j) / Ghi(j) (one lookup each for the lower
//      argument) and the CTA takes the minimum j that reaches the rank
//   4  candidate bins, base and candidate counts (a dozen lookups, every thread)
// Clears the histogram bins it reads and the counters of the OTHER round parity.  Uniform result.
// ---------------------------------------------------------------------------------------------------------------
template <int NT, class SM>
__device__ __forceinline__ S4Loc s4_locate(SM& sm, const S4Win& w, int parity, int area, int nTotal, S4Counters* cnOut, uint32_t* kOut,
                                           bool* needPredOut)
{
    constexpr int NW = NT / 32;
    constexpr int CS = NT + 8;  // stride of cab
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    S4Loc L;
    L.ok = false;
    L.bMin = L.bMax = 0, L.hMb = L.rM = 0, L.j0 = -1, L.j1 = 0, L.base = 0, L.Ll = L.Lh = L.Rl = L.Rh = 0, L.nCandM = L.nCandD = 0;
    // ---- step 1 ----
    const uint32_t* c = sm.cnt() + parity * 16;
    S4Counters cn;
    cn.below[0] = c[0], cn.below[1] = c[1], cn.below[2] = c[2], cn.nvis = c[3], cn.flags = c[4];
    *cnOut = cn;
    const uint32_t k    = cn.nvis * (uint32_t)area / 2u;  // numValid / 2
    const bool needPred = !(nTotal & 1) && k > 0;
    *kOut = k, *needPredOut = needPred;
    uint32_t h[3], incl[3];
#pragma unroll
    for (int X = 0; X < 3; X++) {
        h[X]                    = sm.hist()[X * NT + tid];
        sm.hist()[X * NT + tid] = 0;
    }
    const uint32_t nz = __ballot_sync(S4_FULL, h[0] != 0);
#pragma unroll
    for (int X = 0; X < 3; X++) {
        incl[X] = s4_warp_incl_scan(h[X], lane);
        if (lane == 31) sm.wtot()[X * 16 + warp] = incl[X];
    }
    if (lane == 0) sm.nzm()[warp] = nz;
    if (tid < 16) sm.cnt()[(parity ^ 1) * 16 + tid] = 0;  // last read before the previous round's barriers
    if (tid == 0) {
        sm.misc()[2] = 0xffffffffu;  // no owner of the median yet
        sm.misc()[6] = (uint32_t)S4_NONE;
        sm.misc()[7] = (uint32_t)S4_NONE;
    }
    __syncthreads();
    // ---- step 2 ----
    uint32_t cabs[3];  // keys below the UPPER edge of this thread's bin
#pragma unroll
    for (int X = 0; X < 3; X++) {
        const uint32_t wt = lane < NW ? sm.wtot()[X * 16 + lane] : 0u;
        const uint32_t wi = s4_warp_incl_scan(wt, lane);
        const uint32_t wb = __shfl_sync(S4_FULL, wi - wt, warp);
        cabs[X]           = cn.below[X] + wb + incl[X];
        sm.cab()[X * CS + tid + 1] = cabs[X];
        if (tid == 0) sm.cab()[X * CS] = cn.below[X];
    }
    {
        const uint32_t cex = cabs[0] - h[0];
        if (cex <= k && k < cabs[0]) {  // this thread's M bin holds the median
            const uint32_t rM = k - cex;
            int iP            = -1;
            if (needPred && rM == 0) {  // predecessor: the last non-empty bin before this one
                const uint32_t mine = nz & ((1u << lane) - 1u);
                if (mine)
                    iP = warp * 32 + 31 - __clz(mine);
                else
                    for (int wv = warp - 1; wv >= 0 && iP < 0; wv--) {
                        const uint32_t z = sm.nzm()[wv];
                        if (z) iP = wv * 32 + 31 - __clz(z);
                    }
            }
            sm.misc()[2] = (uint32_t)tid;
            sm.misc()[3] = rM;
            sm.misc()[4] = h[0];
            sm.misc()[5] = (uint32_t)iP;
        }
    }
    __syncthreads();
    // ---- step 3 ----
    const uint32_t iMu = sm.misc()[2];
    if (iMu == 0xffffffffu) return L;  // the median lies outside the M window
    L.rM = sm.misc()[3], L.hMb = sm.misc()[4];
    const int iP = (int)sm.misc()[5];
    L.bMax       = w.gM + (int)iMu;
    L.bMin       = L.bMax;
    bool bad     = false;
    if (needPred && L.rM == 0) {
        if (iP < 0)
            bad = true;  // below the window
        else
            L.bMin = w.gM + iP;
    }
    const int bMin = L.bMin, bMax = L.bMax;
    const int hi1 = min(w.gR + w.nbR - bMin, bMax + 1 - w.gL);
    const int hi0 = min(w.gR + w.nbR - bMax - 1, bMin - w.gL);
    int lo1 = 1, lo0 = 0;
    if (!w.contig) {
        lo1 = max(max(w.gR - bMin, bMax + 1 - w.gL - w.nbL), 1);
        lo0 = max(max(w.gR - bMax - 1, bMin - w.gL - w.nbL), 0);
    }
    const uint32_t kk = needPred ? k : k + 1;
    // Glo(j) = C(bMin + j) - C(bMax - j + 1) and Ghi(j') = C(bMax + j' + 1) - C(bMin - j') are ONE function of the upper
    // argument b:  G(b) = C(b) - C(bMin + bMax + 1 - b),  Glo(j) = G(bMin + j),  Ghi(j') = G(bMax + j' + 1).
    // Thread t probes the bin edges it owns (index t + 1 of a window, thread 0 also index 0) as b: one lookup each.
    {
        const int S   = bMin + bMax + 1;
        int best1 = S4_NONE, best0 = S4_NONE;
        auto probe = [&](int b, uint32_t cb, uint32_t cl) {
            const int g = (int)(cb - cl);
            const int j = b - bMin, jj = b - bMax - 1;
            if (j >= lo1 && j <= hi1 && g >= (int)(k + 1)) best1 = min(best1, j);
            if (jj >= lo0 && jj <= hi0 && g >= (int)kk) best0 = min(best0, jj);
        };
        if (!w.contig) {  // the upper argument is an edge of R, the lower one an edge of L
            if (tid < w.nbR) {
                const int b = w.gR + tid + 1;
                const int i = min(max(S - b - w.gL, 0), NT);
                probe(b, cabs[2], sm.cab()[CS + i]);
            }
            if (tid == 0) {
                const int i = min(max(S - w.gR - w.gL, 0), NT);
                probe(w.gR, cn.below[2], sm.cab()[CS + i]);
            }
        } else {
#pragma unroll
            for (int X = 0; X < 3; X++) {
                const int nb = X == 2 ? w.nbR : (X == 0 ? w.nbM : w.nbL);
                const int g  = X == 2 ? w.gR : (X == 0 ? w.gM : w.gL);
                if (tid < nb) {
                    const int b = g + tid + 1;
                    probe(b, cabs[X], s4_cf<NT>(sm, w, S - b));
                }
            }
            if (tid == 0) probe(w.gL, cn.below[1], s4_cf<NT>(sm, w, S - w.gL));
        }
        best1 = __reduce_min_sync(S4_FULL, best1);
        best0 = __reduce_min_sync(S4_FULL, best0);
        if (lane == 0) {
            if (best1 != S4_NONE) atomicMin(reinterpret_cast<int*>(&sm.misc()[6]), best1);
            if (best0 != S4_NONE) atomicMin(reinterpret_cast<int*>(&sm.misc()[7]), best0);
        }
    }
    __syncthreads();
    // ---- step 4 ----
    if (bad) return L;
    const int j1 = (int)sm.misc()[6];
    if (j1 == S4_NONE) return L;  // the deviation lies beyond the windows
    int jz = (int)sm.misc()[7];
    jz     = min(jz == S4_NONE ? hi0 + 1 : jz, min(hi0, j1) + 1);
    int j0 = jz - 1;
    if (j0 < lo0) {
        if (!w.contig) return L;  // the lower bound lies below the windows
        j0 = -1;
    }
    if (j0 >= 0 && j0 < lo1 && !w.contig) return L;
    L.j0 = j0, L.j1 = j1;
    L.Ll = bMin - j1, L.Lh = bMax - j0, L.Rl = bMin + j0, L.Rh = bMax + j1;
    if (w.contig) {
        if (L.Ll < w.gL || L.Rh >= w.gR + w.nbR) return L;
        uint32_t base = 0;
        if (j0 >= 0) {
            const int g = (int)(s4_cf<NT>(sm, w, bMin + j0) - s4_cf<NT>(sm, w, bMax - j0 + 1));
            base        = g > 0 ? (uint32_t)g : 0u;
        }
        L.base   = base;
        L.nCandM = s4_cf<NT>(sm, w, bMax + 1) - s4_cf<NT>(sm, w, bMin);
        if (L.Lh >= L.Rl - 1)
            L.nCandD = s4_cf<NT>(sm, w, L.Rh + 1) - s4_cf<NT>(sm, w, L.Ll);
        else
            L.nCandD = (s4_cf<NT>(sm, w, L.Lh + 1) - s4_cf<NT>(sm, w, L.Ll)) + (s4_cf<NT>(sm, w, L.Rh + 1) - s4_cf<NT>(sm, w, L.Rl));
    } else {
        if (L.Ll < w.gL || L.Lh >= w.gL + w.nbL || L.Rl < w.gR || L.Rh >= w.gR + w.nbR) return L;
        // separate windows: every argument is known to lie in its window, no window selection
        const uint32_t* cM = sm.cab();
        const uint32_t* cL = sm.cab() + CS;
        const uint32_t* cR = sm.cab() + 2 * CS;
        const uint32_t cLh = cL[L.Lh + 1 - w.gL], cRl = cR[L.Rl - w.gR];
        L.base   = cRl > cLh ? cRl - cLh : 0u;  // = max(0, Glo(j0)): keys strictly between the candidate bins
        L.nCandM = cM[bMax + 1 - w.gM] - cM[bMin - w.gM];
        L.nCandD = (cLh - cL[L.Ll - w.gL]) + (cR[L.Rh + 1 - w.gR] - cRl);
    }
    L.ok = true;
    return L;
}

// windows of the next round: the candidate bins of this round plus S4_MARGIN bins on either side (the bounds of the finer
// round probe up to two bins beyond the targets), on the finest grid s2 <= s where every window has at most nb bins
// (model_select4.refine_windows).  Returns false if not even s does.
__device__ __forceinline__ bool s4_refine(const S4Loc& r, int s, int nb, S4Win* out)
{
    // the finest grid: the widest region (in bins of this round) shifted left by f, plus the margins, fills nb bins
    const int widest = max(max(r.bMax - r.bMin, r.Lh - r.Ll), r.Rh - r.Rl) + 1;
    int f            = min(s, 31 - __clz(max((nb - 2 * S4_MARGIN) / widest, 1)));
#pragma unroll 1
    for (; f >= 0; f--) {  // (the first f fits unless the windows touch and the contiguous triple is too long)
        const int s2        = s - f;
        const long long mlo = ((long long)r.bMin << f) - S4_MARGIN, mhi = (((long long)r.bMax + 1) << f) + S4_MARGIN;
        const long long llo = ((long long)r.Ll << f) - S4_MARGIN, lhi = (((long long)r.Lh + 1) << f) + S4_MARGIN;
        const long long rlo = ((long long)r.Rl << f) - S4_MARGIN, rhi = (((long long)r.Rh + 1) << f) + S4_MARGIN;
        if (mhi - mlo > nb || lhi - llo > nb || rhi - rlo > nb) continue;
        if (lhi <= mlo && mhi <= rlo) {  // separate windows
            *out = s4_make_win(s2, (int)mlo, (int)(mhi - mlo), (int)llo, (int)(lhi - llo), (int)rlo, (int)(rhi - rlo));
            return true;
        }
        // they touch or overlap: a contiguous triple around the M window
        const long long nbL = max(mlo - min(llo, mlo - 1), 1LL), nbR = max(max(rhi, mhi + 1) - mhi, 1LL);
        if (nbL <= nb && nbR <= nb) {
            *out = s4_make_win(s2, (int)mlo, (int)(mhi - mlo), (int)(mlo - nbL), (int)nbL, (int)mhi, (int)nbR);
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ uint32_t s4_key(float rs) { return __float_as_uint(rs + S4_MAGIC); }

// ---------------------------------------------------------------------------------------------------------------
// fills
// ---------------------------------------------------------------------------------------------------------------
// atomic round over ALL keys: branch-free classification (three subtractions, three sign counts, one unconditional
// store whose slot is kept when the key is inside a window), then atomics by the keys on the stack only.
// Returns the number of keys on this thread's stack.
template <int AREA, int NT, class SM>
__device__ __forceinline__ uint32_t s4_fill_atomic(const float (&rs)[AREA], bool vis, const S4Win& w, SM& sm, int parity)
{
    const int lane     = threadIdx.x & 31;
    const int s        = w.s;
    const uint32_t loM = (uint32_t)w.gM << s, loL = (uint32_t)w.gL << s, loR = (uint32_t)w.gR << s;
    const uint32_t wdM = (uint32_t)w.nbM << s, wdL = (uint32_t)w.nbL << s, wdR = (uint32_t)w.nbR << s;
    uint32_t cM = 0, cL = 0, cR = 0;
    uint32_t* col = sm.stack() + threadIdx.x;
    uint32_t n    = 0;
    if (vis) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t key = s4_key(rs[i]);
            const uint32_t tM = key - loM, tL = key - loL, tR = key - loR;  // keys below a window wrap to >= 2^31
            cM += tM >> 31;
            cL += tL >> 31;
            cR += tR >> 31;
            col[n * NT] = key;  // slot n <= i: the stack holds AREA keys, no overflow
            const bool in = (tM < wdM) | (tL < wdL) | (tR < wdR);
            n += in ? 1u : 0u;
        }
    }
    uint32_t* c = sm.cnt() + parity * 16;
    {   // every key on the stack lies in exactly one window: which one follows from two comparisons
        const uint32_t nmax = __reduce_max_sync(S4_FULL, n);
        for (uint32_t j = 0; j < nmax; j++) {
            if (j < n) {
                const uint32_t key = col[j * NT];
                const bool isR = key >= loR, isM = !isR && key >= loM;
                const uint32_t lo  = isR ? loR : (isM ? loM : loL);
                const uint32_t off = isR ? 2u * NT : (isM ? 0u : (uint32_t)NT);
                atomicAdd(&sm.hist()[off + ((key - lo) >> s)], 1u);
            }
        }
    }
    cM = __reduce_add_sync(S4_FULL, cM);
    cL = __reduce_add_sync(S4_FULL, cL);
    cR = __reduce_add_sync(S4_FULL, cR);
    const uint32_t nv = __popc(__ballot_sync(S4_FULL, vis));
    if (lane == 0) {
        if (cM) atomicAdd(&c[0], cM);
        if (cL) atomicAdd(&c[1], cL);
        if (cR) atomicAdd(&c[2], cR);
        if (nv) atomicAdd(&c[3], nv);
    }
    return n;
}

// private round: ONE contiguous span of 64 bins starting at grid bin g0 (L = bins 0..23, M = 24..39, R = 40..63), every
// key counted in this thread's own column of packed byte counters.  The caller synchronises, then s4_private_reduce.
template <int AREA, int NT, class SM>
__device__ __forceinline__ void s4_fill_private(const float (&rs)[AREA], bool vis, int g0, int s, SM& sm, int parity)
{
    const int lane    = threadIdx.x & 31;
    const uint32_t lo = (uint32_t)g0 << s, wd = 64u << s;
    uint32_t* col     = sm.stack() + threadIdx.x;
#pragma unroll
    for (int r = 0; r < 16; r++) col[r * NT] = 0;
    uint32_t below = 0;
    if (vis) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t t = s4_key(rs[i]) - lo;
            below += t >> 31;
            if (t < wd) {
                const uint32_t b = t >> s;
                col[(b >> 2) * NT] += 1u << ((b & 3u) * 8u);
            }
        }
    }
    below             = __reduce_add_sync(S4_FULL, below);
    const uint32_t nv = __popc(__ballot_sync(S4_FULL, vis));
    uint32_t* c       = sm.cnt() + parity * 16;
    if (lane == 0) {
        if (below) {  // below the span = below all three windows (the reduce adds the L and M bins to M's and R's counts)
            atomicAdd(&c[0], below);
            atomicAdd(&c[1], below);
            atomicAdd(&c[2], below);
        }
        if (nv) atomicAdd(&c[3], nv);
    }
}
// column sums of the private counters -> hist (L bins 0..23, M 24..39, R 40..63) and the counts below M and R
template <int NT, class SM>
__device__ __forceinline__ void s4_private_reduce(SM& sm, int parity)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 16; r += NW) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < NT / 32; j++) {
            const uint32_t v = sm.stack()[r * NT + lane + 32 * j];
            lo += v & 0x00ff00ffu;
            hi += (v >> 8) & 0x00ff00ffu;
        }
        lo = __reduce_add_sync(S4_FULL, lo);  // 16-bit fields: at most 25 * NT <= 12,800 each
        hi = __reduce_add_sync(S4_FULL, hi);
        if (lane < 4) {
            const int b      = 4 * r + lane;
            const uint32_t v = lane == 0 ? (lo & 0xffffu) : (lane == 1 ? (hi & 0xffffu) : (lane == 2 ? (lo >> 16) : (hi >> 16)));
            const int X = b < 24 ? 1 : (b < 40 ? 0 : 2), i = b < 24 ? b : (b < 40 ? b - 24 : b - 40);
            sm.hist()[X * NT + i] = v;
            uint32_t* c = sm.cnt() + parity * 16;
            if (v && b < 24) atomicAdd(&c[0], v);
            if (v && b < 40) atomicAdd(&c[2], v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// lists: the keys of the candidate bins, gathered from the stack (after an atomic round) or from the registers
// ---------------------------------------------------------------------------------------------------------------
template <int NT, class SM>
__device__ __forceinline__ void s4_push_candidate(uint32_t key, int s, const S4Loc& r, SM& sm, int parity)
{
    const int gb = (int)(key >> s);
    uint32_t* c  = sm.cnt() + parity * 16;
    if (gb == r.bMax || gb == r.bMin) {
        const uint32_t slot = atomicAdd(&c[5], 1u);
        if (slot < (uint32_t)S4_CAPM) sm.listM()[slot] = key;
    }
    if ((gb >= r.Ll && gb <= r.Lh) || (gb >= r.Rl && gb <= r.Rh)) {
        const uint32_t slot = atomicAdd(&c[6], 1u);
        if (slot < (uint32_t)S4_CAPD) sm.listD()[slot] = key;
    }
}
template <int NT, class SM>
__device__ __forceinline__ void s4_gather_stack(uint32_t n, int s, const S4Loc& r, SM& sm, int parity)
{
    const uint32_t* col = sm.stack() + threadIdx.x;
    const uint32_t nmax = __reduce_max_sync(S4_FULL, n);
    for (uint32_t j = 0; j < nmax; j++)
        if (j < n) s4_push_candidate<NT>(col[j * NT], s, r, sm, parity);
}
template <int AREA, int NT, class SM>
__device__ __forceinline__ void s4_gather_regs(const float (&rs)[AREA], bool vis, int s, const S4Loc& r, SM& sm, int parity)
{
    if (vis) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t key = s4_key(rs[i]);
            const int gb       = (int)(key >> s);
            if (gb == r.bMax || gb == r.bMin || (gb >= r.Ll && gb <= r.Lh) || (gb >= r.Rl && gb <= r.Rh)) s4_push_candidate<NT>(key, s, r, sm, parity);
        }
    }
}

struct S4Out {
    uint32_t kHi, kLo;  // median: elements k and k-1 (kLo = kHi when the even rule does not apply)
    uint32_t dHi, dLo;  // doubled deviations |2 (key - bias) - med2|: elements k and k-1
    bool ok;
};

// Rank of list entry e among n values, by GROUP consecutive threads (a power of two <= 32): each counts the entries
// sub, sub + GROUP, ... that sort before entry e (ties by slot), a shuffle tree adds the counts.  transformed: the list
// holds keys, the values are their doubled deviations from med2.
template <int GROUP>
__device__ __forceinline__ uint32_t s4_group_rank(const uint32_t* list, uint32_t n, uint32_t e, uint32_t sub, bool transformed, int med2, uint32_t* mine)
{
    auto val = [&](uint32_t x) { return transformed ? (uint32_t)abs(2 * (int)(x - S4_BIAS) - med2) : x; };
    const bool live   = e < n;
    const uint32_t me = live ? val(list[e]) : 0xffffffffu;
    uint32_t cnt      = 0;
    if (live)
        for (uint32_t j = sub; j < n; j += GROUP) {
            const uint32_t x = val(list[j]);
            cnt += (x < me || (x == me && j < e)) ? 1u : 0u;
        }
#pragma unroll
    for (int o = GROUP / 2; o >= 1; o >>= 1) cnt += __shfl_xor_sync(S4_FULL, cnt, o);
    *mine = me;
    return cnt;
}

// after the gather barrier: ranks the two lists with all threads (model_select4.lists); two barriers
template <int NT, class SM>
__device__ __forceinline__ S4Out s4_lists(const S4Loc& r, uint32_t k, bool needPred, SM& sm, int parity)
{
    const int tid = threadIdx.x;
    S4Out o;
    o.ok = false, o.kHi = o.kLo = o.dHi = o.dLo = 0;
    const uint32_t* c = sm.cnt() + parity * 16;
    const uint32_t nM = c[5], nD = c[6];
    const uint32_t idxM = nM - r.hMb + r.rM;
    const bool okM      = nM >= 1 && nM <= (uint32_t)S4_CAPM && nD <= (uint32_t)S4_CAPD && idxM < nM && !(needPred && idxM == 0);
    if (okM) {
        constexpr int GM = 16;  // threads per entry
        for (uint32_t e0 = 0; e0 < nM; e0 += NT / GM) {
            uint32_t me;
            const uint32_t e    = e0 + (uint32_t)tid / GM;
            const uint32_t rank = s4_group_rank<GM>(sm.listM(), nM, e, (uint32_t)tid % GM, false, 0, &me);
            if ((tid % GM) == 0 && e < nM) {
                if (rank == idxM) sm.misc()[8] = me;
                if (rank + 1 == idxM) sm.misc()[9] = me;
            }
        }
    }
    __syncthreads();
    if (!okM) {
        __syncthreads();
        if (tid == 0) sm.cnt()[parity * 16 + 5] = 0, sm.cnt()[parity * 16 + 6] = 0;
        return o;
    }
    o.kHi          = sm.misc()[8];
    o.kLo          = needPred ? sm.misc()[9] : o.kHi;
    const int med2 = (int)(o.kHi - S4_BIAS) + (int)(o.kLo - S4_BIAS);
    const uint32_t tD = k - r.base;
    const bool okD    = k >= r.base && tD < nD && !(needPred && tD == 0);
    if (okD) {
        constexpr int GD = 8;
        for (uint32_t e0 = 0; e0 < nD; e0 += NT / GD) {
            uint32_t me;
            const uint32_t e    = e0 + (uint32_t)tid / GD;
            const uint32_t rank = s4_group_rank<GD>(sm.listD(), nD, e, (uint32_t)tid % GD, true, med2, &me);
            if ((tid % GD) == 0 && e < nD) {
                if (rank == tD) sm.misc()[10] = me;
                if (rank + 1 == tD) sm.misc()[11] = me;
            }
        }
    }
    __syncthreads();
    if (tid == 0) sm.cnt()[parity * 16 + 5] = 0, sm.cnt()[parity * 16 + 6] = 0;  // (read by everybody two barriers ago)
    if (!okD) return o;
    o.dHi = sm.misc()[10];
    o.dLo = needPred ? sm.misc()[11] : o.dHi;
    o.ok  = true;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// generic tier: exact k-th smallest (and its predecessor) of 28-bit keys produced by keyOf(i), radix passes of log2(NT)
// bits with atomics by every live key.  Slow and always right.
// ---------------------------------------------------------------------------------------------------------------
template <int AREA, int NT, class KeyOf, class SM>
__device__ __forceinline__ void s4_generic_select(KeyOf keyOf, bool vis, uint32_t k, bool needPred, SM& sm, uint32_t* outHi, uint32_t* outLo)
{
    constexpr int NW   = NT / 32;
    constexpr int BITS = NT == 512 ? 9 : (NT == 256 ? 8 : (NT == 128 ? 7 : 6));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t prefix = 0, mask = 0, kk = k;
#pragma unroll 1
    for (int shift = 28 - BITS;; shift -= BITS) {
        if (shift < 0) shift = 0;
        if (vis) {
#pragma unroll
            for (int i = 0; i < AREA; i++) {  // (static indices: the residuals stay in registers)
                const uint32_t key = keyOf(i);
                if ((key & mask) == prefix) atomicAdd(&sm.hist()[(key >> shift) & (NT - 1)], 1u);
            }
        }
        __syncthreads();
        const uint32_t h    = sm.hist()[tid];
        sm.hist()[tid]      = 0;
        const uint32_t incl = s4_warp_incl_scan(h, lane);
        if (lane == 31) sm.wtot()[warp] = incl;
        __syncthreads();
        const uint32_t wt = lane < NW ? sm.wtot()[lane] : 0u;
        const uint32_t wi = s4_warp_incl_scan(wt, lane);
        const uint32_t ci = __shfl_sync(S4_FULL, wi - wt, warp) + incl;  // keys of the pass in bins <= tid
        if (ci - h <= kk && kk < ci) {
            sm.misc()[12] = (uint32_t)tid;
            sm.misc()[13] = kk - (ci - h);
        }
        __syncthreads();
        const uint32_t bin = sm.misc()[12];
        kk                 = sm.misc()[13];
        prefix |= bin << shift;
        mask |= (uint32_t)(NT - 1) << shift;
        if (shift == 0) break;
    }
    *outHi = prefix;
    *outLo = prefix;
    if (needPred && kk == 0) {  // predecessor: the largest key below
        uint32_t m = 0;
        if (vis) {
#pragma unroll
            for (int i = 0; i < AREA; i++) {
                const uint32_t key = keyOf(i);
                if (key < prefix) m = max(m, key);
            }
        }
        m = __reduce_max_sync(S4_FULL, m);
        if (lane == 0 && m) atomicMax(&sm.misc()[0], m);
        __syncthreads();
        *outLo = sm.misc()[0];
        __syncthreads();
        if (tid == 0) sm.misc()[0] = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bracket tier: the previous evaluation predicts the median and the deviation to a few hundredths of an intensity unit
// (the common case once a level has started to converge).  No histogram at all: ONE pass counts the keys below the median
// bracket [mA, mB) and the keys in the MIDDLE range between the two deviation brackets
//     L = [mA - dB, mB - 1 - dA],  R = [mA + dA, mB - 1 + dB]      (every key whose deviation from ANY median in the
//                                                                     bracket lies in (dA, dB] is inside L or R)
// and pushes the few keys inside M, L, R on the key stacks; they are ranked exactly as lists.  The median is exact by
// construction (rank k - cM of the complete list of [mA, mB)).  The deviation is element k - mid of the L / R list, and it
// is VERIFIED: every middle key must deviate by no more than it, every key outside L and R by no less; then the rank
// arithmetic is a proof, otherwise the tier reports a miss and a wider tier takes over.
// ---------------------------------------------------------------------------------------------------------------
struct S4Bracket {
    uint32_t mA, wM;      // median bracket [mA, mA + wM)
    uint32_t lLo, rLo, wD;  // deviation brackets [lLo, lLo + wD), [rLo, rLo + wD)
    uint32_t midLo, wMid;   // middle range [midLo, midLo + wMid): between the end of L and the start of R
};

template <int AREA, int NT, class SM>
__device__ __forceinline__ uint32_t s4_fill_bracket(const float (&rs)[AREA], bool vis, const S4Bracket& b, SM& sm, int parity)
{
    const int lane = threadIdx.x & 31;
    uint32_t cM = 0, mid = 0, n = 0;
    uint32_t* col = sm.stack() + threadIdx.x;
    if (vis) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t key = s4_key(rs[i]);
            const uint32_t tM = key - b.mA, tL = key - b.lLo, tR = key - b.rLo, tC = key - b.midLo;
            cM += tM >> 31;
            mid += tC < b.wMid ? 1u : 0u;
            col[n * NT] = key;
            const bool in = (tM < b.wM) | (min(tL, tR) < b.wD);
            n += in ? 1u : 0u;
        }
    }
    cM                = __reduce_add_sync(S4_FULL, cM);
    mid               = __reduce_add_sync(S4_FULL, mid);
    const uint32_t nv = __popc(__ballot_sync(S4_FULL, vis));
    uint32_t* c       = sm.cnt() + parity * 16;
    if (lane == 0) {
        if (cM) atomicAdd(&c[0], cM);
        if (mid) atomicAdd(&c[1], mid);
        if (nv) atomicAdd(&c[3], nv);
    }
    // the keys on the stack -> the two candidate lists
    const uint32_t nmax = __reduce_max_sync(S4_FULL, n);
    for (uint32_t j = 0; j < nmax; j++) {
        if (j < n) {
            const uint32_t key = col[j * NT];
            if (key - b.mA < b.wM) {
                const uint32_t slot = atomicAdd(&c[5], 1u);
                if (slot < (uint32_t)S4_CAPM) sm.listM()[slot] = key;
            } else {
                const uint32_t slot = atomicAdd(&c[6], 1u);
                if (slot < (uint32_t)S4_CAPD) sm.listD()[slot] = key;
            }
        }
    }
    return n;
}

// returns 0 on success, else the reason of the miss; always three barriers; uses and then leaves the counters of `parity`
// dirty and clears those of the other parity (the caller flips the parity)
template <int AREA, int NT, class SM>
__device__ __forceinline__ int s4_bracket(const float (&rs)[AREA], bool vis, int nTotal, uint32_t m0, uint32_t d0, uint32_t hm, uint32_t hd,
                                          SM& sm, int parity, S4Out* out, S4Counters* cnOut, uint32_t* kOut, bool* needPredOut)
{
    const int tid = threadIdx.x;
    S4Bracket b;
    const uint32_t dA = d0 - hd, dB = d0 + hd;  // the caller guarantees d0 > hd + 2 hm + 2: M lies inside the middle range
    b.mA = m0 - hm, b.wM = 2u * hm + 1u;
    const uint32_t mB1 = b.mA + b.wM - 1u;       // last key of the median bracket
    b.lLo = b.mA - dB, b.rLo = b.mA + dA, b.wD = (mB1 - dA) - b.lLo + 1u;
    b.midLo = mB1 - dA + 1u, b.wMid = b.rLo - b.midLo;
    s4_fill_bracket<AREA, NT>(rs, vis, b, sm, parity);
    __syncthreads();
    const uint32_t* c = sm.cnt() + parity * 16;
    S4Counters cn;
    cn.below[0] = c[0], cn.below[1] = c[1], cn.below[2] = 0, cn.nvis = c[3], cn.flags = 0;
    const uint32_t nM = c[5], nD = c[6];
    *cnOut = cn;
    const uint32_t k    = cn.nvis * (uint32_t)AREA / 2u;
    const bool needPred = !(nTotal & 1) && k > 0;
    *kOut = k, *needPredOut = needPred;
    if (tid < 16) sm.cnt()[(parity ^ 1) * 16 + tid] = 0;
    S4Out o;
    o.ok = false, o.kHi = o.kLo = o.dHi = o.dLo = 0;
    *out = o;
    int why = 0;
    // ---- median: element k - cM of the complete list of the bracket ----
    const uint32_t idxM = k - cn.below[0];
    if (cn.nvis == 0)
        why = 15;
    else if (nM > (uint32_t)S4_CAPM || nD > (uint32_t)S4_CAPD)
        why = 13;
    else if (k < cn.below[0] || idxM >= nM || (needPred && idxM == 0))
        why = 14;
    if (!why) {
        constexpr int GM = 16;
        for (uint32_t e0 = 0; e0 < nM; e0 += NT / GM) {
            uint32_t me;
            const uint32_t e    = e0 + (uint32_t)tid / GM;
            const uint32_t rank = s4_group_rank<GM>(sm.listM(), nM, e, (uint32_t)tid % GM, false, 0, &me);
            if ((tid % GM) == 0 && e < nM) {
                if (rank == idxM) sm.misc()[8] = me;
                if (rank + 1 == idxM) sm.misc()[9] = me;
            }
        }
    }
    __syncthreads();
    int med2 = 0;
    uint32_t tD = 0;
    if (!why) {
        o.kHi = sm.misc()[8];
        o.kLo = needPred ? sm.misc()[9] : o.kHi;
        med2  = (int)(o.kHi - S4_BIAS) + (int)(o.kLo - S4_BIAS);
        // ---- deviation: element k - mid of the L / R list ----
        tD = k - cn.below[1];
        if (k < cn.below[1] || tD >= nD || (needPred && tD == 0)) why = 14;
    }
    if (!why) {
        constexpr int GD = 8;
        for (uint32_t e0 = 0; e0 < nD; e0 += NT / GD) {
            uint32_t me;
            const uint32_t e    = e0 + (uint32_t)tid / GD;
            const uint32_t rank = s4_group_rank<GD>(sm.listD(), nD, e, (uint32_t)tid % GD, true, med2, &me);
            if ((tid % GD) == 0 && e < nD) {
                if (rank == tD) sm.misc()[10] = me;
                if (rank + 1 == tD) sm.misc()[11] = me;
            }
        }
    }
    __syncthreads();
    if (tid == 0) sm.cnt()[parity * 16 + 5] = 0, sm.cnt()[parity * 16 + 6] = 0;  // (read by everybody two barriers ago)
    if (why) return why;
    o.dHi = sm.misc()[10];
    o.dLo = needPred ? sm.misc()[11] : o.dHi;
    // ---- the proof: middle keys deviate by at most dLo, outside keys by at least dHi (doubled units) ----
    const int midLo2 = 2 * (int)(b.midLo - S4_BIAS), midHi2 = 2 * (int)(b.midLo + b.wMid - 1u - S4_BIAS);
    const int outL2 = 2 * (int)(b.lLo - 1u - S4_BIAS), outR2 = 2 * (int)(b.rLo + b.wD - S4_BIAS);
    const int dMidMax = max(med2 - midLo2, midHi2 - med2), dOutMin = min(med2 - outL2, outR2 - med2);
    if ((int)o.dLo < dMidMax || (int)o.dHi > dOutMin) return 14;
    o.ok = true;
    *out = o;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// driver: sigma = 1.4826 MAD of the keys of the CTA
// ---------------------------------------------------------------------------------------------------------------
struct S4Pred {       // carried from evaluation to evaluation (uniform over the CTA)
    uint32_t m, d;    // last median key, last deviation (key units)
    uint32_t moved;   // how far they moved at the last evaluation (key units)
    bool haveM, haveD, haveMove;
};

__device__ __forceinline__ int s4_ceil_log2(uint32_t x)  // smallest s with 2^s >= x
{
    return x <= 1 ? 0 : 32 - __clz(x - 1);
}

// Returns false when no feature is visible.  *tier: 1 hot (one atomic round + lists), 2 cold (private round(s), atomic
// round, lists), 4 generic.  nTotal = rows of the reference's residual vector (its parity picks the median rule).
template <int AREA, int NT, class SM>
__device__ __forceinline__ bool s4_sigma(const float (&rs)[AREA], bool vis, int nTotal, S4Pred& pr, SM& sm, int& parity, S4Out* res,
                                         uint32_t* nvisOut, int* tier)
{
    constexpr int LOGNT = NT == 512 ? 9 : (NT == 256 ? 8 : (NT == 128 ? 7 : 6));
    constexpr int SMAX  = 16 - LOGNT;  // hot windows are at most one intensity unit wide
    constexpr int SMIN  = 12 - LOGNT;  // and at least 1/16
    S4Out o;
    o.ok = false, o.kHi = o.kLo = o.dHi = o.dLo = 0;
    S4Counters cn;
    cn.below[0] = cn.below[1] = cn.below[2] = cn.nvis = cn.flags = 0;
    uint32_t k    = 0;
    bool needPred = false, known = false;
    *tier  = 1;
    sm.why = 0, sm.shiftUsed = 0;
#ifdef SVO_PROFILE
    sm.tlast = clock64();
#endif
    // the window tiers need LINEAR keys inside their windows: median within 16 units of zero, deviation below 30 units
    // (16 + 30 + margins: every window ends below 64 units); anything else is the generic tier's
    const uint32_t mOff = pr.m > S4_BIAS ? pr.m - S4_BIAS : S4_BIAS - pr.m;
    const bool inRange  = (!pr.haveM || mOff <= (1u << 20)) && (!pr.haveD || pr.d <= S4_DMAX - (1u << 18));
    // ---- bracket: the previous evaluation predicts both statistics to within 1/64 of an intensity unit ----
    bool triedBracket = false;
    if (pr.haveM && pr.haveD && pr.haveMove && inRange) {
        const uint32_t h = max(4u * min(pr.moved, 1u << 20), 512u);  // half width of both brackets
        if (h <= 1024u && pr.d > 3u * h + 4u) {  // (wider brackets hold more keys than the lists rank cheaply)
            triedBracket  = true;
            const int why = s4_bracket<AREA, NT>(rs, vis, nTotal, pr.m, pr.d, h, h, sm, parity, &o, &cn, &k, &needPred);
            parity ^= 1;
            known = true;
            S4_T(12);
            if (cn.nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            *tier = 0;
            if (why) sm.why = why;
        }
    }
    // ---- hot: windows from the previous evaluation ----
    if (!o.ok && pr.haveM && pr.haveD && pr.haveMove && inRange) {
        *tier = 1;
        uint32_t want = 4u * min(pr.moved, 1u << 24) + 64u;  // half width of the windows
        if (triedBracket) want = max(want, 16384u);         // (the brackets were too narrow: the movement is larger than predicted)
        const int s         = max(s4_ceil_log2(want) - (LOGNT - 1), SMIN);
        if (s <= SMAX) {
            const S4Win w    = s4_predicted(pr.m, pr.d, s, NT);
            sm.shiftUsed     = s;
            const uint32_t n = s4_fill_atomic<AREA, NT>(rs, vis, w, sm, parity);
            __syncthreads();
            S4_T(0);
            const S4Loc L = s4_locate<NT>(sm, w, parity, AREA, nTotal, &cn, &k, &needPred);
            parity ^= 1;
            known = true;
            S4_T(1);
            if (cn.nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            if (cn.flags & 1u)
                sm.why = 1;
            else if (!L.ok)
                sm.why = 2;
            else if (L.nCandM > (uint32_t)S4_CAPM || L.nCandD > (uint32_t)S4_CAPD)
                sm.why = 3;
            else {
                s4_gather_stack<NT>(n, s, L, sm, parity);
                __syncthreads();
                o = s4_lists<NT>(L, k, needPred, sm, parity);
                S4_T(3);
                if (!o.ok) sm.why = 4;
            }
        } else
            sm.why = 12;
    }
    // ---- cold: private round(s) over a contiguous span, then one atomic round over the candidate bins ----
    if (!o.ok && inRange) {
        *tier        = 2;
        uint32_t m0  = pr.haveM ? pr.m : S4_BIAS;
        uint32_t dhi = pr.haveD ? pr.d + (pr.d >> 1) + (pr.d >> 2) + 8192u : S4_DMAX;  // 1.75 x the last deviation; 30 units
        // (m0 within 16 units, dhi <= 30 units -> bins of at most one unit, span of 32 bins either side: inside +/- 64 units)
        bool full    = !pr.haveD;
#pragma unroll 1
        for (int attempt = 0; attempt < 4 && !o.ok; attempt++) {
            // span: M = 16 bins centred on m0, 24 bins on either side; the R window must reach m0 + dhi
            dhi           = min(dhi, S4_DMAX);
            const int s   = s4_ceil_log2((dhi + 29u) / 30u);
            const int gM  = (int)(m0 >> s) - 8;
            const S4Win w = s4_make_win(s, gM, 16, gM - 24, 24, gM + 16, 24);
            s4_fill_private<AREA, NT>(rs, vis, w.gL, s, sm, parity);
            __syncthreads();
            S4_T(4);
            s4_private_reduce<NT>(sm, parity);
            __syncthreads();
            S4_T(5);
            const S4Loc L = s4_locate<NT>(sm, w, parity, AREA, nTotal, &cn, &k, &needPred);
            parity ^= 1;
            known = true;
            S4_T(6);
            if (cn.nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            if (!L.ok) {
                sm.why = 5 | (sm.why << 4);
                if (full) break;  // not even the full range holds the targets: generic
                full = true, dhi = S4_DMAX;
                continue;
            }
            if (L.nCandM <= (uint32_t)S4_CAPM && L.nCandD <= (uint32_t)S4_CAPD) {  // few keys already: rank them
                s4_gather_regs<AREA, NT>(rs, vis, s, L, sm, parity);
                __syncthreads();
                o = s4_lists<NT>(L, k, needPred, sm, parity);
                S4_T(3);
                if (!o.ok) sm.why = 11 | (sm.why << 4);
                break;
            }
            if (L.nCandM + L.nCandD > (uint32_t)S4_ATOMIC_LIMIT && L.j1 + 1 <= 12 && s > 0) {
                // too many candidates for atomics, and a narrower span resolves them better: again, around the median bin
                m0   = (uint32_t)((((long long)L.bMin + L.bMax + 1) << s) >> 1);
                dhi  = (uint32_t)(L.j1 + 1) << s;
                full = false;
                continue;
            }
            S4Win w1;
            if (!s4_refine(L, s, NT, &w1)) {
                sm.why = 6 | (sm.why << 4);
                break;
            }
            sm.shiftUsed     = w1.s;
            const uint32_t n = s4_fill_atomic<AREA, NT>(rs, vis, w1, sm, parity);
            __syncthreads();
            S4_T(8);
            S4Counters c1;
            const S4Loc L1 = s4_locate<NT>(sm, w1, parity, AREA, nTotal, &c1, &k, &needPred);
            parity ^= 1;
            S4_T(9);
            if (c1.flags & 1u) {
                sm.why = 7 | (sm.why << 4);
                break;
            }
            if (!L1.ok || L1.nCandM > (uint32_t)S4_CAPM || L1.nCandD > (uint32_t)S4_CAPD) {
                sm.why = (L1.ok ? 8 : 9) | (sm.why << 4);
                break;
            }
            s4_gather_stack<NT>(n, w1.s, L1, sm, parity);
            __syncthreads();
            o = s4_lists<NT>(L1, k, needPred, sm, parity);
            S4_T(3);
            if (!o.ok) sm.why = 10 | (sm.why << 4);
            break;
        }
    }
    // ---- generic: two exact radix selects ----
    if (!o.ok) {
        *tier = 4;
        if (!known) {  // (no round has counted the visible features yet)
            const uint32_t nv = __popc(__ballot_sync(S4_FULL, vis));
            if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&sm.misc()[1], nv);
            __syncthreads();
            cn.nvis = sm.misc()[1];
            __syncthreads();
            if (threadIdx.x == 0) sm.misc()[1] = 0;
            k        = cn.nvis * (uint32_t)AREA / 2u;
            needPred = !(nTotal & 1) && k > 0;
            if (cn.nvis == 0) {
                *nvisOut = 0;
                return false;
            }
        }
        // keys of the generic tier: rint(rs) + 2^25 (27 bits), deviations |2 rint(rs) - med2| (28 bits)
        uint32_t gHi, gLo;
        s4_generic_select<AREA, NT>([&](int i) { return (uint32_t)(__float2int_rn(rs[i]) + (1 << 25)); }, vis, k, needPred, sm, &gHi, &gLo);
        const int med2 = ((int)gHi - (1 << 25)) + ((int)gLo - (1 << 25));
        s4_generic_select<AREA, NT>([&](int i) { return (uint32_t)abs(2 * __float2int_rn(rs[i]) - med2); }, vis, k, needPred, sm, &o.dHi, &o.dLo);
        o.kHi = gHi - (1u << 25) + S4_BIAS;
        o.kLo = gLo - (1u << 25) + S4_BIAS;
        o.ok  = true;
        S4_T(11);
    }
    // ---- prediction for the next evaluation ----
    {
        const uint32_t nm = o.kHi, nd = o.dHi >> 1;
        if (pr.haveM && pr.haveD) {
            const uint32_t mm = nm > pr.m ? nm - pr.m : pr.m - nm;
            const uint32_t md = nd > pr.d ? nd - pr.d : pr.d - nd;
            pr.moved          = max(mm, md);
            pr.haveMove       = true;
        }
        pr.m = nm, pr.d = nd, pr.haveM = pr.haveD = true;
    }
    *res     = o;
    *nvisOut = cn.nvis;
    return true;
}

}  // namespace
