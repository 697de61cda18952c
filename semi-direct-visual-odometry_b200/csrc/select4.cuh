// select4.cuh -- robust scale of the alignment (tukeyWeighting / computeSigma, src/optimizer.cpp:485-507, median rule
// src/algorithm.cpp:834-872, MEDIAN_EXACT of SURVEY 9.3) for ONE CTA of NT threads, each holding AREA residuals in
// registers: the median and the median absolute deviation are found TOGETHER.
//
// A ROUND histograms the keys (fixed point rint(r 2^16) + 2^25) over three windows of one global grid (bin = key >> s):
// M around the predicted median, L and R around median -/+ deviation.  From the three histograms and the number of keys
// below each window, `s4_locate` derives the bin(s) [bMin, bMax] of the median (and of its predecessor when the even
// rule needs it), bounds j0 2^s < d* <= j1 2^s on the k-th smallest deviation by counting the keys that MUST / CAN lie
// within j bins of any median in those bins, the candidate bins L' = [bMin - j1, bMax - j0], R' = [bMin + j0, bMax + j1]
// and `base`, the number of keys closer to the median than every candidate.  `s4_lists` then ranks the few keys of the
// candidate bins exactly (every warp redundantly, no further barrier).  tests/model_select4.py is the executable
// statement of this arithmetic; tests/test_model_select4.py pins it against sorted arrays on the CPU.
//
// Rounds:   atomic   NT bins per window, shared-memory atomics by the FEW keys inside the windows (they are pushed on a
//                    thread-private stack in shared memory by a branch-free pass over the registers); windows come
//                    from the previous evaluation ("hot") or from a coarse round.
//           private  one contiguous span of 64 bins counted with thread-private packed byte counters (no atomics: every
//                    key takes part), summed over the columns by one warp per counter row ("cold": no usable prediction).
//           generic  exact radix select over all bits with atomics by every key: the safety net (degenerate
//                    distributions, overflowing stacks or lists), never the common path.
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
constexpr uint32_t S4_BIAS    = 1u << 25;
constexpr int S4_STACK        = 16;  // key stack slots per thread (slot S4_STACK is the dump slot of the branch-free push)
constexpr int S4_CAPM         = 32;  // candidate list of the median: one key per lane
constexpr int S4_CAPD         = 64;  // candidate list of the deviation: two keys per lane
constexpr int S4_ATOMIC_LIMIT = 3072;  // a round with more keys inside its windows than this is counted privately instead

// shared-memory words of the selection for a CTA of NT threads
template <int NT>
__host__ __device__ constexpr size_t s4_smem_words()
{
    return (size_t)(S4_STACK + 1) * NT + 3 * NT + 3 * NT + 48 + 48 + 16 + 32 + 8 + S4_CAPM + S4_CAPD;
}

// The selection's shared memory lives at byte offset OFF of the kernel's dynamic shared memory.  Every view is derived
// from the `extern __shared__` symbol with compile-time offsets, never from a stored pointer: the compiler keeps the
// shared address space and emits LDS / STS / ATOMS (a pointer that has been through a struct or a call degrades to
// generic loads and generic atomics, which are several times slower).
template <int NT, int OFF>
struct S4Smem {
    int why;  // diagnostics: why the last evaluation left the hot / cold tier (0: it did not)
#ifdef SVO_PROFILE
    long long prof[16], tlast;  // cycles: 0 hot fill 1 hot scan 2 hot locate 3 lists 4 private fill 5 private reduce 6 scan 7 locate
                                //         8 refine + atomic fill 9 scan 10 locate 11 generic
#endif
    __device__ __forceinline__ static uint32_t* base()
    {
        extern __shared__ __align__(128) unsigned char s4_dynamic_smem[];
        return reinterpret_cast<uint32_t*>(s4_dynamic_smem + OFF);
    }
    // [S4_STACK + 1][NT] key stacks; the private counters [16][NT] of a coarse round live here too
    __device__ __forceinline__ static uint32_t* stack() { return base(); }
    // [3][NT] windows M, L, R; zero between rounds (the scan clears what it reads)
    __device__ __forceinline__ static uint32_t* hist() { return base() + (S4_STACK + 1) * NT; }
    // [3][NT] warp-local inclusive prefix of hist
    __device__ __forceinline__ static uint32_t* pin() { return hist() + 3 * NT; }
    // [3][16] warp totals
    __device__ __forceinline__ static uint32_t* wtot() { return pin() + 3 * NT; }
    // [3][16] exclusive warp bases (every warp writes the same values)
    __device__ __forceinline__ static uint32_t* wbs() { return wtot() + 48; }
    // [16] non-empty bins of the M window, one mask per warp
    __device__ __forceinline__ static uint32_t* nzm() { return wbs() + 48; }
    // [2][16] per round parity: 0..2 keys below M / L / R, 3 visible features, 4 flags; [5], [6] list fills
    __device__ __forceinline__ static uint32_t* cnt() { return nzm() + 16; }
    // [8] [0] max-below of the generic tier, [1] visible features of the generic tier
    __device__ __forceinline__ static uint32_t* misc() { return cnt() + 32; }
    __device__ __forceinline__ static uint32_t* listM() { return misc() + 8; }
    __device__ __forceinline__ static uint32_t* listD() { return listM() + S4_CAPM; }
    __device__ __forceinline__ static void clear()
    {
        uint32_t* b = base();
        for (int i = threadIdx.x; i < (int)s4_smem_words<NT>(); i += NT) b[i] = 0;
    }
};

struct S4Win {  // three windows on the grid of bins of width 2^s (starts in bins, may be negative), ordered L < M < R
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
    int gL       = ((int)(m0 - d0) >> s) - (nb - 1) / 2;  // arithmetic shift: m0 - d0 may be negative
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

__device__ __forceinline__ uint32_t s4_warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(S4_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// After the fill barrier: every thread owns bin `tid` of the three windows.  Publishes warp-local prefixes, warp totals
// and the non-empty mask, clears the bins it read and the counters of the OTHER round parity.  Returns the four round
// counters.  The caller synchronises afterwards.
// ---------------------------------------------------------------------------------------------------------------
struct S4Counters {
    uint32_t below[3];  // keys below the M, L, R windows
    uint32_t nvis;      // visible features
    uint32_t flags;     // 1: a key stack was full
};

template <int NT, class SM>
__device__ __forceinline__ S4Counters s4_scan(SM& sm, int parity)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* c = sm.cnt() + parity * 16;
    S4Counters counters;
    counters.below[0] = c[0], counters.below[1] = c[1], counters.below[2] = c[2], counters.nvis = c[3], counters.flags = c[4];
    uint32_t h[3];
#pragma unroll
    for (int X = 0; X < 3; X++) {
        h[X]                 = sm.hist()[X * NT + tid];
        sm.hist()[X * NT + tid] = 0;
    }
    const uint32_t nz = __ballot_sync(S4_FULL, h[0] != 0);
#pragma unroll
    for (int X = 0; X < 3; X++) {
        const uint32_t incl  = s4_warp_incl_scan(h[X], lane);
        sm.pin()[X * NT + tid] = incl;
        if (lane == 31) sm.wtot()[X * 16 + warp] = incl;
    }
    if (lane == 0) sm.nzm()[warp] = nz;
    if (tid < 16) sm.cnt()[(parity ^ 1) * 16 + tid] = 0;  // last read before the previous round's second barrier
    return counters;
}

// keys with grid bin < b; b must be an edge of the window the rule picks (guaranteed by the ranges of s4_locate)
template <int NT, class SM>
__device__ __forceinline__ uint32_t s4_cf(const SM& sm, const S4Win& w, const uint32_t below[3], int b)
{
    const int X = b >= w.gR ? 2 : (b >= w.gM ? 0 : 1);
    const int g = X == 2 ? w.gR : (X == 0 ? w.gM : w.gL);
    int i       = b - g;
    i           = min(max(i, 0), NT);  // (only out of range for probes whose result is discarded)
    uint32_t v  = X == 2 ? below[2] : (X == 0 ? below[0] : below[1]);  // (no dynamic index: stays in registers)
    if (i > 0) v += sm.wbs()[X * 16 + ((i - 1) >> 5)] + sm.pin()[X * NT + (i - 1)];
    return v;
}

// smallest j in [lo, hi] with f(j) >= thr (f monotone non-decreasing), hi + 1 if none: 32 probes at a time
template <class F>
__device__ __forceinline__ int s4_first_ge(int lo, int hi, uint32_t thr, F f)
{
    const int lane = threadIdx.x & 31;
    if (lo > hi) return hi + 1;
#pragma unroll 1
    while (hi - lo + 1 > 32) {
        const int n    = hi - lo + 1;
        const int step = (n + 31) >> 5;
        const int p    = min(lo + (lane + 1) * step - 1, hi);  // last element of this lane's segment
        const uint32_t m = __ballot_sync(S4_FULL, (int)f(p) >= (int)thr);
        if (m == 0) return hi + 1;
        const int seg = __ffs(m) - 1;
        lo            = lo + seg * step;
        hi            = min(lo + step - 1, hi);
    }
    const int p      = min(lo + lane, hi);
    const uint32_t m = __ballot_sync(S4_FULL, (lo + lane <= hi) && (int)f(p) >= (int)thr);
    return m ? lo + __ffs(m) - 1 : hi + 1;
}

// ---------------------------------------------------------------------------------------------------------------
// After the scan barrier: every warp derives the same S4Loc (model_select4.locate).
// ---------------------------------------------------------------------------------------------------------------
template <int NT, class SM>
__device__ __forceinline__ S4Loc s4_locate(SM& sm, const S4Win& w, const S4Counters& counters, uint32_t k, bool needPred)
{
    constexpr int NW = NT / 32;
    const int lane   = threadIdx.x & 31;
    S4Loc L;
    L.ok = false;
    L.bMin = L.bMax = 0, L.hMb = L.rM = 0, L.j0 = -1, L.j1 = 0, L.base = 0, L.Ll = L.Lh = L.Rl = L.Rh = 0, L.nCandM = L.nCandD = 0;
    const uint32_t below[3] = {counters.below[0], counters.below[1], counters.below[2]};
    // exclusive warp bases of the three windows
    uint32_t tot[3];
#pragma unroll
    for (int X = 0; X < 3; X++) {
        const uint32_t wt   = lane < NW ? sm.wtot()[X * 16 + lane] : 0u;
        const uint32_t incl = s4_warp_incl_scan(wt, lane);
        tot[X]              = __shfl_sync(S4_FULL, incl, 31);
        if (lane < 16) sm.wbs()[X * 16 + lane] = incl - wt;
    }
    __syncwarp();
    // ---- median: rank kM of the M window ----
    if (k < below[0]) return L;
    const uint32_t kM = k - below[0];
    if (kM >= tot[0]) return L;
    int iM;
    uint32_t exclM, inclM;
    {
        const uint32_t wincl = lane < NW ? sm.wbs()[lane] + sm.wtot()[lane] : 0xffffffffu;
        const int wM         = __popc(__ballot_sync(S4_FULL, lane < NW && wincl <= kM));
        const uint32_t wb    = sm.wbs()[wM];
        const uint32_t pi    = wb + sm.pin()[wM * 32 + lane];
        const int lM         = __popc(__ballot_sync(S4_FULL, pi <= kM));
        iM                   = wM * 32 + lM;
        inclM                = __shfl_sync(S4_FULL, pi, lM);
        const uint32_t prev  = __shfl_sync(S4_FULL, pi, (lM + 31) & 31);
        exclM                = lM == 0 ? wb : prev;
        L.bMax               = w.gM + iM;
        L.bMin               = L.bMax;
        L.rM                 = kM - exclM;
        L.hMb                = inclM - exclM;
        if (needPred && L.rM == 0) {  // predecessor: the last non-empty bin before iM
            const uint32_t mine = sm.nzm()[wM] & ((1u << lM) - 1u);
            int iP              = -1;
            if (mine)
                iP = wM * 32 + 31 - __clz(mine);
            else {
                const uint32_t z = lane < wM ? sm.nzm()[lane] : 0u;
                const int cand   = z ? lane * 32 + 31 - __clz(z) : -1;
                iP               = __reduce_max_sync(S4_FULL, cand);
            }
            if (iP < 0) return L;  // below the window
            L.bMin = w.gM + iP;
        }
    }
    // ---- deviation bracket ----
    const int bMin = L.bMin, bMax = L.bMax;
    auto Glo = [&](int j) { return s4_cf<NT>(sm, w, below, bMin + j) - s4_cf<NT>(sm, w, below, bMax - j + 1); };
    auto Ghi = [&](int j) { return s4_cf<NT>(sm, w, below, bMax + j + 1) - s4_cf<NT>(sm, w, below, bMin - j); };
    const int hi1 = min(w.gR + w.nbR - bMin, bMax + 1 - w.gL);
    const int hi0 = min(w.gR + w.nbR - bMax - 1, bMin - w.gL);
    int lo1 = 1, lo0 = 0;
    if (!w.contig) {
        lo1 = max(max(w.gR - bMin, bMax + 1 - w.gL - w.nbL), 1);
        lo0 = max(max(w.gR - bMax - 1, bMin - w.gL - w.nbL), 0);
    }
    const uint32_t kk = needPred ? k : k + 1;
    const int j1      = s4_first_ge(lo1, hi1, k + 1, Glo);
    if (j1 > hi1) return L;
    const int jz = s4_first_ge(lo0, min(hi0, j1), kk, Ghi);
    int j0       = jz - 1;
    if (j0 < lo0) {
        if (!w.contig) return L;  // the lower bound lies below the windows
        j0 = -1;
    }
    uint32_t base = 0;
    if (j0 >= 0) {
        if (j0 < lo1 && !w.contig) return L;
        const int g = (int)Glo(j0);
        base        = g > 0 ? (uint32_t)g : 0u;
    }
    L.j0 = j0, L.j1 = j1, L.base = base;
    L.Ll = bMin - j1, L.Lh = bMax - j0, L.Rl = bMin + j0, L.Rh = bMax + j1;
    if (w.contig) {
        if (L.Ll < w.gL || L.Rh >= w.gR + w.nbR) return L;
    } else if (L.Ll < w.gL || L.Lh >= w.gL + w.nbL || L.Rl < w.gR || L.Rh >= w.gR + w.nbR)
        return L;
    // keys in the candidate bins (exact): decides how the next step resolves them
    L.nCandM = s4_cf<NT>(sm, w, below, bMax + 1) - s4_cf<NT>(sm, w, below, bMin);
    if (L.Lh >= L.Rl - 1)
        L.nCandD = s4_cf<NT>(sm, w, below, L.Rh + 1) - s4_cf<NT>(sm, w, below, L.Ll);
    else
        L.nCandD = (s4_cf<NT>(sm, w, below, L.Lh + 1) - s4_cf<NT>(sm, w, below, L.Ll)) +
                   (s4_cf<NT>(sm, w, below, L.Rh + 1) - s4_cf<NT>(sm, w, below, L.Rl));
    L.ok = true;
    return L;
}

// windows of the next round: the candidate bins of this round plus S4_MARGIN bins on either side (the bounds of the finer
// round probe up to two bins beyond the targets), on the finest grid s2 <= s where every window has at most nb bins
// (model_select4.refine_windows).  Returns false if not even s does.
constexpr int S4_MARGIN = 4;
__device__ __forceinline__ bool s4_refine(const S4Loc& r, int s, int nb, S4Win* out)
{
#pragma unroll 1
    for (int s2 = 0; s2 <= s; s2++) {
        const int f = s - s2;
        if (f > 24) continue;  // (grid bins are < 2^27 >> s: 64-bit arithmetic below never overflows)
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

__device__ __forceinline__ uint32_t s4_key(float rs) { return (uint32_t)(__float2int_rn(rs) + (int)S4_BIAS); }

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
            col[n * NT] = key;
            const bool in = (tM < wdM) | (tL < wdL) | (tR < wdR);
            n             = min(n + (in ? 1u : 0u), (uint32_t)S4_STACK);
        }
    }
    uint32_t* c = sm.cnt() + parity * 16;
    if (n >= (uint32_t)S4_STACK) atomicOr(&c[4], 1u);  // (a full stack may have dropped keys)
    {
        const uint32_t nmax = __reduce_max_sync(S4_FULL, n);
        for (uint32_t j = 0; j < nmax; j++) {
            if (j < n) {
                const uint32_t key = col[j * NT];
                const uint32_t tM = key - loM, tL = key - loL, tR = key - loR;
                if (tM < wdM)
                    atomicAdd(&sm.hist()[tM >> s], 1u);
                else if (tL < wdL)
                    atomicAdd(&sm.hist()[NT + (tL >> s)], 1u);
                else if (tR < wdR)
                    atomicAdd(&sm.hist()[2 * NT + (tR >> s)], 1u);
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
            // keys below the M window = below the span + the L bins; below R = ... + the M bins
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

// element `idx` (and idx-1) of the ascending order of n <= 32 * PER values held PER per lane (v[p] belongs to list slot
// lane + 32 p; slots >= n are ignored).  Every lane counts the values smaller than its own.
template <int PER>
__device__ __forceinline__ void s4_rank(const uint32_t (&v)[PER], uint32_t n, uint32_t idx, bool wantPrev, const uint32_t* listSmem,
                                        bool transformed, int med2, uint32_t* outHi, uint32_t* outLo)
{
    const int lane = threadIdx.x & 31;
    uint32_t cnt[PER];
#pragma unroll
    for (int p = 0; p < PER; p++) cnt[p] = 0;
    for (uint32_t j = 0; j < n; j++) {
        uint32_t x = listSmem[j];  // broadcast
        if (transformed) x = (uint32_t)abs(2 * ((int)x - (int)S4_BIAS) - med2);
#pragma unroll
        for (int p = 0; p < PER; p++) {
            const uint32_t me = (uint32_t)lane + 32u * p;
            cnt[p] += (x < v[p] || (x == v[p] && j < me)) ? 1u : 0u;
        }
    }
    uint32_t hi = 0, lo = 0;
#pragma unroll
    for (int p = 0; p < PER; p++) {
        const uint32_t me = (uint32_t)lane + 32u * p;
        const bool live   = me < n;
        const uint32_t mh = __ballot_sync(S4_FULL, live && cnt[p] == idx);
        if (mh) hi = __shfl_sync(S4_FULL, v[p], __ffs(mh) - 1);
        const uint32_t ml = __ballot_sync(S4_FULL, live && wantPrev && cnt[p] + 1 == idx);
        if (ml) lo = __shfl_sync(S4_FULL, v[p], __ffs(ml) - 1);
    }
    *outHi = hi;
    *outLo = wantPrev ? lo : hi;
}

// after the gather barrier: every warp ranks the two lists (model_select4.lists)
template <int NT, class SM>
__device__ __forceinline__ S4Out s4_lists(const S4Loc& r, uint32_t k, bool needPred, SM& sm, int parity)
{
    const int lane = threadIdx.x & 31;
    S4Out o;
    o.ok = false, o.kHi = o.kLo = o.dHi = o.dLo = 0;
    const uint32_t* c = sm.cnt() + parity * 16;
    const uint32_t nM = c[5], nD = c[6];
    if (nM > (uint32_t)S4_CAPM || nD > (uint32_t)S4_CAPD || nM == 0) return o;
    const uint32_t idxM = nM - r.hMb + r.rM;
    if (idxM >= nM || (needPred && idxM == 0)) return o;
    {
        uint32_t v[1] = {lane < (int)nM ? sm.listM()[lane] : 0xffffffffu};
        s4_rank<1>(v, nM, idxM, needPred, sm.listM(), false, 0, &o.kHi, &o.kLo);
    }
    const int med2 = ((int)o.kHi - (int)S4_BIAS) + ((int)o.kLo - (int)S4_BIAS);
    if (k < r.base) return o;
    const uint32_t tD = k - r.base;
    if (tD >= nD || (needPred && tD == 0)) return o;
    {
        uint32_t v[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const uint32_t me = (uint32_t)lane + 32u * p;
            v[p]              = me < nD ? (uint32_t)abs(2 * ((int)sm.listD()[me] - (int)S4_BIAS) - med2) : 0xffffffffu;
        }
        s4_rank<2>(v, nD, tD, needPred, sm.listD(), true, med2, &o.dHi, &o.dLo);
    }
    o.ok = true;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// generic tier: exact k-th smallest (and its predecessor) of 28-bit keys produced by keyOf(i), radix passes of log2(NT)
// bits with atomics by every live key.  Slow and always right.
// ---------------------------------------------------------------------------------------------------------------
template <int AREA, int NT, class KeyOf, class SM>
__device__ __forceinline__ void s4_generic_select(KeyOf keyOf, bool vis, uint32_t k, bool needPred, SM& sm, int& parity, uint32_t* outHi,
                                               uint32_t* outLo)
{
    constexpr int NW   = NT / 32;
    constexpr int BITS = NT == 512 ? 9 : (NT == 256 ? 8 : (NT == 128 ? 7 : 6));
    const int lane     = threadIdx.x & 31;
    uint32_t prefix = 0, mask = 0, kk = k;
    uint32_t rankInKey = 0;
#pragma unroll 1
    for (int shift = 28 - BITS; ; shift -= BITS) {
        if (shift < 0) shift = 0;
        if (vis) {
#pragma unroll
            for (int i = 0; i < AREA; i++) {  // (static indices: the residuals stay in registers)
                const uint32_t key = keyOf(i);
                if ((key & mask) == prefix) atomicAdd(&sm.hist()[(key >> shift) & (NT - 1)], 1u);
            }
        }
        __syncthreads();
        s4_scan<NT>(sm, parity);
        __syncthreads();
        parity ^= 1;
        // every warp: locate kk in the M window's histogram
        const uint32_t wt   = lane < NW ? sm.wtot()[lane] : 0u;
        const uint32_t incl = s4_warp_incl_scan(wt, lane);
        const int wM        = __popc(__ballot_sync(S4_FULL, lane < NW && incl <= kk));
        const uint32_t wb   = __shfl_sync(S4_FULL, incl - wt, wM & 31);
        const uint32_t pi   = wb + sm.pin()[(wM & (NW - 1)) * 32 + lane];
        const int lM        = __popc(__ballot_sync(S4_FULL, pi <= kk));
        const uint32_t prev = __shfl_sync(S4_FULL, pi, (lM + 31) & 31);
        const uint32_t excl = lM == 0 ? wb : prev;
        const uint32_t bin  = (uint32_t)(wM * 32 + lM) & (NT - 1);
        prefix |= bin << shift;
        mask |= (uint32_t)(NT - 1) << shift;
        kk -= excl;
        rankInKey = kk;
        if (shift == 0) break;
    }
    *outHi = prefix;
    *outLo = prefix;
    if (needPred && rankInKey == 0) {  // predecessor: the largest key below
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
        if (threadIdx.x == 0) sm.misc()[0] = 0;
    }
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
    uint32_t nvis = 0, k = 0;
    bool needPred = false, known = false;
    *tier  = 1;
    sm.why = 0;
#ifdef SVO_PROFILE
    sm.tlast = clock64();
#endif
    auto set_counts = [&](const S4Counters& cn) {
        nvis     = cn.nvis;
        k        = nvis * (uint32_t)AREA / 2u;
        needPred = !(nTotal & 1) && k > 0;
        known    = true;
    };
    // ---- hot: windows from the previous evaluation ----
    if (pr.haveM && pr.haveD && pr.haveMove) {
        const uint32_t want = 4u * min(pr.moved, 1u << 24) + 64u;  // half width of the windows
        const int s         = max(s4_ceil_log2(want) - (LOGNT - 1), SMIN);
        if (s <= SMAX) {
            const S4Win w     = s4_predicted(pr.m, pr.d, s, NT);
            const uint32_t n  = s4_fill_atomic<AREA, NT>(rs, vis, w, sm, parity);
            __syncthreads();
            S4_T(0);
            const S4Counters cn = s4_scan<NT>(sm, parity);
            __syncthreads();
            S4_T(1);
            parity ^= 1;
            set_counts(cn);
            if (nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            if (!(cn.flags & 1u)) {
                const S4Loc L = s4_locate<NT>(sm, w, cn, k, needPred);
                S4_T(2);
                if (L.ok && L.nCandM <= (uint32_t)S4_CAPM && L.nCandD <= (uint32_t)S4_CAPD) {
                    s4_gather_stack<NT>(n, s, L, sm, parity);
                    __syncthreads();
                    o = s4_lists<NT>(L, k, needPred, sm, parity);
                    S4_T(3);
                    if (!o.ok) sm.why = 4;
                } else
                    sm.why = L.ok ? 3 : 2;
            } else
                sm.why = 1;
        } else
            sm.why = 12;
    }
    // ---- cold: private round(s) over a contiguous span, then one atomic round over the candidate bins ----
    if (!o.ok) {
        *tier             = 2;
        uint32_t m0       = pr.haveM ? pr.m : S4_BIAS;
        uint32_t dhi      = pr.haveD ? pr.d + (pr.d >> 1) + (pr.d >> 2) + 8192u : (1u << 21);  // 1.75 x the last deviation; 32 units
        bool full         = !pr.haveD;
#pragma unroll 1
        for (int attempt = 0; attempt < 4 && !o.ok; attempt++) {
            // span: M = 16 bins centred on m0, 24 bins on either side; the R window must reach m0 + dhi
            const int s   = s4_ceil_log2((dhi + 29u) / 30u);
            const int gM  = (int)(m0 >> s) - 8;
            const S4Win w = s4_make_win(s, gM, 16, gM - 24, 24, gM + 16, 24);
            s4_fill_private<AREA, NT>(rs, vis, w.gL, s, sm, parity);
            __syncthreads();
            S4_T(4);
            s4_private_reduce<NT>(sm, parity);
            __syncthreads();
            S4_T(5);
            const S4Counters cn = s4_scan<NT>(sm, parity);
            __syncthreads();
            S4_T(6);
            parity ^= 1;
            set_counts(cn);
            if (nvis == 0) {
                *nvisOut = 0;
                return false;
            }
            const S4Loc L = s4_locate<NT>(sm, w, cn, k, needPred);
            S4_T(7);
            if (!L.ok) {
                sm.why = 5 | (sm.why << 4);
                if (full) break;  // not even the full range holds the targets: generic
                full = true, dhi = 1u << 21;
                if (!pr.haveM) m0 = S4_BIAS;
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
                m0   = (uint32_t)(((long long)L.bMin + L.bMax + 1) << s >> 1);
                dhi  = (uint32_t)(L.j1 + 1) << s;
                full = false;
                continue;
            }
            S4Win w1;
            if (!s4_refine(L, s, NT, &w1)) {
                sm.why = 6 | (sm.why << 4);
                break;
            }
            const uint32_t n = s4_fill_atomic<AREA, NT>(rs, vis, w1, sm, parity);
            __syncthreads();
            S4_T(8);
            const S4Counters c1 = s4_scan<NT>(sm, parity);
            __syncthreads();
            S4_T(9);
            parity ^= 1;
            if (c1.flags & 1u) {
                sm.why = 7 | (sm.why << 4);
                break;
            }
            const S4Loc L1 = s4_locate<NT>(sm, w1, c1, k, needPred);
            S4_T(10);
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
            nvis = sm.misc()[1];
            __syncthreads();
            if (threadIdx.x == 0) sm.misc()[1] = 0;
            k        = nvis * (uint32_t)AREA / 2u;
            needPred = !(nTotal & 1) && k > 0;
            if (nvis == 0) {
                *nvisOut = 0;
                return false;
            }
        }
        s4_generic_select<AREA, NT>([&](int i) { return s4_key(rs[i]); }, vis, k, needPred, sm, parity, &o.kHi, &o.kLo);
        const int med2 = ((int)o.kHi - (int)S4_BIAS) + ((int)o.kLo - (int)S4_BIAS);
        s4_generic_select<AREA, NT>([&](int i) { return (uint32_t)abs(2 * ((int)s4_key(rs[i]) - (int)S4_BIAS) - med2); }, vis, k, needPred, sm,
                                    parity, &o.dHi, &o.dLo);
        o.ok = true;
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
    *nvisOut = nvis;
    return true;
}

}  // namespace
