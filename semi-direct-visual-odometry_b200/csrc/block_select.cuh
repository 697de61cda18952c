// block_select.cuh -- exact block-wide order statistics over 27-bit keys held in REGISTERS (AREA keys per thread,
// NT = 512 threads), for the robust scale of the alignment (tukeyWeighting / computeSigma, src/optimizer.cpp:485-507,
// src/algorithm.cpp:834-872).  Measured on B200: shared-memory atomics cost ~2 cycles per lane and the integer pipes
// issue at half rate, so a selection is priced by (sweeps over the 25 keys) x (instructions per key) and by the number
// of block barriers.  Three tiers, all exact on the keys, cheapest first:
//   hot      the result of the previous evaluation brackets this one.  Sweep A counts the keys below the bracket and
//            histograms (shared atomics, 512 bins of 2^shift) the few inside; sweep B resolves the chosen bin to single
//            keys in a window that also contains the predecessor.  Keys outside the bracket cost ~5 instructions and no
//            memory traffic.  Two barriers per sweep: every warp finds the target redundantly from the per-warp totals.
//   cold     one straight-line sweep with thread-private packed counters over 64 coarse bins of 2^16 keys, one over 64
//            sub-bins of 2^10 inside the chosen bin, then the hot machinery on exactly that sub-bin.
//   generic  targets in the clamped outer coarse bins: 4-pass MSD radix select over all 27 bits.
#pragma once
#include <stdint.h>

namespace {

constexpr int SEL_NT        = 512;
constexpr int SEL_NW        = SEL_NT / 32;
constexpr unsigned SEL_FULL = 0xffffffffu;
constexpr uint32_t SEL_NONE = 0xffffffffu;

struct SelSmem {
    uint32_t* priv;  // [16][NT] thread-private packed 8-bit counters (64 bins); zero between uses
    uint32_t* bins;  // [2][512] shared counters, ping-pong: the half a sweep uses is zero when it starts
    uint32_t* tot;   // [64] bin totals of a private pass
    uint32_t* wtot;  // [4][NW] per-warp partials
};
constexpr size_t SEL_SMEM_BYTES = (size_t)16 * SEL_NT * 4 + 2 * 512 * 4 + 64 * 4 + 4 * SEL_NW * 4;

struct Bracket {  // uniform per CTA (every thread holds the same values)
    uint32_t center;
    int shift;    // log2 of the bin width of sweep A; bracket = center -/+ (256 << shift)
    bool valid;
};

struct SelCtx {   // uniform per CTA
    SelSmem s;
    int pp;       // which half of s.bins the next sweep uses
};

// shared-memory counter increment predicated on t < bound: ONE predicated RED, no divergent branch around it
// (25 per-key branches cost ~750 cycles per sweep on B200)
__device__ __forceinline__ void red_shared_inc_if_below(uint32_t smem_addr, uint32_t t, uint32_t bound)
{
    asm volatile(
        "{\n .reg .pred p;\n setp.lt.u32 p, %1, %2;\n @p red.shared.add.u32 [%0], 1;\n}\n" ::"r"(smem_addr), "r"(t), "r"(bound)
        : "memory");
}

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(SEL_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// After a histogram of 512 bins was filled in `cur` and a barrier passed: every thread scans its bin inside its warp,
// publishes the warp total / last non-empty bin, a second barrier, then EVERY warp locates the bin holding rank kin
// (kin < total), the rank inside it and the last non-empty bin before it.  `other` (the idle half) is zeroed.
// prefixBefore: if >= 0, also returns in *prefixOut the number of entries in bins [0, prefixBefore).
__device__ __forceinline__ void locate_in_bins(const uint32_t* cur, uint32_t* other, const SelSmem& s, uint32_t kinBase, int prefixBefore,
                                               bool kinIsRelativeToPrefix, uint32_t* totalOut, uint32_t* binOut, uint32_t* rankOut,
                                               uint32_t* predBinOut, bool* found)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = cur[tid];
    other[tid]       = 0;
    const uint32_t wsum = __reduce_add_sync(SEL_FULL, c);
    const uint32_t nz   = __ballot_sync(SEL_FULL, c != 0);
    if (lane == 0) {
        s.wtot[warp]              = wsum;
        s.wtot[3 * SEL_NW + warp] = nz ? (uint32_t)(warp * 32 + 31 - __clz(nz)) : SEL_NONE;  // last non-empty bin of the warp
    }
    __syncthreads();
    const uint32_t wi    = lane < SEL_NW ? s.wtot[lane] : 0u;
    const uint32_t wlast = lane < SEL_NW ? s.wtot[3 * SEL_NW + lane] : SEL_NONE;
    const uint32_t wincl = warp_incl_scan(wi, lane);
    const uint32_t total = __shfl_sync(SEL_FULL, wincl, 31);
    *totalOut            = total;
    uint32_t kin         = kinBase;
    if (prefixBefore >= 0) {
        // entries in bins [0, prefixBefore): whole warps below + the partial warp
        const int pw = prefixBefore >> 5, pl = prefixBefore & 31;
        const uint32_t wholeBelow = pw > 0 ? __shfl_sync(SEL_FULL, wincl, (pw - 1) & 31) : 0u;
        const uint32_t cp         = pw < SEL_NW ? cur[pw * 32 + lane] : 0u;
        const uint32_t part       = __reduce_add_sync(SEL_FULL, lane < pl ? cp : 0u);
        if (kinIsRelativeToPrefix) kin += wholeBelow + part;
    }
    *found = kin < total;
    if (kin >= total) return;
    // target warp: first warp whose inclusive total exceeds kin
    const int tw          = __popc(__ballot_sync(SEL_FULL, lane < SEL_NW && wincl <= kin));
    const uint32_t wbase  = tw > 0 ? __shfl_sync(SEL_FULL, wincl, tw - 1) : 0u;
    const uint32_t cb     = cur[tw * 32 + lane];
    const uint32_t bincl  = warp_incl_scan(cb, lane) + wbase;
    const int tl          = __popc(__ballot_sync(SEL_FULL, bincl <= kin));
    const uint32_t tIncl  = __shfl_sync(SEL_FULL, bincl, tl);
    const uint32_t tCount = __shfl_sync(SEL_FULL, cb, tl);
    *binOut               = (uint32_t)(tw * 32 + tl);
    *rankOut              = kin - (tIncl - tCount);
    // last non-empty bin before the target: lower lanes of the target warp, else the previous warps' last bins
    const uint32_t nzb   = __ballot_sync(SEL_FULL, cb != 0) & ((1u << tl) - 1u);
    const uint32_t prevW = __ballot_sync(SEL_FULL, lane < tw && wlast != SEL_NONE);
    uint32_t pb          = SEL_NONE;
    if (nzb)
        pb = (uint32_t)(tw * 32 + 31 - __clz(nzb));
    else if (prevW)
        pb = __shfl_sync(SEL_FULL, wlast, 31 - __clz(prevW));
    *predBinOut = pb;
}

// Sweep A: 512 bins of width 2^shift from lo; also counts the keys below lo.  Returns 0 on a hit with (bin, rank in
// bin, previous non-empty bin), -1 / +1 when the k-th key lies below / above the bracket.
template <int AREA>
__device__ __forceinline__ int bracket_sweep_a(const uint32_t (&key)[AREA], bool live, uint32_t lo, int shift, int k, SelCtx& sc,
                                               uint32_t* binOut, uint32_t* rankOut, uint32_t* predBin)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* cur   = sc.s.bins + sc.pp * 512;
    uint32_t* other = sc.s.bins + (sc.pp ^ 1) * 512;
    sc.pp ^= 1;
    uint32_t below       = 0;
    const uint32_t width = 512u << shift;
    const uint32_t cur32 = (uint32_t)__cvta_generic_to_shared(cur);
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t t = key[i] - lo;  // keys below lo wrap to >= 2^31 (keys are < 2^27)
            below += t >> 31;
            red_shared_inc_if_below(cur32 + ((t >> shift) << 2), t, width);  // predicated, no branch per key
        }
    }
    below = __reduce_add_sync(SEL_FULL, below);
    if (lane == 0) sc.s.wtot[SEL_NW + warp] = below;
    __syncthreads();
    // (the below partials are read after the barrier inside locate_in_bins' second phase: load them here, after it)
    uint32_t total, bin = 0, rank = 0, pb = SEL_NONE;
    bool found;
    // rank of the target among the keys inside the bracket = k - (keys below lo): needs the block total of `below`,
    // which is available after the next barrier; locate_in_bins takes it through a callback-free two-step: compute
    // it from wtot[NW..] after its barrier.  To keep one barrier, the totals are summed by every warp here:
    //   (wtot[NW + w] was written before the barrier above)
    const uint32_t wb       = lane < SEL_NW ? sc.s.wtot[SEL_NW + lane] : 0u;
    const uint32_t totBelow = __reduce_add_sync(SEL_FULL, wb);
    const int kin           = k - (int)totBelow;
    locate_in_bins(cur, other, sc.s, kin < 0 ? SEL_NONE : (uint32_t)kin, -1, false, &total, &bin, &rank, &pb, &found);
    if (kin < 0) return -1;
    if (!found) return 1;
    *binOut  = bin;
    *rankOut = rank;
    *predBin = pb;
    return 0;
}

// Sweep B: the chosen bin [lo2, lo2 + W), W <= 128, resolved to single keys.  The 512 unit bins start up to 287 keys
// BELOW lo2, so that the predecessor of the target (needed for the even-count median rule) is normally inside the
// window too.  rankA = rank of the target among the keys >= lo2.
template <int AREA>
__device__ __forceinline__ void bracket_sweep_b(const uint32_t (&key)[AREA], bool live, uint32_t lo2, uint32_t rankA, SelCtx& sc,
                                                uint32_t* keyOut, uint32_t* rankOut, uint32_t* predOut, bool* hasPred)
{
    uint32_t* cur   = sc.s.bins + sc.pp * 512;
    uint32_t* other = sc.s.bins + (sc.pp ^ 1) * 512;
    sc.pp ^= 1;
    const uint32_t ws  = lo2 >= 256u ? ((lo2 - 256u) & ~31u) : 0u;  // window start
    const uint32_t off = lo2 - ws;                                   // 0 .. 287
    const uint32_t cur32 = (uint32_t)__cvta_generic_to_shared(cur);
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t t = key[i] - ws;
            red_shared_inc_if_below(cur32 + (t << 2), t, 512u);
        }
    }
    __syncthreads();
    uint32_t total, bin = 0, rank = 0, pb = SEL_NONE;
    bool found;
    locate_in_bins(cur, other, sc.s, rankA, (int)off, true, &total, &bin, &rank, &pb, &found);
    *keyOut  = ws + bin;
    *rankOut = rank;
    *hasPred = rank > 0 || pb != SEL_NONE;
    *predOut = rank > 0 ? ws + bin : ws + pb;
}

// largest key strictly below bound (0 if none)
template <int AREA>
__device__ __forceinline__ uint32_t block_max_below(const uint32_t (&key)[AREA], bool live, uint32_t bound, SelCtx& sc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t m = 0;
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) m = max(m, key[i] < bound ? key[i] : 0u);
    }
    m = __reduce_max_sync(SEL_FULL, m);
    __syncthreads();  // wtot may still be read by the previous step
    if (lane == 0) sc.s.wtot[2 * SEL_NW + warp] = m;
    __syncthreads();
    const uint32_t wm = lane < SEL_NW ? sc.s.wtot[2 * SEL_NW + lane] : 0u;
    return __reduce_max_sync(SEL_FULL, wm);
}

// Bracketed exact select.  On a hit: *keyOut = the k-th smallest key, *predOut = the (k-1)-th smallest (valid when
// needPred and k > 0).
template <int AREA>
__device__ __forceinline__ int bracket_select(const uint32_t (&key)[AREA], bool live, uint32_t lo, int shift, int k, bool needPred,
                                              SelCtx& sc, uint32_t* keyOut, uint32_t* predOut)
{
    uint32_t bin, rank, pbin;
    const int rc = bracket_sweep_a<AREA>(key, live, lo, shift, k, sc, &bin, &rank, &pbin);
    if (rc != 0) return rc;
    bool hasPred;
    if (shift == 0) {
        *keyOut  = lo + bin;
        hasPred  = rank > 0 || pbin != SEL_NONE;
        *predOut = rank > 0 ? lo + bin : lo + pbin;
    } else {
        uint32_t rank2;
        bracket_sweep_b<AREA>(key, live, lo + (bin << shift), rank, sc, keyOut, &rank2, predOut, &hasPred);
    }
    if (needPred && !hasPred) *predOut = block_max_below<AREA>(key, live, *keyOut, sc);  // rare: predecessor far below
    return 0;
}

// One pass with thread-private packed counters over 64 bins: digit(key) must be < 64 for participating keys; keys
// for which match(key) is false are skipped.  Two barriers; every warp scans the 64 totals redundantly.
// Returns the bin of rank k among the participating keys, *below = participating keys in lower bins.
template <int AREA, class DigitFn>
__device__ __forceinline__ uint32_t private_pass(const uint32_t (&key)[AREA], bool live, DigitFn digit, uint32_t k, SelCtx& sc,
                                                 uint32_t* below)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SelSmem& s = sc.s;
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t d = digit(key[i]);  // >= 64: not participating
            if (d < 64u) s.priv[(d >> 2) * SEL_NT + tid] += 1u << ((d & 3u) * 8u);
        }
    }
    __syncthreads();
    {   // warp w reduces word row w (bins 4w .. 4w+3) over all NT columns
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < SEL_NT / 32; j++) {
            const uint32_t wv                     = s.priv[warp * SEL_NT + lane + 32 * j];
            s.priv[warp * SEL_NT + lane + 32 * j] = 0;
            lo += wv & 0x00ff00ffu;
            hi += (wv >> 8) & 0x00ff00ffu;
        }
        lo = __reduce_add_sync(SEL_FULL, lo);
        hi = __reduce_add_sync(SEL_FULL, hi);
        if (lane == 0) {
            s.tot[warp * 4 + 0] = lo & 0xffffu;
            s.tot[warp * 4 + 1] = hi & 0xffffu;
            s.tot[warp * 4 + 2] = lo >> 16;
            s.tot[warp * 4 + 3] = hi >> 16;
        }
    }
    __syncthreads();
    const uint32_t c0 = s.tot[2 * lane], c1 = s.tot[2 * lane + 1];
    const uint32_t sum  = c0 + c1;
    const uint32_t incl = warp_incl_scan(sum, lane);
    const uint32_t excl = incl - sum;
    uint32_t mine       = SEL_NONE;
    if (k >= excl && k < excl + c0)
        mine = 2 * lane;
    else if (k >= excl + c0 && k < incl)
        mine = 2 * lane + 1;
    const uint32_t bin = __reduce_min_sync(SEL_FULL, mine);
    *below             = __shfl_sync(SEL_FULL, (bin & 1u) ? excl + c0 : excl, (int)((bin >> 1) & 31u));
    return bin;
}

struct DigitCoarse {  // 64 bins of 2^16 keys starting at base << 16, clamped: every key participates
    int base;
    __device__ __forceinline__ uint32_t operator()(uint32_t kk) const { return (uint32_t)min(max((int)(kk >> 16) - base, 0), 63); }
};
struct DigitRange {  // keys in [lo, lo + (64 << shift)) participate; digit = (key - lo) >> shift
    uint32_t lo;
    int shift;
    __device__ __forceinline__ uint32_t operator()(uint32_t kk) const { return (kk - lo) >> shift; }  // >= 64: outside
};

// one generic pass over `bits` (<= 9) bits with the shared atomic counters
template <int AREA>
__device__ __forceinline__ void generic_pass_atomic(const uint32_t (&key)[AREA], bool live, uint32_t& prefix, uint32_t& mask, uint32_t& k,
                                                    int shift, int bits, SelCtx& sc)
{
    uint32_t* cur   = sc.s.bins + sc.pp * 512;
    uint32_t* other = sc.s.bins + (sc.pp ^ 1) * 512;
    sc.pp ^= 1;
    const uint32_t dm = (1u << bits) - 1u;
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++)
            if ((key[i] & mask) == prefix) atomicAdd(&cur[(key[i] >> shift) & dm], 1u);
    }
    __syncthreads();
    uint32_t total, bin = 0, rank = 0, pb;
    bool found;
    locate_in_bins(cur, other, sc.s, k, -1, false, &total, &bin, &rank, &pb, &found);
    prefix |= bin << shift;
    mask |= dm << shift;
    k = rank;
}

// generic tier: k-th smallest over all 27 bits (6 + 6 private, 9 + 6 atomic); *rankInKey = rank among equal keys
template <int AREA>
__device__ __forceinline__ uint32_t generic_select27(const uint32_t (&key)[AREA], bool live, uint32_t k, SelCtx& sc, uint32_t* rankInKey)
{
    uint32_t prefix = 0, mask = 0, below;
    uint32_t b = private_pass<AREA>(key, live, DigitRange{0u, 21}, k, sc, &below);
    k -= below;
    prefix |= b << 21;
    mask |= 63u << 21;
    b = private_pass<AREA>(key, live, DigitRange{prefix, 15}, k, sc, &below);
    k -= below;
    prefix |= b << 15;
    mask |= 63u << 15;
    generic_pass_atomic<AREA>(key, live, prefix, mask, k, 6, 9, sc);
    generic_pass_atomic<AREA>(key, live, prefix, mask, k, 0, 6, sc);
    *rankInKey = k;
    return prefix;
}

// k-th smallest key and (needPred) its predecessor, through the tiers.  coarseBase: the cold tier's 64 coarse bins
// start at key (coarseBase << 16).  *tier: 1 hot, 2 cold, 4 generic.
template <int AREA>
__device__ __forceinline__ uint32_t tiered_select(const uint32_t (&key)[AREA], bool live, int coarseBase, int k, bool needPred, Bracket& br,
                                                  SelCtx& sc, uint32_t* predOut, int* tier)
{
    uint32_t kOut = 0, pred = 0, lo = 0;
    int shift     = 0;
    bool have     = br.valid;
    if (have) {
        const uint32_t half = 256u << br.shift;
        lo                  = br.center > half ? br.center - half : 0u;
        shift               = br.shift;
    }
    *tier = 1;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!have) {
            *tier = 2;
            uint32_t below;
            const uint32_t b = private_pass<AREA>(key, live, DigitCoarse{coarseBase}, (uint32_t)k, sc, &below);
            if (b == 0 || b == 63) {  // clamped outer bins
                *tier = 4;
                uint32_t rank;
                kOut = generic_select27<AREA>(key, live, (uint32_t)k, sc, &rank);
                pred = kOut;
                if (needPred && rank == 0) pred = block_max_below<AREA>(key, live, kOut, sc);
                break;
            }
            // 64 sub-bins of 2^10 keys inside the coarse bin, leaving a few dozen keys for the atomic sweeps
            const uint32_t pfx = ((uint32_t)((int)b + coarseBase)) << 16;
            uint32_t below2;
            const uint32_t b2 = private_pass<AREA>(key, live, DigitRange{pfx, 10}, (uint32_t)k - below, sc, &below2);
            lo                = pfx | (b2 << 10);
            shift             = 1;  // 512 bins of 2 keys = the sub-bin
        }
        if (bracket_select<AREA>(key, live, lo, shift, k, needPred, sc, &kOut, &pred) == 0) break;
        have = false;  // hot miss: go cold
    }
    // next bracket: centred on this result, half-width >= 4x the last movement, at most +/- 2^15 keys (1/2 intensity
    // unit; measured best on B200 among 2^11 .. 2^15 -- wider brackets pay ~2 cycles per key inside them)
    const uint32_t moved = br.valid ? (kOut > br.center ? kOut - br.center : br.center - kOut) : 0u;
    int sh               = 4;
    if (br.valid) {
        const uint32_t want = 4u * min(moved, 1u << 20) + 64u;
        sh                  = 0;
        while ((256u << sh) < want && sh < 8) sh++;
    }
    br.valid  = sh <= 7;
    br.shift  = min(sh, 7);
    br.center = kOut;
    *predOut  = pred;
    return kOut;
}

}  // namespace
