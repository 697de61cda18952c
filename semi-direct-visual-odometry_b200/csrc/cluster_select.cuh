// cluster_select.cuh -- exact order statistics over 27-bit keys held in REGISTERS (AREA keys per thread) by the NT
// threads of each of the C CTAs of a thread-block cluster, for the robust scale of the alignment (tukeyWeighting /
// computeSigma, src/optimizer.cpp:485-507, src/algorithm.cpp:834-872).
//
// Every step is a ROUND: the CTAs fill zeroed round buffers in their own shared memory (a 512-bin histogram through
// shared atomics, 64 totals of thread-private counters, three scalars), meet at ONE barrier (barrier.cluster for
// C > 1, bar.sync for C = 1), then every CTA sums the C copies through distributed shared memory and every warp
// locates the target redundantly -- no second barrier, identical results in all CTAs.  Round buffers rotate through
// three copies: the copy of round r+1 is cleared before the barrier of round r, when its last readers (round r-2)
// are provably done.
//
// Tiers (all exact on the keys), cheapest first:
//   hot      the result of the previous evaluation brackets this one.  Sweep A counts the keys below the bracket
//            and histograms the keys inside (512 bins of 2^shift); sweep B resolves the chosen bin to single keys in
//            a window that also holds the predecessor (needed by the even-count median rule).
//   cold     one sweep with thread-private packed counters over 64 coarse bins of 2^16 keys, then the hot machinery
//            on that coarse bin (512 bins of 2^7).
//   generic  targets in the clamped outer coarse bins: three 9-bit MSD radix passes over all 27 bits.
#pragma once
#include <stdint.h>

namespace {

constexpr unsigned CS_FULL = 0xffffffffu;
constexpr uint32_t CS_NONE = 0xffffffffu;
constexpr int CS_BINS      = 512;
constexpr int CS_XS        = 80;  // small words per round: [0] below [1] aux [2] max, [8..71] private totals

__device__ __forceinline__ uint32_t cs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cs_mapa(uint32_t addr, uint32_t cta)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ uint32_t cs_ld_cluster(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double cs_ld_cluster_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
// sum of the same shared-memory word / double in the CC CTAs of the cluster: all loads are issued before the first
// use (a dependent add after each load would serialise the ~200-cycle DSMEM round trips on the in-order pipeline)
template <int CC>
__device__ __forceinline__ uint32_t cs_gather_sum_cc(uint32_t la)
{
    uint32_t v[CC];
#pragma unroll
    for (int c = 0; c < CC; c++) v[c] = cs_ld_cluster(cs_mapa(la, (uint32_t)c));
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CC; c++) s += v[c];
    return s;
}
template <int CC>
__device__ __forceinline__ uint2 cs_gather_sum2_cc(uint32_t la)  // two consecutive words
{
    uint32_t v[CC], w[CC];
#pragma unroll
    for (int c = 0; c < CC; c++) {
        const uint32_t ra = cs_mapa(la, (uint32_t)c);
        v[c]              = cs_ld_cluster(ra);
        w[c]              = cs_ld_cluster(ra + 4);
    }
    uint2 s = make_uint2(0u, 0u);
#pragma unroll
    for (int c = 0; c < CC; c++) {
        s.x += v[c];
        s.y += w[c];
    }
    return s;
}
template <int CC>
__device__ __forceinline__ double cs_gather_sum_f64_cc(uint32_t la)  // fixed order 0 .. CC-1: identical in every CTA
{
    double v[CC];
#pragma unroll
    for (int c = 0; c < CC; c++) v[c] = cs_ld_cluster_f64(cs_mapa(la, (uint32_t)c));
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CC; c++) s += v[c];
    return s;
}
__device__ __forceinline__ uint32_t cs_gather_sum(uint32_t la, int C)
{
    return C == 2 ? cs_gather_sum_cc<2>(la) : (C == 4 ? cs_gather_sum_cc<4>(la) : cs_gather_sum_cc<8>(la));
}
__device__ __forceinline__ uint2 cs_gather_sum2(uint32_t la, int C)
{
    return C == 2 ? cs_gather_sum2_cc<2>(la) : (C == 4 ? cs_gather_sum2_cc<4>(la) : cs_gather_sum2_cc<8>(la));
}
__device__ __forceinline__ double cs_gather_sum_f64(uint32_t la, int C)
{
    return C == 2 ? cs_gather_sum_f64_cc<2>(la) : (C == 4 ? cs_gather_sum_f64_cc<4>(la) : cs_gather_sum_f64_cc<8>(la));
}

// Cluster-wide barrier with shared-memory visibility.  A release by every warp costs a cluster-scope fence per warp
// (the `membar` stall was 10 % of the samples at C = 2); instead the CTA meets at bar.sync, ONE warp issues the
// cluster-scope fence on behalf of the writes it has observed through that barrier (fence cumulativity), and all
// threads arrive relaxed; the wait keeps its acquire.
__device__ __forceinline__ void cs_cluster_sync()
{
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("fence.acq_rel.cluster;" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cs_cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cs_cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

template <int NT>
__host__ __device__ constexpr size_t cs_smem_bytes()
{
    return ((size_t)16 * NT + 4 * CS_BINS + 32 + 3 * CS_XS) * 4;
}

#ifdef SVO_PROFILE
#define CS_T(i)                          \
    do {                                 \
        const long long t__ = clock64(); \
        sc.prof[i] += t__ - sc.tlast;    \
        sc.tlast = t__;                  \
    } while (0)
#else
#define CS_T(i)
#endif

struct Bracket {  // uniform over the cluster
    uint32_t center;  // last result
    int shift;        // log2 of the bin width of sweep A; bracket = center -/+ (256 << shift)
    bool valid;       // try the bracket first
    bool have;        // center holds a result of this level
};

template <int NT>
struct SelCtx {  // uniform over the cluster
    uint32_t* priv;  // [16][NT] per thread 16 words: packed 8-bit counters of a private pass, or the key stack
    uint32_t* bins;  // [3][512] round histograms
    uint32_t* tot;   // [512]    cluster-wide histogram of the round (C > 1)
    uint32_t* wtot;  // [32]     per-warp totals / last non-empty bins of cs_locate
    uint32_t* xs;    // [3][CS_XS] round scalars
    int C;           // CTAs per cluster
    int cur;         // round buffer in use
    uint32_t tBelow, tAux, tMax;  // cluster totals of the last finished round
    uint32_t auxTotal;            // aux total of the round that fixed a rank (kFromAux)
#ifdef SVO_PROFILE
    long long stat[20];           // hot attempts by shift [0..7], hits by shift [8..15]; first evaluation of a level with a
                                  // carried bracket: [16] median attempts [17] hits [18] MAD attempts [19] hits
    uint32_t lastInside;
    long long prof[12], tlast;    // cycles: 0 push 1 pop A 2 finish A 3 locate A 4 pop B 5 finish B 6 locate B 7 private count
                                  //         8 private reduce 9 private finish + scan 10 other
#endif

    __device__ __forceinline__ void init(unsigned char* base, int clusterSize)
    {
        priv = reinterpret_cast<uint32_t*>(base);
        bins = priv + 16 * NT;
        tot  = bins + 3 * CS_BINS;
        wtot = tot + CS_BINS;
        xs   = wtot + 32;
        C    = clusterSize;
        cur  = 0;
        for (int i = threadIdx.x; i < 16 * NT + 4 * CS_BINS + 32 + 3 * CS_XS; i += NT) priv[i] = 0;
    }
    __device__ __forceinline__ uint32_t* rbins() const { return bins + cur * CS_BINS; }
    __device__ __forceinline__ uint32_t* rxs() const { return xs + cur * CS_XS; }

    // contribution to the `aux` scalar of the NEXT round to finish (the alignment passes the visible-feature count)
    __device__ __forceinline__ void add_aux(uint32_t warpTotal)
    {
        if ((threadIdx.x & 31) == 0 && warpTotal) atomicAdd(&rxs()[1], warpTotal);
    }

    // End of the fill phase of a round: clear the next round's buffers, meet, total the scalars.  Returns this CTA's
    // histogram of the round (cs_locate adds the copies of the other CTAs).  Afterwards `cur` names the next round.
    __device__ __forceinline__ const uint32_t* finish_round()
    {
        const int tid = threadIdx.x, lane = tid & 31;
        const int nxt = cur == 2 ? 0 : cur + 1;
        uint32_t* nb  = bins + nxt * CS_BINS;
#pragma unroll
        for (int i = 0; i < CS_BINS / NT; i++) nb[tid + i * NT] = 0;
        if (NT > CS_BINS && tid < CS_BINS) nb[tid] = 0;
        for (int i = tid; i < CS_XS; i += NT) xs[nxt * CS_XS + i] = 0;
        const uint32_t* hist = rbins();
        const uint32_t* x    = rxs();
        if (C == 1) {
            __syncthreads();
            tBelow = x[0];
            tAux   = x[1];
            tMax   = x[2];
        } else {
            cs_cluster_sync();
            uint32_t b = 0, a = 0, m = 0;
            if (lane < C) {
                const uint32_t ra = cs_mapa(cs_smem_u32(x), (uint32_t)lane);
                b = cs_ld_cluster(ra);
                a = cs_ld_cluster(ra + 4);
                m = cs_ld_cluster(ra + 8);
            }
            tBelow = __reduce_add_sync(CS_FULL, b);
            tAux   = __reduce_add_sync(CS_FULL, a);
            tMax   = __reduce_max_sync(CS_FULL, m);
        }
        cur = nxt;
        return hist;
    }
};

// shared-memory counter increment predicated on t < bound
__device__ __forceinline__ void cs_red_inc_if_below(uint32_t smem_addr, uint32_t t, uint32_t bound)
{
    asm volatile(
        "{\n .reg .pred p;\n setp.lt.u32 p, %1, %2;\n @p red.shared.add.u32 [%0], 1;\n}\n" ::"r"(smem_addr), "r"(t), "r"(bound)
        : "memory");
}

__device__ __forceinline__ uint32_t cs_warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(CS_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

struct Located {
    uint32_t total, bin, rank, pred;  // pred: last non-empty bin before `bin` (CS_NONE if none)
    bool found;
};

__device__ __forceinline__ uint4 cs_ld_cluster_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
template <int CC>
__device__ __forceinline__ uint4 cs_gather_sum4_cc(uint32_t la)
{
    uint4 v[CC];
#pragma unroll
    for (int c = 0; c < CC; c++) v[c] = cs_ld_cluster_v4(cs_mapa(la, (uint32_t)c));
    uint4 s = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c = 0; c < CC; c++) {
        s.x += v[c].x;
        s.y += v[c].y;
        s.z += v[c].z;
        s.w += v[c].w;
    }
    return s;
}

// Locate rank kin in the 512-bin histogram of the round (summed over the C CTAs through distributed shared memory):
// the bin holding it, the rank inside the bin and the last non-empty bin before it.  Two levels, one barrier:
// thread t totals bins [t BPT, (t+1) BPT) (BPT = 512 / NT) and publishes the cluster sums, every warp publishes its
// total and its last non-empty bin; after the barrier EVERY warp scans the NT/32 warp totals and the 32 thread totals of
// the target warp redundantly -- short dependent chains, no second barrier, identical results in all CTAs.
// prefixBefore >= 0: kin is relative to the entries in bins [0, prefixBefore).
template <int NT>
__device__ __forceinline__ Located cs_locate(const uint32_t* hist, uint32_t kin, int prefixBefore, SelCtx<NT>& sc)
{
    constexpr int NW  = NT / 32;
    constexpr int BPT = CS_BINS / NT;
    static_assert(BPT == 1 || BPT == 2 || BPT == 4 || BPT == 8, "NT must be 64 .. 512");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- level 1: this thread's bins ----
    uint32_t c[BPT];
    if (sc.C == 1) {
#pragma unroll
        for (int j = 0; j < BPT; j++) c[j] = hist[tid * BPT + j];
    } else {
        const uint32_t la = cs_smem_u32(hist + tid * BPT);
        if (BPT == 1) {
            c[0] = cs_gather_sum(la, sc.C);
        } else {
#pragma unroll
            for (int j = 0; j < BPT; j += 2) {
                const uint2 q = cs_gather_sum2(la + 4u * j, sc.C);
                c[j]          = q.x;
                c[j + 1]      = q.y;
            }
        }
#pragma unroll
        for (int j = 0; j < BPT; j++) sc.tot[tid * BPT + j] = c[j];
    }
    const uint32_t* T = sc.C == 1 ? hist : sc.tot;  // the cluster-wide histogram, local after the barrier
    uint32_t ssum = 0, lastp1 = 0;                  // lastp1: 1 + last non-empty bin of the thread, 0 if none
#pragma unroll
    for (int j = 0; j < BPT; j++) {
        ssum += c[j];
        if (c[j]) lastp1 = (uint32_t)(tid * BPT + j + 1);
    }
    const uint32_t wsum  = __reduce_add_sync(CS_FULL, ssum);
    const uint32_t wlast = __reduce_max_sync(CS_FULL, lastp1);
    if (lane == 0) {
        sc.wtot[warp]      = wsum;
        sc.wtot[16 + warp] = wlast;
    }
    __syncthreads();
    // ---- level 2: the warp totals ----
    const uint32_t wi    = lane < NW ? sc.wtot[lane] : 0u;
    const uint32_t wl    = lane < NW ? sc.wtot[16 + lane] : 0u;
    const uint32_t wincl = cs_warp_incl_scan(wi, lane);
    Located L;
    L.total = __shfl_sync(CS_FULL, wincl, 31);
    L.bin = 0, L.rank = 0, L.pred = CS_NONE;
    if (prefixBefore >= 0) {
        // entries in bins [0, prefixBefore): whole warps, whole threads of the partial warp, bins of the partial thread
        const int pt = prefixBefore / BPT, pe = prefixBefore % BPT;
        const int pw = pt >> 5, pl = pt & 31;
        const uint32_t wholeBelow = pw > 0 ? __shfl_sync(CS_FULL, wincl, (pw - 1) & 31) : 0u;
        uint32_t part = 0;
        if (pw < NW) {
            uint32_t ts = 0, tp = 0;  // thread (pw, lane): its total, and its first pe bins
#pragma unroll
            for (int j = 0; j < BPT; j++) {
                const uint32_t v = T[(pw * 32 + lane) * BPT + j];
                ts += v;
                tp += j < pe ? v : 0u;
            }
            part = __reduce_add_sync(CS_FULL, lane < pl ? ts : (lane == pl ? tp : 0u));
        }
        kin += wholeBelow + part;
    }
    L.found = kin < L.total;
    if (!L.found) return L;
    // target warp: first warp whose inclusive total exceeds kin
    const int tw         = __popc(__ballot_sync(CS_FULL, lane < NW && wincl <= kin));
    const uint32_t wbase = tw > 0 ? __shfl_sync(CS_FULL, wincl, tw - 1) : 0u;
    uint32_t tc[BPT], ts = 0, tlastp1 = 0;  // thread (tw, lane)
#pragma unroll
    for (int j = 0; j < BPT; j++) {
        tc[j] = T[(tw * 32 + lane) * BPT + j];
        ts += tc[j];
        if (tc[j]) tlastp1 = (uint32_t)((tw * 32 + lane) * BPT + j + 1);
    }
    const uint32_t tincl = cs_warp_incl_scan(ts, lane) + wbase;
    const int tl         = __popc(__ballot_sync(CS_FULL, tincl <= kin));  // target thread of the warp
    // inside the target thread: every lane evaluates its own bins, lane tl's answer counts
    const uint32_t kk = kin - (tincl - ts);
    uint32_t run = 0, be = 0, brk = 0, predIn = 0, seen = 0;
#pragma unroll
    for (int j = 0; j < BPT; j++) {
        if (kk >= run && kk < run + tc[j]) {
            be     = j;
            brk    = kk - run;
            predIn = seen;
        }
        run += tc[j];
        if (tc[j]) seen = (uint32_t)((tw * 32 + lane) * BPT + j + 1);
    }
    L.bin  = (uint32_t)((tw * 32 + tl) * BPT) + __shfl_sync(CS_FULL, be, tl);
    L.rank = __shfl_sync(CS_FULL, brk, tl);
    // last non-empty bin before the target: inside the thread, else lower threads of the warp, else lower warps
    uint32_t p1 = __shfl_sync(CS_FULL, predIn, tl);
    if (p1 == 0) p1 = __reduce_max_sync(CS_FULL, lane < tl ? tlastp1 : 0u);
    if (p1 == 0) p1 = __reduce_max_sync(CS_FULL, lane < tw ? wl : 0u);
    L.pred = p1 ? p1 - 1u : CS_NONE;
    return L;
}

__device__ __forceinline__ void cs_sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint32_t cs_lds_u16(uint32_t addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return (uint32_t)v;
}

// The keys of a thread that fall inside the bracket [lo, lo + 65536) at most, as 16-bit offsets from lo, in a
// thread-private stack of 32 slots in shared memory (the 16 words per thread of sc.priv, [slot][NT] halfwords).
// Registers cannot be indexed dynamically; the stack can, so the atomics of both sweeps loop over the FEW keys inside
// the bracket instead of branching around a predicated atomic for each of the AREA keys.
struct KeyStack {
    uint32_t base;  // shared-memory byte address of slot 0 of this thread
    uint32_t n;     // keys on the stack
    uint32_t lo;    // bracket start the offsets refer to
};

// Sweep A: 512 bins of width 2^shift from lo; also counts the keys below lo.  Returns 0 on a hit, -1 / +1 when the
// k-th key lies below / above the bracket.  kFromAux: k = (cluster total of aux) * AREA / 2, known after the barrier.
// Push pass: branch-free, one unconditional 16-bit store per key whose slot is kept only when the key is inside.
template <int AREA, int NT>
__device__ __forceinline__ int cs_sweep_a(const uint32_t (&key)[AREA], bool live, uint32_t lo, int shift, int& k, bool kFromAux,
                                          SelCtx<NT>& sc, Located* out, KeyStack& ks)
{
    static_assert(AREA <= 32, "the key stack holds 32 halfwords per thread");
    const int lane = threadIdx.x & 31;
    uint32_t below       = 0;
    const uint32_t width = 512u << shift;  // <= 65536: offsets fit 16 bits
    CS_T(10);
    ks.base = cs_smem_u32(sc.priv) + 2u * threadIdx.x;
    ks.lo   = lo;
    uint32_t ptr = ks.base;
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t t = key[i] - lo;  // keys below lo wrap to >= 2^31 (keys are < 2^27)
            below += t >> 31;
            cs_sts_u16(ptr, t);
            ptr += t < width ? 2u * NT : 0u;
        }
    }
    ks.n = (ptr - ks.base) / (2u * NT);
    CS_T(0);
    {   // histogram the keys on the stack
        uint32_t* cur        = sc.rbins();
        const uint32_t nmax = __reduce_max_sync(CS_FULL, ks.n);
        for (uint32_t j = 0; j < nmax; j++)
            if (j < ks.n) atomicAdd(&cur[cs_lds_u16(ks.base + j * 2u * NT) >> shift], 1u);
    }
    below = __reduce_add_sync(CS_FULL, below);
    if (lane == 0 && below) atomicAdd(&sc.rxs()[0], below);
    CS_T(1);
    const uint32_t* hist = sc.finish_round();
    CS_T(2);
    if (kFromAux) {
        sc.auxTotal = sc.tAux;
        k           = (int)(sc.tAux * (uint32_t)AREA / 2u);
    }
    const int kin = k - (int)sc.tBelow;
    if (kin < 0) return -1;
    *out = cs_locate<NT>(hist, (uint32_t)kin, -1, sc);
    CS_T(3);
#ifdef SVO_PROFILE
    sc.lastInside = out->total;
#endif
    return out->found ? 0 : 1;
}

// Sweep B: the chosen bin [lo2, lo2 + W), W <= 128, resolved to single keys.  The 512 unit bins start up to 287
// keys BELOW lo2, so that the predecessor of the target is normally inside the window too (keys below the bracket are
// not on the stack: a predecessor there is found by cs_max_below).  rankA = rank of the target among the keys >= lo2.
template <int NT>
__device__ __forceinline__ void cs_sweep_b(const KeyStack& ks, uint32_t lo2, uint32_t rankA, SelCtx<NT>& sc, uint32_t* keyOut,
                                           uint32_t* predOut, bool* hasPred)
{
    const uint32_t ws  = lo2 >= 256u ? ((lo2 - 256u) & ~31u) : 0u;  // window start
    const uint32_t off = lo2 - ws;                                   // 0 .. 287
    const uint32_t rel = ks.lo - ws;                                 // stack offset -> window offset (wraps when ws > lo)
    CS_T(10);
    {
        uint32_t* cur        = sc.rbins();
        const uint32_t nmax = __reduce_max_sync(CS_FULL, ks.n);
        for (uint32_t j = 0; j < nmax; j++) {
            if (j < ks.n) {
                const uint32_t u = cs_lds_u16(ks.base + j * 2u * NT) + rel;
                if (u < 512u) atomicAdd(&cur[u], 1u);
            }
        }
    }
    CS_T(4);
    const uint32_t* hist = sc.finish_round();
    CS_T(5);
    const Located L      = cs_locate<NT>(hist, rankA, (int)off, sc);
    CS_T(6);
    *keyOut  = ws + L.bin;
    *hasPred = L.rank > 0 || L.pred != CS_NONE;
    *predOut = L.rank > 0 ? ws + L.bin : ws + L.pred;
}

// largest key strictly below bound (0 if none)
template <int AREA, int NT>
__device__ __forceinline__ uint32_t cs_max_below(const uint32_t (&key)[AREA], bool live, uint32_t bound, SelCtx<NT>& sc)
{
    const int lane = threadIdx.x & 31;
    uint32_t m = 0;
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) m = max(m, key[i] < bound ? key[i] : 0u);
    }
    m = __reduce_max_sync(CS_FULL, m);
    if (lane == 0 && m) atomicMax(&sc.rxs()[2], m);
    sc.finish_round();
    return sc.tMax;
}

// Bracketed exact select.  On a hit: *keyOut = the k-th smallest key, *predOut = the (k-1)-th smallest (valid when
// needPred and k > 0).
template <int AREA, int NT>
__device__ __forceinline__ int cs_bracket_select(const uint32_t (&key)[AREA], bool live, uint32_t lo, int shift, int& k, bool kFromAux,
                                                 bool needPred, SelCtx<NT>& sc, uint32_t* keyOut, uint32_t* predOut)
{
    Located A;
    KeyStack ks;
    const int rc = cs_sweep_a<AREA, NT>(key, live, lo, shift, k, kFromAux, sc, &A, ks);
    if (rc != 0) return rc;
    bool hasPred;
    if (shift == 0) {
        *keyOut  = lo + A.bin;
        hasPred  = A.rank > 0 || A.pred != CS_NONE;
        *predOut = A.rank > 0 ? lo + A.bin : lo + A.pred;
    } else {
        cs_sweep_b<NT>(ks, lo + (A.bin << shift), A.rank, sc, keyOut, predOut, &hasPred);
    }
    if (needPred && k > 0 && !hasPred) *predOut = cs_max_below<AREA, NT>(key, live, *keyOut, sc);  // rare: predecessor far below
    return 0;
}

// One pass with thread-private packed counters over 64 clamped coarse bins of 2^16 keys starting at coarseBase << 16
// (every key takes part).  Returns the bin of rank k, *below = keys in lower bins.
template <int AREA, int NT>
__device__ __forceinline__ uint32_t cs_private_coarse(const uint32_t (&key)[AREA], bool live, int coarseBase, int& k, bool kFromAux,
                                                      SelCtx<NT>& sc, uint32_t* below)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    CS_T(10);
#pragma unroll
    for (int j = 0; j < 16; j++) sc.priv[j * NT + tid] = 0;  // this thread's column: the key stacks live here too
    if (live) {
#pragma unroll
        for (int i = 0; i < AREA; i++) {
            const uint32_t d = (uint32_t)min(max((int)(key[i] >> 16) - coarseBase, 0), 63);
            sc.priv[(d >> 2) * NT + tid] += 1u << ((d & 3u) * 8u);
        }
    }
    CS_T(7);
    __syncthreads();
    uint32_t* x = sc.rxs();
    for (int row = warp; row < 16; row += NW) {  // word row `row` = bins 4 row .. 4 row + 3, reduced over all NT columns
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < NT / 32; j++) {
            const uint32_t wv = sc.priv[row * NT + lane + 32 * j];
            lo += wv & 0x00ff00ffu;
            hi += (wv >> 8) & 0x00ff00ffu;
        }
        lo = __reduce_add_sync(CS_FULL, lo);
        hi = __reduce_add_sync(CS_FULL, hi);
        if (lane == 0) {
            x[8 + row * 4 + 0] = lo & 0xffffu;
            x[8 + row * 4 + 1] = hi & 0xffffu;
            x[8 + row * 4 + 2] = lo >> 16;
            x[8 + row * 4 + 3] = hi >> 16;
        }
    }
    CS_T(8);
    sc.finish_round();
    if (kFromAux) {
        sc.auxTotal = sc.tAux;
        k           = (int)(sc.tAux * (uint32_t)AREA / 2u);
    }
    uint32_t c0 = 0, c1 = 0;
    if (sc.C == 1) {
        c0 = x[8 + 2 * lane];
        c1 = x[8 + 2 * lane + 1];
    } else {
        const uint2 cc = cs_gather_sum2(cs_smem_u32(x + 8 + 2 * lane), sc.C);
        c0             = cc.x;
        c1             = cc.y;
    }
    const uint32_t sum  = c0 + c1;
    const uint32_t incl = cs_warp_incl_scan(sum, lane);
    const uint32_t excl = incl - sum;
    const uint32_t kk   = (uint32_t)k;
    uint32_t mine       = CS_NONE;
    if (kk >= excl && kk < excl + c0)
        mine = 2 * lane;
    else if (kk >= excl + c0 && kk < incl)
        mine = 2 * lane + 1;
    const uint32_t bin = __reduce_min_sync(CS_FULL, mine);
    *below             = __shfl_sync(CS_FULL, (bin & 1u) ? excl + c0 : excl, (int)((bin >> 1) & 31u));
    CS_T(9);
    return bin;
}

// generic tier: k-th smallest over all 27 bits in three 9-bit passes; *rankInKey = rank among equal keys
template <int AREA, int NT>
__device__ __forceinline__ uint32_t cs_generic_select27(const uint32_t (&key)[AREA], bool live, uint32_t k, SelCtx<NT>& sc, uint32_t* rankInKey)
{
    uint32_t prefix = 0, mask = 0;
#pragma unroll 1
    for (int shift = 18; shift >= 0; shift -= 9) {
        uint32_t* cur = sc.rbins();
        if (live) {
#pragma unroll
            for (int i = 0; i < AREA; i++)
                if ((key[i] & mask) == prefix) atomicAdd(&cur[(key[i] >> shift) & 511u], 1u);
        }
        const uint32_t* hist = sc.finish_round();
        const Located L      = cs_locate<NT>(hist, k, -1, sc);
        prefix |= L.bin << shift;
        mask |= 511u << shift;
        k = L.rank;
    }
    *rankInKey = k;
    return prefix;
}

// k-th smallest key and (needPred) its predecessor, through the tiers.  coarseBase: the cold tier's 64 coarse bins
// start at key (coarseBase << 16).  kFromAux: the rank is (cluster total of aux) * AREA / 2 and becomes known at
// the first barrier; on return k holds it.  *tier: 1 hot, 2 cold, 4 generic.  Returns false when kFromAux and the
// aux total is zero (nothing to select).
template <int AREA, int NT>
__device__ __forceinline__ bool cs_tiered_select(const uint32_t (&key)[AREA], bool live, int coarseBase, int& k, bool kFromAux,
                                                 int nTotal, Bracket& br, SelCtx<NT>& sc, uint32_t* keyOut, uint32_t* predOut, int* tier)
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
    // needPred: even-count median rule, mean of elements k-1 and k; evaluated once k is known
    auto needPred = [&]() { return !(nTotal & 1) && k > 0; };
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!have) {
            *tier = 2;
            uint32_t below;
            const uint32_t b = cs_private_coarse<AREA, NT>(key, live, coarseBase, k, kFromAux, sc, &below);
            if (kFromAux && sc.auxTotal == 0) return false;
            kFromAux = false;
            // clamped outer bins (bin 0 clamps nothing when the coarse range starts at key 0, as for the MAD keys)
            if ((b == 0 && coarseBase > 0) || b == 63) {
                *tier = 4;
                uint32_t rank;
                kOut = cs_generic_select27<AREA, NT>(key, live, (uint32_t)k, sc, &rank);
                pred = kOut;
                if (needPred() && rank == 0) pred = cs_max_below<AREA, NT>(key, live, kOut, sc);
                break;
            }
            lo    = ((uint32_t)((int)b + coarseBase)) << 16;
            shift = 7;  // 512 bins of 2^7 keys = the coarse bin
        }
        const int rc = cs_bracket_select<AREA, NT>(key, live, lo, shift, k, kFromAux, !(nTotal & 1), sc, &kOut, &pred);
#ifdef SVO_PROFILE
        if (have) {
            sc.stat[shift]++;
            if (rc == 0) {
                sc.stat[8 + shift]++;
            }
        }
#endif
        if (kFromAux && sc.auxTotal == 0) return false;
        kFromAux = false;
        if (rc == 0) break;
        have = false;  // hot miss: go cold
    }
    // next bracket: centred on this result, half-width >= 4x the last movement, at most +/- 2^15 keys (1/2 intensity
    // unit; wider brackets pay ~2 cycles per key inside them).  The evaluation after the first one of a level follows
    // the largest step of the level and nothing is known about the movement yet: the widest bracket.
    if (br.have) {
        const uint32_t moved = kOut > br.center ? kOut - br.center : br.center - kOut;
        const uint32_t want  = 4u * min(moved, 1u << 20) + 64u;
        int sh               = 0;
        while ((256u << sh) < want && sh < 8) sh++;
        br.valid = sh <= 7;
        br.shift = min(sh, 7);
    } else {
        br.valid = true;
        br.shift = 7;
    }
    br.have   = true;
    br.center = kOut;
    *keyOut   = kOut;
    *predOut  = needPred() ? pred : kOut;
    return true;
}

}  // namespace
