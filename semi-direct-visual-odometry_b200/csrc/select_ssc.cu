// select_ssc.cu -- FeatureSelection::gradientMagnitudeWithSSC (src/feature_selection.cpp:27-89) with
// FeatureSelection::SSC (:165-248): what System calls on every keyframe (src/system.cpp:81,253,429).
//
// The reference thresholds every pixel, SORTS the ~10^5 keypoints by response and then, for each width of a binary
// search, walks them in order, keeping a keypoint iff its cell (side width/2) is not yet covered by the 5x5-cell square
// of an earlier keeper.  Two observations make this a parallel problem with the SAME result:
//   * only the first keypoint (in sorted order) of a cell can ever be kept: if it is kept the cell is covered, if it is
//     refused the cell was covered already.  The first keypoint of a cell is its pixel with the largest gradient, the
//     earliest in raster order among equals (the sort is taken as stable: the reference's std::sort leaves the order of
//     equal responses unspecified) -- a per-cell arg-max, no sort;
//   * the greedy walk over those cell champions is the lexicographically-first maximal independent set of the graph
//     "cells within Chebyshev distance 2", priorities = (gradient desc, raster asc).  It is computed by rounds: a champion
//     is kept once every higher-priority champion in its 5x5 neighbourhood has been refused, refused once one of them
//     has been kept; decisions only ever move from undecided to final, so rounds need no double buffering.
// One launch runs the whole binary search (count, per-width champions, rounds, decision) without host round trips, sorts
// the <= 4,096 keepers by priority (the order the reference emits them in) and applies the first-come bucketing of
// :62-78.  The launch is ONE CLUSTER of 8 CTAs: all of them scan pixels (one thread per cell scans the cell's pixel
// rectangle, no atomics) and write the champions straight into CTA 0's shared memory (distributed shared memory); CTA 0
// owns the cell arrays (36,864 cells fit: every width >= 7 on a 1241x376 frame; finer grids use global arrays), runs the
// rounds -- a cell's whole 5x5 neighbourhood is loaded before it is looked at, one barrier per round -- and everything
// after the search.  Every CTA keeps an identical copy of the search state, so the only exchange per width is the keeper
// count.  Phase counters (cycles per width, 1241x376, ~1,500 cells): champions 215 K with an atomicMax per keypoint into
// global memory -> 140 K shared-memory atomics, branch-free two-segment words -> 80-100 K per-cell scan -> ~12 K over the
// cluster; rounds 240 K with dependent global loads -> 45 K; count pass 47 K -> 14 K (16 loads in flight) -> 2 K; final sort
// 80 K -> 16 K over the power of two that holds the keepers.  What remains is CTA 0's rounds.
#include <cstdlib>

#include "ctx.h"

namespace {

constexpr int SSC_NT  = 1024;
constexpr int SSC_CAP = 4096;  // keepers (Kmax = 1.1 K: numberCandidate up to ~3,700)
constexpr int SSC_SMEM_CELLS = 36864;  // cells held in shared memory: 4-byte key + 1-byte state each = 180 KB
constexpr size_t SSC_DYN_SMEM = (size_t)SSC_SMEM_CELLS * 5;

struct SscArgs {
    const uint8_t* grad;
    int w, h, pitch;
    uint32_t thr;
    int K;     // numberCandidate
    int cell;  // m_cellSize of the occupancy grid
    int gridRows, gridCols;
    const uint8_t* occ;  // nullable
    int useBucketing;
    uint32_t* cellKey;    // champion of each SSC cell: gradient << 24 | (0xFFFFFF - raster index); 0 = no keypoint
    uint8_t* cellState;   // 0 undecided, 1 kept, 2 refused
    long long cellCap;
    int smemCells;        // cell grids up to this size live in CTA 0's shared memory (SSC_SMEM_CELLS; 0 forces the global arrays)
    uint32_t* bucket;     // per occupancy-grid cell: position of its first keeper
    svo_feature_px* out;
    int maxOut;
    int32_t* count;  // features written
    int32_t* info;   // [0] keypoints above thr [1] last width [2] iterations [3] keepers before bucketing [4] error
};

__device__ __forceinline__ uint32_t ldcg(const uint32_t* p) { return __ldcg(p); }

// ---- thread-block cluster plumbing: CTA 0 owns the cell arrays and the search state, the other CTAs of the cluster only
// scan pixels and write their champions into CTA 0's shared memory (distributed shared memory) ----
__device__ __forceinline__ uint32_t ssc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ssc_cta0(uint32_t addr)  // the same shared-memory address in CTA 0 of the cluster
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(0u));
    return r;
}
__device__ __forceinline__ void ssc_st_cluster_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void ssc_st_cluster_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int ssc_ld_cluster_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ssc_red_cluster_add(uint32_t addr, int v) { asm volatile("red.shared::cluster.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ssc_cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t ssc_cluster_size()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs; orders shared-memory (local and distributed) and global accesses across the cluster
__device__ __forceinline__ void ssc_sync(int C)
{
    if (C == 1)
        __syncthreads();
    else
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The cell arrays carry a ring of two empty cells (key 0) around the grid: the 5x5 neighbourhood of every real cell is
// inside the array and the rounds need no bounds test.  Pitch and padded index of cell (r, cc):
__device__ __forceinline__ int ssc_pitch(int ncc) { return ncc + 5; }
__device__ __forceinline__ int ssc_pidx(int r, int cc, int ncc) { return (r + 2) * (ncc + 5) + cc + 2; }
__device__ __forceinline__ long long ssc_padded_cells(int ncr, int ncc) { return (long long)(ncr + 5) * (ncc + 5); }

// One width of the binary search: clear, per-cell champions, rounds of the lexicographically-first maximal independent
// set, number of keepers.  SM: the cell arrays are the CTA's shared memory (the compiler sees the address space).
template <bool SM>
__device__ __forceinline__ void ssc_champions(const SscArgs& a, uint32_t* __restrict__ key, uint8_t* state, int width, int ncc, int ncr,
                                              int C, int crank)
{
    const int tid = threadIdx.x;
    const int w = a.w, h = a.h;
    const int cells = (ncr + 1) * (ncc + 1);
    // SM: the arrays are CTA 0's shared memory -- written through the cluster window (a plain store when C == 1)
    const uint32_t keyA = SM && C > 1 ? ssc_cta0(ssc_smem_u32(key)) : 0u, stateA = SM && C > 1 ? ssc_cta0(ssc_smem_u32(state)) : 0u;
    // champion of every cell: largest gradient above the threshold, earliest raster position among equals.  The cell of a
    // keypoint is static_cast<int>(kp.pt.x / c) with pt a Point2f and c = width / 2.0 (:199-205): for integer coordinates
    // that is the integer quotient (2 x) / width exactly (a correctly rounded quotient of two integers below 2^53 that is
    // not an integer lies at least 1 / width away from one).  So cell (r, cc) owns the pixel rectangle
    //   x in [ceil(cc width / 2), ceil((cc + 1) width / 2)),  y in [ceil(r width / 2), ceil((r + 1) width / 2))
    // and ONE THREAD PER CELL scans its rectangle in raster order, word-wise: no division, no atomics, no clearing pass;
    // neighbouring lanes read neighbouring segments of the same image rows.  (A pass over the pixels with an atomicMax per
    // cell was 8x slower: ~100 instructions per 4-pixel word on one SM, phase counters in profiles/README.md.)
    {
        const int pw4       = a.pitch >> 2;
        const uint32_t* g4  = reinterpret_cast<const uint32_t*>(a.grad);
        const uint32_t thr4 = a.thr * 0x01010101u;
        for (int i = crank * SSC_NT + tid; i < cells; i += C * SSC_NT) {
            const int r = i / (ncc + 1), cc = i - r * (ncc + 1);
            const int ys = (int)(((long long)r * width + 1) >> 1), ye = (int)min((long long)h, ((long long)(r + 1) * width + 1) >> 1);
            const int xs = (int)(((long long)cc * width + 1) >> 1), xe = (int)min((long long)w, ((long long)(cc + 1) * width + 1) >> 1);
            uint32_t bestM = 0, bestPos = 0;
            if (xs < xe) {
                const int w0 = xs >> 2, w1 = (xe - 1) >> 2;  // words of a row that hold the segment
                const uint32_t mFirst = 0xffffffffu << (8 * (xs & 3));
                const uint32_t mLast  = 0xffffffffu >> (8 * (3 - ((xe - 1) & 3)));
                for (int y = ys; y < ye; y++) {
                    const uint32_t* grow = g4 + (long long)y * pw4;
                    // eight words of the row in flight before the first is looked at: the scan is bound by load latency
                    for (int xb = w0; xb <= w1; xb += 8) {
                        uint32_t q8[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) q8[u] = xb + u <= w1 ? __ldg(grow + xb + u) : 0u;
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int xw = xb + u;
                            uint32_t q   = q8[u];
                            uint32_t msk = __vcmpgtu4(q, thr4);  // 0xff in every byte above the threshold (thr <= 255)
                            if (xw == w0) msk &= mFirst;
                            if (xw == w1) msk &= mLast;
                            q &= msk;  // words beyond w1 were loaded as zeros
                            const uint32_t t2 = __vmaxu4(q, q >> 8);  // byte 0 = max(b0, b1), byte 2 = max(b2, b3)
                            const uint32_t m  = max(t2 & 0xffu, (t2 >> 16) & 0xffu);
                            if (m > bestM) {  // strictly: the first maximal pixel in raster order
                                bestM   = m;
                                bestPos = (uint32_t)(y * w + 4 * xw + ((__ffs(__vcmpeq4(q, m * 0x01010101u)) - 1) >> 3));
                            }
                        }
                    }
                }
            }
            const uint32_t kv = bestM ? (bestM << 24) | (0xFFFFFFu - bestPos) : 0u;
            const int pi      = ssc_pidx(r, cc, ncc);
            if (SM && C > 1) {
                ssc_st_cluster_u32(keyA + 4u * (uint32_t)pi, kv);
                ssc_st_cluster_u8(stateA + (uint32_t)pi, 0u);
            } else {
                key[pi]   = kv;
                state[pi] = 0;
            }
        }
    }
    if (crank == 0) {  // the ring: two rows above and below, two columns left and right of every row
        const int P = ssc_pitch(ncc), R = ncr + 5;
        for (int t = tid; t < 4 * P + 4 * (R - 4); t += SSC_NT) {
            int pi;
            if (t < 2 * P)
                pi = t;
            else if (t < 4 * P)
                pi = (R - 2) * P + (t - 2 * P);
            else {
                const int q = t - 4 * P, row = 2 + (q >> 2), k = q & 3;
                pi = row * P + (k < 2 ? k : P - 4 + k);
            }
            key[pi]   = 0;
            state[pi] = 2;
        }
    }
}

// rounds of the lexicographically-first maximal independent set over the champions, and the number of keepers (one CTA)
template <bool SM>
__device__ __forceinline__ int ssc_rounds(uint32_t* __restrict__ key, uint8_t* state, int width, int ncc, int ncr, int* s_kept)
{
    const int tid   = threadIdx.x;
    const int cells = (ncr + 1) * (ncc + 1);
    const int P = ssc_pitch(ncc);
    while (true) {
        int pending = 0;
        for (int i = tid; i < cells; i += SSC_NT) {
            const int row = i / (ncc + 1), col = i - row * (ncc + 1);
            const int pi = ssc_pidx(row, col, ncc);
            if ((SM ? state[pi] : __ldcg(state + pi)) != 0) continue;
            const uint32_t mine = key[pi];
            if (mine == 0) {
                state[pi] = 2;
                continue;
            }
            // the whole 5 x 5 neighbourhood (reach = width / c = 2 exactly) is loaded before any of it is looked at:
            // independent loads at constant offsets instead of 25 dependent load -> branch steps
            uint32_t nk[25], ns[25];
#pragma unroll
            for (int dr = -2; dr <= 2; dr++)
#pragma unroll
                for (int dc = -2; dc <= 2; dc++) {
                    const int q = (dr + 2) * 5 + dc + 2, j = pi + dr * P + dc;
                    nk[q]       = key[j];
                    ns[q]       = SM ? *(volatile uint8_t*)(state + j) : __ldcg(state + j);
                }
            bool refused = false, blocked = false;
#pragma unroll
            for (int q = 0; q < 25; q++) {
                const bool higher = nk[q] > mine;
                refused |= higher && ns[q] == 1;
                blocked |= higher && ns[q] == 0;
            }
            if (refused)
                state[pi] = 2;
            else if (!blocked)
                state[pi] = 1;
            else
                pending++;
        }
        if (!__syncthreads_or(pending)) break;  // one barrier per round: it also carries "anything left?"
    }
    // result.size()
    if (tid == 0) *s_kept = 0;
    __syncthreads();
    {
        int k = 0;
        for (int i = tid; i < cells; i += SSC_NT) {
            const int row = i / (ncc + 1), pi = ssc_pidx(row, i - row * (ncc + 1), ncc);
            k += (SM ? state[pi] : __ldcg(state + pi)) == 1;
        }
        k = __reduce_add_sync(0xffffffffu, k);
        if ((tid & 31) == 0 && k) atomicAdd(s_kept, k);
    }
    __syncthreads();
    return *s_kept;
}

__global__ void __launch_bounds__(SSC_NT) k_select_ssc(const SscArgs a)
{
    __shared__ uint32_t list[SSC_CAP];
    __shared__ uint32_t keep[SSC_CAP];
    __shared__ int s_n, s_kept, s_low, s_high, s_prev, s_done, s_width, s_iters, s_err, s_nsel;
    __shared__ uint32_t s_kmin, s_kmax;
    __shared__ int s_scan[SSC_NT / 32];
    const int tid = threadIdx.x;
    const int w = a.w, h = a.h;
    const int C = (int)ssc_cluster_size(), crank = (int)ssc_cluster_rank();
    if (tid == 0) s_n = 0, s_iters = 0, s_err = 0, s_done = 0, s_prev = -1, s_width = -1, s_nsel = 0, s_kept = 0;
    ssc_sync(C);  // CTA 0's counters are zero before any CTA adds to them
    // ---- keypoints above the threshold, :41-51 ----
    {
        int n = 0;
        const int pw4 = a.pitch >> 2;  // rows are padded to 16 bytes with zeros: four pixels per load
        const uint32_t* g4 = reinterpret_cast<const uint32_t*>(a.grad);
        const int words    = h * pw4;   // the image as one flat array of words (row padding is zeros: never above thr >= 0)
        for (int i0 = crank * SSC_NT * 16; i0 < words; i0 += C * SSC_NT * 16) {  // sixteen loads per thread in flight
            uint32_t q16[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int i = i0 + u * SSC_NT + tid;
                q16[u]      = i < words ? __ldg(g4 + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 16; u++) n += __popc(__vcmpgtu4(q16[u], a.thr * 0x01010101u)) >> 3;
        }
        n = __reduce_add_sync(0xffffffffu, n);
        if ((tid & 31) == 0 && n) {
            if (C == 1)
                atomicAdd(&s_n, n);
            else
                ssc_red_cluster_add(ssc_cta0(ssc_smem_u32(&s_n)), n);
        }
    }
    ssc_sync(C);
    if (C > 1 && crank != 0 && tid == 0) s_n = ssc_ld_cluster_s32(ssc_cta0(ssc_smem_u32(&s_n)));
    // (CTA 0's s_n is not written again: the read needs no further barrier)
    if (tid == 0) {  // SSC :172-192 -- every CTA keeps its own copy of the search state; the copies stay identical
        const int rows = h, cols = w, K = a.K;
        const int exp1       = rows + cols + 2 * K;
        const long long exp2 = ((long long)4 * cols + (long long)4 * K + (long long)4 * rows * K + (long long)rows * rows +
                                (long long)cols * cols - (long long)2 * rows * cols + (long long)4 * rows * cols * K);
        const double exp3 = sqrt((double)exp2);
        const double exp4 = (2 * (K - 1));
        const double sol1 = -round((exp1 + exp3) / exp4);
        const double sol2 = -round((exp1 - exp3) / exp4);
        s_high = (sol1 > sol2) ? (int)sol1 : (int)sol2;
        s_low  = (int)sqrt((double)s_n / K);
        const float Kf = (float)K, tol = 0.1f;
        s_kmin = (uint32_t)roundf(Kf - (Kf * tol));
        s_kmax = (uint32_t)roundf(Kf + (Kf * tol));
    }
    __syncthreads();
    extern __shared__ __align__(16) unsigned char ssc_dyn[];
    uint32_t* sKey  = reinterpret_cast<uint32_t*>(ssc_dyn);
    uint8_t* sState = ssc_dyn + (size_t)SSC_SMEM_CELLS * 4;
    bool inSmem     = false;
    int ncc = 0, ncr = 0;
    while (true) {
        if (tid == 0) {
            const int width = s_low + (s_high - s_low) / 2;
            // width <= 0 divides by zero in the reference: defined as "stop with the previous result" (as the oracle)
            if (width == s_prev || s_low > s_high || width <= 0)
                s_done = 1;
            else {
                s_width = width;
                s_iters++;
            }
        }
        __syncthreads();
        if (s_done) break;
        const int width = s_width;
        const double c  = width / 2.0;
        ncc             = (int)(w / c);
        ncr             = (int)(h / c);
        const long long cells = ssc_padded_cells(ncr, ncc);
        if (cells > a.cellCap) {
            if (tid == 0) s_err = 1, s_done = 1, s_iters--;
            __syncthreads();
            break;
        }
        inSmem = cells <= a.smemCells;
        if (inSmem)
            ssc_champions<true>(a, sKey, sState, width, ncc, ncr, C, crank);
        else
            ssc_champions<false>(a, a.cellKey, a.cellState, width, ncc, ncr, C, crank);
        ssc_sync(C);  // the champions of every CTA are in CTA 0's shared memory (or in global memory)
        if (crank == 0) {
            if (inSmem)
                ssc_rounds<true>(sKey, sState, width, ncc, ncr, &s_kept);
            else
                ssc_rounds<false>(a.cellKey, a.cellState, width, ncc, ncr, &s_kept);
        }
        if (C > 1) {
            ssc_sync(C);  // CTA 0's keeper count is final
            if (crank != 0 && tid == 0) s_kept = ssc_ld_cluster_s32(ssc_cta0(ssc_smem_u32(&s_kept)));
            // (CTA 0 resets it inside the rounds of the next width, i.e. behind the next cluster barrier, which every
            // CTA reaches only after this read)
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t sz = (uint32_t)s_kept;
            if (sz >= s_kmin && sz <= s_kmax)
                s_done = 1;
            else if (sz < s_kmin)
                s_high = s_width - 1;
            else
                s_low = s_width + 1;
            s_prev = s_width;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (crank != 0) return;  // no cluster barrier and no access to another CTA's shared memory follows
    // ---- the keepers of the last width walked, in the order the reference emits them (sorted-keypoint order) ----
    for (int i = tid; i < SSC_CAP; i += SSC_NT) list[i] = 0;
    __syncthreads();
    if (s_iters > 0 && !s_err) {
        const int cells = (ncr + 1) * (ncc + 1);
        for (int i = tid; i < cells; i += SSC_NT) {
            const int row = i / (ncc + 1), pi = ssc_pidx(row, i - row * (ncc + 1), ncc);
            if ((inSmem ? sState[pi] : __ldcg(a.cellState + pi)) == 1) {
                const int pos = atomicAdd(&s_nsel, 1);
                if (pos < SSC_CAP) list[pos] = inSmem ? sKey[pi] : a.cellKey[pi];
            }
        }
    }
    __syncthreads();
    if (s_nsel > SSC_CAP) {
        if (tid == 0) s_err = 2;
        __syncthreads();
    }
    const int nsel = min(s_nsel, SSC_CAP);
    // bitonic sort, descending (larger key = earlier in the reference's sorted keypoint list), over the power of two that
    // holds the keepers (the zeros behind them sort last)
    int sortN = 32;
    while (sortN < nsel) sortN <<= 1;
    for (int k = 2; k <= sortN; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < sortN; i += SSC_NT) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t x = list[i], y = list[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? x < y : x > y) {
                        list[i]   = y;
                        list[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    // ---- bucketing, :62-78: the first keeper of an occupancy-grid cell wins, occupied cells take none ----
    const int gcells = a.gridRows * a.gridCols;
    for (int i = tid; i < gcells; i += SSC_NT) a.bucket[i] = 0xFFFFFFFFu;
    __syncthreads();
    for (int e = tid; e < nsel; e += SSC_NT) {
        const uint32_t raster = 0xFFFFFFu - (list[e] & 0xFFFFFFu);
        const int y = (int)(raster / (uint32_t)w), x = (int)(raster - (uint32_t)y * (uint32_t)w);
        const int b = (y / a.cell) * a.gridCols + x / a.cell;
        if (a.useBucketing && !(a.occ && a.occ[b])) atomicMin(&a.bucket[b], (uint32_t)e);
    }
    __syncthreads();
    for (int e = tid; e < SSC_CAP; e += SSC_NT) {
        uint32_t kp = 0;
        if (e < nsel) {
            if (!a.useBucketing)
                kp = 1;
            else {
                const uint32_t raster = 0xFFFFFFu - (list[e] & 0xFFFFFFu);
                const int y = (int)(raster / (uint32_t)w), x = (int)(raster - (uint32_t)y * (uint32_t)w);
                kp = ldcg(&a.bucket[(y / a.cell) * a.gridCols + x / a.cell]) == (uint32_t)e;
            }
        }
        keep[e] = kp;
    }
    __syncthreads();
    // ordered compaction: thread t owns entries 4 t .. 4 t + 3
    {
        const int e0 = 4 * tid;
        const uint32_t k0 = keep[e0], k1 = keep[e0 + 1], k2 = keep[e0 + 2], k3 = keep[e0 + 3];
        const int mine = (int)(k0 + k1 + k2 + k3);
        int incl       = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        for (int i = 0; i < (tid >> 5); i++) base += s_scan[i];
        int pos = base + incl - mine;
        const uint32_t ks[4] = {k0, k1, k2, k3};
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (ks[q]) {
                if (pos < a.maxOut) {
                    const uint32_t key    = list[e0 + q];
                    const uint32_t raster = 0xFFFFFFu - (key & 0xFFFFFFu);
                    svo_feature_px f;
                    f.y         = (int)(raster / (uint32_t)w);
                    f.x         = (int)(raster - (uint32_t)f.y * (uint32_t)w);
                    f.magnitude = (int)(key >> 24);
                    a.out[pos]  = f;
                }
                pos++;
            }
        if (tid == SSC_NT - 1) *a.count = pos;
    }
    if (tid == 0) {
        a.info[0] = s_n;
        a.info[1] = s_width;
        a.info[2] = s_iters;
        a.info[3] = s_nsel;
        a.info[4] = s_err;
    }
}

}  // namespace

svo_status launch_select_ssc(svo_ctx* ctx, int slot, uint32_t thr, int numCandidates, int cell, int rows, int cols, bool useOcc,
                             bool useBucketing, int maxOut)
{
    const LevelGeom& g = ctx->arena.geom[0];
    if ((int64_t)g.w * g.h > (1 << 24)) SVO_FAIL(SVO_ERR_UNSUPPORTED, "svo_select_ssc: images above 2^24 pixels are not supported");
    const long long cap = (long long)(2 * g.h + 6) * (2 * g.w + 6);  // cells at the smallest width (1 pixel: side 0.5) + the ring
    if (!ctx->d_ssc_key) {
        SVO_CUDA(cudaMalloc(&ctx->d_ssc_key, sizeof(uint32_t) * cap));
        SVO_CUDA(cudaMalloc(&ctx->d_ssc_state, sizeof(uint32_t) * cap));
        // results travel zero-copy: the kernel writes records, count and info straight to mapped page-locked memory
        SVO_CUDA(cudaHostAlloc(&ctx->h_ssc, sizeof(svo_feature_px) * SSC_CAP + 64, cudaHostAllocMapped));
        unsigned char* d = nullptr;
        SVO_CUDA(cudaHostGetDevicePointer(&d, ctx->h_ssc, 0));
        ctx->d_ssc_out   = reinterpret_cast<svo_feature_px*>(d);
        ctx->d_ssc_count = reinterpret_cast<int32_t*>(d + sizeof(svo_feature_px) * SSC_CAP);
        ctx->d_ssc_info  = ctx->d_ssc_count + 4;
        SVO_CUDA(cudaFuncSetAttribute(k_select_ssc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SSC_DYN_SMEM));
        // a cluster of 8 CTAs scans the pixels (SVO_SSC_CLUSTER=1|2|4|8 overrides); CTA 0 owns the cell arrays
        int C = 8;
        if (const char* e = getenv("SVO_SSC_CLUSTER")) C = atoi(e);
        if (C != 1 && C != 2 && C != 4 && C != 8) C = 8;
        if (C > 1) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(C, 1, 1), q.blockDim = dim3(SSC_NT, 1, 1), q.dynamicSmemBytes = SSC_DYN_SMEM;
            cudaLaunchAttribute qa[1];
            qa[0].id               = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = C, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = 1;
            q.attrs = qa, q.numAttrs = 1;
            int maxClusters = 0;
            if (cudaOccupancyMaxActiveClusters(&maxClusters, k_select_ssc, &q) != cudaSuccess || maxClusters < 1) {
                cudaGetLastError();  // the device cannot co-schedule the cluster: one CTA does everything
                C = 1;
            }
        }
        ctx->ssc_cluster = C;
    }
    SscArgs a;
    a.grad = ctx->arena.grad[0] + (int64_t)slot * g.plane_stride;
    a.w = g.w, a.h = g.h, a.pitch = g.pitch;
    a.thr = thr, a.K = numCandidates, a.cell = cell, a.gridRows = rows, a.gridCols = cols;
    a.occ          = useOcc ? ctx->d_occupancy : nullptr;
    a.useBucketing = useBucketing ? 1 : 0;
    a.cellKey = ctx->d_ssc_key, a.cellState = reinterpret_cast<uint8_t*>(ctx->d_ssc_state), a.cellCap = cap;
    a.smemCells = SSC_SMEM_CELLS;
    if (const char* e = getenv("SVO_SSC_SMEM_CELLS")) a.smemCells = std::min(std::max(atoi(e), 0), SSC_SMEM_CELLS);  // tests: global-array path
    a.bucket = ctx->d_cell_best;
    a.out = ctx->d_ssc_out, a.maxOut = std::min(maxOut, SSC_CAP), a.count = ctx->d_ssc_count, a.info = ctx->d_ssc_info;
    const int C            = ctx->ssc_cluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(C, 1, 1);
    cfg.blockDim           = dim3(SSC_NT, 1, 1);
    cfg.dynamicSmemBytes   = SSC_DYN_SMEM;
    cfg.stream             = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id               = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs                = attr;
    cfg.numAttrs             = C > 1 ? 1 : 0;
    SVO_CUDA(cudaLaunchKernelEx(&cfg, k_select_ssc, a));
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
