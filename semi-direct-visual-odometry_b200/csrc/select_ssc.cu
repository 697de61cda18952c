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
// One CTA runs the whole binary search (count, per-width arg-max by atomicMax, rounds, decision) without host round
// trips, sorts the <= 4,096 keepers by priority (the order the reference emits them in) and applies the first-come
// bucketing of :62-78.
#include "ctx.h"

namespace {

constexpr int SSC_NT  = 1024;
constexpr int SSC_CAP = 4096;  // keepers (Kmax = 1.1 K: numberCandidate up to ~3,700)

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
    uint32_t* cellState;  // 0 undecided, 1 kept, 2 refused
    long long cellCap;
    uint32_t* bucket;     // per occupancy-grid cell: position of its first keeper
    svo_feature_px* out;
    int maxOut;
    int32_t* count;  // features written
    int32_t* info;   // [0] keypoints above thr [1] last width [2] iterations [3] keepers before bucketing [4] error
};

__device__ __forceinline__ uint32_t ldcg(const uint32_t* p) { return __ldcg(p); }

__global__ void __launch_bounds__(SSC_NT) k_select_ssc(const SscArgs a)
{
    __shared__ uint32_t list[SSC_CAP];
    __shared__ uint32_t keep[SSC_CAP];
    __shared__ int s_n, s_undecided, s_kept, s_low, s_high, s_prev, s_done, s_width, s_iters, s_err, s_nsel;
    __shared__ uint32_t s_kmin, s_kmax;
    __shared__ int s_scan[SSC_NT / 32];
    const int tid = threadIdx.x;
    const int w = a.w, h = a.h;
    if (tid == 0) s_n = 0, s_iters = 0, s_err = 0, s_done = 0, s_prev = -1, s_width = -1, s_nsel = 0;
    __syncthreads();
    // ---- keypoints above the threshold, :41-51 ----
    {
        int n = 0;
        const int pw4 = a.pitch >> 2;  // rows are padded to 16 bytes with zeros: four pixels per load
        const uint32_t* g4 = reinterpret_cast<const uint32_t*>(a.grad);
        const int words    = h * pw4;   // the image as one flat array of words: every thread busy, loads in flight
#pragma unroll 4
        for (int i = tid; i < words; i += SSC_NT) {
            const uint32_t q = g4[i];
            const int xw     = i % pw4;
#pragma unroll
            for (int k = 0; k < 4; k++) n += (4 * xw + k < w) && ((q >> (8 * k)) & 0xffu) > a.thr;
        }
        n = __reduce_add_sync(0xffffffffu, n);
        if ((tid & 31) == 0 && n) atomicAdd(&s_n, n);
    }
    __syncthreads();
    if (tid == 0) {  // SSC :172-192
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
        const long long cells = (long long)(ncr + 1) * (ncc + 1);
        if (cells > a.cellCap) {
            if (tid == 0) s_err = 1, s_done = 1, s_iters--;
            __syncthreads();
            break;
        }
        for (long long i = tid; i < cells; i += SSC_NT) {
            a.cellKey[i]   = 0;
            a.cellState[i] = 0;
        }
        __syncthreads();
        // champion of every cell: largest gradient, earliest raster position among equals
        {
            const int pw4      = a.pitch >> 2;
            const uint32_t* g4 = reinterpret_cast<const uint32_t*>(a.grad);
            const int words    = h * pw4;
#pragma unroll 2
            for (int i = tid; i < words; i += SSC_NT) {
                const uint32_t q = g4[i];
                if (__vcmpgtu4(q, a.thr * 0x01010101u) == 0) continue;  // no byte above the threshold (thr <= 255)
                const int y = i / pw4, xw = i - y * pw4;
                const int row  = (int)((double)(float)y / c);  // static_cast<int32_t>(kp.pt.y / c), pt is Point2f
                uint32_t* krow = a.cellKey + (long long)row * (ncc + 1);
                int curCol      = -1;
                uint32_t curKey = 0;  // four consecutive pixels span at most a few cells: one atomic per cell
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x      = 4 * xw + k;
                    const uint32_t v = (q >> (8 * k)) & 0xffu;
                    if (x < w && v > a.thr) {
                        const int col      = (int)((double)(float)x / c);
                        const uint32_t key = (v << 24) | (0xFFFFFFu - (uint32_t)(y * w + x));
                        if (col != curCol) {
                            if (curKey) atomicMax(&krow[curCol], curKey);
                            curCol = col;
                            curKey = key;
                        } else
                            curKey = max(curKey, key);
                    }
                }
                if (curKey) atomicMax(&krow[curCol], curKey);
            }
        }
        __syncthreads();
        // rounds of the lexicographically-first maximal independent set
        const int reach = (int)(width / c);  // 2
        while (true) {
            if (tid == 0) s_undecided = 0;
            __syncthreads();
            int pending = 0;
            for (long long i = tid; i < cells; i += SSC_NT) {
                if (ldcg(&a.cellState[i]) != 0) continue;
                const uint32_t key = a.cellKey[i];
                if (key == 0) {
                    a.cellState[i] = 2;
                    continue;
                }
                const int row = (int)(i / (ncc + 1)), col = (int)(i - (long long)row * (ncc + 1));
                const int r0 = max(row - reach, 0), r1 = min(row + reach, ncr);
                const int c0 = max(col - reach, 0), c1 = min(col + reach, ncc);
                bool refused = false, blocked = false;
                for (int r = r0; r <= r1 && !refused; r++)
                    for (int cc = c0; cc <= c1; cc++) {
                        const long long j = (long long)r * (ncc + 1) + cc;
                        if (a.cellKey[j] > key) {
                            const uint32_t st = ldcg(&a.cellState[j]);
                            if (st == 1) {
                                refused = true;
                                break;
                            }
                            blocked |= st == 0;
                        }
                    }
                if (refused)
                    a.cellState[i] = 2;
                else if (!blocked)
                    a.cellState[i] = 1;
                else
                    pending++;
            }
            if (pending) atomicAdd(&s_undecided, pending);
            __syncthreads();
            if (s_undecided == 0) break;
            __syncthreads();
        }
        // result.size() and the binary-search step, :233-245
        if (tid == 0) s_kept = 0;
        __syncthreads();
        {
            int k = 0;
            for (long long i = tid; i < cells; i += SSC_NT) k += ldcg(&a.cellState[i]) == 1;
            k = __reduce_add_sync(0xffffffffu, k);
            if ((tid & 31) == 0 && k) atomicAdd(&s_kept, k);
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
    // ---- the keepers of the last width walked, in the order the reference emits them (sorted-keypoint order) ----
    for (int i = tid; i < SSC_CAP; i += SSC_NT) list[i] = 0;
    __syncthreads();
    if (s_iters > 0 && !s_err) {
        const long long cells = (long long)(ncr + 1) * (ncc + 1);
        for (long long i = tid; i < cells; i += SSC_NT)
            if (ldcg(&a.cellState[i]) == 1) {
                const int pos = atomicAdd(&s_nsel, 1);
                if (pos < SSC_CAP) list[pos] = a.cellKey[i];
            }
    }
    __syncthreads();
    if (s_nsel > SSC_CAP) {
        if (tid == 0) s_err = 2;
        __syncthreads();
    }
    const int nsel = min(s_nsel, SSC_CAP);
    // bitonic sort, descending (larger key = earlier in the reference's sorted keypoint list)
    for (int k = 2; k <= SSC_CAP; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < SSC_CAP; i += SSC_NT) {
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
    const long long cap = (long long)(2 * g.h + 2) * (2 * g.w + 2);  // cells at the smallest width (1 pixel: side 0.5)
    if (!ctx->d_ssc_key) {
        SVO_CUDA(cudaMalloc(&ctx->d_ssc_key, sizeof(uint32_t) * cap));
        SVO_CUDA(cudaMalloc(&ctx->d_ssc_state, sizeof(uint32_t) * cap));
        SVO_CUDA(cudaMalloc(&ctx->d_ssc_info, sizeof(int32_t) * 8));
    }
    SscArgs a;
    a.grad = ctx->arena.grad[0] + (int64_t)slot * g.plane_stride;
    a.w = g.w, a.h = g.h, a.pitch = g.pitch;
    a.thr = thr, a.K = numCandidates, a.cell = cell, a.gridRows = rows, a.gridCols = cols;
    a.occ          = useOcc ? ctx->d_occupancy : nullptr;
    a.useBucketing = useBucketing ? 1 : 0;
    a.cellKey = ctx->d_ssc_key, a.cellState = ctx->d_ssc_state, a.cellCap = cap;
    a.bucket = ctx->d_cell_best;
    a.out = ctx->d_sel_out, a.maxOut = maxOut, a.count = ctx->d_sel_count, a.info = ctx->d_ssc_info;
    k_select_ssc<<<1, SSC_NT, 0, ctx->stream>>>(a);
    ctx->launches++;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
