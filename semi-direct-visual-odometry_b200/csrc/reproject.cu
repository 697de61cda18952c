// reproject.cu -- Map::reprojectMap (src/map.cpp:260-489) as one batched device pass: reprojectPoint (:492-504) for every
// candidate, the per-cell choice of reprojectCell (:506-579) and ONE FeatureAlignment launch for the chosen candidates.
//
// What the reference executes (much of the function is commented out there): every feature with a 3D point of the
// reference frame and of its last keyframe is projected with the new frame's pose and, if inside the image minus 3
// pixels, appended to its grid cell; the cells are visited in a (shuffled, fixed at start-up) order; a cell's
// candidates are sorted by Point::m_type DESCENDING (UNKNOWN 3 > CANDIDATE 2 > DELETED 1 > GOOD 0 -- the enum order,
// not the comment's intent), DELETED ones are skipped, and the FIRST remaining candidate is aligned with
// FeatureAlignment::align and accepted whatever the error (the match test is commented out, :528-549); the walk stops
// once m_matches exceeds 150.  Equal types keep their insertion order here (the reference's std::sort is unstable).
//   k_reproject_bin   one thread per candidate: world2image, in-frame test, atomicMax of (type << 24 | ~index) per cell
//   k_reproject_emit  one CTA: ordered walk over the cell order (ballot / popc scan), first max_matches + 1 non-empty
//                     cells -> FeatureAlignment items
//   k_feature_align   (feature_align.cu) on those items
#include "ctx.h"
#include "math.cuh"

namespace {

struct ReprojBinArgs {
    const svo_reproj_candidate* cands;
    int n;
    double T[7];
    double K[4];
    int w, h, cell, gridCols, nCells;
    uint32_t* cellBest;   // per cell: (type << 24) | (0xFFFFFF - candidate index), 0 = empty
    double* px;           // [n][2] projected pixel
    uint8_t* projected;   // [n] reprojectPoint's return value
};

__global__ void __launch_bounds__(128) k_reproject_bin(const ReprojBinArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const svo_reproj_candidate c = a.cands[i];
    svo::Pose T;
    for (int k = 0; k < 4; k++) T.q[k] = a.T[k];
    for (int k = 0; k < 3; k++) T.t[k] = a.T[4 + k];
    double pc[3];
    svo::quat_rotate(T.q, c.point, pc);  // Frame::world2image, src/frame.cpp:84-92
    pc[0] += T.t[0], pc[1] += T.t[1], pc[2] += T.t[2];
    const double u = a.K[0] * (pc[0] / pc[2]) + a.K[2];
    const double v = a.K[1] * (pc[1] / pc[2]) + a.K[3];
    a.px[2 * i]     = u;
    a.px[2 * i + 1] = v;
    const bool in = u >= 3.0 && v >= 3.0 && u < a.w - 3.0 && v < a.h - 3.0;  // isInFrame(pixel, 3), :495
    a.projected[i] = in ? 1 : 0;
    if (in && c.type != 1 /* Point::PointType::DELETED is skipped by reprojectCell, :520-524 */) {
        const int k = (int)v / a.cell * a.gridCols + (int)u / a.cell;  // :497-498
        if (k >= 0 && k < a.nCells) atomicMax(&a.cellBest[k], ((uint32_t)c.type << 24) | (0xFFFFFFu - (uint32_t)i));
    }
}

struct ReprojEmitArgs {
    const svo_reproj_candidate* cands;
    const uint32_t* cellBest;
    const int32_t* cellOrder;
    const double* px;
    int nCells, maxItems, curSlot;
    svo_fa_item* items;          // [maxItems]
    svo_reproj_match* matches;   // [maxItems]: cell and candidate filled here, the rest after the alignment
    int32_t* count;
};

__global__ void __launch_bounds__(1024) k_reproject_emit(const ReprojEmitArgs a)
{
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    for (int i = threadIdx.x; i < a.maxItems; i += blockDim.x) a.items[i].ref_slot = -1;  // unused slots: skipped by the kernel
    __syncthreads();
    for (int start = 0; start < a.nCells && base < a.maxItems; start += blockDim.x) {
        const int i   = start + threadIdx.x;
        uint32_t key  = 0;
        int cell      = -1;
        if (i < a.nCells) {
            cell = a.cellOrder[i];
            key  = (cell >= 0 && cell < a.nCells) ? a.cellBest[cell] : 0u;
        }
        const bool emit    = key != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int pos = base;
        for (int w = 0; w < wid; w++) pos += warp_tot[w];
        pos += __popc(bal & ((1u << lane) - 1u));
        if (emit && pos < a.maxItems) {
            const int ci = (int)(0xFFFFFFu - (key & 0xFFFFFFu));
            const svo_reproj_candidate c = a.cands[ci];
            svo_fa_item it;
            it.ref_slot  = c.ref_slot;
            it.cur_slot  = a.curSlot;
            it.ref_px[0] = c.ref_px[0], it.ref_px[1] = c.ref_px[1];
            it.px[0] = a.px[2 * ci], it.px[1] = a.px[2 * ci + 1];
            it.A[0] = it.A[3] = 1.0;
            it.A[1] = it.A[2] = 0.0;
            it.use_affine = 0, it.reserved = 0;
            a.items[pos] = it;
            svo_reproj_match m;
            m.cell = cell, m.candidate = ci;
            m.px[0] = m.px[1] = m.rmse = 0.0;
            m.status = 0, m.reserved = 0;
            a.matches[pos] = m;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += warp_tot[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.count = min(base, a.maxItems);
}

// the records go straight to the caller-visible mapped host buffer (out, outCount): no copy behind the last kernel
__global__ void k_reproject_finish(const svo_fa_result* fa, const svo_reproj_match* matches, const int32_t* count, svo_reproj_match* out,
                                   int32_t* outCount)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *outCount = *count;
    if (i >= *count) return;
    svo_reproj_match m = matches[i];
    m.px[0]  = fa[i].px[0];
    m.px[1]  = fa[i].px[1];
    m.rmse   = fa[i].rmse;
    m.status = fa[i].status;
    out[i]   = m;
}

}  // namespace

svo_status launch_reproject_map(svo_ctx* ctx, int curSlot, const double T[7], int n, int cell, int nCells, int gridCols, int maxItems,
                                const svo_fa_params& fa)
{
    const LevelGeom& g = ctx->arena.geom[0];
    SVO_CUDA(cudaMemsetAsync(ctx->d_cell_best, 0, sizeof(uint32_t) * nCells, ctx->stream));
    ReprojBinArgs b;
    b.cands = ctx->d_rp_cands, b.n = n;
    for (int i = 0; i < 7; i++) b.T[i] = T[i];
    for (int i = 0; i < 4; i++) b.K[i] = ctx->cfg.K[i];
    b.w = g.w, b.h = g.h, b.cell = cell, b.gridCols = gridCols, b.nCells = nCells;
    b.cellBest = ctx->d_cell_best, b.px = ctx->d_rp_px, b.projected = ctx->d_rp_projected;
    if (n > 0) k_reproject_bin<<<(n + 127) / 128, 128, 0, ctx->stream>>>(b);
    ReprojEmitArgs e;
    e.cands = ctx->d_rp_cands, e.cellBest = ctx->d_cell_best, e.cellOrder = ctx->d_rp_order, e.px = ctx->d_rp_px;
    e.nCells = nCells, e.maxItems = maxItems, e.curSlot = curSlot;
    e.items = ctx->d_fa_items, e.matches = ctx->d_rp_matches, e.count = ctx->d_sel_count;
    k_reproject_emit<<<1, 1024, 0, ctx->stream>>>(e);
    ctx->staged_fa        = maxItems;
    ctx->staged_fa_params = fa;
    const svo_status st   = launch_feature_align(ctx);
    if (st != SVO_OK) return st;
    k_reproject_finish<<<(maxItems + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_fa_results, ctx->d_rp_matches, ctx->d_sel_count, ctx->d_rp_out,
                                                                       ctx->d_rp_count);
    ctx->launches += 3;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
