// select.cu -- FeatureSelection::gradientMagnitudeByValue with bucketing
// (src/feature_selection.cpp:91-146): per grid cell the first (raster order) pixel with the largest
// gradient magnitude, emitted in cell raster order iff magnitude > threshold and the cell is free.
//
// k_cell_argmax: one warp per cell.  Each lane scans whole cell rows (aligned 32-bit loads) and keeps the packed
//   key (value << 24) | (0xFFFFFF - index_in_cell); the maximum key is the largest value and, among
//   ties, the smallest raster index -- exactly the strict `>` scan of the reference.  The warp maximum
//   is one __reduce_max_sync.
// k_cell_compact: one block; ballot/popc scan over the cells in raster order, writes the feature list.
#include "ctx.h"

namespace {

__global__ void __launch_bounds__(128) k_cell_argmax(const uint8_t* __restrict__ grad, int w, int h, int pitch, int cell,
                                                     int rows, int cols, uint32_t* __restrict__ best)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows * cols) return;
    const int r = warp / cols, c = warp - r * cols;
    // ROI clipped to the image, src/feature_selection.cpp:114-116 (may be empty)
    const int cw = (c + 1) * cell < w ? cell : w - c * cell;
    const int ch = (r + 1) * cell < h ? cell : h - r * cell;
    uint32_t key = 0;
    if (cw > 0 && ch > 0) {
        // lane i takes cell rows i, i + 32, ...: the row segment is read as aligned 32-bit words (the pitched rows are
        // padded to 16 bytes, so the last word of a row exists), no division and no byte load per pixel
        const uint8_t* base = grad + (long long)(r * cell) * pitch + c * cell;
        for (int i = lane; i < ch; i += 32) {
            const uintptr_t ad  = reinterpret_cast<uintptr_t>(base + (long long)i * pitch);
            const uint32_t* w4  = reinterpret_cast<const uint32_t*>(ad & ~uintptr_t(3));
            const int off       = (int)(ad & 3u);
            const int nwords    = (off + cw + 3) >> 2;
            const uint32_t tidx = 0xFFFFFFu - (uint32_t)(i * cw) + (uint32_t)off;  // 0xFFFFFF - (i cw + j) for byte q 4 + k: minus (4 q + k)
            for (int q = 0; q < nwords; q++) {
                const uint32_t wv = __ldg(w4 + q);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int j = 4 * q + k - off;  // column inside the cell
                    if (j >= 0 && j < cw) {
                        const uint32_t v = (wv >> (8 * k)) & 0xffu;
                        key              = max(key, (v << 24) | (tidx - (uint32_t)(4 * q + k)));
                    }
                }
            }
        }
    }
    key = __reduce_max_sync(0xffffffffu, key);
    if (lane == 0) best[warp] = key;
}

__global__ void __launch_bounds__(1024) k_cell_compact(const uint32_t* __restrict__ best, const uint8_t* __restrict__ occupancy,
                                                       int w, int h, int cell, int rows, int cols, uint32_t thr,
                                                       svo_feature_px* __restrict__ out, int32_t* __restrict__ count)
{
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int n    = rows * cols;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        const int cellIdx = start + threadIdx.x;
        bool emit         = false;
        uint32_t key      = 0;
        if (cellIdx < n) {
            key              = best[cellIdx];
            const uint32_t v = key >> 24;
            emit             = v > thr && !(occupancy && occupancy[cellIdx]);  // max = 0 never passes (thr is unsigned)
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int offset = base;
        for (int i = 0; i < wid; i++) offset += warp_tot[i];
        if (emit) {
            const int pos = offset + __popc(bal & ((1u << lane) - 1u));
            const int r = cellIdx / cols, c = cellIdx - r * cols;
            const int cw = (c + 1) * cell < w ? cell : w - c * cell;
            const int t  = (int)(0xFFFFFFu - (key & 0xFFFFFFu));
            const int i = t / cw, j = t - i * cw;
            svo_feature_px f;
            f.x         = c * cell + j;
            f.y         = r * cell + i;
            f.magnitude = (int)(key >> 24);
            out[pos]    = f;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += warp_tot[i];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

}  // namespace

svo_status launch_grid_select(svo_ctx* ctx, int slot, int cell, uint32_t thr, int rows, int cols, svo_feature_px* out, int32_t* count)
{
    if (!out) out = ctx->d_sel_out, count = ctx->d_sel_count;  // default: device buffers (front-end graph); else mapped host memory
    const LevelGeom& g  = ctx->arena.geom[0];
    const uint8_t* grad = ctx->arena.grad[0] + (int64_t)slot * g.plane_stride;
    const int n         = rows * cols;
    const int blocks    = (n * 32 + 127) / 128;
    k_cell_argmax<<<blocks, 128, 0, ctx->stream>>>(grad, g.w, g.h, g.pitch, cell, rows, cols, ctx->d_cell_best);
    k_cell_compact<<<1, 1024, 0, ctx->stream>>>(ctx->d_cell_best, ctx->sel_use_occupancy ? ctx->d_occupancy : nullptr, g.w, g.h, cell, rows, cols, thr,
                                                out, count);
    ctx->launches += 2;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}
