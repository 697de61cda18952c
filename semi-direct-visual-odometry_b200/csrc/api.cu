// api.cu -- the C ABI of libsvo_b200.so (include/svo_b200.h): context, pyramid arena, pinned staging and the
// stage -> H2D -> launch -> D2H -> fetch plumbing around the kernels in pyramid.cu / select.cu /
// sparse_align.cu / feature_align.cu.  There is no CPU path: without a CUDA device svo_create fails.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.h"

namespace {

thread_local std::string g_create_error;

inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// staging ring for pageable host images: two halves so the CPU copy of chunk i+1 overlaps the DMA of chunk i
constexpr int kStageFrames = 32;

void release(svo_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    frontend_release(ctx);
    for (int l = 0; l < SVO_MAX_LEVELS; l++) {
        if (ctx->arena.img[l]) cudaFree(ctx->arena.img[l]);
        if (ctx->arena.grad[l]) cudaFree(ctx->arena.grad[l]);
    }
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; i++) {
        if (ctx->h_img_stage[i]) cudaFreeHost(ctx->h_img_stage[i]);
        if (ctx->d_img_stage[i]) cudaFree(ctx->d_img_stage[i]);
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ingest_stream) {
        cudaStreamSynchronize(ctx->ingest_stream);
        cudaStreamDestroy(ctx->ingest_stream);
    }
    if (ctx->ev_ingest_done) cudaEventDestroy(ctx->ev_ingest_done);
    for (int i = 0; i < 2; i++) {
        if (ctx->d_pf_stage[i]) cudaFree(ctx->d_pf_stage[i]);
        if (ctx->ev_pf_consumed[i]) cudaEventDestroy(ctx->ev_pf_consumed[i]);
    }
    for (int i = 0; i < 64; i++)
        if (ctx->ev_pf_chunk[i]) cudaEventDestroy(ctx->ev_pf_chunk[i]);
    if (ctx->ev_jobs_h2d) cudaEventDestroy(ctx->ev_jobs_h2d);
    cudaFree(ctx->d_ssc_key);
    cudaFree(ctx->d_ssc_state);
    cudaFreeHost(ctx->h_ssc);  // d_ssc_out, d_ssc_count, d_ssc_info are views of it
    cudaFree(ctx->d_cell_best);
    cudaFree(ctx->d_occupancy);
    cudaFree(ctx->d_sel_out);
    cudaFree(ctx->d_sel_count);
    cudaFreeHost(ctx->h_sel_out);
    cudaFreeHost(ctx->h_sel_count);
    cudaFreeHost(ctx->h_grid);
    cudaFreeHost(ctx->h_occupancy);
    cudaFreeHost(ctx->h_jobs);
    cudaFreeHost(ctx->h_feats);
    cudaFreeHost(ctx->h_results);
    cudaFreeHost(ctx->h_stats);
    cudaFree(ctx->d_jobs);
    cudaFree(ctx->d_feats);
    cudaFree(ctx->d_results);
    cudaFree(ctx->d_stats);
    cudaFree(ctx->d_scratch_tpl);
    cudaFree(ctx->d_scratch_jac);
    cudaFree(ctx->d_scratch2);
    cudaFree(ctx->d_dbg);
    cudaFreeHost(ctx->h_rp);  // d_rp_cands, d_rp_order, d_rp_out, d_rp_projected, d_rp_count are views of it
    cudaFree(ctx->d_rp_matches);
    cudaFree(ctx->d_rp_px);
    cudaFreeHost(ctx->h_klt);  // d_klt_* are device views of this mapped allocation
    cudaFreeHost(ctx->h_epi_items);
    cudaFreeHost(ctx->h_epi_results);
    cudaFree(ctx->d_epi_items);
    cudaFreeHost(ctx->h_fa_items);
    cudaFreeHost(ctx->h_fa_results);
    cudaFree(ctx->d_fa_items);
    cudaFree(ctx->d_fa_results);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

svo_status init(svo_ctx* ctx)
{
    const svo_config& c = ctx->cfg;
    SVO_CUDA(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    SVO_CUDA(cudaGetDeviceProperties(&prop, c.device));
    ctx->sm_count       = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (c.stream) {
        ctx->stream     = (cudaStream_t)c.stream;
        ctx->own_stream = false;
    } else {
        SVO_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }

    // ---- pyramid arena: [level][slot][h][pitch], pitch a multiple of 16 B, planes 256-B aligned ----
    PyramidArena& a = ctx->arena;
    a.levels        = c.levels;
    int w = c.width, h = c.height;
    for (int l = 0; l < c.levels; l++) {
        LevelGeom& g   = a.geom[l];
        g.w            = w;
        g.h            = h;
        g.pitch        = (int)round_up(w, 16);
        g.plane_stride = round_up((int64_t)g.pitch * h, 256);
        const size_t bytes = (size_t)g.plane_stride * c.max_frames;
        SVO_CUDA(cudaMalloc(&a.img[l], bytes));
        SVO_CUDA(cudaMalloc(&a.grad[l], bytes));
        SVO_CUDA(cudaMemsetAsync(a.img[l], 0, bytes, ctx->stream));
        SVO_CUDA(cudaMemsetAsync(a.grad[l], 0, bytes, ctx->stream));
        w = (w + 1) / 2;  // cv::pyrDown default size, src/image_pyramid.cpp:49-50
        h = (h + 1) / 2;
    }
    SVO_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    {   // the ingest kernels are short and feed the DMA pipeline: highest priority, so that their CTAs are scheduled as
        // soon as an SM has room even while a long alignment launch owns the device
        int lo = 0, hi = 0;
        SVO_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SVO_CUDA(cudaStreamCreateWithPriority(&ctx->ingest_stream, cudaStreamNonBlocking, hi));
    }
    SVO_CUDA(cudaEventCreateWithFlags(&ctx->ev_ingest_done, cudaEventDisableTiming));
    ctx->pyr_stream     = ctx->stream;
    ctx->ingest_pending = false;
    for (int i = 0; i < 2; i++) SVO_CUDA(cudaEventCreateWithFlags(&ctx->ev_pf_consumed[i], cudaEventDisableTiming));
    for (int i = 0; i < 64; i++) SVO_CUDA(cudaEventCreateWithFlags(&ctx->ev_pf_chunk[i], cudaEventDisableTiming));
    SVO_CUDA(cudaEventCreateWithFlags(&ctx->ev_jobs_h2d, cudaEventDisableTiming));
    ctx->stage_frames = std::min(kStageFrames, c.max_frames);
    const size_t stage_bytes = (size_t)c.width * c.height * ctx->stage_frames + 64;
    for (int i = 0; i < 2; i++) {
        SVO_CUDA(cudaHostAlloc(&ctx->h_img_stage[i], stage_bytes, cudaHostAllocDefault));
        SVO_CUDA(cudaMalloc(&ctx->d_img_stage[i], stage_bytes));
        SVO_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        SVO_CUDA(cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
    }
    ctx->stage_next = 0;

    // ---- selection: cells of >= 4 px ----
    ctx->sel_cap_cells = (c.height / 4 + 1) * (c.width / 4 + 1);
    SVO_CUDA(cudaMalloc(&ctx->d_cell_best, sizeof(uint32_t) * ctx->sel_cap_cells));
    SVO_CUDA(cudaMalloc(&ctx->d_occupancy, ctx->sel_cap_cells));
    SVO_CUDA(cudaMalloc(&ctx->d_sel_out, sizeof(svo_feature_px) * ctx->sel_cap_cells));
    SVO_CUDA(cudaMalloc(&ctx->d_sel_count, sizeof(int32_t)));
    SVO_CUDA(cudaHostAlloc(&ctx->h_sel_out, sizeof(svo_feature_px) * ctx->sel_cap_cells, cudaHostAllocDefault));
    SVO_CUDA(cudaHostAlloc(&ctx->h_sel_count, sizeof(int32_t), cudaHostAllocDefault));
    SVO_CUDA(cudaHostAlloc(&ctx->h_occupancy, ctx->sel_cap_cells, cudaHostAllocDefault));

    // ---- sparse alignment ----
    const int64_t nj = std::max(1, c.max_jobs);
    ctx->feats_cap   = nj * std::max(1, c.max_features);
    SVO_CUDA(cudaHostAlloc(&ctx->h_jobs, sizeof(svo_align_job) * nj, cudaHostAllocDefault));
    SVO_CUDA(cudaHostAlloc(&ctx->h_feats, sizeof(svo_align_feature) * ctx->feats_cap, cudaHostAllocDefault));
    SVO_CUDA(cudaHostAlloc(&ctx->h_results, sizeof(svo_align_result) * nj, cudaHostAllocDefault));
    SVO_CUDA(cudaHostAlloc(&ctx->h_stats, sizeof(svo_align_level_stats) * nj * c.levels, cudaHostAllocDefault));
    SVO_CUDA(cudaMalloc(&ctx->d_jobs, sizeof(svo_align_job) * nj));
    SVO_CUDA(cudaMalloc(&ctx->d_feats, sizeof(svo_align_feature) * ctx->feats_cap));
    SVO_CUDA(cudaMalloc(&ctx->d_results, sizeof(svo_align_result) * nj));
    SVO_CUDA(cudaMalloc(&ctx->d_stats, sizeof(svo_align_level_stats) * nj * c.levels));
    SVO_CUDA(cudaMalloc(&ctx->d_scratch_jac, sizeof(float) * 12 * ctx->feats_cap));
    ctx->scratch_area = 0;
    SVO_CUDA(cudaMalloc(&ctx->d_dbg, 64 * sizeof(long long)));
    SVO_CUDA(cudaMemsetAsync(ctx->d_dbg, 0, 64 * sizeof(long long), ctx->stream));

    // ---- feature alignment ----
    const int64_t nf = std::max(1, c.max_fa_items);
    SVO_CUDA(cudaHostAlloc(&ctx->h_fa_items, sizeof(svo_fa_item) * nf, cudaHostAllocDefault));
    SVO_CUDA(cudaHostAlloc(&ctx->h_fa_results, sizeof(svo_fa_result) * nf, cudaHostAllocDefault));
    SVO_CUDA(cudaMalloc(&ctx->d_fa_items, sizeof(svo_fa_item) * nf));
    SVO_CUDA(cudaMalloc(&ctx->d_fa_results, sizeof(svo_fa_result) * nf));
    SVO_CUDA(cudaHostAlloc(&ctx->h_epi_items, sizeof(svo_epi_item) * nf, cudaHostAllocDefault));
    // mapped: k_epipolar_match writes its one record per seed straight to the host (no device->host copy behind the kernel)
    SVO_CUDA(cudaHostAlloc(&ctx->h_epi_results, sizeof(svo_epi_result) * nf, cudaHostAllocMapped));
    SVO_CUDA(cudaMalloc(&ctx->d_epi_items, sizeof(svo_epi_item) * nf));
    SVO_CUDA(cudaHostGetDevicePointer(&ctx->d_epi_results, ctx->h_epi_results, 0));  // a view, not an allocation
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

// template scratch grows with the patch area of the request (allocation happens outside any warm loop)
svo_status ensure_scratch(svo_ctx* ctx, int area)
{
    if (area <= ctx->scratch_area) return SVO_OK;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_scratch_tpl) SVO_CUDA(cudaFree(ctx->d_scratch_tpl));
    ctx->d_scratch_tpl = nullptr;
    ctx->scratch_area  = 0;
    SVO_CUDA(cudaMalloc(&ctx->d_scratch_tpl, sizeof(float) * 3 * (size_t)ctx->feats_cap * area));
    ctx->scratch_area = area;
    return SVO_OK;
}

inline bool bad_slot(const svo_ctx* ctx, int s) { return s < 0 || s >= ctx->cfg.max_frames; }

// work on the main stream that touches frame slots runs after the last svo_frames_prefetch
svo_status wait_ingest(svo_ctx* ctx)
{
    if (ctx->ingest_pending) {
        SVO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_ingest_done, 0));
        ctx->ingest_pending = false;
    }
    return SVO_OK;
}

}  // namespace

extern "C" {

const char* svo_version(void) { return "svo_b200 0.1 (sm_100a)"; }

svo_status svo_create(const svo_config* cfg, svo_ctx** out)
{
    if (!cfg || !out) return SVO_ERR_INVALID;
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->levels < 1 || cfg->levels > SVO_MAX_LEVELS || cfg->max_frames < 1 ||
        cfg->max_jobs < 0 || cfg->max_features < 0 || cfg->max_fa_items < 0) {
        g_create_error = "svo_create: invalid configuration";
        return SVO_ERR_INVALID;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || cfg->device < 0 || cfg->device >= n) {
        cudaGetLastError();
        g_create_error = "svo_create: no usable CUDA device (this library has no CPU path)";
        return SVO_ERR_NO_DEVICE;
    }
    svo_ctx* ctx = new (std::nothrow) svo_ctx();
    if (!ctx) return SVO_ERR_INVALID;
    ctx->cfg = *cfg;
    const svo_status st = init(ctx);
    if (st != SVO_OK) {
        g_create_error = ctx->err;
        release(ctx);
        return st;
    }
    *out = ctx;
    return SVO_OK;
}

void svo_destroy(svo_ctx* ctx) { release(ctx); }

const char* svo_last_error(const svo_ctx* ctx)
{
    if (!ctx) return g_create_error.c_str();
    std::lock_guard<std::recursive_mutex> guard(const_cast<svo_ctx*>(ctx)->mu);  // (the string is replaced under this lock)
    return ctx->err.c_str();
}

svo_status svo_sync(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    SVO_CUDA(cudaStreamSynchronize(ctx->ingest_stream));
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

int64_t svo_launch_count(const svo_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* svo_stream(const svo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

svo_status svo_level_dims(const svo_ctx* ctx, int level, int* w, int* h, int* pitch)
{
    if (!ctx || level < 0 || level >= ctx->arena.levels) return SVO_ERR_INVALID;  // immutable after svo_create: no lock
    if (w) *w = ctx->arena.geom[level].w;
    if (h) *h = ctx->arena.geom[level].h;
    if (pitch) *pitch = ctx->arena.geom[level].pitch;
    return SVO_OK;
}

svo_status svo_host_alloc(svo_ctx* ctx, int64_t bytes, void** out)
{
    if (!ctx || !out || bytes <= 0) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    SVO_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return SVO_OK;
}

svo_status svo_host_free(svo_ctx* ctx, void* p)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (p) SVO_CUDA(cudaFreeHost(p));
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------
// frames
// ------------------------------------------------------------------------------------------------
static svo_status frames_upload_impl(svo_ctx* ctx, int first_slot, int n, const uint8_t* imgs, int pitch, int64_t frame_stride,
                                     bool overlap)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (n == 0) return SVO_OK;
    const LevelGeom& g = ctx->arena.geom[0];
    if (!imgs || n < 0 || bad_slot(ctx, first_slot) || bad_slot(ctx, first_slot + n - 1) || pitch < g.w ||
        (n > 1 && frame_stride < (int64_t)pitch * g.h))
        SVO_FAIL(SVO_ERR_INVALID, "svo_frames_upload: bad slot range, pitch or stride");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    // the stream the repack + pyramid kernels run on: the main stream (ordered with everything else), or the ingest
    // stream (svo_frames_prefetch: concurrent with work already enqueued on the main stream)
    cudaStream_t ks = overlap ? ctx->ingest_stream : ctx->stream;
    if (!overlap) {
        const svo_status ws = wait_ingest(ctx);
        if (ws != SVO_OK) return ws;
    }
    ctx->pyr_stream = ks;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, imgs) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const bool dense       = pitch == g.w && (n == 1 || frame_stride == (int64_t)g.w * g.h);
    const int64_t frame_sz = (int64_t)g.w * g.h;
    if (overlap && pinned && dense) {
        // whole-batch staging: the copy engine streams the batch without ever waiting for a kernel; repack + pyramid
        // kernels follow chunk by chunk on the high-priority ingest stream whenever SMs are free
        const int b = ctx->pf_next;
        ctx->pf_next ^= 1;
        const size_t need = (size_t)n * frame_sz + 64;
        if (need > ctx->pf_stage_bytes[b]) {
            SVO_CUDA(cudaEventSynchronize(ctx->ev_pf_consumed[b]));
            if (ctx->d_pf_stage[b]) SVO_CUDA(cudaFree(ctx->d_pf_stage[b]));
            ctx->d_pf_stage[b]     = nullptr;
            ctx->pf_stage_bytes[b] = 0;
            SVO_CUDA(cudaMalloc(&ctx->d_pf_stage[b], need));
            ctx->pf_stage_bytes[b] = need;
        }
        SVO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pf_consumed[b], 0));
        const int chunk = 64;
        for (int i0 = 0; i0 < n; i0 += chunk) {
            const int m      = std::min(chunk, n - i0);
            uint8_t* dstage = ctx->d_pf_stage[b] + (int64_t)i0 * frame_sz;
            cudaEvent_t ev  = ctx->ev_pf_chunk[ctx->pf_chunk_next];
            ctx->pf_chunk_next = (ctx->pf_chunk_next + 1) & 63;
            SVO_CUDA(cudaMemcpyAsync(dstage, imgs + (int64_t)i0 * frame_stride, (size_t)m * frame_sz, cudaMemcpyHostToDevice,
                                     ctx->copy_stream));
            SVO_CUDA(cudaEventRecord(ev, ctx->copy_stream));
            SVO_CUDA(cudaStreamWaitEvent(ks, ev, 0));
            svo_status st = launch_repack(ctx, dstage, g.w, frame_sz, first_slot + i0, m);
            if (st != SVO_OK) return st;
            st = launch_pyramid_build(ctx, first_slot + i0, m);
            if (st != SVO_OK) return st;
        }
        SVO_CUDA(cudaEventRecord(ctx->ev_pf_consumed[b], ks));
        ctx->pyr_stream = ctx->stream;
        SVO_CUDA(cudaEventRecord(ctx->ev_ingest_done, ctx->ingest_stream));
        ctx->ingest_pending = true;
        return SVO_OK;
    }
    // Chunked pipeline: the DMA of chunk i+1 (copy stream) overlaps k_repack + the pyramid kernels of chunk i (main
    // stream).  PCIe moves DENSE bytes in one 1-D transfer per chunk; the pitched arena layout is made on the device.
    for (int i0 = 0; i0 < n; i0 += ctx->stage_frames) {
        const int m   = std::min(ctx->stage_frames, n - i0);
        const int buf = ctx->stage_next;
        ctx->stage_next ^= 1;
        const uint8_t* src = imgs + (int64_t)i0 * frame_stride;
        SVO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[buf], 0));  // device buffer free again
        if (pinned && dense) {
            SVO_CUDA(cudaMemcpyAsync(ctx->d_img_stage[buf], src, (size_t)m * frame_sz, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else if (pinned) {
            for (int i = 0; i < m; i++)
                SVO_CUDA(cudaMemcpy2DAsync(ctx->d_img_stage[buf] + (int64_t)i * frame_sz, g.w, src + (int64_t)i * frame_stride, pitch,
                                           g.w, g.h, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
            SVO_CUDA(cudaEventSynchronize(ctx->ev_h2d[buf]));  // the pinned half is no longer being read
            uint8_t* st = ctx->h_img_stage[buf];
            for (int i = 0; i < m; i++) {
                const uint8_t* fs = src + (int64_t)i * frame_stride;
                uint8_t* fd       = st + (int64_t)i * frame_sz;
                if (pitch == g.w)
                    std::memcpy(fd, fs, (size_t)frame_sz);
                else
                    for (int y = 0; y < g.h; y++) std::memcpy(fd + (int64_t)y * g.w, fs + (int64_t)y * pitch, g.w);
            }
            SVO_CUDA(cudaMemcpyAsync(ctx->d_img_stage[buf], st, (size_t)m * frame_sz, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        SVO_CUDA(cudaEventRecord(ctx->ev_h2d[buf], ctx->copy_stream));
        SVO_CUDA(cudaStreamWaitEvent(ks, ctx->ev_h2d[buf], 0));
        svo_status st = launch_repack(ctx, ctx->d_img_stage[buf], g.w, frame_sz, first_slot + i0, m);
        if (st != SVO_OK) return st;
        SVO_CUDA(cudaEventRecord(ctx->ev_consumed[buf], ks));
        st = launch_pyramid_build(ctx, first_slot + i0, m);
        if (st != SVO_OK) return st;
    }
    ctx->pyr_stream = ctx->stream;
    if (overlap) {
        SVO_CUDA(cudaEventRecord(ctx->ev_ingest_done, ctx->ingest_stream));
        ctx->ingest_pending = true;
    }
    return SVO_OK;
}

svo_status svo_frames_upload(svo_ctx* ctx, int first_slot, int n, const uint8_t* imgs, int pitch, int64_t frame_stride)
{
    return frames_upload_impl(ctx, first_slot, n, imgs, pitch, frame_stride, false);
}

svo_status svo_frames_prefetch(svo_ctx* ctx, int first_slot, int n, const uint8_t* imgs, int pitch, int64_t frame_stride)
{
    return frames_upload_impl(ctx, first_slot, n, imgs, pitch, frame_stride, true);
}


svo_status svo_frames_upload_device(svo_ctx* ctx, int first_slot, int n, const void* dptr, int pitch, int64_t frame_stride)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (n == 0) return SVO_OK;
    const LevelGeom& g = ctx->arena.geom[0];
    if (!dptr || n < 0 || bad_slot(ctx, first_slot) || bad_slot(ctx, first_slot + n - 1) || pitch < g.w)
        SVO_FAIL(SVO_ERR_INVALID, "svo_frames_upload_device: bad slot range or pitch");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    {
        const svo_status ws = wait_ingest(ctx);
        if (ws != SVO_OK) return ws;
        const svo_status st = launch_repack(ctx, (const uint8_t*)dptr, pitch, frame_stride, first_slot, n);
        if (st != SVO_OK) return st;
    }
    return launch_pyramid_build(ctx, first_slot, n);
}

svo_status svo_frames_rebuild(svo_ctx* ctx, int first_slot, int n)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (n == 0) return SVO_OK;
    if (n < 0 || bad_slot(ctx, first_slot) || bad_slot(ctx, first_slot + n - 1))
        SVO_FAIL(SVO_ERR_INVALID, "svo_frames_rebuild: bad slot range");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    {
        const svo_status ws = wait_ingest(ctx);
        if (ws != SVO_OK) return ws;
    }
    return launch_pyramid_build(ctx, first_slot, n);
}

svo_status svo_frame_download(svo_ctx* ctx, int slot, int level, int which, uint8_t* dst, int dst_pitch)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (!dst || bad_slot(ctx, slot) || level < 0 || level >= ctx->arena.levels || (which != 0 && which != 1))
        SVO_FAIL(SVO_ERR_INVALID, "svo_frame_download: bad slot / level / which");
    const LevelGeom& g = ctx->arena.geom[level];
    if (dst_pitch < g.w) SVO_FAIL(SVO_ERR_INVALID, "svo_frame_download: dst_pitch below the level width");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    {
        const svo_status ws = wait_ingest(ctx);
        if (ws != SVO_OK) return ws;
    }
    const uint8_t* src = (which ? ctx->arena.grad[level] : ctx->arena.img[level]) + (int64_t)slot * g.plane_stride;
    SVO_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost, ctx->stream));
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------
// grid selection
// ------------------------------------------------------------------------------------------------
svo_status svo_select_grid(svo_ctx* ctx, int slot, int cell, uint32_t thr, const uint8_t* occupancy, svo_feature_px* out,
                           int max_out, int* n_out)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (!out || !n_out || max_out < 0 || bad_slot(ctx, slot) || cell < 4)
        SVO_FAIL(SVO_ERR_INVALID, "svo_select_grid: bad arguments (cell must be >= 4)");
    const LevelGeom& g = ctx->arena.geom[0];
    const int rows = g.h / cell + 1, cols = g.w / cell + 1;  // src/feature_selection.cpp:19-25
    if (rows * cols > ctx->sel_cap_cells) SVO_FAIL(SVO_ERR_CAPACITY, "svo_select_grid: too many cells");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    {
        const svo_status ws = wait_ingest(ctx);
        if (ws != SVO_OK) return ws;
    }
    if (occupancy) {
        std::memcpy(ctx->h_occupancy, occupancy, (size_t)rows * cols);
        SVO_CUDA(cudaMemcpyAsync(ctx->d_occupancy, ctx->h_occupancy, (size_t)rows * cols, cudaMemcpyHostToDevice,
                                 ctx->stream));
    }
    ctx->sel_use_occupancy = occupancy != nullptr;
    if (!ctx->h_grid)  // zero-copy results: the compaction kernel writes count | records to mapped page-locked memory
        SVO_CUDA(cudaHostAlloc(&ctx->h_grid, 16 + sizeof(svo_feature_px) * ctx->sel_cap_cells, cudaHostAllocMapped));
    unsigned char* dGrid = nullptr;
    SVO_CUDA(cudaHostGetDevicePointer(&dGrid, ctx->h_grid, 0));
    const svo_status st = launch_grid_select(ctx, slot, cell, thr, rows, cols, reinterpret_cast<svo_feature_px*>(dGrid + 16),
                                             reinterpret_cast<int32_t*>(dGrid));
    if (st != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = *reinterpret_cast<const int32_t*>(ctx->h_grid);
    *n_out      = n;
    std::memcpy(out, ctx->h_grid + 16, sizeof(svo_feature_px) * std::min(n, max_out));
    return SVO_OK;
}

svo_status svo_select_ssc(svo_ctx* ctx, int slot, uint32_t thr, int num_candidates, int cell, const uint8_t* occupancy,
                          int use_bucketing, svo_feature_px* out, int max_out, int* n_out, int32_t* info)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (!out || !n_out || max_out < 0 || bad_slot(ctx, slot) || cell < 4 || num_candidates < 2)
        SVO_FAIL(SVO_ERR_INVALID, "svo_select_ssc: bad arguments (cell >= 4, num_candidates >= 2)");
    const LevelGeom& g = ctx->arena.geom[0];
    const int rows = g.h / cell + 1, cols = g.w / cell + 1;
    if (rows * cols > ctx->sel_cap_cells) SVO_FAIL(SVO_ERR_CAPACITY, "svo_select_ssc: too many cells");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    {
        const svo_status ws = wait_ingest(ctx);
        if (ws != SVO_OK) return ws;
    }
    if (occupancy) {
        std::memcpy(ctx->h_occupancy, occupancy, (size_t)rows * cols);
        SVO_CUDA(cudaMemcpyAsync(ctx->d_occupancy, ctx->h_occupancy, (size_t)rows * cols, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int cap = 4096;  // records in the mapped result block (SSC_CAP of select_ssc.cu)
    const svo_status st = launch_select_ssc(ctx, slot, thr, num_candidates, cell, rows, cols, occupancy != nullptr, use_bucketing != 0, cap);
    if (st != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    // the kernel wrote records | count | info to the mapped block
    const int32_t* hinfo = reinterpret_cast<const int32_t*>(ctx->h_ssc + sizeof(svo_feature_px) * 4096) + 4;
    if (info)
        for (int i = 0; i < 4; i++) info[i] = hinfo[i];
    if (hinfo[4] == 1) SVO_FAIL(SVO_ERR_CAPACITY, "svo_select_ssc: more SSC cells than the scratch holds");
    if (hinfo[4] == 2) SVO_FAIL(SVO_ERR_CAPACITY, "svo_select_ssc: more than 4,096 points survive the suppression (num_candidates too large)");
    const int n = *reinterpret_cast<const int32_t*>(ctx->h_ssc + sizeof(svo_feature_px) * 4096);
    *n_out      = n;
    std::memcpy(out, ctx->h_ssc, sizeof(svo_feature_px) * std::min(n, std::min(max_out, cap)));
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------
// sparse image alignment
// ------------------------------------------------------------------------------------------------
svo_status svo_sparse_align_stage(svo_ctx* ctx, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats,
                                  int n_feats, const svo_align_params* prm, int want_stats)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (!jobs || !prm || n_jobs < 0 || n_feats < 0 || (n_feats > 0 && !feats))
        SVO_FAIL(SVO_ERR_INVALID, "svo_sparse_align: null argument");
    if (n_jobs > ctx->cfg.max_jobs || n_feats > ctx->feats_cap)
        SVO_FAIL(SVO_ERR_CAPACITY, "svo_sparse_align: more jobs / features than the context was created for");
    if (prm->patch_size < 1 || prm->patch_size > 16 || prm->min_level < 0 || prm->max_level < prm->min_level ||
        prm->max_level >= ctx->arena.levels || prm->mode < SVO_LM_FAITHFUL || prm->mode > SVO_GN)
        SVO_FAIL(SVO_ERR_INVALID, "svo_sparse_align: bad parameters");
    for (int j = 0; j < n_jobs; j++) {
        const svo_align_job& J = jobs[j];
        if (bad_slot(ctx, J.ref_slot) || bad_slot(ctx, J.kf_slot) || bad_slot(ctx, J.cur_slot))
            SVO_FAIL(SVO_ERR_INVALID, "svo_sparse_align: frame slot out of range (a last keyframe is required, "
                                      "src/image_alignment.cpp:30-31)");
        if (J.n_ref < 0 || J.n_kf < 0 || J.n_ref + J.n_kf > ctx->cfg.max_features)
            SVO_FAIL(SVO_ERR_CAPACITY, "svo_sparse_align: job has more features than max_features");
        if (J.feat_offset < 0 || (int64_t)J.feat_offset + J.n_ref + J.n_kf > n_feats)
            SVO_FAIL(SVO_ERR_INVALID, "svo_sparse_align: feature range outside the feats array");
    }
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    const svo_status st = ensure_scratch(ctx, prm->patch_size * prm->patch_size);
    if (st != SVO_OK) return st;
    // the pinned buffers may still be read by the previous batch's H2D
    SVO_CUDA(cudaEventSynchronize(ctx->ev_jobs_h2d));
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    // page-locked caller buffers (svo_host_alloc) are DMA'd in place -- they must stay untouched until the fetch;
    // pageable ones are copied to the context's pinned mirrors first
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes at;
        const bool ok = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        return ok;
    };
    ctx->src_jobs  = jobs;
    ctx->src_feats = feats;
    if (!is_pinned(jobs)) {
        std::memcpy(ctx->h_jobs, jobs, sizeof(svo_align_job) * n_jobs);
        ctx->src_jobs = ctx->h_jobs;
    }
    if (n_feats && !is_pinned(feats)) {
        std::memcpy(ctx->h_feats, feats, sizeof(svo_align_feature) * n_feats);
        ctx->src_feats = ctx->h_feats;
    }
    ctx->staged_jobs       = n_jobs;
    ctx->staged_feats      = n_feats;
    ctx->staged_levels     = prm->max_level - prm->min_level + 1;
    ctx->staged_want_stats = want_stats;
    ctx->staged_params     = *prm;
    return SVO_OK;
}

svo_status svo_sparse_align_h2d(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (ctx->staged_jobs == 0) return SVO_OK;
    // On the COPY stream, i.e. in call order with the frame DMA: the copy engine drains one stream's queued transfers
    // before it turns to another, so a small copy on the main stream would sit behind every frame batch that a
    // pipelined caller has already queued (measured: the alignment then starts a whole batch late).
    SVO_CUDA(cudaMemcpyAsync(ctx->d_jobs, ctx->src_jobs, sizeof(svo_align_job) * ctx->staged_jobs, cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    if (ctx->staged_feats)
        SVO_CUDA(cudaMemcpyAsync(ctx->d_feats, ctx->src_feats, sizeof(svo_align_feature) * ctx->staged_feats,
                                 cudaMemcpyHostToDevice, ctx->copy_stream));
    SVO_CUDA(cudaEventRecord(ctx->ev_jobs_h2d, ctx->copy_stream));
    SVO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_jobs_h2d, 0));
    return SVO_OK;
}

svo_status svo_sparse_align_launch(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    const svo_status ws = wait_ingest(ctx);
    if (ws != SVO_OK) return ws;
    return launch_sparse_align(ctx);
}

svo_status svo_sparse_align_d2h(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (ctx->staged_jobs == 0) return SVO_OK;
    SVO_CUDA(cudaMemcpyAsync(ctx->h_results, ctx->d_results, sizeof(svo_align_result) * ctx->staged_jobs,
                             cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->staged_want_stats)
        SVO_CUDA(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats,
                                 sizeof(svo_align_level_stats) * ctx->staged_jobs * ctx->staged_levels,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    return SVO_OK;
}

svo_status svo_sparse_align_fetch(svo_ctx* ctx, svo_align_result* results, svo_align_level_stats* stats)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (results) std::memcpy(results, ctx->h_results, sizeof(svo_align_result) * ctx->staged_jobs);
    if (stats) {
        if (!ctx->staged_want_stats) SVO_FAIL(SVO_ERR_INVALID, "svo_sparse_align_fetch: stats were not requested at stage");
        std::memcpy(stats, ctx->h_stats, sizeof(svo_align_level_stats) * ctx->staged_jobs * ctx->staged_levels);
    }
    return SVO_OK;
}

svo_status svo_debug_cycles(svo_ctx* ctx, int64_t* out64)
{
    if (!ctx || !out64) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    SVO_CUDA(cudaMemcpy(out64, ctx->d_dbg, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
    return SVO_OK;
}

const void* svo_sparse_align_results_device(const svo_ctx* ctx) { return ctx ? ctx->d_results : nullptr; }

svo_status svo_sparse_align(svo_ctx* ctx, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats, int n_feats,
                            const svo_align_params* prm, svo_align_result* results, svo_align_level_stats* stats)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);  // the five phases share the staged state: one caller at a time runs all of them (recursive lock)
    svo_status st = svo_sparse_align_stage(ctx, jobs, n_jobs, feats, n_feats, prm, stats != nullptr);
    if (st != SVO_OK) return st;
    if ((st = svo_sparse_align_h2d(ctx)) != SVO_OK) return st;
    if ((st = svo_sparse_align_launch(ctx)) != SVO_OK) return st;
    if ((st = svo_sparse_align_d2h(ctx)) != SVO_OK) return st;
    return svo_sparse_align_fetch(ctx, results, stats);
}

// ------------------------------------------------------------------------------------------------
// feature alignment
// ------------------------------------------------------------------------------------------------
svo_status svo_feature_align_stage(svo_ctx* ctx, const svo_fa_item* items, int n, const svo_fa_params* prm)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (n < 0 || !prm || (n > 0 && !items)) SVO_FAIL(SVO_ERR_INVALID, "svo_feature_align: null argument");
    if (n > ctx->cfg.max_fa_items) SVO_FAIL(SVO_ERR_CAPACITY, "svo_feature_align: more items than max_fa_items");
    if (prm->patch_size < 1 || prm->patch_size > 8 || prm->mode < SVO_LM_FAITHFUL || prm->mode > SVO_GN)
        SVO_FAIL(SVO_ERR_INVALID, "svo_feature_align: patch_size must be 1..8 and mode valid");
    for (int i = 0; i < n; i++)
        if (bad_slot(ctx, items[i].ref_slot) || bad_slot(ctx, items[i].cur_slot))
            SVO_FAIL(SVO_ERR_INVALID, "svo_feature_align: frame slot out of range");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n) std::memcpy(ctx->h_fa_items, items, sizeof(svo_fa_item) * n);
    ctx->staged_fa        = n;
    ctx->staged_fa_params = *prm;
    return SVO_OK;
}

svo_status svo_feature_align_h2d(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (ctx->staged_fa)
        SVO_CUDA(cudaMemcpyAsync(ctx->d_fa_items, ctx->h_fa_items, sizeof(svo_fa_item) * ctx->staged_fa,
                                 cudaMemcpyHostToDevice, ctx->stream));
    return SVO_OK;
}

svo_status svo_feature_align_launch(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    const svo_status ws = wait_ingest(ctx);
    if (ws != SVO_OK) return ws;
    return launch_feature_align(ctx);
}

svo_status svo_feature_align_d2h(svo_ctx* ctx)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (ctx->staged_fa)
        SVO_CUDA(cudaMemcpyAsync(ctx->h_fa_results, ctx->d_fa_results, sizeof(svo_fa_result) * ctx->staged_fa,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    return SVO_OK;
}

svo_status svo_feature_align_fetch(svo_ctx* ctx, svo_fa_result* results)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (results && ctx->staged_fa) std::memcpy(results, ctx->h_fa_results, sizeof(svo_fa_result) * ctx->staged_fa);
    return SVO_OK;
}

svo_status svo_feature_align(svo_ctx* ctx, const svo_fa_item* items, int n, const svo_fa_params* prm, svo_fa_result* results)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);  // (as svo_sparse_align)
    svo_status st = svo_feature_align_stage(ctx, items, n, prm);
    if (st != SVO_OK) return st;
    if ((st = svo_feature_align_h2d(ctx)) != SVO_OK) return st;
    if ((st = svo_feature_align_launch(ctx)) != SVO_OK) return st;
    if ((st = svo_feature_align_d2h(ctx)) != SVO_OK) return st;
    return svo_feature_align_fetch(ctx, results);
}

// ------------------------------------------------------------------------------------------------
// Map::reprojectMap
// ------------------------------------------------------------------------------------------------
svo_status svo_reproject_map(svo_ctx* ctx, int cur_slot, const double T_cur[7], const svo_reproj_candidate* cands, int n, int cell_size,
                             const int32_t* cell_order, int n_cells, int max_matches, const svo_fa_params* fa,
                             svo_reproj_match* matches, int* n_matches, uint8_t* projected)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (!T_cur || n < 0 || (n > 0 && !cands) || !cell_order || !fa || !matches || !n_matches || cell_size < 4 || max_matches < 0 ||
        bad_slot(ctx, cur_slot))
        SVO_FAIL(SVO_ERR_INVALID, "svo_reproject_map: bad arguments");
    const LevelGeom& g = ctx->arena.geom[0];
    const int gridCols = (g.w + cell_size - 1) / cell_size, gridRows = (g.h + cell_size - 1) / cell_size;  // Map::initializeGrid, :226-227
    if (n_cells != gridCols * gridRows) SVO_FAIL(SVO_ERR_INVALID, "svo_reproject_map: cell_order must list ceil(w/cell) * ceil(h/cell) cells");
    const int maxItems = max_matches + 1;  // the walk stops once m_matches EXCEEDS max_matches (:484-487)
    if (n > ctx->cfg.max_fa_items || maxItems > ctx->cfg.max_fa_items || n_cells > ctx->sel_cap_cells)
        SVO_FAIL(SVO_ERR_CAPACITY, "svo_reproject_map: more candidates / matches than max_fa_items, or more cells than the context holds");
    if (fa->patch_size < 1 || fa->patch_size > 8 || fa->mode < SVO_LM_FAITHFUL || fa->mode > SVO_GN)
        SVO_FAIL(SVO_ERR_INVALID, "svo_reproject_map: bad FeatureAlignment parameters");
    for (int i = 0; i < n; i++)
        if (bad_slot(ctx, cands[i].ref_slot)) SVO_FAIL(SVO_ERR_INVALID, "svo_reproject_map: candidate frame slot out of range");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    const size_t nf = (size_t)std::max(1, ctx->cfg.max_fa_items), nc = (size_t)std::max(1, ctx->sel_cap_cells);
    auto up16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    const size_t offOrder = up16(sizeof(svo_reproj_candidate) * nf), offOut = up16(offOrder + sizeof(int32_t) * nc);
    const size_t offProj = up16(offOut + sizeof(svo_reproj_match) * nf), offCount = up16(offProj + nf);
    if (!ctx->h_rp) {
        // Zero-copy: ~50 KB in, ~8 KB out.  The kernels read the candidates and the cell order from, and write the
        // matches, flags and count to, mapped page-locked host memory; no DMA round trip on either side of four tiny
        // launches (every input word is read once, coalesced).
        SVO_CUDA(cudaHostAlloc(&ctx->h_rp, offCount + 16, cudaHostAllocMapped));
        unsigned char* d = nullptr;
        SVO_CUDA(cudaHostGetDevicePointer(&d, ctx->h_rp, 0));
        ctx->d_rp_cands     = reinterpret_cast<svo_reproj_candidate*>(d);
        ctx->d_rp_order     = reinterpret_cast<int32_t*>(d + offOrder);
        ctx->d_rp_out       = reinterpret_cast<svo_reproj_match*>(d + offOut);
        ctx->d_rp_projected = d + offProj;
        ctx->d_rp_count     = reinterpret_cast<int32_t*>(d + offCount);
        SVO_CUDA(cudaMalloc(&ctx->d_rp_matches, sizeof(svo_reproj_match) * nf));
        SVO_CUDA(cudaMalloc(&ctx->d_rp_px, sizeof(double) * 2 * nf));
    }
    svo_status st = wait_ingest(ctx);
    if (st != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n) std::memcpy(ctx->h_rp, cands, sizeof(svo_reproj_candidate) * n);
    std::memcpy(ctx->h_rp + offOrder, cell_order, sizeof(int32_t) * n_cells);
    if ((st = launch_reproject_map(ctx, cur_slot, T_cur, n, cell_size, n_cells, gridCols, maxItems, *fa)) != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_matches = *reinterpret_cast<const int32_t*>(ctx->h_rp + offCount);
    std::memcpy(matches, ctx->h_rp + offOut, sizeof(svo_reproj_match) * *n_matches);
    if (projected && n) std::memcpy(projected, ctx->h_rp + offProj, (size_t)n);
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------
// epipolar search (depth filter)
// ------------------------------------------------------------------------------------------------
svo_status svo_epipolar_match(svo_ctx* ctx, const svo_epi_item* items, int n, const svo_epi_params* prm, svo_epi_result* results)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (n < 0 || !prm || (n > 0 && (!items || !results))) SVO_FAIL(SVO_ERR_INVALID, "svo_epipolar_match: null argument");
    if (n > ctx->cfg.max_fa_items) SVO_FAIL(SVO_ERR_CAPACITY, "svo_epipolar_match: more seeds than max_fa_items");
    if (prm->patch_size < 1 || prm->patch_size > 8 || !(prm->patch_size & 1) || prm->mean_mode < SVO_MEAN_EIGEN_U8 ||
        prm->mean_mode > SVO_MEAN_EXACT)
        SVO_FAIL(SVO_ERR_INVALID, "svo_epipolar_match: patch_size must be odd, 1..7; mean_mode valid");
    for (int i = 0; i < n; i++)
        if (bad_slot(ctx, items[i].ref_slot) || bad_slot(ctx, items[i].cur_slot))
            SVO_FAIL(SVO_ERR_INVALID, "svo_epipolar_match: frame slot out of range");
    if (n == 0) return SVO_OK;
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(ctx->h_epi_items, items, sizeof(svo_epi_item) * n);
    svo_status st = wait_ingest(ctx);
    if (st != SVO_OK) return st;
    SVO_CUDA(cudaMemcpyAsync(ctx->d_epi_items, ctx->h_epi_items, sizeof(svo_epi_item) * n, cudaMemcpyHostToDevice, ctx->stream));
    if ((st = launch_epipolar_match(ctx, n, *prm)) != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(results, ctx->h_epi_results, sizeof(svo_epi_result) * n);
    return SVO_OK;
}

// ------------------------------------------------------------------------------------------------
// pyramidal Lucas-Kanade (initialisation)
// ------------------------------------------------------------------------------------------------
svo_status svo_klt_track(svo_ctx* ctx, int ref_slot, int cur_slot, const float* prev_pts, float* next_pts, int n,
                         const svo_klt_params* prm, uint8_t* status, float* err)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (n < 0 || !prm || (n > 0 && (!prev_pts || !next_pts || !status))) SVO_FAIL(SVO_ERR_INVALID, "svo_klt_track: null argument");
    if (bad_slot(ctx, ref_slot) || bad_slot(ctx, cur_slot)) SVO_FAIL(SVO_ERR_INVALID, "svo_klt_track: frame slot out of range");
    if (prm->win < 3 || prm->win > 21 || prm->max_level < 0) SVO_FAIL(SVO_ERR_INVALID, "svo_klt_track: win must be 3..21, max_level >= 0");
    if (n > ctx->cfg.max_fa_items) SVO_FAIL(SVO_ERR_CAPACITY, "svo_klt_track: more points than max_fa_items");
    // buildOpticalFlowPyramid stops before the first level that is not larger than the window
    int top = 0;
    {
        int w = ctx->arena.geom[0].w, h = ctx->arena.geom[0].h;
        if (w <= prm->win || h <= prm->win) SVO_FAIL(SVO_ERR_INVALID, "svo_klt_track: the image must be larger than the window");
        for (int l = 1; l <= prm->max_level; l++) {
            w = (w + 1) / 2, h = (h + 1) / 2;
            if (w <= prm->win || h <= prm->win) break;
            top = l;
        }
    }
    if (top >= ctx->cfg.levels) SVO_FAIL(SVO_ERR_UNSUPPORTED, "svo_klt_track: the context holds fewer pyramid levels than max_level needs");
    if (n == 0) return SVO_OK;
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    const size_t nf = (size_t)std::max(1, ctx->cfg.max_fa_items);
    if (!ctx->h_klt) {
        // Zero-copy: a few KB in, a few KB out.  The kernel reads the points from and writes the results to mapped
        // page-locked host memory directly (one PCIe transaction per warp each way), which saves the two DMA round trips
        // (~10 us each) that cudaMemcpyAsync would put in front of and behind a ~40 us kernel.
        SVO_CUDA(cudaHostAlloc(&ctx->h_klt, nf * 21, cudaHostAllocMapped));
        unsigned char* d = nullptr;
        SVO_CUDA(cudaHostGetDevicePointer(&d, ctx->h_klt, 0));
        ctx->d_klt_prev   = reinterpret_cast<float2*>(d);
        ctx->d_klt_next   = reinterpret_cast<float2*>(d + 8 * nf);
        ctx->d_klt_err    = reinterpret_cast<float*>(d + 16 * nf);
        ctx->d_klt_status = d + 20 * nf;
    }
    svo_status st = wait_ingest(ctx);
    if (st != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    float* hPrev     = reinterpret_cast<float*>(ctx->h_klt);
    float* hNext     = hPrev + 2 * nf;
    float* hErr      = hNext + 2 * nf;
    uint8_t* hStatus = reinterpret_cast<uint8_t*>(hErr + nf);
    std::memcpy(hPrev, prev_pts, sizeof(float2) * n);
    std::memcpy(hNext, next_pts, sizeof(float2) * n);
    if ((st = launch_klt_track(ctx, ref_slot, cur_slot, n, *prm, top)) != SVO_OK) return st;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(next_pts, hNext, sizeof(float2) * n);
    std::memcpy(status, hStatus, (size_t)n);
    if (err) std::memcpy(err, hErr, sizeof(float) * n);
    return SVO_OK;
}

}  // extern "C"
