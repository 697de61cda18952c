// frontend.cu -- the per-frame front end as ONE CUDA graph launch (BASELINE config 3): what System::processNewFrame
// does with a new camera image between Frame::Frame (src/frame.cpp:26, src/system.cpp:36) and the end of
// Map::reprojectMap (src/system.cpp:313-330):
//     H2D image -> k_repack -> pyramids of both stacks           ImagePyramid::createImagePyramid
//     -> grid argmax on the new frame's gradient                  FeatureSelection::gradientMagnitudeByValue
//     -> sparse image alignment of (ref, lastKF) -> new frame     ImageAlignment::align
//     -> k_reproject_items: world2image of every tracked point with the ALIGNED pose (Frame::world2image,
//        src/frame.cpp:84-92) and the in-frame test of Map::addCandidateToFrame (src/map.cpp:601-602)
//     -> per-feature 2D alignment of those candidates             FeatureAlignment::align (src/map.cpp:538,608)
//     -> the pose, the new features and the refined pixel positions are written by those kernels straight to mapped
//        page-locked host memory (no device->host copy node in the chain; all host->device copies come first).
// The graph is captured once per configuration (frame slots + parameters) and cached; a call copies the inputs into
// the context's pinned mirrors, launches the graph and reads the pinned outputs.  No host round trip between stages.
#include <cstring>

#include "ctx.h"
#include "math.cuh"

namespace {

struct ReprojArgs {
    const svo_align_job* job;
    const svo_align_feature* feats;
    const svo_align_result* aligned;
    svo_align_result* alignedOut;  // mapped host copy of the alignment's result (thread 0 writes it)
    svo_fa_item* items;
    int capacity;
    int w, h;
    int border;
    double K[4];
};

// one thread per feature slot: item = (owner frame's gradient, feature pixel) -> (new frame, projected pixel)
__global__ void __launch_bounds__(128) k_reproject_items(const ReprojArgs a)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f == 0 && a.alignedOut) *a.alignedOut = *a.aligned;
    if (f >= a.capacity) return;
    const svo_align_job J = *a.job;
    svo_fa_item it;
    it.ref_slot   = -1;
    it.cur_slot   = J.cur_slot;
    it.ref_px[0] = it.ref_px[1] = it.px[0] = it.px[1] = 0.0;
    it.A[0] = it.A[3] = 1.0;
    it.A[1] = it.A[2] = 0.0;
    it.use_affine = 0;
    it.reserved   = 0;
    if (f < J.n_ref + J.n_kf) {
        const svo_align_feature ft = a.feats[J.feat_offset + f];
        it.ref_px[0]               = ft.px[0];
        it.ref_px[1]               = ft.px[1];
        if (ft.has_point) {
            svo::Pose T;
            for (int i = 0; i < 4; i++) T.q[i] = a.aligned->T_cur[i];
            for (int i = 0; i < 3; i++) T.t[i] = a.aligned->T_cur[4 + i];
            double pc[3];
            svo::quat_rotate(T.q, ft.point, pc);  // Frame::world2camera: R p + t
            pc[0] += T.t[0];
            pc[1] += T.t[1];
            pc[2] += T.t[2];
            const double u = a.K[0] * (pc[0] / pc[2]) + a.K[2];  // PinholeCamera::project2d, src/pinhole_camera.cpp:55-56
            const double v = a.K[1] * (pc[1] / pc[2]) + a.K[3];
            it.px[0]       = u;
            it.px[1]       = v;
            const double b = (double)a.border;
            // in front of the camera (Frame::isVisible, src/frame.cpp:72) and PinholeCamera::isInFrame(px, 3)
            if (pc[2] > 0.0 && u >= b && v >= b && u < a.w - b && v < a.h - b) it.ref_slot = f < J.n_ref ? J.ref_slot : J.kf_slot;
        }
    }
    a.items[f] = it;
}

bool same_params(const svo_frontend_params& a, const svo_frontend_params& b) { return std::memcmp(&a, &b, sizeof(a)) == 0; }

// everything of one frame on the main stream, in order; used eagerly once (allocations, attributes) and then captured
svo_status frontend_enqueue(svo_ctx* ctx, const svo_frontend_params& p)
{
    const LevelGeom& g = ctx->arena.geom[0];
    const int64_t frame_sz = (int64_t)g.w * g.h;
    const int rows = g.h / p.cell + 1, cols = g.w / p.cell + 1;
    cudaStream_t st = ctx->stream;
    svo_status rc;
    // every input first: one queue of host->device copies, none of them between two kernels of the chain
    SVO_CUDA(cudaMemcpyAsync(ctx->d_img_stage[0], ctx->h_img_stage[0], (size_t)frame_sz, cudaMemcpyHostToDevice, st));
    SVO_CUDA(cudaMemcpyAsync(ctx->d_occupancy, ctx->h_occupancy, (size_t)rows * cols, cudaMemcpyHostToDevice, st));
    SVO_CUDA(cudaMemcpyAsync(ctx->d_jobs, ctx->h_jobs, sizeof(svo_align_job), cudaMemcpyHostToDevice, st));
    SVO_CUDA(cudaMemcpyAsync(ctx->d_feats, ctx->h_feats, sizeof(svo_align_feature) * p.max_features, cudaMemcpyHostToDevice, st));
    // the results travel zero-copy: the kernels write them to mapped page-locked memory, so no device->host copy sits
    // between (or behind) the kernels either
    unsigned char* dSel       = nullptr;
    svo_align_result* dAlign  = nullptr;
    svo_fa_result* dFa        = nullptr;
    SVO_CUDA(cudaHostGetDevicePointer(&dSel, ctx->h_fe_sel, 0));
    SVO_CUDA(cudaHostGetDevicePointer(&dAlign, ctx->h_fe_align, 0));
    SVO_CUDA(cudaHostGetDevicePointer(&dFa, ctx->h_fe_fa, 0));
    // new frame
    if ((rc = launch_repack(ctx, ctx->d_img_stage[0], g.w, frame_sz, p.cur_slot, 1)) != SVO_OK) return rc;
    if ((rc = launch_pyramid_build(ctx, p.cur_slot, 1)) != SVO_OK) return rc;
    // new features on it
    ctx->sel_use_occupancy = true;
    if ((rc = launch_grid_select(ctx, p.cur_slot, p.cell, p.thr, rows, cols, reinterpret_cast<svo_feature_px*>(dSel + 16),
                                 reinterpret_cast<int32_t*>(dSel))) != SVO_OK)
        return rc;
    // pose of the new frame
    ctx->staged_jobs       = 1;
    ctx->staged_feats      = p.max_features;
    ctx->staged_levels     = p.align.max_level - p.align.min_level + 1;
    ctx->staged_want_stats = 0;
    ctx->staged_params     = p.align;
    // (<= 512 features: the single-CTA kernel, which owns no scratch memory; beyond: the cluster kernel, whose scratch
    // pointer is a captured argument -- frontend_invalidate_graphs drops the graphs whenever it is reallocated)
    if (sparse_align_v5_supported(ctx, p.max_features))
        rc = launch_sparse_align_v5(ctx, p.max_features);
    else if (sparse_align_v3_supported(ctx, p.max_features))
        rc = launch_sparse_align_v3(ctx, p.max_features);
    else
        SVO_FAIL(SVO_ERR_CAPACITY, "svo_frontend_run: more features than the alignment kernels hold");
    if (rc != SVO_OK) return rc;
    // candidates with the aligned pose, refined per feature
    ReprojArgs ra;
    ra.job        = ctx->d_jobs;
    ra.feats      = ctx->d_feats;
    ra.aligned    = ctx->d_results;
    ra.alignedOut = dAlign;
    ra.items      = ctx->d_fa_items;
    ra.capacity   = p.max_features;
    ra.w          = g.w;
    ra.h          = g.h;
    ra.border     = 3;
    for (int i = 0; i < 4; i++) ra.K[i] = ctx->cfg.K[i];
    k_reproject_items<<<(p.max_features + 127) / 128, 128, 0, st>>>(ra);
    ctx->launches++;
    ctx->staged_fa        = p.max_features;
    ctx->staged_fa_params = p.fa;
    if ((rc = launch_feature_align(ctx, dFa)) != SVO_OK) return rc;
    SVO_CUDA(cudaGetLastError());
    return SVO_OK;
}

// *ranEagerly: the configuration was new, its eager pass has already processed the staged inputs
svo_status frontend_graph(svo_ctx* ctx, const svo_frontend_params& p, cudaGraphExec_t* out, bool* ranEagerly)
{
    *ranEagerly = false;
    for (int i = 0; i < ctx->fe_count; i++)
        if (same_params(ctx->fe_graphs[i].prm, p)) {
            *out = ctx->fe_graphs[i].exec;
            return SVO_OK;
        }
    if (!ctx->h_fe_align) {
        SVO_CUDA(cudaHostAlloc(&ctx->h_fe_align, sizeof(svo_align_result), cudaHostAllocMapped));
        SVO_CUDA(cudaHostAlloc(&ctx->h_fe_fa, sizeof(svo_fa_result) * std::max(1, ctx->cfg.max_fa_items), cudaHostAllocMapped));
        SVO_CUDA(cudaHostAlloc(&ctx->h_fe_sel, 16 + sizeof(svo_feature_px) * ctx->sel_cap_cells, cudaHostAllocMapped));
    }
    // eager pass: scratch allocation and function attributes happen outside the capture
    svo_status rc = frontend_enqueue(ctx, p);
    if (rc != SVO_OK) return rc;
    SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->fe_count == 8) {  // evict the oldest
        cudaGraphExecDestroy(ctx->fe_graphs[0].exec);
        cudaGraphDestroy(ctx->fe_graphs[0].graph);
        for (int i = 1; i < 8; i++) ctx->fe_graphs[i - 1] = ctx->fe_graphs[i];
        ctx->fe_count = 7;
    }
    FrontendGraph& G = ctx->fe_graphs[ctx->fe_count];
    G.prm            = p;
    const int64_t launches_before = ctx->launches;
    SVO_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    rc = frontend_enqueue(ctx, p);
    G.graph       = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &G.graph);
    const int64_t eager_launches = ctx->launches - launches_before;  // kernels of one pass
    ctx->launches                = launches_before;                  // capturing is not launching
    auto drop = [&]() {  // no graph object outlives a failure
        if (G.graph) cudaGraphDestroy(G.graph);
        G.graph = nullptr;
    };
    if (rc != SVO_OK) {
        drop();
        return rc;
    }
    if (e != cudaSuccess) {
        drop();
        SVO_CUDA(e);
    }
    e = cudaGraphInstantiate(&G.exec, G.graph, 0);
    if (e != cudaSuccess) {
        drop();
        SVO_CUDA(e);
    }
    ctx->fe_kernel_nodes = 0;
    {
        size_t n = 0;
        std::vector<cudaGraphNode_t> nodes;
        e = cudaGraphGetNodes(G.graph, nullptr, &n);
        if (e == cudaSuccess) {
            nodes.resize(n);
            e = cudaGraphGetNodes(G.graph, nodes.data(), &n);
        }
        for (size_t i = 0; e == cudaSuccess && i < n; i++) {
            cudaGraphNodeType t;
            e = cudaGraphNodeGetType(nodes[i], &t);
            if (e == cudaSuccess && t == cudaGraphNodeTypeKernel) ctx->fe_kernel_nodes++;
        }
        if (e != cudaSuccess) {
            cudaGraphExecDestroy(G.exec);
            drop();
            SVO_CUDA(e);
        }
    }
    (void)eager_launches;
    ctx->fe_count++;
    *out        = G.exec;
    *ranEagerly = true;
    return SVO_OK;
}

}  // namespace

void frontend_invalidate_graphs(svo_ctx* ctx)
{
    for (int i = 0; i < ctx->fe_count; i++) {
        cudaGraphExecDestroy(ctx->fe_graphs[i].exec);
        cudaGraphDestroy(ctx->fe_graphs[i].graph);
    }
    ctx->fe_count = 0;
}

void frontend_release(svo_ctx* ctx)
{
    frontend_invalidate_graphs(ctx);
    if (ctx->h_fe_align) cudaFreeHost(ctx->h_fe_align);
    if (ctx->h_fe_fa) cudaFreeHost(ctx->h_fe_fa);
    if (ctx->h_fe_sel) cudaFreeHost(ctx->h_fe_sel);
    ctx->h_fe_align = nullptr;
    ctx->h_fe_fa    = nullptr;
    ctx->h_fe_sel   = nullptr;
}

extern "C" {

uint8_t* svo_frontend_image_buffer(svo_ctx* ctx) { return ctx ? ctx->h_img_stage[0] : nullptr; }

svo_status svo_frontend_run(svo_ctx* ctx, const svo_frontend_params* prm, const uint8_t* img, int pitch, const svo_align_job* job,
                            const svo_align_feature* feats, int n_feats, const uint8_t* occupancy, svo_frontend_result* result,
                            svo_feature_px* selected, int max_selected, svo_fa_result* refined)
{
    if (!ctx) return SVO_ERR_INVALID;
    SVO_LOCK(ctx);
    if (!prm || !img || !job || !result || n_feats < 0 || (n_feats > 0 && !feats) || max_selected < 0 || (max_selected > 0 && !selected))
        SVO_FAIL(SVO_ERR_INVALID, "svo_frontend_run: null argument");
    const LevelGeom& g = ctx->arena.geom[0];
    const svo_frontend_params& p = *prm;
    auto bad = [&](int s) { return s < 0 || s >= ctx->cfg.max_frames; };
    if (bad(p.ref_slot) || bad(p.kf_slot) || bad(p.cur_slot) || p.cell < 4 || pitch < g.w)
        SVO_FAIL(SVO_ERR_INVALID, "svo_frontend_run: bad slot, cell or pitch");
    if (p.max_features < 1 || p.max_features > ctx->cfg.max_features || p.max_features > ctx->cfg.max_fa_items ||
        n_feats > p.max_features || job->n_ref < 0 || job->n_kf < 0 || job->n_ref + job->n_kf > n_feats)
        SVO_FAIL(SVO_ERR_CAPACITY, "svo_frontend_run: feature counts above max_features / the context's capacities");
    if (p.align.patch_size != 4 && p.align.patch_size != 5)
        SVO_FAIL(SVO_ERR_UNSUPPORTED, "svo_frontend_run: the captured alignment is the cluster fast path (patch size 4 or 5)");
    if (p.align.min_level < 0 || p.align.max_level < p.align.min_level || p.align.max_level >= ctx->arena.levels ||
        p.align.mode < SVO_LM_FAITHFUL || p.align.mode > SVO_GN || p.fa.patch_size < 1 || p.fa.patch_size > 8 ||
        p.fa.mode < SVO_LM_FAITHFUL || p.fa.mode > SVO_GN)
        SVO_FAIL(SVO_ERR_INVALID, "svo_frontend_run: bad alignment parameters");
    const int rows = g.h / p.cell + 1, cols = g.w / p.cell + 1;
    if (rows * cols > ctx->sel_cap_cells) SVO_FAIL(SVO_ERR_CAPACITY, "svo_frontend_run: too many cells");
    SVO_CUDA(cudaSetDevice(ctx->cfg.device));
    {   // previous work that touches frame slots or the pinned mirrors
        SVO_CUDA(cudaStreamSynchronize(ctx->ingest_stream));
        ctx->ingest_pending = false;
        SVO_CUDA(cudaEventSynchronize(ctx->ev_jobs_h2d));
        SVO_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    // ---- inputs -> pinned mirrors ----
    if (img != ctx->h_img_stage[0]) {
        if (pitch == g.w)
            std::memcpy(ctx->h_img_stage[0], img, (size_t)g.w * g.h);
        else
            for (int y = 0; y < g.h; y++) std::memcpy(ctx->h_img_stage[0] + (int64_t)y * g.w, img + (int64_t)y * pitch, g.w);
    }
    svo_align_job j = *job;
    j.ref_slot      = p.ref_slot;
    j.kf_slot       = p.kf_slot;
    j.cur_slot      = p.cur_slot;
    j.feat_offset   = 0;
    ctx->h_jobs[0]  = j;
    if (n_feats) std::memcpy(ctx->h_feats, feats, sizeof(svo_align_feature) * n_feats);
    if (n_feats < p.max_features) std::memset(ctx->h_feats + n_feats, 0, sizeof(svo_align_feature) * (p.max_features - n_feats));
    if (occupancy)
        std::memcpy(ctx->h_occupancy, occupancy, (size_t)rows * cols);
    else
        std::memset(ctx->h_occupancy, 0, (size_t)rows * cols);
    // ---- one graph launch ----
    cudaGraphExec_t exec;
    bool ranEagerly = false;
    const svo_status rc = frontend_graph(ctx, p, &exec, &ranEagerly);
    if (rc != SVO_OK) return rc;
    if (!ranEagerly) {  // (a new configuration has just processed this frame while it was set up: no second pass)
        SVO_CUDA(cudaGraphLaunch(exec, ctx->stream));
        SVO_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->launches += ctx->fe_kernel_nodes;
    // ---- outputs ----
    result->align      = *ctx->h_fe_align;
    result->n_selected = *reinterpret_cast<const int32_t*>(ctx->h_fe_sel);
    int nfa = 0;
    for (int f = 0; f < n_feats; f++) nfa += ctx->h_fe_fa[f].status != SVO_ST_FAILED || ctx->h_fe_fa[f].iterations != 0;
    result->n_candidates = nfa;
    if (selected) std::memcpy(selected, ctx->h_fe_sel + 16, sizeof(svo_feature_px) * std::min(result->n_selected, max_selected));
    if (refined && n_feats) std::memcpy(refined, ctx->h_fe_fa, sizeof(svo_fa_result) * n_feats);
    return SVO_OK;
}

}  // extern "C"
