// ctx.h -- internal context of libsvo_b200.so (not part of the C ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/svo_b200.h"

struct LevelGeom {
    int w, h, pitch;       // pitch: multiple of 16 B so that TMA descriptors over a plane are legal
    int64_t plane_stride;  // bytes between the same level of consecutive slots (multiple of 256)
};

// One arena per stack (image / gradient) and level: [slot][h][pitch] u8.
struct PyramidArena {
    int levels;
    LevelGeom geom[SVO_MAX_LEVELS];
    uint8_t* img[SVO_MAX_LEVELS];
    uint8_t* grad[SVO_MAX_LEVELS];
};

// Device-side view handed to kernels by value.
struct ArenaView {
    const uint8_t* img[SVO_MAX_LEVELS];
    const uint8_t* grad[SVO_MAX_LEVELS];
    int w[SVO_MAX_LEVELS], h[SVO_MAX_LEVELS], pitch[SVO_MAX_LEVELS];
    long long plane_stride[SVO_MAX_LEVELS];
};

// svo_frontend_run: one captured CUDA graph per (slots, parameters) configuration
struct FrontendGraph {
    svo_frontend_params prm;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
};

// every entry point that takes a context holds its lock for the duration of the call: calls from several threads (the
// reference's depth-filter thread next to the tracker) are serialised, never interleaved.  Recursive: the one-call forms
// (svo_sparse_align, svo_feature_align) are built from the phase-split entry points.
#define SVO_LOCK(ctx) std::lock_guard<std::recursive_mutex> svo_lock_guard_((ctx)->mu)

struct svo_ctx {
    std::recursive_mutex mu;
    svo_config cfg;
    cudaStream_t stream;
    bool own_stream;
    int sm_count;
    int max_smem_optin;
    std::string err;
    int64_t launches;

    PyramidArena arena;

    // pinned staging for image upload
    // frame ingest pipeline: host -> (pinned staging, pageable sources only) -> dense device staging on the copy
    // stream -> k_repack + pyramid kernels on the main stream; two buffers so the DMA of chunk i+1 overlaps the
    // kernels of chunk i
    cudaStream_t copy_stream;
    cudaStream_t ingest_stream;  // svo_frames_prefetch: repack + pyramid kernels, concurrent with the main stream
    cudaStream_t pyr_stream;     // where launch_repack / launch_pyramid_build enqueue (main or ingest stream)
    cudaEvent_t ev_ingest_done;  // ingest stream: the last prefetch is complete
    bool ingest_pending;         // the main stream has not waited for ev_ingest_done yet
    // svo_frames_prefetch from page-locked memory: two whole-batch dense staging buffers, so that the DMA of batch
    // k+1 never waits for a kernel (it needs no SM while the alignment of batch k owns them)
    uint8_t* d_pf_stage[2];
    size_t pf_stage_bytes[2];
    cudaEvent_t ev_pf_consumed[2];  // ingest stream: the repack kernels have read d_pf_stage[b]
    cudaEvent_t ev_pf_chunk[64];    // copy stream: a chunk landed
    int pf_next, pf_chunk_next;
    cudaEvent_t ev_jobs_h2d;        // copy stream: the jobs / features of the staged alignment batch are on the device
    uint8_t* h_img_stage[2];     // pinned, dense frames
    uint8_t* d_img_stage[2];     // device, dense frames (+ slack)
    int stage_frames;            // frames per buffer
    cudaEvent_t ev_h2d[2];       // copy stream: chunk landed in d_img_stage[b]
    cudaEvent_t ev_consumed[2];  // main stream: k_repack has read d_img_stage[b]
    int stage_next;

    // selection
    uint32_t* d_cell_best;       // per cell packed (value << 24 | ~index)
    uint8_t* d_occupancy;        // per cell
    svo_feature_px* d_sel_out;   // compacted features
    int32_t* d_sel_count;
    svo_feature_px* h_sel_out;   // pinned
    int32_t* h_sel_count;        // pinned
    unsigned char* h_grid;       // svo_select_grid: mapped page-locked block the compaction writes to (count | records)
    uint8_t* h_occupancy;        // pinned
    int sel_cap_cells;
    uint32_t* d_ssc_key;         // svo_select_ssc: champion key / state per SSC cell (allocated on first use)
    uint32_t* d_ssc_state;
    unsigned char* h_ssc;        // mapped page-locked block the kernel writes its results to: records | count | info
    svo_feature_px* d_ssc_out;   // device views of its parts
    int32_t* d_ssc_count;
    int32_t* d_ssc_info;
    int ssc_cluster;             // CTAs per cluster of k_select_ssc (decided on first use)
    bool sel_use_occupancy;

    // sparse alignment batch
    svo_align_job* h_jobs;       // pinned
    svo_align_feature* h_feats;  // pinned
    svo_align_result* h_results; // pinned
    svo_align_level_stats* h_stats;  // pinned
    svo_align_job* d_jobs;
    svo_align_feature* d_feats;
    svo_align_result* d_results;
    svo_align_level_stats* d_stats;
    float* d_scratch_tpl;        // [job][3][Nmax]  template intensity, gx, gy per patch pixel
    float* d_scratch_jac;        // [job][F][12]   image Jacobian rows per feature
    int64_t feats_cap;
    int scratch_area;            // patch area d_scratch_tpl is sized for
    long long* d_dbg;            // 64 per-phase cycle counters of job 0 (diagnostics)
    float* d_scratch2;           // fast path: per job world points + per level template blocks
    size_t scratch2_bytes;
    int staged_jobs, staged_feats, staged_levels, staged_want_stats;
    const svo_align_job* src_jobs;      // where the H2D of the staged batch reads from (pinned user memory or h_jobs)
    const svo_align_feature* src_feats;
    svo_align_params staged_params;
    int last_align_nt, last_align_c;    // shape of the last cluster launch (threads per CTA, CTAs per pair)

    // front-end graphs
    FrontendGraph fe_graphs[8];
    int fe_count;
    int fe_kernel_nodes;           // kernel nodes of the last instantiated front-end graph
    svo_align_result* h_fe_align;
    svo_fa_result* h_fe_fa;
    unsigned char* h_fe_sel;       // count (16 bytes) | selected features; all three are MAPPED: the graph's kernels write them

    // Map::reprojectMap batch (svo_reproject_map), capacity max_fa_items candidates / sel_cap_cells cells
    unsigned char* h_rp;  // mapped page-locked block: candidates | cell order | finished matches | projected flags | count
    svo_reproj_match* d_rp_out;   // device views of its parts (d_rp_cands, d_rp_order, d_rp_projected are views too)
    int32_t* d_rp_count;
    svo_reproj_candidate* d_rp_cands;
    svo_reproj_match* d_rp_matches;
    int32_t* d_rp_order;
    double* d_rp_px;
    uint8_t* d_rp_projected;

    // epipolar search batch (depth-filter seeds), capacity max_fa_items
    // svo_klt_track
    float2* d_klt_prev;
    float2* d_klt_next;
    uint8_t* d_klt_status;
    float* d_klt_err;
    unsigned char* h_klt;  // pinned: next, err, status
    svo_epi_item* h_epi_items;      // pinned
    svo_epi_result* h_epi_results;  // pinned
    svo_epi_item* d_epi_items;
    svo_epi_result* d_epi_results;

    // feature alignment batch
    svo_fa_item* h_fa_items;     // pinned
    svo_fa_result* h_fa_results; // pinned
    svo_fa_item* d_fa_items;
    svo_fa_result* d_fa_results;
    int staged_fa;
    svo_fa_params staged_fa_params;
};

#define SVO_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +     \
                       std::to_string(__LINE__) + ")";                                                   \
            return SVO_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

#define SVO_FAIL(code, msg)  \
    do {                     \
        ctx->err = (msg);    \
        return (code);       \
    } while (0)

inline ArenaView make_view(const PyramidArena& a)
{
    ArenaView v{};
    for (int l = 0; l < a.levels; l++) {
        v.img[l]          = a.img[l];
        v.grad[l]         = a.grad[l];
        v.w[l]            = a.geom[l].w;
        v.h[l]            = a.geom[l].h;
        v.pitch[l]        = a.geom[l].pitch;
        v.plane_stride[l] = a.geom[l].plane_stride;
    }
    return v;
}

// kernel launchers (defined in the .cu files)
svo_status launch_pyramid_build(svo_ctx* ctx, int first_slot, int n);
svo_status launch_repack(svo_ctx* ctx, const uint8_t* dsrc, long long src_pitch, long long src_frame_stride, int first_slot, int n);
svo_status launch_grid_select(svo_ctx* ctx, int slot, int cell, uint32_t thr, int rows, int cols, svo_feature_px* out = nullptr,
                              int32_t* count = nullptr);
svo_status launch_select_ssc(svo_ctx* ctx, int slot, uint32_t thr, int numCandidates, int cell, int rows, int cols, bool useOcc,
                             bool useBucketing, int maxOut);
svo_status launch_sparse_align(svo_ctx* ctx);
svo_status launch_feature_align(svo_ctx* ctx, svo_fa_result* results = nullptr);
svo_status launch_reproject_map(svo_ctx* ctx, int curSlot, const double T[7], int n, int cell, int nCells, int gridCols, int maxItems,
                                const svo_fa_params& fa);
svo_status launch_epipolar_match(svo_ctx* ctx, int n, const svo_epi_params& prm);
svo_status launch_klt_track(svo_ctx* ctx, int refSlot, int curSlot, int n, const svo_klt_params& prm, int topLevel);
size_t klt_smem_bytes(int win);
void frontend_release(svo_ctx* ctx);
void frontend_invalidate_graphs(svo_ctx* ctx);  // destroys the captured front-end graphs (their kernel arguments went stale)
size_t sparse_align_smem_bytes(int nthreads, int max_features, int patch_area);
bool sparse_align_v3_supported(const svo_ctx* ctx, int maxF);
svo_status launch_sparse_align_v3(svo_ctx* ctx, int maxF);
bool sparse_align_v5_supported(const svo_ctx* ctx, int maxF);
svo_status launch_sparse_align_v5(svo_ctx* ctx, int maxF);
