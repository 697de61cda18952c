// multi.cu -- one process, several B200s: the batch of independent frame pairs is cut into contiguous shards, one per
// device; every device has its own context (svo_ctx) and its own host worker thread, so the per-device calls of a batch
// (jobs / features H2D, the alignment kernel, results D2H) run concurrently.  There is no exchange between the shards
// (a VO sequence is sequential, the batch is over independent pairs -- SURVEY 8e): the "final gather" is every device
// writing its block of ONE caller-owned result array.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>

#include "ctx.h"

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<svo_status()> task;
    bool pending = false, stop = false;
    svo_status result = SVO_OK;

    void loop()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return pending || stop; });
            if (stop) return;
            std::function<svo_status()> t = std::move(task);
            lk.unlock();
            const svo_status r = t();
            lk.lock();
            result  = r;
            pending = false;
            cv.notify_all();
        }
    }
    void post(std::function<svo_status()> t)
    {
        std::lock_guard<std::mutex> lk(mu);
        task    = std::move(t);
        pending = true;
        cv.notify_all();
    }
    svo_status wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !pending; });
        return result;
    }
};

}  // namespace

struct svo_multi {
    std::vector<svo_ctx*> ctx;
    std::vector<int> dev;
    std::vector<Worker*> workers;
    std::vector<std::vector<svo_align_job>> jobs;  // per device: its shard with rebased feature offsets
    std::vector<int> lo, hi;                       // shard of the staged batch
    std::vector<cudaEvent_t> ev0, ev1;
    int staged_levels = 0;
    std::string err;
    std::mutex mu;
};

namespace {

// every device runs f(i) on its worker; the first failure is reported
svo_status run_all(svo_multi* m, const std::function<svo_status(int)>& f)
{
    const int n = (int)m->ctx.size();
    for (int i = 0; i < n; i++) m->workers[i]->post([&f, i] { return f(i); });
    svo_status first = SVO_OK;
    for (int i = 0; i < n; i++) {
        const svo_status r = m->workers[i]->wait();
        if (r != SVO_OK && first == SVO_OK) {
            first  = r;
            m->err = "device " + std::to_string(m->dev[i]) + ": " + svo_last_error(m->ctx[i]);
        }
    }
    return first;
}

}  // namespace

extern "C" {

void svo_multi_shard(int n_items, int n_parts, int part, int* lo, int* hi)
{
    // contiguous blocks whose sizes differ by at most one (the first n_items % n_parts blocks are the larger ones)
    const int q = n_items / n_parts, r = n_items % n_parts;
    const int a = part * q + (part < r ? part : r);
    if (lo) *lo = a;
    if (hi) *hi = a + q + (part < r ? 1 : 0);
}

svo_status svo_multi_create(const svo_config* cfg, const int32_t* devices, int n_devices, svo_multi** out)
{
    if (!cfg || !out || n_devices < 1 || n_devices > 64) return SVO_ERR_INVALID;
    *out = nullptr;
    svo_multi* m = new svo_multi();
    for (int i = 0; i < n_devices; i++) {
        svo_config c = *cfg;
        c.device     = devices ? devices[i] : i;
        c.stream     = nullptr;  // every context owns its stream
        svo_ctx* x   = nullptr;
        const svo_status st = svo_create(&c, &x);
        if (st != SVO_OK) {
            for (svo_ctx* y : m->ctx) svo_destroy(y);
            delete m;
            return st;
        }
        m->ctx.push_back(x);
        m->dev.push_back(c.device);
    }
    m->jobs.resize(n_devices);
    m->lo.assign(n_devices, 0);
    m->hi.assign(n_devices, 0);
    m->ev0.resize(n_devices);
    m->ev1.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        Worker* w = new Worker();
        w->th     = std::thread([w] { w->loop(); });
        m->workers.push_back(w);
    }
    // the worker binds its device once; events live on that device
    const svo_status st = run_all(m, [m](int i) {
        if (cudaSetDevice(m->dev[i]) != cudaSuccess) return (svo_status)SVO_ERR_CUDA;
        if (cudaEventCreate(&m->ev0[i]) != cudaSuccess || cudaEventCreate(&m->ev1[i]) != cudaSuccess) return (svo_status)SVO_ERR_CUDA;
        return (svo_status)SVO_OK;
    });
    if (st != SVO_OK) {
        svo_multi_destroy(m);
        return st;
    }
    *out = m;
    return SVO_OK;
}

void svo_multi_destroy(svo_multi* m)
{
    if (!m) return;
    for (size_t i = 0; i < m->workers.size(); i++) {
        Worker* w = m->workers[i];
        w->wait();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->stop = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
    }
    for (size_t i = 0; i < m->ctx.size(); i++) {
        cudaSetDevice(m->dev[i]);
        if (i < m->ev0.size() && m->ev0[i]) cudaEventDestroy(m->ev0[i]);
        if (i < m->ev1.size() && m->ev1[i]) cudaEventDestroy(m->ev1[i]);
        svo_destroy(m->ctx[i]);
    }
    delete m;
}

int svo_multi_devices(const svo_multi* m) { return m ? (int)m->ctx.size() : 0; }
svo_ctx* svo_multi_ctx(svo_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
const char* svo_multi_last_error(const svo_multi* m) { return m ? m->err.c_str() : "null svo_multi"; }

svo_status svo_multi_frames_upload(svo_multi* m, int first_slot, int n, const uint8_t* imgs, int pitch, int64_t frame_stride, int prefetch)
{
    if (!m || !imgs || n < 0) return SVO_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    const int D = (int)m->ctx.size();
    return run_all(m, [=](int i) {
        int lo, hi;
        svo_multi_shard(n, D, i, &lo, &hi);
        if (hi == lo) return (svo_status)SVO_OK;
        const uint8_t* p = imgs + (int64_t)lo * frame_stride;
        return prefetch ? svo_frames_prefetch(m->ctx[i], first_slot, hi - lo, p, pitch, frame_stride)
                        : svo_frames_upload(m->ctx[i], first_slot, hi - lo, p, pitch, frame_stride);
    });
}

svo_status svo_multi_sparse_align_stage(svo_multi* m, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats, int n_feats,
                                        const svo_align_params* prm, int want_stats)
{
    if (!m || !jobs || !prm || n_jobs < 0 || n_feats < 0 || (n_feats > 0 && !feats)) return SVO_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    const int D      = (int)m->ctx.size();
    m->staged_levels = prm->max_level - prm->min_level + 1;
    return run_all(m, [=](int i) {
        int lo, hi;
        svo_multi_shard(n_jobs, D, i, &lo, &hi);
        m->lo[i] = lo, m->hi[i] = hi;
        std::vector<svo_align_job>& J = m->jobs[i];
        J.assign(jobs + lo, jobs + hi);
        // the shard's slice of the global feature array, offsets rebased to it
        int64_t fmin = n_feats, fmax = 0;
        for (const svo_align_job& j : J) {
            fmin = std::min<int64_t>(fmin, j.feat_offset);
            fmax = std::max<int64_t>(fmax, (int64_t)j.feat_offset + j.n_ref + j.n_kf);
        }
        if (J.empty() || fmax <= fmin) fmin = fmax = 0;
        if (fmax > n_feats) return (svo_status)SVO_ERR_INVALID;
        for (svo_align_job& j : J) j.feat_offset -= (int32_t)fmin;
        svo_status st = svo_sparse_align_stage(m->ctx[i], J.data(), hi - lo, feats + fmin, (int)(fmax - fmin), prm, want_stats);
        if (st != SVO_OK) return st;
        return svo_sparse_align_h2d(m->ctx[i]);
    });
}

svo_status svo_multi_sparse_align_launch(svo_multi* m)
{
    if (!m) return SVO_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    return run_all(m, [=](int i) { return m->hi[i] > m->lo[i] ? svo_sparse_align_launch(m->ctx[i]) : (svo_status)SVO_OK; });
}

svo_status svo_multi_sparse_align_fetch(svo_multi* m, svo_align_result* results, svo_align_level_stats* stats)
{
    if (!m || !results) return SVO_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    const int L = m->staged_levels;
    return run_all(m, [=](int i) {
        if (m->hi[i] == m->lo[i]) return (svo_status)SVO_OK;
        svo_status st = svo_sparse_align_d2h(m->ctx[i]);
        if (st != SVO_OK) return st;
        // every device writes its own block of the caller's arrays: this is the gather
        return svo_sparse_align_fetch(m->ctx[i], results + m->lo[i], stats ? stats + (int64_t)m->lo[i] * L : nullptr);
    });
}

svo_status svo_multi_sparse_align(svo_multi* m, const svo_align_job* jobs, int n_jobs, const svo_align_feature* feats, int n_feats,
                                  const svo_align_params* prm, svo_align_result* results, svo_align_level_stats* stats)
{
    svo_status st = svo_multi_sparse_align_stage(m, jobs, n_jobs, feats, n_feats, prm, stats != nullptr);
    if (st != SVO_OK) return st;
    if ((st = svo_multi_sparse_align_launch(m)) != SVO_OK) return st;
    return svo_multi_sparse_align_fetch(m, results, stats);
}

svo_status svo_multi_sync(svo_multi* m)
{
    if (!m) return SVO_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    return run_all(m, [=](int i) { return svo_sync(m->ctx[i]); });
}

// `steps` launches of the staged batch on every device, timed on each device with CUDA events on its stream after
// `warmup` untimed ones; *ms_max = the slowest device's time for all steps (what a multi-GPU number must be)
svo_status svo_multi_time_launches(svo_multi* m, int warmup, int steps, double* ms_max)
{
    if (!m || !ms_max || steps < 1 || warmup < 0) return SVO_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<float> ms(m->ctx.size(), 0.f);
    const svo_status st = run_all(m, [&](int i) {
        if (m->hi[i] == m->lo[i]) return (svo_status)SVO_OK;
        svo_ctx* c = m->ctx[i];
        for (int k = 0; k < warmup; k++) {
            const svo_status s = svo_sparse_align_launch(c);
            if (s != SVO_OK) return s;
        }
        if (svo_sync(c) != SVO_OK) return (svo_status)SVO_ERR_CUDA;
        cudaStream_t str = (cudaStream_t)svo_stream(c);
        if (cudaEventRecord(m->ev0[i], str) != cudaSuccess) return (svo_status)SVO_ERR_CUDA;
        for (int k = 0; k < steps; k++) {
            const svo_status s = svo_sparse_align_launch(c);
            if (s != SVO_OK) return s;
        }
        if (cudaEventRecord(m->ev1[i], str) != cudaSuccess) return (svo_status)SVO_ERR_CUDA;
        if (cudaEventSynchronize(m->ev1[i]) != cudaSuccess) return (svo_status)SVO_ERR_CUDA;
        if (cudaEventElapsedTime(&ms[i], m->ev0[i], m->ev1[i]) != cudaSuccess) return (svo_status)SVO_ERR_CUDA;
        return (svo_status)SVO_OK;
    });
    double mx = 0;
    for (float v : ms) mx = std::max<double>(mx, v);
    *ms_max = mx;
    return st;
}

}  // extern "C"
