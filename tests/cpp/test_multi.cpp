// test_multi.cpp -- the single-process multi-GPU layer (svo_multi_*, csrc/multi.cu) driven from C++ as a host of many
// sequences would: a batch of independent frame pairs is cut into contiguous shards over n devices; the result must equal,
// bit for bit, the same batch run on one context.  Usage: test_multi <dir> <n_pairs> <w> <h> <n_devices>
// <dir> holds ref.u8, cur.u8 (n_pairs frames each), jobs.bin (svo_align_job[n_pairs], slots 0..n-1 / n..2n-1),
// feats.bin (svo_align_feature[]).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/svo_b200.h"

static std::vector<unsigned char> slurp(const std::string& p)
{
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", p.c_str());
        exit(2);
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> b(n);
    if (fread(b.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return b;
}

#define CHECK(x)                                                              \
    do {                                                                      \
        svo_status s__ = (x);                                                 \
        if (s__ != SVO_OK) {                                                  \
            fprintf(stderr, "%s failed: %d (%s)\n", #x, s__, err());          \
            return 1;                                                         \
        }                                                                     \
    } while (0)

int main(int argc, char** argv)
{
    if (argc < 6) return 2;
    const std::string dir = argv[1];
    const int n = atoi(argv[2]), w = atoi(argv[3]), h = atoi(argv[4]), D = atoi(argv[5]);
    auto ref = slurp(dir + "/ref.u8"), cur = slurp(dir + "/cur.u8"), jb = slurp(dir + "/jobs.bin"), fb = slurp(dir + "/feats.bin");
    std::vector<svo_align_job> jobs(n);
    memcpy(jobs.data(), jb.data(), sizeof(svo_align_job) * n);
    const int nf = (int)(fb.size() / sizeof(svo_align_feature));
    const svo_align_feature* feats = reinterpret_cast<const svo_align_feature*>(fb.data());
    svo_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.width = w, cfg.height = h, cfg.levels = 4, cfg.max_frames = 2 * n, cfg.max_jobs = n, cfg.max_features = 512, cfg.max_fa_items = 16;
    cfg.K[0] = cfg.K[1] = 721.5377, cfg.K[2] = 609.5593, cfg.K[3] = 172.854;
    svo_align_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.patch_size = 5, prm.min_level = 0, prm.max_level = 3, prm.mode = SVO_GN, prm.max_iter = 30;

    // ---- one context, one device: the whole batch ----
    svo_ctx* one = nullptr;
    svo_multi* m = nullptr;
    auto err = [&]() { return m ? svo_multi_last_error(m) : svo_last_error(one); };
    CHECK(svo_create(&cfg, &one));
    CHECK(svo_frames_upload(one, 0, n, ref.data(), w, (int64_t)w * h));
    CHECK(svo_frames_upload(one, n, n, cur.data(), w, (int64_t)w * h));
    std::vector<svo_align_result> want(n), got(n);
    std::vector<svo_align_level_stats> wstats((size_t)n * 4), gstats((size_t)n * 4);
    CHECK(svo_sparse_align(one, jobs.data(), n, feats, nf, &prm, want.data(), wstats.data()));
    svo_destroy(one);
    one = nullptr;

    // ---- D devices: contiguous shards; pair j lives in LOCAL slots (j - lo) and per + (j - lo) of its device ----
    CHECK(svo_multi_create(&cfg, nullptr, D, &m));
    if (svo_multi_devices(m) != D) return 1;
    int per = 0;
    for (int i = 0; i < D; i++) {
        int lo, hi;
        svo_multi_shard(n, D, i, &lo, &hi);
        per = hi - lo > per ? hi - lo : per;
    }
    CHECK(svo_multi_frames_upload(m, 0, n, ref.data(), w, (int64_t)w * h, 0));
    CHECK(svo_multi_frames_upload(m, per, n, cur.data(), w, (int64_t)w * h, 1));
    std::vector<svo_align_job> mj = jobs;
    for (int i = 0; i < D; i++) {
        int lo, hi;
        svo_multi_shard(n, D, i, &lo, &hi);
        for (int j = lo; j < hi; j++) mj[j].ref_slot = mj[j].kf_slot = j - lo, mj[j].cur_slot = per + (j - lo);
    }
    CHECK(svo_multi_sparse_align(m, mj.data(), n, feats, nf, &prm, got.data(), gstats.data()));
    int bad = 0;
    for (int j = 0; j < n; j++) bad += memcmp(&want[j], &got[j], sizeof(svo_align_result)) != 0;
    bad += memcmp(wstats.data(), gstats.data(), sizeof(svo_align_level_stats) * wstats.size()) != 0;
    // the phase-split form and the device-timed launches
    CHECK(svo_multi_sparse_align_stage(m, mj.data(), n, feats, nf, &prm, 0));
    double ms = 0;
    CHECK(svo_multi_time_launches(m, 1, 3, &ms));
    std::vector<svo_align_result> again(n);
    CHECK(svo_multi_sparse_align_fetch(m, again.data(), nullptr));
    for (int j = 0; j < n; j++) bad += memcmp(&want[j], &again[j], sizeof(svo_align_result)) != 0;
    CHECK(svo_multi_sync(m));
    printf("devices %d pairs %d: %d mismatches, 3 launches %.3f ms (slowest device)\n", D, n, bad, ms);
    svo_multi_destroy(m);
    if (bad) return 1;
    printf("MULTI-GPU LAYER OK\n");
    return 0;
}
