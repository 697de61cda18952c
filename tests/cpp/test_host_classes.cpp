// test_host_classes.cpp -- exercises the reference-shaped host classes (host/svo_host.hpp) end to end on the GPU and
// checks them against the CPU oracle.  Shaped after the reference's own tests (tests/test_image_pyramid.cpp:20-60:
// level count, sizes, base image identity) plus the numeric checks the reference never had.
// usage: test_host_classes <ref.u8> <cur.u8> <w> <h> <Tcur_true 7 doubles...>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "svo_host.hpp"
#include "svo_oracle.h"

using namespace svo;

static int g_fail = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            g_fail++;                                                      \
        }                                                                  \
    } while (0)

static Mat8 readImage(const char* path, int w, int h)
{
    Mat8 m(h, w);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char*>(m.ptr()), (std::streamsize)w * h);
    if (!f) {
        std::printf("cannot read %s\n", path);
        std::exit(2);
    }
    return m;
}

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
    const Mat8 refImg = readImage(argv[1], w, h), curImg = readImage(argv[2], w, h);
    const double K[4] = {721.5377, 721.5377, 609.5593, 172.8540};  // resource/kitti.yaml:7-8

    Device::current() = std::make_shared<Device>(w, h, K, /*levels*/ 4, /*maxFrames*/ 8);
    auto camera       = std::make_shared<PinholeCamera>(w, h, K[0], K[1], K[2], K[3], 0, 0, 0, 0, 0);

    // ---- ImagePyramid (tests/test_image_pyramid.cpp) ----
    {
        ImagePyramid pyr(4);
        CHECK(pyr.getSizeImagePyramid() == 0);
        pyr.createImagePyramid(refImg, 4);
        CHECK(pyr.getSizeImagePyramid() == 4);
        CHECK(pyr.getBaseImage().ptr() == refImg.ptr());  // level 0 shares the input buffer
        CHECK(pyr.getBaseImageSize().width == w && pyr.getBaseImageSize().height == h);
        CHECK(pyr.getImageSizeAtLevel(1).width == (w + 1) / 2 && pyr.getImageSizeAtLevel(1).height == (h + 1) / 2);
        CHECK(pyr.getImageSizeAtLevel(7).width == 0);
        std::vector<uint8_t> ip(orc_pyramid_bytes(w, h, 4)), gp(ip.size());
        orc_build_pyramid(refImg.ptr(), w, h, w, 4, ip.data(), gp.data());
        size_t off = 0;
        for (int l = 0; l < 4; l++) {
            const Mat8& a = pyr.getImageAtLevel(l);
            const Mat8& g = pyr.getGradientAtLevel(l);
            CHECK(std::memcmp(a.ptr(), ip.data() + off, (size_t)a.rows * a.cols) == 0);
            CHECK(std::memcmp(g.ptr(), gp.data() + off, (size_t)g.rows * g.cols) == 0);
            off += (size_t)a.rows * a.cols;
        }
    }
    // ---- Frame: corrupted image throws (src/frame.cpp:20-24) ----
    {
        bool threw = false;
        try {
            Frame bad(camera, Mat8(10, 10), 4, 0, nullptr);
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()) == "Image Corrupted";
        }
        CHECK(threw);
    }

    auto kf  = std::make_shared<Frame>(camera, refImg, 4, 0, nullptr);  // an empty last keyframe
    auto ref = std::make_shared<Frame>(camera, refImg, 4, 1, kf);
    auto cur = std::make_shared<Frame>(camera, curImg, 4, 2, kf);

    // ---- FeatureSelection::gradientMagnitudeByValue ----
    FeatureSelection selector(w, h, 30);
    CHECK(selector.m_gridRows == h / 30 + 1 && selector.m_gridCols == w / 30 + 1);
    selector.setCellInGridOccupancy(Vec2(45.0, 40.0));  // cell (1, 1) is taken
    selector.gradientMagnitudeByValue(ref, 50, true);
    {
        std::vector<uint8_t> occ((size_t)selector.m_gridRows * selector.m_gridCols, 0);
        occ[(size_t)1 * selector.m_gridCols + 1] = 1;
        std::vector<int32_t> want(3 * occ.size());
        const Mat8& g = ref->m_imagePyramid.getBaseGradientImage();
        const int n   = orc_grid_select(g.ptr(), w, h, w, 30, 50, occ.data(), want.data(), (int)occ.size());
        CHECK((int)ref->numberObservation() == n);
        for (int i = 0; i < n && i < (int)ref->numberObservation(); i++) {
            const auto& f = ref->m_features[i];
            CHECK(f->m_pixelPosition.x() == want[3 * i] && f->m_pixelPosition.y() == want[3 * i + 1] &&
                  f->m_gradientMagnitude == want[3 * i + 2]);
            CHECK(std::fabs(f->m_bearingVec.norm() - 1.0) < 1e-12);
        }
        bool occupancyReset = true;
        for (bool b : selector.m_occupancyGrid) occupancyReset &= !b;
        CHECK(occupancyReset);  // src/feature_selection.cpp:145
        bool threw = false;
        try {
            selector.gradientMagnitudeByValue(ref, 50, false);
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        CHECK(threw);
    }
    // ---- FeatureSelection::gradientMagnitudeWithSSC (what System calls on keyframes) ----
    {
        auto probe = std::make_shared<Frame>(camera, refImg, 4, 7, kf);
        FeatureSelection ssc(w, h, 30);
        ssc.setCellInGridOccupancy(Vec2(45.0, 40.0));
        ssc.gradientMagnitudeWithSSC(probe, 50, 250, true);
        std::vector<uint8_t> occ((size_t)ssc.m_gridRows * ssc.m_gridCols, 0);
        occ[(size_t)1 * ssc.m_gridCols + 1] = 1;
        std::vector<int32_t> want(3 * 4096);
        int32_t info[4];
        const Mat8& g = probe->m_imagePyramid.getBaseGradientImage();
        const int n   = orc_select_ssc(g.ptr(), w, h, w, 50, 250, 30, occ.data(), 1, want.data(), 4096, info);
        CHECK((int)probe->numberObservation() == n && n > 0);
        for (int i = 0; i < n && i < (int)probe->numberObservation(); i++) {
            const auto& f = probe->m_features[i];
            CHECK(f->m_pixelPosition.x() == want[3 * i] && f->m_pixelPosition.y() == want[3 * i + 1] && f->m_gradientMagnitude == want[3 * i + 2]);
        }
        std::printf("gradientMagnitudeWithSSC: %d features from %d keypoints (width %d, %d iterations)\n", n, info[0], info[1], info[2]);
    }
    // 3D points on the plane z = 15 m (the scene the images were rendered from); every 7th feature has no point
    for (size_t i = 0; i < ref->m_features.size(); i++) {
        if (i % 7 == 3) continue;
        auto& f      = ref->m_features[i];
        const Vec3 b = f->m_bearingVec;
        auto p       = std::make_shared<Point>(ref->camera2world(b * (15.0 / b.z())));
        f->setPoint(p);
    }

    // ---- ImageAlignment::align, faithful mode == the reference's behaviour ----
    std::vector<orc_feature> of;
    for (const auto& f : ref->m_features) {
        orc_feature a{};
        a.px[0] = f->m_pixelPosition.x();
        a.px[1] = f->m_pixelPosition.y();
        for (int i = 0; i < 3; i++) a.bearing[i] = f->m_bearingVec[i];
        a.has_point = f->m_point != nullptr;
        if (f->m_point)
            for (int i = 0; i < 3; i++) a.point[i] = f->m_point->m_position[i];
        of.push_back(a);
    }
    std::vector<uint8_t> rp(orc_pyramid_bytes(w, h, 4)), rg(rp.size()), cp(rp.size()), cg(rp.size());
    orc_build_pyramid(refImg.ptr(), w, h, w, 4, rp.data(), rg.data());
    orc_build_pyramid(curImg.ptr(), w, h, w, 4, cp.data(), cg.data());
    const double I7[7] = {0, 0, 0, 1, 0, 0, 0};
    for (int mode = 0; mode < 3; mode++) {
        ImageAlignment aligner(5, 0, 3, 6);
        aligner.m_mode         = mode;
        aligner.m_maxIteration = 30;
        cur->m_absPose         = ref->m_absPose;  // prior
        const double err       = aligner.align(ref, cur);
        orc_align_params prm{5, 0, 3, mode, 30, ORC_MEDIAN_EXACT};
        double T[7];
        std::memcpy(T, I7, sizeof(T));
        int32_t st         = 0;
        const double oerr  = orc_sparse_align(rp.data(), rp.data(), cp.data(), w, h, of.data(), (int)of.size(), 0, I7, I7, K, &prm,
                                              T, nullptr, &st);
        double got[7];
        cur->m_absPose.params(got);
        double dq = 0, dt = 0;
        for (int i = 0; i < 4; i++) dq = std::fmax(dq, std::fabs(got[i] - T[i]));
        for (int i = 4; i < 7; i++) dt = std::fmax(dt, std::fabs(got[i] - T[i]));
        std::printf("ImageAlignment mode %d: rmse gpu %.6f oracle %.6f  |dq| %.2e |dt| %.2e status %d/%d\n", mode, err, oerr, dq,
                    dt, aligner.m_status, st);
        CHECK(dq < 5e-6 && dt < 1e-4);  // 1e-5 rad = 5e-6 in quaternion units
        CHECK(std::fabs(err - oerr) <= 1e-4 * oerr);
        if (mode == 0) CHECK(aligner.m_status == st);
        if (mode != 0 && argc >= 12) {  // iterated modes recover the rendered motion
            double e = 0;
            for (int i = 4; i < 7; i++) e = std::fmax(e, std::fabs(got[i] - std::atof(argv[5 + i])));
            CHECK(e < 5e-3);
        }
    }
    {
        auto empty = std::make_shared<Frame>(camera, refImg, 4, 3, kf);
        ImageAlignment aligner(5, 0, 3, 6);
        CHECK(aligner.align(empty, cur) == 0.0);  // src/image_alignment.cpp:27-28
    }

    // ---- FeatureAlignment::align (single calls and the batched form agree with the oracle) ----
    {
        FeatureAlignment fa(7, 0, 3);
        std::vector<FeatureAlignment::Item> items;
        for (size_t i = 50; i < 90 && i < ref->m_features.size(); i++)
            items.push_back({ref->m_features[i], cur, Vec2(ref->m_features[i]->m_pixelPosition.x() + 0.7,
                                                           ref->m_features[i]->m_pixelPosition.y() - 0.4)});
        std::vector<Vec2> px;
        std::vector<double> err;
        fa.alignBatch(items, px, err);
        for (size_t i = 0; i < items.size(); i++) {
            const auto& f       = items[i].refFeature;
            const double rpx[2] = {f->m_pixelPosition.x(), f->m_pixelPosition.y()};
            double p[2]         = {items[i].pixelPos.x(), items[i].pixelPos.y()};
            orc_fa_params prm{7, ORC_LM_FAITHFUL, 20, ORC_MEDIAN_EXACT};
            int32_t st = 0, it = 0;
            const double oerr = orc_feature_align(rg.data(), cg.data(), w, h, rpx, nullptr, p, &prm, &st, &it);
            CHECK(std::fabs(px[i].x() - p[0]) < 1e-7 && std::fabs(px[i].y() - p[1]) < 1e-7);
            CHECK((std::isnan(oerr) && std::isnan(err[i])) || std::fabs(err[i] - oerr) < 1e-7);
            if (i == 0) {
                Vec2 single = items[i].pixelPos;
                const double e1 = fa.align(f, cur, single);
                CHECK(single.x() == px[i].x() && single.y() == px[i].y() && (e1 == err[i] || (std::isnan(e1) && std::isnan(err[i]))));
            }
        }
    }
    // ---- FrontEnd: the same stages as ONE graph launch == the calls above made one by one ----
    {
        ImageAlignment aligner(5, 0, 3, 6);
        cur->m_absPose   = ref->m_absPose;
        const double err = aligner.align(ref, cur);
        double want[7];
        cur->m_absPose.params(want);
        auto probe = std::make_shared<Frame>(camera, curImg, 4, 9, kf);
        FeatureSelection sel2(w, h, 30);
        sel2.gradientMagnitudeByValue(probe, 50, true);
        FeatureAlignment fa(7, 0, 3);
        std::vector<FeatureAlignment::Item> items;
        std::vector<size_t> which;
        for (size_t i = 0; i < ref->m_features.size(); i++) {
            const auto& f = ref->m_features[i];
            if (!f->m_point) continue;
            const Vec3 pc = cur->m_absPose * f->m_point->m_position;
            const Vec2 px = camera->project2d(pc);
            if (pc.z() > 0 && camera->isInFrame(px, 3.0)) {
                items.push_back({f, cur, px});
                which.push_back(i);
            }
        }
        std::vector<Vec2> px;
        std::vector<double> e;
        fa.alignBatch(items, px, e);

        auto cur2       = std::make_shared<Frame>(camera, curImg, 4, 10, kf);
        cur2->m_absPose = ref->m_absPose;
        FrontEnd fe(5, 0, 3, 30, 50, 7, 2048);
        const FrontEnd::Result r = fe.run(ref, cur2);
        double got[7];
        cur2->m_absPose.params(got);
        for (int i = 0; i < 7; i++) CHECK(got[i] == want[i]);
        CHECK(r.alignError == err);
        CHECK(r.newFeatures.size() == probe->numberObservation());
        for (size_t i = 0; i < r.newFeatures.size() && i < probe->numberObservation(); i++)
            CHECK(r.newFeatures[i].pixelPosition.x() == probe->m_features[i]->m_pixelPosition.x() &&
                  r.newFeatures[i].pixelPosition.y() == probe->m_features[i]->m_pixelPosition.y());
        size_t nMatched = 0;
        for (bool b : r.matched) nMatched += b;
        CHECK(nMatched == items.size());
        for (size_t k = 0; k < which.size(); k++) {
            const size_t i = which[k];
            CHECK(r.matched[i]);
            CHECK(std::fabs(r.pixelPosition[i].x() - px[k].x()) < 1e-6 && std::fabs(r.pixelPosition[i].y() - px[k].y()) < 1e-6);
            CHECK((std::isnan(r.matchError[i]) && std::isnan(e[k])) || std::fabs(r.matchError[i] - e[k]) < 1e-6 * (1.0 + std::fabs(e[k])));
        }
        std::printf("FrontEnd: %zu new features, %zu of %zu tracked features matched, rmse %.6f\n", r.newFeatures.size(), nMatched,
                    r.matched.size(), r.alignError);
    }
    // ---- algorithm::matchEpipolarConstraint (depth filter), single call and batch vs the oracle ----
    {
        cur->m_absPose = SE3::fromParams(&std::vector<double>{std::atof(argv[5]), std::atof(argv[6]), std::atof(argv[7]), std::atof(argv[8]),
                                                              std::atof(argv[9]), std::atof(argv[10]), std::atof(argv[11])}[0]);
        double Trel[7];
        (cur->m_absPose * ref->m_absPose.inverse()).params(Trel);
        std::vector<algorithm::EpipolarSeed> seeds;
        for (size_t i = 0; i < ref->m_features.size() && seeds.size() < 64; i++) {
            const auto& f = ref->m_features[i];
            if (!f->m_point) continue;
            const double d = (ref->m_absPose * f->m_point->m_position).norm();
            seeds.push_back({f, d * 1.1, d * 0.6, d * 1.7});
        }
        std::vector<bool> found;
        std::vector<double> depth;
        algorithm::matchEpipolarConstraintBatch(cur, seeds, 7, found, depth);
        size_t nFound = 0;
        for (size_t i = 0; i < seeds.size(); i++) {
            const auto& f = seeds[i].refFeature;
            const double px[2] = {f->m_pixelPosition.x(), f->m_pixelPosition.y()};
            const double b[3]  = {f->m_bearingVec[0], f->m_bearingVec[1], f->m_bearingVec[2]};
            orc_epi_params prm{7, ORC_MEAN_EIGEN_U8};
            orc_epi_result o{};
            orc_epipolar_match(refImg.ptr(), curImg.ptr(), w, h, K, Trel, px, b, seeds[i].initialDepth, seeds[i].minDepth,
                               seeds[i].maxDepth, &prm, &o);
            CHECK(found[i] == (o.found != 0));
            if (o.found) CHECK(std::fabs(depth[i] - o.depth) <= 1e-9 * o.depth);
            nFound += found[i];
        }
        double est = -1.0;
        auto f0    = seeds[0].refFeature;
        const bool ok = algorithm::matchEpipolarConstraint(ref, cur, f0, 7, seeds[0].initialDepth, seeds[0].minDepth, seeds[0].maxDepth, est);
        CHECK(ok == found[0] && (!ok || est == depth[0]));
        std::printf("matchEpipolarConstraint: %zu of %zu seeds matched\n", nFound, seeds.size());
    }
    // ---- Map::reprojectMap (one batched pass) vs the oracle: same cells, same candidates, same refined positions ----
    {
        std::vector<ReprojectionCandidate> cands;
        std::vector<orc_reproj_candidate> oc;
        for (size_t i = 0; i < ref->m_features.size(); i++) {
            const auto& f = ref->m_features[i];
            if (!f->m_point) continue;
            const uint32_t type = (uint32_t)((i * 7) % 4);  // GOOD / DELETED / CANDIDATE / UNKNOWN mixed
            cands.push_back({f, type});
            orc_reproj_candidate c{};
            c.ref_slot = 0;
            c.type     = (int32_t)type;
            c.ref_px[0] = f->m_pixelPosition.x(), c.ref_px[1] = f->m_pixelPosition.y();
            for (int k = 0; k < 3; k++) c.point[k] = f->m_point->m_position[k];
            oc.push_back(c);
        }
        const int cell = 30, cols = (w + cell - 1) / cell, rows = (h + cell - 1) / cell;
        std::vector<int32_t> order(cols * rows);
        for (size_t i = 0; i < order.size(); i++) order[i] = (int32_t)((i * 37) % order.size());  // 37 is coprime to the cell count below
        CHECK(order.size() % 37 != 0);
        std::vector<bool> projected;
        const auto matches = reprojectMap(cur, cands, cell, order, 150, &projected);
        const Mat8& gr = ref->m_imagePyramid.getGradientAtLevel(0);
        const Mat8& gc = cur->m_imagePyramid.getGradientAtLevel(0);
        const uint8_t* grads[1] = {gr.ptr()};
        double T[7];
        cur->m_absPose.params(T);
        orc_fa_params fa{7, ORC_LM_FAITHFUL, 20, ORC_MEDIAN_EXACT};
        std::vector<double> om(151 * 6);
        std::vector<uint8_t> op(oc.size());
        const int m = orc_reproject_map(grads, gc.ptr(), w, h, K, T, oc.data(), (int)oc.size(), cell, order.data(), (int)order.size(), 150, &fa,
                                        om.data(), op.data());
        CHECK((size_t)m == matches.size() && m > 10);
        for (size_t i = 0; i < oc.size(); i++) CHECK(projected[i] == (op[i] != 0));
        for (int i = 0; i < m && (size_t)i < matches.size(); i++) {
            CHECK(matches[i].cell == (int32_t)om[i * 6] && matches[i].candidate == (size_t)om[i * 6 + 1]);
            CHECK(std::fabs(matches[i].pixelPosition.x() - om[i * 6 + 2]) < 1e-7 && std::fabs(matches[i].pixelPosition.y() - om[i * 6 + 3]) < 1e-7);
            CHECK(cands[matches[i].candidate].pointType != 1);
        }
        std::printf("reprojectMap: %zu candidates -> %zu matches\n", cands.size(), matches.size());
    }
    // ---- algorithm::computeOpticalFlowSparse (System's initialisation) vs the oracle's calcOpticalFlowPyrLK ----
    {
        const size_t n0 = ref->numberObservation(), c0 = cur->numberObservation();
        std::vector<float> prev(2 * n0), next(2 * n0), oerr(n0);
        for (size_t i = 0; i < n0; i++) {
            prev[2 * i] = next[2 * i] = (float)ref->m_features[i]->m_pixelPosition.x();
            prev[2 * i + 1] = next[2 * i + 1] = (float)ref->m_features[i]->m_pixelPosition.y();
        }
        std::vector<uint8_t> ost(n0);
        orc_klt_params kp{11, 3, 30, 1, 1e-4, 1e-4};
        CHECK(orc_klt_track(refImg.ptr(), curImg.ptr(), w, h, prev.data(), next.data(), (int)n0, &kp, ost.data(), oerr.data()) == 3);
        size_t tracked = 0;
        for (size_t i = 0; i < n0; i++) tracked += ost[i];
        auto r2 = ref, c2 = cur;
        CHECK(!algorithm::computeOpticalFlowSparse(r2, c2, 11, 1e9));  // median disparity below the threshold: refFrame untouched
        CHECK(ref->numberObservation() == n0);
        const size_t added = cur->numberObservation() - c0;            // ... but the tracked features were added already (:66-76)
        CHECK(added + 2 >= tracked && added <= tracked + 2);
        size_t k = 0, close = 0;
        for (size_t i = 0; i < n0 && c0 + k < cur->numberObservation(); i++) {
            if (!ost[i]) continue;
            const auto& f = cur->m_features[c0 + k++];
            close += std::fabs(f->m_pixelPosition.x() - next[2 * i]) < 2e-3 && std::fabs(f->m_pixelPosition.y() - next[2 * i + 1]) < 2e-3;
        }
        CHECK(close + 4 >= tracked);
        CHECK(algorithm::computeOpticalFlowSparse(r2, c2, 11, 0.0));
        CHECK(ref->numberObservation() + 2 >= tracked && ref->numberObservation() <= tracked + 2);  // untracked features erased (:86-102)
        CHECK(algorithm::computeMedian({3.0, 1.0, 2.0}) == 2.0 && algorithm::computeMedian({4.0, 1.0, 3.0, 2.0}) == 2.5);
        std::printf("computeOpticalFlowSparse: %zu of %zu features tracked\n", tracked, n0);
    }
    Device::current().reset();
    std::printf(g_fail ? "FAILED (%d checks)\n" : "ALL HOST-CLASS CHECKS PASSED\n", g_fail);
    return g_fail ? 1 : 0;
}
