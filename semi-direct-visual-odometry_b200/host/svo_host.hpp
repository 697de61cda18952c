// svo_host.hpp -- host-side C++ mirror of the reference's hot-path classes over the C ABI of libsvo_b200.so.
//
// Same class names, constructor arguments, method names, argument meaning and error behaviour as the reference
// (citations relative to the reference tree), so System / Map / Estimator code compiles against them unchanged
// apart from the value types (svo_types.hpp stands in for Eigen / Sophus / cv::Mat, which this image lacks):
//   ImagePyramid      include/image_pyramid.hpp:23-149, src/image_pyramid.cpp
//   PinholeCamera     include/pinhole_camera.hpp, src/pinhole_camera.cpp:50-101,123-176 (no-distortion branch)
//   Point, Feature    include/point.hpp:26-40, include/feature.hpp:27-38, src/feature.cpp:6-45
//   Frame             include/frame.hpp:82-205, src/frame.cpp
//   FeatureSelection  include/feature_selection.hpp:35-86, src/feature_selection.cpp:19-25,91-146,269-287
//   ImageAlignment    include/image_alignment.hpp:15-73, src/image_alignment.cpp:25-67
//   FeatureAlignment  include/feature_alignment.hpp:15-44, src/feature_alignment.cpp:25-62
// All arithmetic of the hot path runs on the GPU; these classes only marshal.  No CPU fallback exists.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>

#include "svo_device.hpp"
#include "svo_types.hpp"

namespace svo {

// ------------------------------------------------------------------------------------------------------------
class PinholeCamera final
{
public:
    explicit PinholeCamera(int32_t width, int32_t height, double fx, double fy, double cx, double cy, double d0 = 0,
                           double d1 = 0, double d2 = 0, double d3 = 0, double d4 = 0)
        : m_width(width), m_height(height), m_fx(fx), m_fy(fy), m_cx(cx), m_cy(cy)
    {
        // the distortion branch (src/pinhole_camera.cpp:58-72,90-99) is out of scope: KITTI / denso have d = 0
        if (d0 != 0 || d1 != 0 || d2 != 0 || d3 != 0 || d4 != 0)
            throw std::invalid_argument("PinholeCamera: lens distortion is not supported on the GPU path");
    }
    PinholeCamera(const PinholeCamera&)            = delete;
    PinholeCamera& operator=(const PinholeCamera&) = delete;

    Vec2 project2d(double x, double y, double z) const { return Vec2(m_fx * (x / z) + m_cx, m_fy * (y / z) + m_cy); }
    Vec2 project2d(const Vec3& p) const { return project2d(p.x(), p.y(), p.z()); }
    Vec3 inverseProject2d(double x, double y) const { return Vec3((x - m_cx) / m_fx, (y - m_cy) / m_fy, 1.0).normalized(); }
    Vec3 inverseProject2d(const Vec2& p) const { return inverseProject2d(p.x(), p.y()); }
    double fx() const { return m_fx; }
    double fy() const { return m_fy; }
    double cx() const { return m_cx; }
    double cy() const { return m_cy; }
    int32_t width() const { return m_width; }
    int32_t height() const { return m_height; }
    bool isInFrame(const Vec2& p, double boundary = 0.0) const
    {
        return p.x() >= boundary && p.y() >= boundary && p.x() < m_width - boundary && p.y() < m_height - boundary;
    }

private:
    int32_t m_width, m_height;
    double m_fx, m_fy, m_cx, m_cy;
};

// ------------------------------------------------------------------------------------------------------------
// ImagePyramid: the stacks live on the device in one frame slot; cv::Mat-like host copies are fetched lazily.
class ImagePyramid final
{
public:
    explicit ImagePyramid(std::size_t maxImagePyramid, std::shared_ptr<Device> dev = Device::current())
        : m_dev(std::move(dev))
    {
        m_vecImages.reserve(maxImagePyramid);
        m_vecGradientImages.reserve(maxImagePyramid);
    }
    explicit ImagePyramid(const Mat8& baseImage, std::size_t maxImagePyramid, std::shared_ptr<Device> dev = Device::current())
        : ImagePyramid(maxImagePyramid, std::move(dev))
    {
        createImagePyramid(baseImage, maxImagePyramid);
    }
    ImagePyramid(const ImagePyramid&)            = delete;
    ImagePyramid& operator=(const ImagePyramid&) = delete;
    ~ImagePyramid() { clear(); }

    // src/image_pyramid.cpp:36-52 -- one upload, abs-gradient + pyrDown kernels for both stacks (asynchronous)
    void createImagePyramid(const Mat8& baseImage, std::size_t maxPyramidLevel)
    {
        if (!m_dev) throw std::runtime_error("ImagePyramid: no svo::Device (construct one and set Device::current())");
        if ((int)maxPyramidLevel > m_dev->levels()) throw std::invalid_argument("ImagePyramid: more levels than the device arena");
        clear();
        m_slot            = m_dev->acquireSlot();
        m_levels          = maxPyramidLevel;
        m_baseImageWidth  = baseImage.cols;
        m_baseImageHeight = baseImage.rows;
        m_dev->check(svo_frames_upload(m_dev->ctx(), m_slot, 1, baseImage.ptr(), baseImage.cols, 0), "svo_frames_upload");
        m_vecImages.assign(maxPyramidLevel, Mat8());
        m_vecGradientImages.assign(maxPyramidLevel, Mat8());
        m_vecImages[0] = baseImage;  // level 0 shares the caller's buffer, as the reference (no clone)
    }
    const Mat8& getImageAtLevel(std::size_t level) const { return fetch(level, 0); }
    const Mat8& getGradientAtLevel(std::size_t level) const { return fetch(level, 1); }
    const Mat8& getBaseImage() const { return fetch(0, 0); }
    const Mat8& getBaseGradientImage() const { return fetch(0, 1); }
    const std::vector<Mat8>& getAllImages() const
    {
        for (std::size_t l = 0; l < m_levels; l++) fetch(l, 0);
        return m_vecImages;
    }
    std::size_t getSizeImagePyramid() const { return m_vecImages.size(); }
    Size getImageSizeAtLevel(std::size_t level) const
    {
        if (level >= m_levels) return Size(0, 0);  // src/image_pyramid.cpp:113-119
        int w = 0, h = 0;
        svo_level_dims(m_dev->ctx(), (int)level, &w, &h, nullptr);
        return Size(w, h);
    }
    Size getBaseImageSize() const { return Size((int)m_baseImageWidth, (int)m_baseImageHeight); }
    void clear()
    {
        if (m_slot >= 0 && m_dev) {
            svo_sync(m_dev->ctx());
            m_dev->releaseSlot(m_slot);
        }
        m_slot   = -1;
        m_levels = 0;
        m_vecImages.clear();
        m_vecGradientImages.clear();
    }
    int slot() const { return m_slot; }  // device handle of this pyramid (what the alignment calls take)
    const std::shared_ptr<Device>& device() const { return m_dev; }

private:
    const Mat8& fetch(std::size_t level, int which) const
    {
        if (level >= m_levels) throw std::out_of_range("ImagePyramid: level out of range");
        Mat8& m = which ? m_vecGradientImages[level] : m_vecImages[level];
        if (m.empty()) {
            const Size s = getImageSizeAtLevel(level);
            m            = Mat8(s.height, s.width);
            m_dev->check(svo_frame_download(m_dev->ctx(), m_slot, (int)level, which, m.ptr(), s.width), "svo_frame_download");
        }
        return m;
    }
    std::shared_ptr<Device> m_dev;
    int m_slot                    = -1;
    std::size_t m_levels          = 0;
    std::size_t m_baseImageWidth  = 0;
    std::size_t m_baseImageHeight = 0;
    mutable std::vector<Mat8> m_vecImages;
    mutable std::vector<Mat8> m_vecGradientImages;
};

// ------------------------------------------------------------------------------------------------------------
class Frame;
class Feature;

class Point final
{
public:
    explicit Point(const Vec3& point3D) : m_id(counter()++), m_position(point3D) {}
    uint32_t m_id;
    Vec3 m_position;
    std::vector<std::shared_ptr<Feature>> m_features;

private:
    static uint32_t& counter()
    {
        static uint32_t c = 0;
        return c;
    }
};

class Frame final
{
public:
    // include/frame.hpp:82-86; throws "Image Corrupted" as src/frame.cpp:20-24
    explicit Frame(const std::shared_ptr<PinholeCamera>& camera, const Mat8& img, uint32_t maxImagePyramid, uint64_t timestamp,
                   const std::shared_ptr<Frame> lastKeyframe, std::shared_ptr<Device> dev = Device::current())
        : m_id(counter()++), m_camera(camera), m_imagePyramid(maxImagePyramid, std::move(dev)), m_keyFrame(false),
          m_timestamp(timestamp), m_lastKeyframe(lastKeyframe)
    {
        if (img.empty() || img.cols != camera->width() || img.rows != camera->height())
            throw std::runtime_error("Image Corrupted");
        m_imagePyramid.createImagePyramid(img, maxImagePyramid);
    }
    Frame(const Frame&)            = delete;
    Frame& operator=(const Frame&) = delete;

    void setKeyframe() { m_keyFrame = true; }
    bool isKeyframe() const { return m_keyFrame; }
    void addFeature(std::shared_ptr<Feature>& feature) { m_features.emplace_back(feature); }
    std::size_t numberObservation() const { return m_features.size(); }
    uint32_t numberObservationWithPoints() const;
    Vec3 world2camera(const Vec3& p) const { return m_absPose * p; }
    Vec3 camera2world(const Vec3& p) const { return m_absPose.inverse() * p; }           // src/frame.cpp:94-97
    Vec2 camera2image(const Vec3& p) const { return m_camera->project2d(p); }            // :99-102
    Vec2 world2image(const Vec3& p) const { return camera2image(world2camera(p)); }
    Vec3 image2camera(const Vec2& px, double depth) const { return m_camera->inverseProject2d(px) * depth; }
    Vec3 image2world(const Vec2& px, double depth) const { return camera2world(image2camera(px, depth)); }
    Vec3 cameraInWorld() const { return m_absPose.inverse().translation(); }  // -R^T t, src/frame.cpp:116-120
    bool isVisible(const Vec3& p) const
    {
        const Vec3 c = world2camera(p);
        return c.z() >= 0.0 && m_camera->isInFrame(camera2image(c));
    }

    uint64_t m_id;
    std::shared_ptr<PinholeCamera> m_camera;
    SE3 m_absPose;  // world -> camera, include/frame.hpp:198
    ImagePyramid m_imagePyramid;
    std::vector<std::shared_ptr<Feature>> m_features;
    bool m_keyFrame;
    uint64_t m_timestamp;
    std::shared_ptr<Frame> m_lastKeyframe;

private:
    static uint64_t& counter()
    {
        static uint64_t c = 0;
        return c;
    }
};

class Feature final
{
public:
    enum class FeatureType : uint32_t { EDGE, EDGELET, CORNER };
    // src/feature.cpp:6-45
    explicit Feature(const std::shared_ptr<Frame>& frame, const Vec2& pixelPosition, uint8_t level,
                     const FeatureType& type = FeatureType::EDGE)
        : Feature(frame, pixelPosition, 1.0, 0.0, level, type)
    {
    }
    explicit Feature(const std::shared_ptr<Frame>& frame, const Vec2& pixelPosition, double gradientMagnitude,
                     double gradientOrientation, uint8_t level, const FeatureType& type = FeatureType::EDGE)
        : m_id(counter()++), m_frame(frame), m_type(type), m_pixelPosition(pixelPosition),
          m_homogenous(pixelPosition.x(), pixelPosition.y(), 1.0), m_bearingVec(frame->m_camera->inverseProject2d(pixelPosition)),
          m_gradientMagnitude(gradientMagnitude), m_gradientOrientation(gradientOrientation), m_level(level), m_point(nullptr)
    {
    }
    void setPoint(std::shared_ptr<Point>& point) { m_point = point; }

    uint64_t m_id;
    std::shared_ptr<Frame> m_frame;
    FeatureType m_type;
    Vec2 m_pixelPosition;
    Vec3 m_homogenous;
    Vec3 m_bearingVec;
    double m_gradientMagnitude;
    double m_gradientOrientation;
    uint8_t m_level;
    std::shared_ptr<Point> m_point;

private:
    static uint64_t& counter()
    {
        static uint64_t c = 0;
        return c;
    }
};

inline uint32_t Frame::numberObservationWithPoints() const
{
    uint32_t n = 0;
    for (const auto& f : m_features) n += f->m_point != nullptr;
    return n;
}

// ------------------------------------------------------------------------------------------------------------
class FeatureSelection final
{
public:
    explicit FeatureSelection(int32_t width, int32_t height, int32_t cellSize)  // src/feature_selection.cpp:19-25
        : m_cellSize(cellSize), m_gridRows(height / cellSize + 1), m_gridCols(width / cellSize + 1),
          m_occupancyGrid((size_t)m_gridRows * m_gridCols, false)
    {
    }
    FeatureSelection(const FeatureSelection&)            = delete;
    FeatureSelection& operator=(const FeatureSelection&) = delete;

    // src/feature_selection.cpp:91-146.  Appends Features to frame->m_features in cell raster order.
    void gradientMagnitudeByValue(std::shared_ptr<Frame>& frame, uint32_t detectionThreshold, bool useBucketing = true)
    {
        if (!useBucketing)  // the reference's non-bucketing branch reads the u8 image as float (:153): rejected, SURVEY 9.7
            throw std::invalid_argument("gradientMagnitudeByValue(useBucketing=false) is not supported");
        const auto& dev = frame->m_imagePyramid.device();
        std::vector<uint8_t> occ(m_occupancyGrid.begin(), m_occupancyGrid.end());
        std::vector<svo_feature_px> out(occ.size());
        int n = 0;
        dev->check(svo_select_grid(dev->ctx(), frame->m_imagePyramid.slot(), m_cellSize, detectionThreshold, occ.data(),
                                   out.data(), (int)out.size(), &n), "svo_select_grid");
        for (int i = 0; i < n; i++) {
            auto f = std::make_shared<Feature>(frame, Vec2(out[i].x, out[i].y), (double)out[i].magnitude, 0.0, 0);
            frame->addFeature(f);
        }
        m_imgGradientMagnitude = frame->m_imagePyramid.getBaseGradientImage();  // computeImageGradient, :250-267
        resetGridOccupancy();                                                     // :145
    }
    // src/feature_selection.cpp:27-89 (SSC :165-248): what System calls on every keyframe (src/system.cpp:81,253,429).
    // Appends Features in the order the reference does (sorted-keypoint order; equal responses in raster order).  The
    // reference resets the occupancy grid only in the bucketing branch (:79).
    void gradientMagnitudeWithSSC(std::shared_ptr<Frame>& frame, uint32_t detectionThreshold, uint32_t numberCandidate,
                                  bool useBucketing = true)
    {
        const auto& dev = frame->m_imagePyramid.device();
        std::vector<uint8_t> occ(m_occupancyGrid.begin(), m_occupancyGrid.end());
        std::vector<svo_feature_px> out(4096);
        int n = 0;
        dev->check(svo_select_ssc(dev->ctx(), frame->m_imagePyramid.slot(), detectionThreshold, (int)numberCandidate, m_cellSize,
                                  occ.data(), useBucketing ? 1 : 0, out.data(), (int)out.size(), &n, nullptr),
                   "svo_select_ssc");
        for (int i = 0; i < n; i++) {
            auto f = std::make_shared<Feature>(frame, Vec2(out[i].x, out[i].y), (double)out[i].magnitude, 0.0, 0);
            frame->addFeature(f);
        }
        m_imgGradientMagnitude = frame->m_imagePyramid.getBaseGradientImage();  // computeImageGradient, :250-267
        if (useBucketing) resetGridOccupancy();
    }
    void setExistingFeatures(const std::vector<std::shared_ptr<Feature>>& features)  // :269-275
    {
        for (const auto& f : features) setCellInGridOccupancy(f->m_pixelPosition);
    }
    void setCellInGridOccupancy(const Vec2& px)  // :277-281
    {
        m_occupancyGrid[(size_t)((int32_t)(px.y() / m_cellSize) * m_gridCols + (int32_t)(px.x() / m_cellSize))] = true;
    }

    Mat8 m_imgGradientMagnitude;
    int32_t m_cellSize;
    int32_t m_gridRows;
    int32_t m_gridCols;
    std::vector<bool> m_occupancyGrid;

private:
    void resetGridOccupancy() { std::fill(m_occupancyGrid.begin(), m_occupancyGrid.end(), false); }
};

// ------------------------------------------------------------------------------------------------------------
class ImageAlignment final
{
public:
    // include/image_alignment.hpp:18; System passes (5, 0, 3, 6), src/system.cpp:26-27
    explicit ImageAlignment(uint32_t patchSize, int32_t minLevel, int32_t maxLevel, uint32_t numParameters)
        : m_patchSize(patchSize), m_halfPatchSize(patchSize / 2), m_patchArea(patchSize * patchSize), m_minLevel(minLevel),
          m_maxLevel(maxLevel)
    {
        if (numParameters != 6) throw std::invalid_argument("ImageAlignment optimises an SE3 pose: numParameters must be 6");
    }
    ImageAlignment(const ImageAlignment&)            = delete;
    ImageAlignment& operator=(const ImageAlignment&) = delete;

    // src/image_alignment.cpp:25-67.  Mutates curFrame->m_absPose in place, returns the RMSE of the last level
    // (0 when refFrame has no features).  refFrame->m_lastKeyframe must be non-null (the reference dereferences it).
    double align(std::shared_ptr<Frame>& refFrame, std::shared_ptr<Frame>& curFrame)
    {
        if (refFrame->numberObservation() == 0) return 0;
        const auto& lastKF = refFrame->m_lastKeyframe;
        if (!lastKF) throw std::invalid_argument("ImageAlignment::align: refFrame->m_lastKeyframe is null");
        const auto& dev = curFrame->m_imagePyramid.device();
        m_feats.clear();
        pack(refFrame->m_features);
        pack(lastKF->m_features);
        svo_align_job job{};
        job.ref_slot    = refFrame->m_imagePyramid.slot();
        job.kf_slot     = lastKF->m_imagePyramid.slot();
        job.cur_slot    = curFrame->m_imagePyramid.slot();
        job.n_ref       = (int32_t)refFrame->numberObservation();
        job.n_kf        = (int32_t)lastKF->numberObservation();
        job.feat_offset = 0;
        refFrame->m_absPose.params(job.T_ref);
        lastKF->m_absPose.params(job.T_kf);
        curFrame->m_absPose.params(job.T_cur);
        svo_align_params prm{};
        prm.patch_size = (int32_t)m_patchSize;
        prm.min_level  = m_minLevel;
        prm.max_level  = m_maxLevel;
        prm.mode       = m_mode;
        prm.max_iter   = (int32_t)m_maxIteration;
        svo_align_result res{};
        m_levelStats.assign((size_t)(m_maxLevel - m_minLevel + 1), svo_align_level_stats{});
        dev->check(svo_sparse_align(dev->ctx(), &job, 1, m_feats.data(), (int)m_feats.size(), &prm, &res, m_levelStats.data()),
                   "svo_sparse_align");
        curFrame->m_absPose = SE3::fromParams(res.T_cur);
        m_status            = res.status;
        return res.rmse;
    }

    int32_t m_mode          = SVO_LM_FAITHFUL;  // what the reference executes (one damped step per level)
    uint32_t m_maxIteration = 20;               // Optimizer::m_maxIteration, src/optimizer.cpp:18
    int32_t m_status        = SVO_ST_SUCCESS;   // Optimizer::Status of the last level (dropped by the reference)
    std::vector<svo_align_level_stats> m_levelStats;

private:
    void pack(const std::vector<std::shared_ptr<Feature>>& features)
    {
        for (const auto& f : features) {
            svo_align_feature a{};
            a.px[0] = f->m_pixelPosition.x();
            a.px[1] = f->m_pixelPosition.y();
            for (int i = 0; i < 3; i++) a.bearing[i] = f->m_bearingVec[i];
            a.has_point = f->m_point != nullptr;
            if (f->m_point)
                for (int i = 0; i < 3; i++) a.point[i] = f->m_point->m_position[i];
            m_feats.push_back(a);
        }
    }
    uint32_t m_patchSize;
    int32_t m_halfPatchSize;
    int32_t m_patchArea;
    int32_t m_minLevel;
    int32_t m_maxLevel;
    std::vector<svo_align_feature> m_feats;
};

// ------------------------------------------------------------------------------------------------------------
class FeatureAlignment final
{
public:
    // include/feature_alignment.hpp:18; Map passes (7, 0, 3), src/map.cpp:18
    explicit FeatureAlignment(uint32_t patchSize, int32_t level, uint32_t numParameters) : m_patchSize(patchSize), m_level(level)
    {
        if (level != 0) throw std::invalid_argument("FeatureAlignment runs on gradient level 0 (src/feature_alignment.cpp:69,118)");
        if (numParameters != 3) throw std::invalid_argument("FeatureAlignment optimises (x, y, offset): numParameters must be 3");
    }
    FeatureAlignment(const FeatureAlignment&)            = delete;
    FeatureAlignment& operator=(const FeatureAlignment&) = delete;

    // src/feature_alignment.cpp:25-62.  pixelPos is in/out; returns the pre-step RMSE (NaN when the start is out of frame).
    double align(const std::shared_ptr<Feature>& refFeature, const std::shared_ptr<Frame>& curFrame, Vec2& pixelPos)
    {
        Item it{refFeature, curFrame, pixelPos};
        std::vector<double> err;
        std::vector<Vec2> px{pixelPos};
        alignBatch({it}, px, err);
        pixelPos = px[0];
        return err[0];
    }

    // The serial per-candidate loop of Map::reprojectCell / addCandidateToFrame (src/map.cpp:538,608) as ONE launch.
    struct Item {
        std::shared_ptr<Feature> refFeature;
        std::shared_ptr<Frame> curFrame;
        Vec2 pixelPos;
    };
    void alignBatch(const std::vector<Item>& items, std::vector<Vec2>& pixelPosOut, std::vector<double>& errorOut)
    {
        pixelPosOut.resize(items.size());
        errorOut.resize(items.size());
        if (items.empty()) return;
        const auto& dev = items[0].curFrame->m_imagePyramid.device();
        std::vector<svo_fa_item> in(items.size());
        for (size_t i = 0; i < items.size(); i++) {
            svo_fa_item a{};
            a.ref_slot  = items[i].refFeature->m_frame->m_imagePyramid.slot();
            a.cur_slot  = items[i].curFrame->m_imagePyramid.slot();
            a.ref_px[0] = items[i].refFeature->m_pixelPosition.x();
            a.ref_px[1] = items[i].refFeature->m_pixelPosition.y();
            a.px[0]     = items[i].pixelPos.x();
            a.px[1]     = items[i].pixelPos.y();
            a.A[0] = a.A[3] = 1.0;
            in[i]           = a;
        }
        svo_fa_params prm{};
        prm.patch_size = (int32_t)m_patchSize;
        prm.mode       = m_mode;
        prm.max_iter   = (int32_t)m_maxIteration;
        std::vector<svo_fa_result> res(items.size());
        dev->check(svo_feature_align(dev->ctx(), in.data(), (int)in.size(), &prm, res.data()), "svo_feature_align");
        for (size_t i = 0; i < items.size(); i++) {
            pixelPosOut[i] = Vec2(res[i].px[0], res[i].px[1]);
            errorOut[i]    = res[i].rmse;
        }
    }

    int32_t m_mode          = SVO_LM_FAITHFUL;
    uint32_t m_maxIteration = 20;

private:
    uint32_t m_patchSize;
    int32_t m_level;
};

// ------------------------------------------------------------------------------------------------------------
// The hot section of System::processNewFrame as ONE CUDA graph launch (svo_frontend_run): the new frame's pyramids
// (Frame::Frame, src/system.cpp:36), gradientMagnitudeByValue on it (src/system.cpp:252-253), ImageAlignment::align
// (src/system.cpp:313) and, with the aligned pose, the projection + FeatureAlignment::align of every tracked point
// (Map::addCandidateToFrame, src/map.cpp:595-610).  The reference makes these calls one after the other, each with its
// own host round trip; results are written where the reference's calls would put them.
class FrontEnd final
{
public:
    FrontEnd(uint32_t patchSize, int32_t minLevel, int32_t maxLevel, int32_t cellSize, uint32_t gradientThreshold,
             uint32_t featurePatchSize = 7, int32_t maxFeatures = 2048)
        : m_patchSize(patchSize), m_minLevel(minLevel), m_maxLevel(maxLevel), m_cellSize(cellSize), m_threshold(gradientThreshold),
          m_featurePatchSize(featurePatchSize), m_maxFeatures(maxFeatures)
    {
    }
    FrontEnd(const FrontEnd&)            = delete;
    FrontEnd& operator=(const FrontEnd&) = delete;

    struct NewFeature {  // what the loop of src/feature_selection.cpp:135-141 turns into a Feature
        Vec2 pixelPosition;
        double gradientMagnitude;
    };
    struct Result {
        double alignError   = 0;  // return value of ImageAlignment::align
        int32_t alignStatus = SVO_ST_SUCCESS;
        std::vector<NewFeature> newFeatures;  // cell raster order
        // one entry per tracked feature (refFrame's, then its last keyframe's)
        std::vector<bool> matched;            // had a point, reprojected into the frame and was aligned
        std::vector<Vec2> pixelPosition;      // FeatureAlignment::align's in/out pixelPos
        std::vector<double> matchError;       // its return value (the caller tests `< 50.0`, src/map.cpp:609)
    };

    // curFrame->m_absPose holds the prior on entry (src/system.cpp:309) and the aligned pose on return.  The graph
    // rebuilds curFrame's pyramid from its base image in the slot the Frame already owns.
    Result run(std::shared_ptr<Frame>& refFrame, std::shared_ptr<Frame>& curFrame, const std::vector<bool>* occupancy = nullptr)
    {
        const auto& lastKF = refFrame->m_lastKeyframe;
        if (!lastKF) throw std::invalid_argument("FrontEnd::run: refFrame->m_lastKeyframe is null");
        const auto& dev = curFrame->m_imagePyramid.device();
        std::vector<svo_align_feature> feats;
        auto pack = [&feats](const std::vector<std::shared_ptr<Feature>>& features) {
            for (const auto& f : features) {
                svo_align_feature a{};
                a.px[0] = f->m_pixelPosition.x();
                a.px[1] = f->m_pixelPosition.y();
                for (int i = 0; i < 3; i++) a.bearing[i] = f->m_bearingVec[i];
                a.has_point = f->m_point != nullptr;
                if (f->m_point)
                    for (int i = 0; i < 3; i++) a.point[i] = f->m_point->m_position[i];
                feats.push_back(a);
            }
        };
        pack(refFrame->m_features);
        pack(lastKF->m_features);
        svo_align_job job{};
        job.n_ref = (int32_t)refFrame->numberObservation();
        job.n_kf  = (int32_t)lastKF->numberObservation();
        refFrame->m_absPose.params(job.T_ref);
        lastKF->m_absPose.params(job.T_kf);
        curFrame->m_absPose.params(job.T_cur);
        svo_frontend_params prm{};
        prm.ref_slot     = refFrame->m_imagePyramid.slot();
        prm.kf_slot      = lastKF->m_imagePyramid.slot();
        prm.cur_slot     = curFrame->m_imagePyramid.slot();
        prm.cell         = m_cellSize;
        prm.thr          = m_threshold;
        prm.max_features = m_maxFeatures;
        prm.align        = svo_align_params{(int32_t)m_patchSize, m_minLevel, m_maxLevel, m_mode, (int32_t)m_maxIteration, 0};
        prm.fa           = svo_fa_params{(int32_t)m_featurePatchSize, m_featureMode, (int32_t)m_maxIteration, 0};
        const Mat8& img  = curFrame->m_imagePyramid.getBaseImage();
        const int rows = img.rows / m_cellSize + 1, cols = img.cols / m_cellSize + 1;
        std::vector<uint8_t> occ;
        if (occupancy) occ.assign(occupancy->begin(), occupancy->end());
        std::vector<svo_feature_px> sel((size_t)rows * cols);
        std::vector<svo_fa_result> refined(feats.size() ? feats.size() : 1);
        svo_frontend_result out{};
        dev->check(svo_frontend_run(dev->ctx(), &prm, img.ptr(), img.cols, &job, feats.data(), (int)feats.size(),
                                    occupancy ? occ.data() : nullptr, &out, sel.data(), (int)sel.size(), refined.data()),
                   "svo_frontend_run");
        Result r;
        r.alignError        = out.align.rmse;
        r.alignStatus       = out.align.status;
        curFrame->m_absPose = SE3::fromParams(out.align.T_cur);
        for (int i = 0; i < out.n_selected; i++)
            r.newFeatures.push_back({Vec2((double)sel[i].x, (double)sel[i].y), (double)sel[i].magnitude});
        for (size_t i = 0; i < feats.size(); i++) {
            r.matched.push_back(!(refined[i].status == SVO_ST_FAILED && refined[i].iterations == 0));
            r.pixelPosition.push_back(Vec2(refined[i].px[0], refined[i].px[1]));
            r.matchError.push_back(refined[i].rmse);
        }
        return r;
    }

    int32_t m_mode          = SVO_LM_FAITHFUL;
    int32_t m_featureMode   = SVO_LM_FAITHFUL;
    uint32_t m_maxIteration = 20;

private:
    uint32_t m_patchSize;
    int32_t m_minLevel, m_maxLevel, m_cellSize;
    uint32_t m_threshold, m_featurePatchSize;
    int32_t m_maxFeatures;
};

// ------------------------------------------------------------------------------------------------------------
// algorithm::matchEpipolarConstraint (src/algorithm.cpp:412-551), the depth filter's epipolar search
// (DepthEstimator::updateFilters, src/depth_estimator.cpp:245): same signature for one seed, plus the batched form
// that replaces the serial loop over the depth filters of a frame.
namespace algorithm
{
struct EpipolarSeed {
    std::shared_ptr<Feature> refFeature;  // depthFilter.m_feature (its frame is the reference frame)
    double initialDepth, minDepth, maxDepth;
};

inline void matchEpipolarConstraintBatch(const std::shared_ptr<Frame>& curFrame, const std::vector<EpipolarSeed>& seeds,
                                         uint32_t patchSize, std::vector<bool>& found, std::vector<double>& estimatedDepth,
                                         int32_t meanMode = SVO_MEAN_EIGEN_U8)
{
    found.assign(seeds.size(), false);
    estimatedDepth.assign(seeds.size(), 0.0);
    if (seeds.empty()) return;
    const auto& dev = curFrame->m_imagePyramid.device();
    std::vector<svo_epi_item> items(seeds.size());
    for (size_t i = 0; i < seeds.size(); i++) {
        const auto& f        = seeds[i].refFeature;
        const auto& refFrame = f->m_frame;
        svo_epi_item it{};
        it.ref_slot = refFrame->m_imagePyramid.slot();
        it.cur_slot = curFrame->m_imagePyramid.slot();
        (curFrame->m_absPose * refFrame->m_absPose.inverse()).params(it.T_rel);  // computeRelativePose, :705-709
        it.px[0] = f->m_pixelPosition.x();
        it.px[1] = f->m_pixelPosition.y();
        for (int k = 0; k < 3; k++) it.bearing[k] = f->m_bearingVec[k];
        it.depth     = seeds[i].initialDepth;
        it.min_depth = seeds[i].minDepth;
        it.max_depth = seeds[i].maxDepth;
        items[i]     = it;
    }
    svo_epi_params prm{(int32_t)patchSize, meanMode};
    std::vector<svo_epi_result> res(seeds.size());
    dev->check(svo_epipolar_match(dev->ctx(), items.data(), (int)items.size(), &prm, res.data()), "svo_epipolar_match");
    for (size_t i = 0; i < seeds.size(); i++) {
        found[i]          = res[i].found != 0;
        estimatedDepth[i] = res[i].depth;
    }
}

inline bool matchEpipolarConstraint(const std::shared_ptr<Frame>& refFrame, const std::shared_ptr<Frame>& curFrame,
                                    std::shared_ptr<Feature>& refFeature, const uint32_t patchSize, const double initialDepth,
                                    const double minDepth, const double maxDepth, double& estimatedDepth)
{
    if (refFeature->m_frame != refFrame) throw std::invalid_argument("matchEpipolarConstraint: refFeature does not belong to refFrame");
    std::vector<bool> found;
    std::vector<double> depth;
    matchEpipolarConstraintBatch(curFrame, {EpipolarSeed{refFeature, initialDepth, minDepth, maxDepth}}, patchSize, found, depth);
    if (found[0]) estimatedDepth = depth[0];  // the reference leaves it untouched on failure
    return found[0];
}
// algorithm::computeMedian (src/algorithm.cpp:813-832); the mean of the two middle elements for an even count is taken
// from true order statistics (the reference reads vec[mid - 1] after one nth_element: SURVEY 9.3)
inline double computeMedian(std::vector<double> vec)
{
    if (vec.empty()) return std::numeric_limits<double>::quiet_NaN();
    const size_t mid = vec.size() / 2;
    std::nth_element(vec.begin(), vec.begin() + mid, vec.end());
    if (vec.size() % 2) return vec[mid];
    return (*std::max_element(vec.begin(), vec.begin() + mid) + vec[mid]) / 2.0;
}

// algorithm::computeOpticalFlowSparse (src/algorithm.cpp:29-107), System's initialisation step (src/system.cpp:129): every
// feature of refFrame is tracked into curFrame with cv::calcOpticalFlowPyrLK(..., Size(patchSize, patchSize), 3,
// TermCriteria(COUNT + EPS, 30, 1e-4), OPTFLOW_USE_INITIAL_FLOW) -- here svo_klt_track on the two frames' pyramids -- the
// tracked positions become features of curFrame (:66-76, before the disparity test, as in the reference), the median
// disparity gates the result (:78-84) and the features that were not tracked leave refFrame (:86-102).
inline bool computeOpticalFlowSparse(std::shared_ptr<Frame>& refFrame, std::shared_ptr<Frame>& curFrame, const uint32_t patchSize,
                                     const double disparityThreshold)
{
    const auto& dev = curFrame->m_imagePyramid.device();
    const size_t n  = refFrame->numberObservation();
    std::vector<float> refPoints(2 * n), curPoints(2 * n);
    for (size_t i = 0; i < n; i++) {
        refPoints[2 * i] = curPoints[2 * i] = (float)refFrame->m_features[i]->m_pixelPosition.x();
        refPoints[2 * i + 1] = curPoints[2 * i + 1] = (float)refFrame->m_features[i]->m_pixelPosition.y();
    }
    std::vector<uint8_t> status(n ? n : 1);
    svo_klt_params prm{(int32_t)patchSize, 3, 30, 1, 1e-4, 1e-4};
    dev->check(svo_klt_track(dev->ctx(), refFrame->m_imagePyramid.slot(), curFrame->m_imagePyramid.slot(), refPoints.data(),
                             curPoints.data(), (int)n, &prm, status.data(), nullptr),
               "svo_klt_track");
    std::vector<double> disparity;
    for (size_t i = 0; i < n; i++) {
        if (!status[i]) continue;
        auto f = std::make_shared<Feature>(curFrame, Vec2(curPoints[2 * i], curPoints[2 * i + 1]), 0.0, 0.0, 0);
        curFrame->addFeature(f);
        disparity.push_back(std::hypot((double)refPoints[2 * i] - curPoints[2 * i], (double)refPoints[2 * i + 1] - curPoints[2 * i + 1]));
    }
    const double medianDisparity = computeMedian(disparity);
    if (medianDisparity < disparityThreshold) return false;
    size_t cnt = 0;
    auto& fs   = refFrame->m_features;
    fs.erase(std::remove_if(fs.begin(), fs.end(), [&](const auto& f) { return !(f != nullptr && status[cnt++] == 1); }), fs.end());
    return true;
}
}  // namespace algorithm

// ------------------------------------------------------------------------------------------------------------
// Map::reprojectMap (src/map.cpp:260-489) as one batched pass: reprojectPoint for every candidate, reprojectCell's choice
// per grid cell, one FeatureAlignment launch.  The caller (the reference's Map) keeps its Point bookkeeping: it lists
// the candidates in insertion order (features with a point of refFrame, then of its last keyframe, :462-476) and applies
// the matches (:551-576: new Feature at `pixelPosition`, m_succeededProjection, type promotion).
struct ReprojectionCandidate {
    std::shared_ptr<Feature> feature;  // candidate.m_feature (its m_point is candidate.m_point)
    uint32_t pointType;                // Point::PointType: 0 GOOD, 1 DELETED, 2 CANDIDATE, 3 UNKNOWN
};
struct ReprojectionMatch {
    int32_t cell;
    size_t candidate;     // index into the candidate list
    Vec2 pixelPosition;   // where the new Feature goes (:564)
    double error;         // FeatureAlignment::align's return value (ignored by the reference)
};
inline std::vector<ReprojectionMatch> reprojectMap(const std::shared_ptr<Frame>& curFrame, const std::vector<ReprojectionCandidate>& candidates,
                                                   uint32_t cellSize, const std::vector<int32_t>& cellOrders, uint32_t maxMatches = 150,
                                                   std::vector<bool>* projected = nullptr)
{
    const auto& dev = curFrame->m_imagePyramid.device();
    std::vector<svo_reproj_candidate> in(candidates.size());
    for (size_t i = 0; i < candidates.size(); i++) {
        const auto& f = candidates[i].feature;
        if (!f->m_point) throw std::invalid_argument("reprojectMap: candidate feature without a 3D point");
        svo_reproj_candidate c{};
        c.ref_slot  = f->m_frame->m_imagePyramid.slot();
        c.type      = (int32_t)candidates[i].pointType;
        c.ref_px[0] = f->m_pixelPosition.x();
        c.ref_px[1] = f->m_pixelPosition.y();
        for (int k = 0; k < 3; k++) c.point[k] = f->m_point->m_position[k];
        in[i] = c;
    }
    double T[7];
    curFrame->m_absPose.params(T);
    svo_fa_params fa{7, SVO_LM_FAITHFUL, 20, 0};  // Map::m_alignment = FeatureAlignment(7, 0, 3), src/map.cpp:18
    std::vector<svo_reproj_match> out(maxMatches + 1);
    std::vector<uint8_t> proj(candidates.size() ? candidates.size() : 1);
    int n = 0;
    dev->check(svo_reproject_map(dev->ctx(), curFrame->m_imagePyramid.slot(), T, in.data(), (int)in.size(), (int)cellSize,
                                 cellOrders.data(), (int)cellOrders.size(), (int)maxMatches, &fa, out.data(), &n, proj.data()),
               "svo_reproject_map");
    if (projected) projected->assign(proj.begin(), proj.begin() + candidates.size());
    std::vector<ReprojectionMatch> res;
    for (int i = 0; i < n; i++) res.push_back({out[i].cell, (size_t)out[i].candidate, Vec2(out[i].px[0], out[i].px[1]), out[i].rmse});
    return res;
}

}  // namespace svo
