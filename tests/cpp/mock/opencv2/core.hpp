// Minimal stand-in for <opencv2/core.hpp>: the cv::Mat members the bindings touch (8-bit single channel only).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8UC1 0
namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    std::size_t step = 0;
    uint8_t* data    = nullptr;
    std::shared_ptr<std::vector<uint8_t>> own;
    Mat() = default;
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), step((std::size_t)c), own(std::make_shared<std::vector<uint8_t>>((std::size_t)r * c)) { data = own->data(); }
    Mat(int r, int c, int /*type*/, void* d, std::size_t s = 0) : rows(r), cols(c), step(s ? s : (std::size_t)c), data((uint8_t*)d) {}
    int type() const { return CV_8UC1; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (std::size_t)cols; }
    template <class T>
    T* ptr(int r = 0) { return (T*)(data + (std::size_t)r * step); }
    template <class T>
    const T* ptr(int r = 0) const { return (const T*)(data + (std::size_t)r * step); }
    Mat clone() const
    {
        Mat m(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (std::size_t)r * m.step, data + (std::size_t)r * step, (std::size_t)cols);
        return m;
    }
};
}  // namespace cv
