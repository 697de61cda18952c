// svo_device.hpp -- the one device context a System owns: wraps svo_ctx (include/svo_b200.h) and hands out frame
// slots of its pyramid arena.  The reference constructs its hot-path objects in System::System
// (src/system.cpp:25-31); a port constructs ONE svo::Device next to them.
#pragma once
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/svo_b200.h"

namespace svo {

class Device
{
public:
    // width/height/K: the PinholeCamera; levels: max_level_image_pyramid + 1 (src/system.cpp:36)
    Device(int width, int height, const double K[4], int levels = 4, int maxFrames = 16, int maxFeatures = 2048,
           int maxFaItems = 4096, int maxJobs = 1, int device = 0)
    {
        svo_config cfg{};
        cfg.device       = device;
        cfg.width        = width;
        cfg.height       = height;
        cfg.levels       = levels;
        cfg.max_frames   = maxFrames;
        cfg.max_jobs     = maxJobs;
        cfg.max_features = maxFeatures;
        cfg.max_fa_items = maxFaItems;
        for (int i = 0; i < 4; i++) cfg.K[i] = K[i];
        const svo_status st = svo_create(&cfg, &m_ctx);
        if (st != SVO_OK) throw std::runtime_error(std::string("svo_create: ") + svo_last_error(nullptr));
        for (int s = maxFrames - 1; s >= 0; s--) m_free.push_back(s);
        m_levels = levels;
    }
    ~Device() { svo_destroy(m_ctx); }
    Device(const Device&)            = delete;
    Device& operator=(const Device&) = delete;

    svo_ctx* ctx() const { return m_ctx; }
    int levels() const { return m_levels; }
    void check(svo_status st, const char* what) const
    {
        if (st != SVO_OK) throw std::runtime_error(std::string(what) + ": " + svo_last_error(m_ctx));
    }
    int acquireSlot()
    {
        std::lock_guard<std::mutex> g(m_mutex);
        if (m_free.empty()) throw std::runtime_error("svo::Device: out of frame slots (raise maxFrames)");
        const int s = m_free.back();
        m_free.pop_back();
        return s;
    }
    void releaseSlot(int s)
    {
        std::lock_guard<std::mutex> g(m_mutex);
        m_free.push_back(s);
    }

    // the process-wide default used by classes constructed without an explicit device (mirrors how the reference's
    // classes take no context argument)
    static std::shared_ptr<Device>& current()
    {
        static std::shared_ptr<Device> d;
        return d;
    }

private:
    svo_ctx* m_ctx = nullptr;
    int m_levels   = 0;
    std::mutex m_mutex;
    std::vector<int> m_free;
};

}  // namespace svo
