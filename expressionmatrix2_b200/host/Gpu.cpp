#include "Gpu.hpp"

#include <cstdlib>
#include <sstream>
#include <vector>

using namespace ChanZuckerberg::ExpressionMatrix2;

Gpu::Gpu()
{
    const char* e = std::getenv("EM2_DEVICE");
    const int device = e ? std::atoi(e) : 0;
    const int rc = em2_create(device, &ctx_);
    if (rc != EM2_OK) throw std::runtime_error(std::string("GPU initialization failed: ") + em2_last_error(nullptr));
}

Gpu::~Gpu()
{
    if (ctx_) em2_destroy(ctx_);
}

Gpu& Gpu::instance()
{
    static Gpu gpu;
    return gpu;
}

std::string Gpu::name()
{
    char buf[256];
    check(em2_device_name(ctx_, buf, sizeof(buf)), "em2_device_name");
    return buf;
}

em2_stats Gpu::stats()
{
    em2_stats s;
    check(em2_get_stats(ctx_, &s), "em2_get_stats");
    return s;
}

void Gpu::check(int status, const char* what)
{
    if (status != EM2_OK)
        throw std::runtime_error(std::string("GPU error ") + std::to_string(status) + " from " + what + ": " +
                                 em2_last_error(ctx_));
}

GpuSet::GpuSet()
{
    std::vector<int> devices;
    if (const char* e = std::getenv("EM2_DEVICES")) {
        std::stringstream ss(e);
        std::string item;
        while (std::getline(ss, item, ','))
            if (!item.empty()) devices.push_back(std::atoi(item.c_str()));
    } else if (const char* e1 = std::getenv("EM2_DEVICE")) {
        devices.push_back(std::atoi(e1));
    }
    const int rc = em2_multi_create(devices.empty() ? nullptr : devices.data(), int(devices.size()), &multi_);
    if (rc != EM2_OK) throw std::runtime_error(std::string("GPU initialization failed: ") + em2_multi_last_error(nullptr));
}

GpuSet::~GpuSet()
{
    if (multi_) em2_multi_destroy(multi_);
}

GpuSet& GpuSet::instance()
{
    static GpuSet gpus;
    return gpus;
}

int GpuSet::deviceCount() const { return em2_multi_device_count(multi_); }

std::string GpuSet::name()
{
    char buf[256];
    if (em2_device_name(em2_multi_context(multi_, 0), buf, sizeof(buf)) != EM2_OK) return "GPU";
    const int n = deviceCount();
    return n > 1 ? std::to_string(n) + " x " + buf : std::string(buf);
}

em2_stats GpuSet::stats()
{
    em2_stats s;
    check(em2_multi_get_stats(multi_, -1, &s), "em2_multi_get_stats");
    return s;
}

void GpuSet::check(int status, const char* what)
{
    if (status != EM2_OK)
        throw std::runtime_error(std::string("GPU error ") + std::to_string(status) + " from " + what + ": " +
                                 em2_multi_last_error(multi_));
}
