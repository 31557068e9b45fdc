#include "Gpu.hpp"

#include <cstdlib>

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
