// Process-wide handle on the CUDA engine (libem2b200, include/em2b200.h).  The host classes call the
// C-ABI through this and turn a non-zero status into std::runtime_error, like the reference turns
// cl::Error into runtime_error (reference src/LshGpu.cpp:38-43,200-204).  There is no CPU fallback.
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/em2b200.h"

namespace ChanZuckerberg {
namespace ExpressionMatrix2 {

class Gpu {
public:
    static Gpu& instance();              // device = $EM2_DEVICE or 0; throws without an sm_100 GPU
    em2_context* context() { return ctx_; }
    std::string name();
    em2_stats stats();
    void check(int status, const char* what);
    ~Gpu();

private:
    Gpu();
    em2_context* ctx_ = nullptr;
};

// Every GPU of the box behind one blocking call (em2_multi): what findSimilarPairs4 runs on.  The devices are
// $EM2_DEVICES (comma separated indices) or, by default, all visible ones; with a single device it is the same
// code path minus the collectives.
class GpuSet {
public:
    static GpuSet& instance();
    em2_multi* handle() { return multi_; }
    int deviceCount() const;
    std::string name();                  // "8 x NVIDIA B200"
    em2_stats stats();                   // of the whole job (times: max over the devices)
    void check(int status, const char* what);
    ~GpuSet();

private:
    GpuSet();
    em2_multi* multi_ = nullptr;
};

}  // namespace ExpressionMatrix2
}  // namespace ChanZuckerberg
