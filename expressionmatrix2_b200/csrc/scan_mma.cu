// tcgen05 int8 variant of the Hamming scan -- placeholder until the kernel lands (fails loudly).
#include "common.cuh"
namespace em2 {
int launchScanMma(em2_context* ctx, const uint64_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, int64_t,
                  const float*, em2_pair*, uint32_t*, cudaStream_t)
{
    return fail(ctx, EM2_ERR_INVALID, "EM2_VARIANT_MMA_I8 is not available in this build");
}
int launchMismatchBlockMma(em2_context* ctx, const uint64_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint16_t*,
                           cudaStream_t)
{
    return fail(ctx, EM2_ERR_INVALID, "EM2_VARIANT_MMA_I8 is not available in this build");
}
}  // namespace em2
