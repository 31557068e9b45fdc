// Exact (Pearson) all-pairs path -- placeholder until the kernel lands (fails loudly).
#include "common.cuh"
namespace em2 {
int launchExact(em2_context* ctx, uint64_t, uint64_t, const uint64_t*, const em2_count*, const double*,
                const double*, uint64_t, double, em2_pair*, uint32_t*, cudaStream_t)
{
    return fail(ctx, EM2_ERR_INVALID, "the exact path is not available in this build");
}
}  // namespace em2
