// ExpressionMatrixSubset construction on the device.
//
// Replaces the constructor loop of reference src/ExpressionMatrixSubset.cpp:9-42: for every cell of the (sorted)
// cell set, keep the stored counts whose gene belongs to the gene set, in stored order, with the global gene id
// replaced by GeneSet::getLocalGeneId (src/GeneSet.hpp:70-77).  The reference does this with one mmap `append`
// per element (and an msync + ftruncate + re-mmap on every capacity overflow); here the selected rows cross
// PCIe once, a count pass + exclusive scan + fill pass (warp per cell, ballot compaction keeps stored order)
// builds the local CSR in HBM, and the signature / scan stages consume it where it lies.
#include "common.cuh"

#include <cub/device/device_scan.cuh>

#include <algorithm>

namespace em2 {

namespace {

constexpr uint32_t kInvalidGene = 0xffffffffu;

__global__ void __launch_bounds__(256)
subsetCountKernel(uint64_t cellCount, const uint64_t* __restrict__ srcToc, const em2_count* __restrict__ src,
                  const uint32_t* __restrict__ geneLocalId, uint64_t globalGeneCount, uint64_t* __restrict__ kept)
{
    const uint64_t c = blockIdx.x * uint64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c > cellCount) return;
    if (c == cellCount) {          // one extra slot so that the exclusive scan also yields the total
        if (lane == 0) kept[c] = 0;
        return;
    }
    uint32_t n = 0;
    const uint64_t end = srcToc[c + 1];
    for (uint64_t e = srcToc[c] + lane; e < end; e += 32) {
        const uint32_t g = src[e].gene;
        n += (g < globalGeneCount && geneLocalId[g] != kInvalidGene);
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) kept[c] = n;
}

__global__ void __launch_bounds__(256)
subsetFillKernel(uint64_t cellCount, const uint64_t* __restrict__ srcToc, const em2_count* __restrict__ src,
                 const uint32_t* __restrict__ geneLocalId, uint64_t globalGeneCount, const uint64_t* __restrict__ dstToc,
                 em2_count* __restrict__ dst)
{
    const uint64_t c = blockIdx.x * uint64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (c >= cellCount) return;
    const uint64_t begin = srcToc[c], end = srcToc[c + 1];
    uint64_t out = dstToc[c];
    const uint32_t lt = (1u << lane) - 1u;
    for (uint64_t base = begin; base < end; base += 32) {
        const uint64_t e = base + lane;
        em2_count p;
        p.gene = kInvalidGene;
        p.count = 0.f;
        uint32_t local = kInvalidGene;
        if (e < end) {
            p = src[e];
            if (p.gene < globalGeneCount) local = geneLocalId[p.gene];
        }
        const bool keep = local != kInvalidGene;
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            em2_count q;
            q.gene = local;
            q.count = p.count;
            dst[out + __popc(mask & lt)] = q;       // stored order is preserved
        }
        out += __popc(mask);
    }
}

}  // namespace

// srcToc: uint64[cellCount + 1] offsets of the selected cells' rows inside `src`; dstToc: uint64[cellCount + 1] out;
// dst: capacity >= srcToc[cellCount] entries.  All device pointers; enqueued on `s`.
int launchSubset(em2_context* ctx, uint64_t cellCount, const uint64_t* srcToc, const em2_count* src,
                 const uint32_t* geneLocalId, uint64_t globalGeneCount, uint64_t* dstToc, em2_count* dst, cudaStream_t s)
{
    void* kept = nullptr;
    size_t scanBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, static_cast<uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                  int(cellCount + 1), s);
    const size_t keptBytes = roundUp((cellCount + 1) * sizeof(uint64_t), 256);
    EM2_TRY(reserve(ctx, em2_context::S_ROWPERM, keptBytes + scanBytes, &kept));
    const unsigned blocks = unsigned((cellCount + 1 + 7) / 8);
    subsetCountKernel<<<blocks, 256, 0, s>>>(cellCount, srcToc, src, geneLocalId, globalGeneCount, static_cast<uint64_t*>(kept));
    EM2_CUDA(ctx, cudaGetLastError());
    EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(static_cast<uint8_t*>(kept) + keptBytes, scanBytes, static_cast<uint64_t*>(kept),
                                                dstToc, int(cellCount + 1), s));
    subsetFillKernel<<<blocks, 256, 0, s>>>(cellCount, srcToc, src, geneLocalId, globalGeneCount, dstToc, dst);
    EM2_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 2;
    return EM2_OK;
}

}  // namespace em2
