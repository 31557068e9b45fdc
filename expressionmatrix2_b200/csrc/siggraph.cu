// SignatureGraph construction on the device.
//
// Replaces the grouping and the edge loop of ExpressionMatrix::createSignatureGraph (reference
// src/ExpressionMatrixSignatureGraph.cpp:69-125) and SignatureGraph::createEdges (src/SignatureGraph.cpp:23-48):
//   * cells with identical signatures form one vertex.  The reference collects them in a
//     std::map<BitSetPointer, vector<CellId>>, whose order is the lexicographic order of the 64-bit words
//     (src/BitSet.hpp:157-160; bit 0 is the most significant bit of word 0, :57-63) with the cells of a vertex in
//     ascending id -- here a stable LSD radix sort of the cell ids by word W-1, ..., word 0;
//   * signatures with fewer than minCellCount cells get no vertex; vertices are numbered in map order;
//   * for every vertex, for every ZERO bit of its signature in bit order, the signature with that bit set is looked
//     up (std::map::find there, a binary search over the sorted vertex signatures here) and, if it is a vertex, the
//     edge (vertex, found vertex) is added -- each undirected edge exactly once, from its lower-signature end, in
//     (vertex, bit) order.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace em2 {

namespace {

__global__ void iotaU32Kernel(uint64_t n, uint32_t* __restrict__ out)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = uint32_t(i);
}

// keys[i] = word w of the signature of cell order[i]
__global__ void gatherWordKernel(uint64_t n, const uint64_t* __restrict__ sig, uint32_t W, uint32_t w,
                                 const uint32_t* __restrict__ order, unsigned long long* __restrict__ keys)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n) keys[i] = sig[uint64_t(order[i]) * W + w];
}

__device__ __forceinline__ bool sameSignature(const uint64_t* __restrict__ sig, uint32_t W, uint32_t a, uint32_t b)
{
    for (uint32_t w = 0; w < W; w++)
        if (sig[uint64_t(a) * W + w] != sig[uint64_t(b) * W + w]) return false;
    return true;
}

// head[i] = 1 where a new signature starts in the sorted order (element n is a sentinel head)
__global__ void headFlagsKernel(uint64_t n, const uint64_t* __restrict__ sig, uint32_t W, const uint32_t* __restrict__ order,
                                uint32_t* __restrict__ head)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i > n) return;
    head[i] = (i == 0 || i == n || !sameSignature(sig, W, order[i], order[i - 1])) ? 1u : 0u;
}

// groupStart[g] = sorted index of the first cell of group g (groupOf = exclusive scan of head, so a head at i opens
// group groupOf[i]); the sentinel writes groupStart[groupCount] = n.
__global__ void groupStartsKernel(uint64_t n, const uint32_t* __restrict__ head, const uint32_t* __restrict__ groupOf,
                                  uint32_t* __restrict__ groupStart)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i <= n && head[i]) groupStart[groupOf[i]] = uint32_t(i);
}

// keptCells[g] = size of group g if it becomes a vertex, else 0; isVertex[g] likewise 1 / 0 (one extra zero entry each)
__global__ void keepGroupsKernel(uint32_t groups, const uint32_t* __restrict__ groupStart, uint32_t minCellCount,
                                 uint32_t* __restrict__ keptCells, uint32_t* __restrict__ isVertex)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > groups) return;
    uint32_t size = 0;
    if (g < groups) size = groupStart[g + 1] - groupStart[g];
    const bool keep = g < groups && size >= minCellCount;
    keptCells[g] = keep ? size : 0;
    isVertex[g] = keep ? 1u : 0u;
}

// Per kept group: its offset into the compacted cell list, its representative cell, and the compacted cells.
__global__ void emitVerticesKernel(uint32_t groups, const uint32_t* __restrict__ groupStart, const uint32_t* __restrict__ isVertex,
                                   const uint32_t* __restrict__ vertexOf, const uint32_t* __restrict__ cellOffset,
                                   const uint32_t* __restrict__ order, uint64_t* __restrict__ vertexOffsets,
                                   uint32_t* __restrict__ vertexCell, uint32_t* __restrict__ cellOrder, uint32_t vertexCount,
                                   uint32_t keptTotal)
{
    const uint32_t g = blockIdx.x;          // one block per group: groups are few compared with cells when this matters
    if (g == groups) {
        if (threadIdx.x == 0) vertexOffsets[vertexCount] = keptTotal;
        return;
    }
    if (!isVertex[g]) return;
    const uint32_t v = vertexOf[g], begin = groupStart[g], size = groupStart[g + 1] - begin, out = cellOffset[g];
    if (threadIdx.x == 0) {
        vertexOffsets[v] = out;
        vertexCell[v] = order[begin];
    }
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) cellOrder[out + i] = order[begin + i];
}

// lexicographic compare of the signature of cell a, with bit `flip` (MSB-first numbering) set, against cell b's
__device__ __forceinline__ int compareFlipped(const uint64_t* __restrict__ sig, uint32_t W, uint32_t a, uint32_t flip, uint32_t b)
{
    for (uint32_t w = 0; w < W; w++) {
        uint64_t x = sig[uint64_t(a) * W + w];
        if (w == (flip >> 6)) x |= 1ull << (63 - (flip & 63));
        const uint64_t y = sig[uint64_t(b) * W + w];
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

// One thread per vertex.  edges == nullptr: count the vertex's edges; else write them at edgeOffset[v] in bit order.
__global__ void signatureEdgesKernel(uint32_t vertexCount, uint32_t lshCount, const uint64_t* __restrict__ sig, uint32_t W,
                                     const uint32_t* __restrict__ vertexCell, uint64_t* __restrict__ counts,
                                     const uint64_t* __restrict__ edgeOffset, em2_signature_edge* __restrict__ edges,
                                     uint64_t capacity)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > vertexCount) return;
    if (v == vertexCount) {
        if (counts) counts[v] = 0;
        return;
    }
    const uint32_t cell = vertexCell[v];
    uint64_t n = 0, out = edgeOffset ? edgeOffset[v] : 0;
    for (uint32_t bit = 0; bit < lshCount; bit++) {
        if ((sig[uint64_t(cell) * W + (bit >> 6)] >> (63 - (bit & 63))) & 1ull) continue;
        // the flipped signature is larger than this vertex's: search the vertices after v
        uint32_t lo = v + 1, hi = vertexCount;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (compareFlipped(sig, W, cell, bit, vertexCell[mid]) > 0) lo = mid + 1;
            else hi = mid;
        }
        if (lo < vertexCount && compareFlipped(sig, W, cell, bit, vertexCell[lo]) == 0) {
            if (edges && out < capacity) {
                em2_signature_edge e;
                e.vertex0 = v;
                e.vertex1 = lo;
                edges[out] = e;
            }
            out++;
            n++;
        }
    }
    if (counts) counts[v] = n;
}

}  // namespace

int launchSignatureGraph(em2_context* ctx, const uint64_t* sig, uint64_t cellCount, uint64_t lshCount, uint64_t minCellCount,
                         uint32_t* cellOrder, uint64_t* vertexOffsets, uint64_t vertexCapacity, uint64_t* vertexCountHost,
                         uint64_t* keptCellsHost, em2_signature_edge* edges, uint64_t edgeCapacity, uint64_t* edgeCountHost,
                         cudaStream_t s)
{
    const uint64_t n = cellCount;
    const uint32_t W = uint32_t(wordCount(lshCount));
    if (n > 0x7fffff00ull) return fail(ctx, EM2_ERR_INVALID, "em2_signature_graph: too many cells");
    const unsigned blocks = unsigned((n + 1 + 255) / 256);

    // scratch: order[2][n], keys[2][n], head[n+1], groupOf[n+1], groupStart[n+1], keptCells[n+1], isVertex[n+1],
    //          vertexOf[n+1], cellOffset[n+1], vertexCell[n], counts/offsets u64 [n+1] x 2, cub temp
    size_t cubSort = 0, cubScan32 = 0, cubScan64 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cubSort, static_cast<const unsigned long long*>(nullptr),
                                    static_cast<unsigned long long*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                    static_cast<uint32_t*>(nullptr), int(n), 0, 64, s);
    cub::DeviceScan::ExclusiveSum(nullptr, cubScan32, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  int(n + 1), s);
    cub::DeviceScan::ExclusiveSum(nullptr, cubScan64, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                  int(n + 1), s);
    const size_t cubBytes = roundUp(std::max(cubSort, std::max(cubScan32, cubScan64)), 256);
    const size_t n1 = roundUp(n + 1, 64);
    void* scratch = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_MISC, n1 * (10 * sizeof(uint32_t) + 4 * sizeof(uint64_t)) + cubBytes, &scratch));
    uint8_t* base = static_cast<uint8_t*>(scratch);
    auto take = [&](size_t bytes) {
        uint8_t* p = base;
        base += roundUp(bytes, 256);
        return p;
    };
    auto* keysA = reinterpret_cast<unsigned long long*>(take(n1 * 8));
    auto* keysB = reinterpret_cast<unsigned long long*>(take(n1 * 8));
    auto* counts = reinterpret_cast<uint64_t*>(take(n1 * 8));
    auto* offsets = reinterpret_cast<uint64_t*>(take(n1 * 8));
    auto* orderA = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* orderB = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* head = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* groupOf = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* groupStart = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* keptCells = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* isVertex = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* vertexOf = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* cellOffset = reinterpret_cast<uint32_t*>(take(n1 * 4));
    auto* vertexCell = reinterpret_cast<uint32_t*>(take(n1 * 4));
    void* cubTemp = base;

    // 1. cells sorted by signature (stable: ascending id inside a signature)
    iotaU32Kernel<<<blocks, 256, 0, s>>>(n, orderA);
    uint32_t* order = orderA;
    uint32_t* other = orderB;
    for (uint32_t w = W; w-- > 0;) {
        gatherWordKernel<<<blocks, 256, 0, s>>>(n, sig, W, w, order, keysA);
        size_t bytes = cubSort;
        EM2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(cubTemp, bytes, keysA, keysB, order, other, int(n), 0, 64, s));
        std::swap(order, other);
    }
    ctx->stats.kernel_launches += 1 + W;
    // 2. groups
    headFlagsKernel<<<blocks, 256, 0, s>>>(n, sig, W, order, head);
    {
        size_t bytes = cubScan32;
        EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cubTemp, bytes, head, groupOf, int(n + 1), s));
    }
    groupStartsKernel<<<blocks, 256, 0, s>>>(n, head, groupOf, groupStart);
    uint32_t groups = 0;
    EM2_CUDA(ctx, cudaMemcpyAsync(&groups, groupOf + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaStreamSynchronize(s));          // groupOf[n] = number of heads before the sentinel
    // 3. vertices
    const uint32_t minCells = uint32_t(std::min<uint64_t>(minCellCount, 0xffffffffu));
    keepGroupsKernel<<<(groups + 1 + 255) / 256, 256, 0, s>>>(groups, groupStart, minCells, keptCells, isVertex);
    {
        size_t bytes = cubScan32;
        EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cubTemp, bytes, isVertex, vertexOf, int(groups + 1), s));
        bytes = cubScan32;
        EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cubTemp, bytes, keptCells, cellOffset, int(groups + 1), s));
    }
    uint32_t tail[2] = {0, 0};
    EM2_CUDA(ctx, cudaMemcpyAsync(&tail[0], vertexOf + groups, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaMemcpyAsync(&tail[1], cellOffset + groups, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaStreamSynchronize(s));
    const uint32_t vertexCount = tail[0], keptTotal = tail[1];
    *vertexCountHost = vertexCount;
    *keptCellsHost = keptTotal;
    *edgeCountHost = 0;
    ctx->stats.kernel_launches += 4;
    if (vertexCount > vertexCapacity) return fail(ctx, EM2_ERR_INVALID, "em2_signature_graph: vertex capacity too small");
    emitVerticesKernel<<<groups + 1, 128, 0, s>>>(groups, groupStart, isVertex, vertexOf, cellOffset, order, vertexOffsets, vertexCell,
                                                  cellOrder, vertexCount, keptTotal);
    EM2_CUDA(ctx, cudaGetLastError());
    // 4. edges: count, scan, fill
    const unsigned vBlocks = (vertexCount + 1 + 127) / 128;
    signatureEdgesKernel<<<vBlocks, 128, 0, s>>>(vertexCount, uint32_t(lshCount), sig, W, vertexCell, counts, nullptr, nullptr, 0);
    {
        size_t bytes = cubScan64;
        EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cubTemp, bytes, counts, offsets, int(vertexCount + 1), s));
    }
    uint64_t edgeCount = 0;
    EM2_CUDA(ctx, cudaMemcpyAsync(&edgeCount, offsets + vertexCount, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaStreamSynchronize(s));
    *edgeCountHost = edgeCount;
    ctx->stats.kernel_launches += 2;
    if (edgeCount > edgeCapacity) return fail(ctx, EM2_ERR_INVALID, "em2_signature_graph: edge capacity too small");
    if (edgeCount) {
        signatureEdgesKernel<<<vBlocks, 128, 0, s>>>(vertexCount, uint32_t(lshCount), sig, W, vertexCell, nullptr, offsets, edges,
                                                     edgeCapacity);
        ctx->stats.kernel_launches++;
    }
    EM2_CUDA(ctx, cudaGetLastError());
    return EM2_OK;
}

}  // namespace em2
