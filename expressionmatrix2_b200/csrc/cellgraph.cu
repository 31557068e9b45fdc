// CellGraph edge construction on the device.
//
// Replaces the edge loop of CellGraph::CellGraph (reference src/CellGraph.cpp:60-107): for every cell of the graph's
// cell set, in order, walk its SimilarPairs row (stored by decreasing similarity), stop at the first similarity below
// the threshold, skip neighbours that are not vertices, keep at most maxConnectivity of them, and add an undirected
// edge unless it already exists (the reference asks boost::edge(v0, v1) per candidate: O(degree) each).
// The result is a function of the rows only: the edge set is the union of the per-cell selections, an edge carries
// the similarity of its FIRST insertion, and edges appear in insertion order -- i.e. sorted by
// (vertex index of the inserting cell, rank inside that cell's row), minimised over the edge's one or two
// occurrences.  Here: one thread per cell emits its <= maxConnectivity records, a radix sort brings the two
// occurrences of an edge together, a segmented pass keeps the earlier one, a second sort restores insertion order.
#include "common.cuh"

#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace em2 {

namespace {

constexpr uint32_t kNoVertex = 0xffffffffu;

// pass 1: count the records of each cell;  pass 2 (records != nullptr): write them at offsets[cell]
__global__ void edgeRecordsKernel(uint64_t cellCount, uint64_t k, const em2_pair* __restrict__ pairs,
                                  const uint32_t* __restrict__ usedCount, const uint32_t* __restrict__ vertexOf,
                                  double similarityThreshold, uint32_t maxConnectivity, uint64_t* __restrict__ counts,
                                  const uint64_t* __restrict__ offsets, unsigned long long* __restrict__ edgeKeys,
                                  unsigned long long* __restrict__ orderKeys, float* __restrict__ sims)
{
    const uint64_t c = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (c > cellCount) return;
    if (c == cellCount) {
        if (counts) counts[c] = 0;
        return;
    }
    const uint32_t v0 = vertexOf[c];
    uint32_t n = 0;
    if (v0 != kNoVertex) {
        const em2_pair* row = pairs + c * k;
        const uint32_t used = usedCount[c];
        uint64_t out = offsets ? offsets[c] : 0;
        for (uint32_t i = 0; i < used && n < maxConnectivity; i++) {
            const em2_pair p = row[i];
            if (double(p.similarity) < similarityThreshold) break;    // float promoted to double, as in the reference (CellGraph.cpp:84); rows are sorted by decreasing similarity
            const uint32_t v1 = vertexOf[p.cell];
            if (v1 == kNoVertex) continue;
            if (edgeKeys) {
                const uint32_t a = v0 < v1 ? v0 : v1, b = v0 < v1 ? v1 : v0;
                edgeKeys[out] = (uint64_t(a) << 32) | b;
                orderKeys[out] = (uint64_t(v0) << 32) | n;             // rank among the KEPT neighbours: insertion order
                sims[out] = p.similarity;
                out++;
            }
            n++;
        }
    }
    if (counts) counts[c] = n;
}

// After sorting by edge key: the first record of each run of equal keys survives, with the smallest order key of
// the run (a run has one or two records).  flags[i] = 1 for survivors.
__global__ void edgeUniqueKernel(uint64_t records, const unsigned long long* __restrict__ edgeKeys,
                                 const uint32_t* __restrict__ index, const unsigned long long* __restrict__ orderKeys,
                                 unsigned long long* __restrict__ bestOrder, uint64_t* __restrict__ flags)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i > records) return;
    if (i == records) {
        flags[i] = 0;
        return;
    }
    const bool first = i == 0 || edgeKeys[i - 1] != edgeKeys[i];
    flags[i] = first ? 1 : 0;
    if (first) {
        unsigned long long o = orderKeys[index[i]];
        if (i + 1 < records && edgeKeys[i + 1] == edgeKeys[i]) {
            const unsigned long long o2 = orderKeys[index[i + 1]];
            o = o2 < o ? o2 : o;
        }
        bestOrder[i] = o;
    }
}

__global__ void edgeCompactKernel(uint64_t records, const uint64_t* __restrict__ flags, const uint64_t* __restrict__ pos,
                                  const unsigned long long* __restrict__ bestOrder, unsigned long long* __restrict__ outOrder,
                                  uint32_t* __restrict__ outRecord)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= records || !flags[i]) return;
    outOrder[pos[i]] = bestOrder[i];
    outRecord[pos[i]] = uint32_t(i);
}

__global__ void edgeWriteKernel(uint64_t edges, const uint32_t* __restrict__ sortedRecord, const unsigned long long* __restrict__ sortedOrder,
                                const unsigned long long* __restrict__ edgeKeys, const uint32_t* __restrict__ index,
                                const unsigned long long* __restrict__ orderKeys, const float* __restrict__ sims,
                                em2_edge* __restrict__ out)
{
    const uint64_t e = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (e >= edges) return;
    const uint32_t i = sortedRecord[e];                  // position in the edge-key-sorted record array
    const unsigned long long key = edgeKeys[i], order = sortedOrder[e];
    // the occurrence that was inserted first decides orientation and similarity
    uint32_t rec = index[i];
    if (orderKeys[rec] != order) rec = index[i + 1];
    const uint32_t v0 = uint32_t(order >> 32);
    const uint32_t a = uint32_t(key >> 32), b = uint32_t(key);
    em2_edge r;
    r.vertex0 = v0;
    r.vertex1 = v0 == a ? b : a;
    r.similarity = sims[rec];
    out[e] = r;
}

__global__ void iotaKernel(uint64_t n, uint32_t* p)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = uint32_t(i);
}

}  // namespace

int launchCellGraphEdges(em2_context* ctx, uint64_t cellCount, uint64_t k, const em2_pair* pairs, const uint32_t* usedCount,
                         const uint32_t* vertexOf, double similarityThreshold, uint64_t maxConnectivity, em2_edge* edges,
                         uint64_t capacity, uint64_t* edgeCountHost, cudaStream_t s)
{
    if (cellCount == 0) {
        *edgeCountHost = 0;
        return EM2_OK;
    }
    // maxConnectivity == 0 never stops the reference's loop (its `pairs.size() == maxConnectivity` test follows a
    // push_back, CellGraph.cpp:93-95): every stored neighbour counts, i.e. k of them
    if (maxConnectivity == 0) maxConnectivity = k;
    if (cellCount * std::min<uint64_t>(maxConnectivity, k) > 0x7fffffffull)
        return fail(ctx, EM2_ERR_INVALID, "em2_cell_graph_edges: more than 2^31 candidate edges");
    const uint32_t maxConn = uint32_t(std::min<uint64_t>(maxConnectivity, k));
    const double thr = similarityThreshold;             // the reference compares the stored float, promoted, with its double parameter
    const unsigned blocks = unsigned((cellCount + 1 + 255) / 256);
    // offsets
    size_t scanBytes = 0, sortBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, static_cast<uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                  int(cellCount + 1), s);
    void* cnt = nullptr;
    const size_t cntBytes = roundUp((cellCount + 1) * 8, 256);
    EM2_TRY(reserve(ctx, em2_context::S_ROWPERM, 2 * cntBytes + scanBytes, &cnt));
    uint64_t* counts = static_cast<uint64_t*>(cnt);
    uint64_t* offsets = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(cnt) + cntBytes);
    void* scanTemp = static_cast<uint8_t*>(cnt) + 2 * cntBytes;
    edgeRecordsKernel<<<blocks, 256, 0, s>>>(cellCount, k, pairs, usedCount, vertexOf, thr, maxConn, counts, nullptr, nullptr,
                                             nullptr, nullptr);
    EM2_CUDA(ctx, cudaGetLastError());
    EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(scanTemp, scanBytes, counts, offsets, int(cellCount + 1), s));
    uint64_t records = 0;
    EM2_CUDA(ctx, cudaMemcpyAsync(&records, offsets + cellCount, 8, cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->stats.kernel_launches += 1;
    if (records == 0) {
        *edgeCountHost = 0;
        return EM2_OK;
    }
    // record arrays + sort scratch in S_CAND
    cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, static_cast<const unsigned long long*>(nullptr),
                                    static_cast<unsigned long long*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                    static_cast<uint32_t*>(nullptr), int(records), 0, 64, s);
    const size_t r8 = roundUp((records + 1) * 8, 256), r4 = roundUp((records + 1) * 4, 256);
    size_t scan2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan2, static_cast<uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr), int(records + 1), s);
    void* buf = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_CAND, 7 * r8 + 4 * r4 + std::max(sortBytes, scan2) + 4096, &buf));
    uint8_t* q = static_cast<uint8_t*>(buf);
    auto take = [&](size_t bytes) { uint8_t* r = q; q += bytes; return r; };
    auto* edgeKeys = reinterpret_cast<unsigned long long*>(take(r8));
    auto* edgeKeysSorted = reinterpret_cast<unsigned long long*>(take(r8));
    auto* orderKeys = reinterpret_cast<unsigned long long*>(take(r8));
    auto* bestOrder = reinterpret_cast<unsigned long long*>(take(r8));
    auto* flags = reinterpret_cast<uint64_t*>(take(r8));
    auto* pos = reinterpret_cast<uint64_t*>(take(r8));
    auto* uniqOrder = reinterpret_cast<unsigned long long*>(take(r8));
    auto* sims = reinterpret_cast<float*>(take(r4));
    auto* iota = reinterpret_cast<uint32_t*>(take(r4));
    auto* index = reinterpret_cast<uint32_t*>(take(r4));
    auto* uniqRecord = reinterpret_cast<uint32_t*>(take(r4));
    void* sortTemp = take(std::max(sortBytes, scan2) + 256);

    edgeRecordsKernel<<<blocks, 256, 0, s>>>(cellCount, k, pairs, usedCount, vertexOf, thr, maxConn, nullptr, offsets, edgeKeys,
                                             orderKeys, sims);
    EM2_CUDA(ctx, cudaGetLastError());
    iotaKernel<<<unsigned((records + 255) / 256), 256, 0, s>>>(records, iota);
    EM2_CUDA(ctx, cudaGetLastError());
    EM2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sortTemp, sortBytes, edgeKeys, edgeKeysSorted, iota, index, int(records), 0, 64, s));
    const unsigned rblocks = unsigned((records + 1 + 255) / 256);
    edgeUniqueKernel<<<rblocks, 256, 0, s>>>(records, edgeKeysSorted, index, orderKeys, bestOrder, flags);
    EM2_CUDA(ctx, cudaGetLastError());
    EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(sortTemp, scan2, flags, pos, int(records + 1), s));
    uint64_t edgeCount = 0;
    EM2_CUDA(ctx, cudaMemcpyAsync(&edgeCount, pos + records, 8, cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaStreamSynchronize(s));
    *edgeCountHost = edgeCount;
    if (edgeCount > capacity) return fail(ctx, EM2_ERR_INVALID, "em2_cell_graph_edges: edge capacity is smaller than the edge count");
    edgeCompactKernel<<<rblocks, 256, 0, s>>>(records, flags, pos, bestOrder, uniqOrder, uniqRecord);
    EM2_CUDA(ctx, cudaGetLastError());
    // insertion order: sort the unique edges by their order key (reuse the now free arrays)
    auto* sortedOrder = edgeKeys;                        // free after the first sort
    auto* sortedRecord = iota;
    size_t sort2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort2, uniqOrder, sortedOrder, uniqRecord, sortedRecord, int(edgeCount), 0, 64, s);
    EM2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sortTemp, sort2, uniqOrder, sortedOrder, uniqRecord, sortedRecord, int(edgeCount), 0, 64, s));
    edgeWriteKernel<<<unsigned((edgeCount + 255) / 256), 256, 0, s>>>(edgeCount, sortedRecord, sortedOrder, edgeKeysSorted, index,
                                                                      orderKeys, sims, edges);
    EM2_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 6;
    return EM2_OK;
}

}  // namespace em2
