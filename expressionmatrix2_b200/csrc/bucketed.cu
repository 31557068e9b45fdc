// Bucketed LSH search on the device: ExpressionMatrix::findSimilarPairs7 and its bucket
// assignment (reference src/ExpressionMatrixLsh.cpp:507-687, 707-827) -- the reference's route past O(N^2).
//
// Reference semantics, all order dependent and reproduced exactly:
//   * tables: for every slice length (decreasing) and every slice of that length (bits [sliceId*len, (sliceId+1)*len),
//     packed first-bit-highest, BitSet::getBits src/BitSet.hpp:111-119) a cell falls into bucket = slice value if
//     len < log2BucketCount, else MurmurHash64A(&value, 8, 231) & (2^log2BucketCount - 1); a bucket lists its cells
//     in ascending id (they are pushed in id order);
//   * search, per cell0: walk the tables in order and, in each, the cells of cell0's bucket in order; skip cell0 and
//     cells already looked at; every NEW cell counts as a candidate; keep it as a neighbour if its mismatch count is
//     < the threshold (strict, :658); stop everything at maxCheck candidates;
//   * keep the k best neighbours by (mismatch, id) (keepBest + sort, :676-677) and store float(similarity).
// Here: one radix sort per table (bucket, cell id) plus an inclusive max-scan give every cell the start of its bucket;
// one WARP per cell0 then walks its buckets 32 cells at a time -- a ballot prefix reproduces the sequential
// "new candidates in order until maxCheck" rule -- with the "already looked at" set in a private bitmap in global
// memory (cleared through the candidate list afterwards, like the reference's cellMap).
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <vector>

namespace em2 {

namespace {

// MurmurHash64A (Austin Appleby, public domain) of one 8-byte value; the reference calls it with seed 231
// (src/ExpressionMatrixLsh.cpp:640-643, src/MurmurHash2.cpp).
__device__ __forceinline__ uint64_t murmur64a8(uint64_t k, uint64_t seed)
{
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = seed ^ (8ull * m);
    k *= m;
    k ^= k >> r;
    k *= m;
    h ^= k;
    h *= m;
    h ^= h >> r;
    h *= m;
    h ^= h >> r;
    return h;
}

// bits [bitStart, bitStart + len) of a signature, first bit highest
__device__ __forceinline__ uint64_t sliceValue(const uint64_t* __restrict__ s, uint32_t W, uint32_t bitStart, uint32_t len)
{
    const uint32_t w0 = bitStart >> 6, o = bitStart & 63;
    const uint64_t hi = s[w0];
    const uint64_t lo = (w0 + 1 < W) ? s[w0 + 1] : 0ull;
    const uint64_t x = o ? ((hi << o) | (lo >> (64 - o))) : hi;
    return len == 64 ? x : (x >> (64 - len));
}

__device__ __forceinline__ uint32_t bucketOf(const uint64_t* __restrict__ s, uint32_t W, uint32_t bitStart, uint32_t len,
                                             uint32_t log2Buckets)
{
    const uint64_t v = sliceValue(s, W, bitStart, len);
    if (len < log2Buckets) return uint32_t(v);
    return uint32_t(murmur64a8(v, 231) & ((1ull << log2Buckets) - 1ull));
}

__global__ void bucketKeysKernel(uint64_t n, const uint64_t* __restrict__ sig, uint32_t W, uint32_t bitStart, uint32_t len,
                                 uint32_t log2Buckets, uint32_t* __restrict__ keys, uint32_t* __restrict__ cells)
{
    const uint64_t c = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (c >= n) return;
    keys[c] = bucketOf(sig + c * W, W, bitStart, len, log2Buckets);
    cells[c] = uint32_t(c);
}

__global__ void bucketHeadsKernel(uint64_t n, const uint32_t* __restrict__ sortedKeys, uint32_t* __restrict__ head)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n) head[i] = (i == 0 || sortedKeys[i] != sortedKeys[i - 1]) ? uint32_t(i) : 0u;
}

__global__ void scatterBeginKernel(uint64_t n, const uint32_t* __restrict__ sortedCells, const uint32_t* __restrict__ beginOfPos,
                                   uint32_t* __restrict__ cellBegin)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < n) cellBegin[sortedCells[i]] = beginOfPos[i];
}

struct BucketedParams {
    uint64_t cellCount;
    uint32_t W, k, maxCheck, mismatchThreshold, tables, bitmapWords;
    const uint64_t* sig;
    const uint32_t* sortedKeys;      // [tables][N]
    const uint32_t* sortedCells;     // [tables][N]
    const uint32_t* cellBegin;       // [tables][N]
    uint32_t* bitmaps;               // [warps][bitmapWords], all zero between cells
    uint32_t* candidates;            // [warps][maxCheck]
    unsigned long long* neighbours;  // [warps][maxCheck]
    const float* lut;
    em2_pair* pairs;
    uint32_t* usedCount;
};

constexpr int kBucketedWarps = 4;

__global__ void __launch_bounds__(kBucketedWarps * 32)
bucketedSearchKernel(const BucketedParams p)
{
    extern __shared__ __align__(16) unsigned long long smem[];      // per warp: W signature words, then k keys
    const uint32_t lane = threadIdx.x & 31, warpInBlock = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * kBucketedWarps + warpInBlock, warps = gridDim.x * kBucketedWarps;
    unsigned long long* sig0 = smem + size_t(warpInBlock) * (p.W + p.k);
    unsigned long long* best = sig0 + p.W;
    uint32_t* bitmap = p.bitmaps + size_t(warp) * p.bitmapWords;
    uint32_t* cand = p.candidates + size_t(warp) * p.maxCheck;
    unsigned long long* nbr = p.neighbours + size_t(warp) * p.maxCheck;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t N = p.cellCount;

    for (uint64_t cell0 = warp; cell0 < N; cell0 += warps) {
        for (uint32_t w = lane; w < p.W; w += 32) sig0[w] = p.sig[cell0 * p.W + w];
        __syncwarp();
        uint32_t count = 0, nn = 0;
        for (uint32_t t = 0; t < p.tables && count < p.maxCheck; t++) {
            const uint32_t* keys = p.sortedKeys + size_t(t) * N;
            const uint32_t* cells = p.sortedCells + size_t(t) * N;
            const uint32_t begin = p.cellBegin[size_t(t) * N + cell0];
            const uint32_t key0 = keys[begin];
            for (uint64_t i = begin; count < p.maxCheck; i += 32) {
                const uint64_t idx = i + lane;
                const bool inBucket = idx < N && keys[idx] == key0;
                const uint32_t c1 = inBucket ? cells[idx] : 0xffffffffu;
                bool fresh = inBucket && c1 != uint32_t(cell0);
                if (fresh) fresh = ((__ldcg(bitmap + (c1 >> 5)) >> (c1 & 31)) & 1u) == 0;      // not looked at yet
                const uint32_t freshMask = __ballot_sync(0xffffffffu, fresh);
                const uint32_t room = p.maxCheck - count;
                const uint32_t rank = __popc(freshMask & lt);
                const bool accept = fresh && rank < room;                                       // in bucket order, up to maxCheck
                uint32_t ham = 0xffffffffu;
                if (accept) {
                    atomicOr(bitmap + (c1 >> 5), 1u << (c1 & 31));
                    cand[count + rank] = c1;
                    const uint64_t* s1 = p.sig + uint64_t(c1) * p.W;
                    ham = 0;
                    for (uint32_t w = 0; w < p.W; w++) ham += __popcll(sig0[w] ^ s1[w]);
                }
                const bool near = accept && ham < p.mismatchThreshold;
                const uint32_t nearMask = __ballot_sync(0xffffffffu, near);
                if (near) nbr[nn + __popc(nearMask & lt)] = (uint64_t(ham) << 32) | c1;
                nn += __popc(nearMask);
                count += min(uint32_t(__popc(freshMask)), room);
                __syncwarp();
                if (__ballot_sync(0xffffffffu, inBucket) != 0xffffffffu) break;                 // the bucket ended in this batch
            }
        }
        __syncwarp();
        // ---- the k best neighbours by (mismatch, id)
        uint32_t used = nn < p.k ? nn : p.k;
        unsigned long long cut = ~0ull;              // keep keys <= cut
        if (nn > p.k) {
            unsigned long long lo = 0, hi = (uint64_t(p.mismatchThreshold) << 32);
            while (lo < hi) {
                const unsigned long long mid = lo + ((hi - lo) >> 1);
                uint32_t c = 0;
                for (uint32_t e = lane; e < nn; e += 32) c += (nbr[e] <= mid);
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= p.k) hi = mid;
                else lo = mid + 1;
            }
            cut = lo;                                // keys are unique: exactly k keys are <= cut
        }
        uint32_t m = 0;
        for (uint32_t base = 0; base < nn; base += 32) {
            const uint32_t e = base + lane;
            const unsigned long long key = e < nn ? nbr[e] : ~0ull;
            const bool keep = e < nn && key <= cut;
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (keep) best[m + __popc(mask & lt)] = key;
            m += __popc(mask);
        }
        __syncwarp();
        for (uint32_t e = lane; e < used; e += 32) {
            const unsigned long long key = best[e];
            uint32_t rank = 0;
            for (uint32_t f = 0; f < used; f++) rank += (best[f] < key);
            em2_pair pr;
            pr.cell = uint32_t(key);
            pr.similarity = p.lut[uint32_t(key >> 32)];
            p.pairs[cell0 * p.k + rank] = pr;
        }
        for (uint32_t e = used + lane; e < p.k; e += 32) {
            em2_pair z;
            z.cell = 0;
            z.similarity = 0.f;
            p.pairs[cell0 * p.k + e] = z;
        }
        if (lane == 0) p.usedCount[cell0] = used;
        // ---- forget the cells looked at (the reference clears its cellMap the same way, :690-692)
        for (uint32_t e = lane; e < count; e += 32) atomicAnd(bitmap + (cand[e] >> 5), ~(1u << (cand[e] & 31)));
        __syncwarp();
    }
}

}  // namespace

int launchBucketedSearch(em2_context* ctx, const uint64_t* sig, uint64_t cellCount, uint64_t lshCount, uint64_t k,
                         uint32_t mismatchThreshold, const float* lut, const int32_t* sliceLengths, uint64_t sliceLengthCount,
                         uint32_t maxCheck, uint32_t log2BucketCount, em2_pair* pairs, uint32_t* usedCount, cudaStream_t s)
{
    // The reference's `size() == maxCheck` stop never fires for 0 or for values beyond the cell count: both mean "no
    // limit", which is N - 1 candidates at most -- the workspace below is sized by it.
    if (cellCount > 0x7fffffffull) return fail(ctx, EM2_ERR_INVALID, "em2_find_similar_pairs7: more than 2^31 - 1 cells");
    {
        const uint64_t most = cellCount > 1 ? cellCount - 1 : 1;
        if (maxCheck == 0 || maxCheck > most) maxCheck = uint32_t(most);
    }
    const uint64_t N = cellCount;
    const uint32_t W = uint32_t(wordCount(lshCount));
    struct Table {
        uint32_t bitStart, len;
    };
    std::vector<Table> tables;
    for (uint64_t l = 0; l < sliceLengthCount; l++) {
        const uint32_t len = uint32_t(sliceLengths[l]);
        const uint32_t slices = uint32_t(lshCount / len);
        for (uint32_t sl = 0; sl < slices; sl++) tables.push_back(Table{sl * len, len});
    }
    const uint64_t T = tables.size();
    if (T == 0) return fail(ctx, EM2_ERR_INVALID, "em2_find_similar_pairs7: no signature slice fits into lshCount bits");
    if (N * T * 12 > (uint64_t(96) << 30)) return fail(ctx, EM2_ERR_INVALID, "em2_find_similar_pairs7: bucket tables exceed 96 GB");

    // tables
    void* tab = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_INBOX, N * T * 12, &tab));
    uint32_t* sortedKeys = static_cast<uint32_t*>(tab);
    uint32_t* sortedCells = sortedKeys + N * T;
    uint32_t* cellBegin = sortedCells + N * T;
    size_t cubSort = 0, cubScan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cubSort, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(N), 0, 32, s);
    cub::DeviceScan::InclusiveScan(nullptr, cubScan, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), cub::Max(),
                                   int(N), s);
    const size_t cubBytes = roundUp(std::max(cubSort, cubScan), 256);
    const size_t n1 = roundUp(N, 64);
    void* scratch = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_MISC, n1 * 4 * sizeof(uint32_t) + cubBytes, &scratch));
    uint32_t* keys = static_cast<uint32_t*>(scratch);
    uint32_t* cells = keys + n1;
    uint32_t* head = cells + n1;
    uint32_t* beginOfPos = head + n1;
    void* cubTemp = beginOfPos + n1;
    const unsigned blocks = unsigned((N + 255) / 256);
    for (uint64_t t = 0; t < T; t++) {
        bucketKeysKernel<<<blocks, 256, 0, s>>>(N, sig, W, tables[t].bitStart, tables[t].len, log2BucketCount, keys, cells);
        size_t bytes = cubSort;
        const int endBit = int(std::min<uint32_t>(32, std::min<uint32_t>(tables[t].len, log2BucketCount)));
        EM2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(cubTemp, bytes, keys, sortedKeys + t * N, cells, sortedCells + t * N, int(N), 0,
                                                      std::max(endBit, 1), s));
        bucketHeadsKernel<<<blocks, 256, 0, s>>>(N, sortedKeys + t * N, head);
        bytes = cubScan;
        EM2_CUDA(ctx, cub::DeviceScan::InclusiveScan(cubTemp, bytes, head, beginOfPos, cub::Max(), int(N), s));
        scatterBeginKernel<<<blocks, 256, 0, s>>>(N, sortedCells + t * N, beginOfPos, cellBegin + t * N);
    }
    EM2_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 3 * T;

    // search
    const size_t smemPerWarp = (size_t(W) + k) * sizeof(unsigned long long);
    const size_t smem = smemPerWarp * kBucketedWarps;
    if (smem > 200 * 1024) return fail(ctx, EM2_ERR_INVALID, "em2_find_similar_pairs7: k or lshCount too large");
    EM2_CUDA(ctx, cudaFuncSetAttribute(bucketedSearchKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const uint32_t ctas = uint32_t(std::min<uint64_t>((N + kBucketedWarps - 1) / kBucketedWarps, uint64_t(ctx->smCount) * 8));
    const uint64_t warps = uint64_t(ctas) * kBucketedWarps;
    const uint32_t bitmapWords = uint32_t((N + 31) / 32);
    void* ws = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_COLLOG, warps * (size_t(bitmapWords) * 4 + size_t(maxCheck) * 12) + 256, &ws));
    BucketedParams p{};
    p.cellCount = N;
    p.W = W;
    p.k = uint32_t(k);
    p.maxCheck = maxCheck;
    p.mismatchThreshold = mismatchThreshold;
    p.tables = uint32_t(T);
    p.bitmapWords = bitmapWords;
    p.sig = sig;
    p.sortedKeys = sortedKeys;
    p.sortedCells = sortedCells;
    p.cellBegin = cellBegin;
    p.neighbours = static_cast<unsigned long long*>(ws);
    p.candidates = reinterpret_cast<uint32_t*>(p.neighbours + warps * maxCheck);
    p.bitmaps = p.candidates + warps * maxCheck;
    p.lut = lut;
    p.pairs = pairs;
    p.usedCount = usedCount;
    EM2_CUDA(ctx, cudaMemsetAsync(p.bitmaps, 0, warps * size_t(bitmapWords) * 4, s));
    bucketedSearchKernel<<<ctas, kBucketedWarps * 32, smem, s>>>(p);
    EM2_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    return EM2_OK;
}

}  // namespace em2
