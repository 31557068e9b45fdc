// All-pairs Hamming scan with fused per-cell top-k -- tcgen05 int8 tensor-core variant (sm_100a).
//
// Hamming distance is a dense contraction: with signature bits encoded as +-1 int8,
//     dot(x, y) = K - 2 * hamming(x, y)      (K = bit count padded to a multiple of 128; pad bits agree)
// so the all-pairs scan is the int8 GEMM  D = E * E^T  with exact s32 accumulation, and the top-k
// selection is its epilogue.  Same selection semantics and candidate machinery as the XOR/POPC variant
// (scan_popc.cu, topk.cuh); Hamming distances are bit-exact because every partial sum is an integer
// far below 2^31.
//
// Replaces (reference): countMismatches src/BitSet.hpp:277-288 + the findSimilarPairs4 pair loop
// src/ExpressionMatrixLsh.cpp:218-269.  Nothing here is derived from src/Lsh.cl.
//
// Kernel structure (one persistent CTA per SM, 192 threads, warp specialised):
//   warps 0-3  epilogue : thread t owns query row t of the work item == TMEM lane t.  At the start of an
//                         item it writes its row's K encoded bytes into TMEM (tcgen05.st): the A operand
//                         is ROW STATIONARY IN TENSOR MEMORY (128 lanes x 256 columns) for the whole
//                         sweep and never touches shared memory again.  Per column tile it reads the
//                         accumulator with tcgen05.ld, 32 columns at a time: one compare per candidate
//                         against the row's running bound (dot > K - 2*tau  <=>  hamming < tau), rare
//                         survivors appended to the row's candidate buffer (topk.cuh).
//   warp 4     producer : TMA (cp.async.bulk.tensor, 128B swizzle) streams the B operand -- 128 columns
//                         x 128-byte K-chunks (16 KB) -- through a deep ring (all of shared memory).
//   warp 5     MMA      : one elected thread issues tcgen05.mma.cta_group::1.kind::i8 (A from TMEM, B from
//                         shared memory), M=128 N=128 K=32; accumulators in TMEM, double buffered
//                         (2 x 128 columns), so the epilogue of tile t overlaps the MMAs of tile t+1.
//                         tcgen05.commit releases shared-memory stages and publishes accumulators
//                         through mbarriers.  TMEM map: [0,128) acc0, [128,256) acc1, [256,512) A.
// Work item = (128-row block, column segment); items are dealt round-robin to the persistent CTAs.
#include "common.cuh"
#include "tc05.cuh"
#include "topk.cuh"

#include <algorithm>
#include <cstdlib>

namespace em2 {

namespace {

using namespace tc05;

constexpr int kRowsPerItem = 128;     // UMMA M
constexpr int kTileN = 128;           // UMMA N
constexpr int kChunkBytes = 128;      // K bytes per TMA box / swizzle atom
constexpr int kUmmaK = 32;            // K per tcgen05.mma for 8-bit operands
constexpr int kEpiWarps = 8;          // two column sub-streams per row: 2 epilogue warps per SM sub-partition
constexpr int kSubStreams = kEpiWarps / 4;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr uint32_t kRingBytes = 32 * kEpiThreads * 4;          // per-thread ring of 32 passing columns
constexpr int kMaxPanels = 8;         // K <= 1024
constexpr uint32_t kChunkTileBytes = kTileN * kChunkBytes;     // 16 KB: one K-chunk of a column tile
constexpr int kChunksPerStage = 2;                             // a ring stage carries two K-chunks (32 KB)
constexpr uint32_t kStageBytes = kChunksPerStage * kChunkTileBytes;
constexpr int kStages = 5;                                     // 160 KB ring
constexpr uint32_t kTmemA = 256;      // first TMEM column of the A operand

// v[j] for a run-time j without spilling the array to local memory: a 5-level select tree (31 SEL).
__device__ __forceinline__ int32_t pick32(const uint32_t (&v)[32], int j)
{
    uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (j & 16) ? v[16 + i] : v[i];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = (j & 8) ? a[8 + i] : a[i];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i] = (j & 4) ? b[4 + i] : b[i];
#pragma unroll
    for (int i = 0; i < 2; i++) d[i] = (j & 2) ? c[2 + i] : c[i];
    return int32_t((j & 1) ? d[1] : d[0]);
}

// Instruction descriptor: kind::i8, A/B signed 8-bit K-major, D s32, M=128, N=256.
constexpr uint32_t kInstrDesc = (2u << 4)                        // c_format = S32
                                | (1u << 7)                      // a_format = signed 8-bit
                                | (1u << 10)                     // b_format = signed 8-bit
                                | (uint32_t(kTileN >> 3) << 17)  // n_dim
                                | (uint32_t(kRowsPerItem >> 4) << 24);   // m_dim

struct MmaParams {
    uint64_t cellCount;      // N (columns)
    uint64_t rowBegin, rows; // scanned rows [rowBegin, rowBegin + rows)
    uint32_t K;              // padded bit count == bytes per encoded row
    uint32_t panels;         // K / 128
    uint32_t stages;         // B ring depth
    uint32_t segments;
    uint64_t segmentCols;
    uint32_t rowBlocks;
    uint32_t k, cap, tau0;
    uint64_t* cand;
    uint32_t* candCount;
    unsigned long long* appendedTotal;
    uint16_t* dump;          // optional: all distances of the scanned rows (tests)
};

template <bool DUMP, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
scanMmaKernel(const __grid_constant__ CUtensorMap mapB, const uint8_t* __restrict__ enc, const MmaParams p)
{
    extern __shared__ uint8_t smemRaw[];
    // carve: [B stages][barriers]; 1024-byte alignment for the 128B swizzle
    uint8_t* smB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + size_t(kStages) * kStageBytes);
    uint64_t* aFull = bars + 0;      // epilogue threads -> MMA: the A operand of this item is in TMEM
    uint64_t* accFull = bars + 2;    // [2]
    uint64_t* accEmpty = bars + 4;   // [2]
    uint64_t* bFull = bars + 6;      // [stages]
    uint64_t* bEmpty = bFull + kStages;
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bEmpty + kStages);
    uint32_t* ring = reinterpret_cast<uint32_t*>(bars) + 128;   // [32][kEpiThreads], after 512 B of barriers

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbarInit(aFull, kEpiWarps * 32);
        for (int i = 0; i < 2; i++) {
            mbarInit(accFull + i, 1);
            mbarInit(accEmpty + i, kEpiWarps * 32);
        }
        for (uint32_t i = 0; i < kStages; i++) {
            mbarInit(bFull + i, 1);
            mbarInit(bEmpty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smemAddr(tmemSlot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fenceBefore();
    __syncthreads();
    fenceAfter();
    const uint32_t tmemBase = *tmemSlot;

    const uint32_t items = p.rowBlocks * p.segments;

    if (warp == kEpiWarps) {
        // ===================== TMA producer (B operand) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t stagesPerTile = (p.panels + kChunksPerStage - 1) / kChunksPerStage;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
                const uint32_t seg = item % p.segments;
                const uint64_t colBegin = uint64_t(seg) * p.segmentCols;
                const uint64_t colEnd = min(colBegin + p.segmentCols, p.cellCount);
                const uint32_t tiles = uint32_t((colEnd - colBegin + kTileN - 1) / kTileN);
                for (uint32_t t = 0; t < tiles; t++) {
                    const int32_t col0 = int32_t(colBegin + uint64_t(t) * kTileN);
                    for (uint32_t j = 0; j < stagesPerTile; j++) {
                        const uint32_t kc0 = j * kChunksPerStage;
                        const uint32_t chunks = min(uint32_t(kChunksPerStage), p.panels - kc0);
                        mbarWait(bEmpty + stage, phase ^ 1);
                        mbarExpectTx(bFull + stage, chunks * kChunkTileBytes);
                        uint8_t* dst = smB + size_t(stage) * kStageBytes;
                        for (uint32_t c = 0; c < chunks; c++)
                            tmaLoad2d(dst + c * kChunkTileBytes, &mapB, bFull + stage, int32_t((kc0 + c) * kChunkBytes), col0);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t itemIter = 0, tileIter = 0, stage = 0, phase = 0;
            const uint32_t stagesPerTile = (p.panels + kChunksPerStage - 1) / kChunksPerStage;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x, itemIter++) {
                const uint32_t seg = item % p.segments;
                const uint64_t colBegin = uint64_t(seg) * p.segmentCols;
                const uint64_t colEnd = min(colBegin + p.segmentCols, p.cellCount);
                const uint32_t tiles = uint32_t((colEnd - colBegin + kTileN - 1) / kTileN);
                mbarWait(aFull, itemIter & 1);
                fenceAfter();
                for (uint32_t t = 0; t < tiles; t++, tileIter++) {
                    const uint32_t buf = tileIter & 1;
                    mbarWait(accEmpty + buf, ((tileIter >> 1) & 1) ^ 1);
                    fenceAfter();
                    const uint32_t tmemD = tmemBase + buf * kTileN;
                    uint32_t aCol = tmemBase + kTmemA;
                    uint32_t first = 0;                      // 0 on the tile's first MMA: overwrite the accumulator
                    for (uint32_t j = 0; j < stagesPerTile; j++) {
                        const uint32_t chunks = min(uint32_t(kChunksPerStage), p.panels - j * kChunksPerStage);
                        mbarWait(bFull + stage, phase);
                        fenceAfter();
                        uint32_t bAddr = smemAddr(smB + size_t(stage) * kStageBytes);
                        for (uint32_t c = 0; c < chunks; c++, bAddr += kChunkTileBytes) {
#pragma unroll
                            for (int ks = 0; ks < kChunkBytes / kUmmaK; ks++, aCol += kUmmaK / 4) {
                                mmaI8Ts(tmemD, aCol, makeSmemDesc(bAddr + ks * kUmmaK), kInstrDesc, first);
                                first = 1;
                            }
                        }
                        commit(bEmpty + stage);       // stage reusable once these MMAs have read it
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    commit(accFull + buf);            // accumulator complete (and, on the item's last
                }                                            // tile, every read of the A operand is done)
            }
        }
    } else {
        // ===================== epilogue: thread == query row == TMEM lane =====================
        // thread -> (row = TMEM lane, column sub-stream): warps 0-3 take columns [0,64) of every tile,
        // warps 4-7 columns [64,128); each (row, sub-stream) has its own bound and candidate buffer, so
        // ids stay increasing within a stream (topk.cuh) and the finalize kernel merges the streams.
        const uint32_t dotK = p.K;
        const uint32_t rowInItem = threadIdx.x & (kRowsPerItem - 1);
        const uint32_t sub = threadIdx.x / kRowsPerItem;
        constexpr int kSubCols = kTileN / kSubStreams;
        const uint32_t laneField = uint32_t((warp & 3) * 32) << 16;
        uint32_t tileIter = 0;
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const uint32_t rb = item / p.segments, seg = item % p.segments;
            const uint64_t localRow = uint64_t(rb) * kRowsPerItem + rowInItem;
            const bool valid = localRow < p.rows;
            const uint64_t colBegin = uint64_t(seg) * p.segmentCols;
            const uint64_t colEndLong = min(colBegin + p.segmentCols, p.cellCount);
            const uint32_t colEnd = uint32_t(colEndLong);
            const uint32_t tiles = uint32_t((colEndLong - colBegin + kTileN - 1) / kTileN);

            // A operand: this thread's encoded row -> TMEM lane, columns [kTmemA, kTmemA + K/4).
            // The previous item's MMAs have all completed (its last accFull was waited on below).
            {
                const uint4* src = reinterpret_cast<const uint4*>(enc + (p.rowBegin + (valid ? localRow : 0)) * uint64_t(p.K));
                for (uint32_t c = sub * 32; c < p.K / 4; c += 32 * kSubStreams) {   // sub-streams share the copy
                    uint32_t v[32];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint4 x = valid ? __ldg(src + c / 4 + q) : make_uint4(0, 0, 0, 0);
                        v[4 * q] = x.x;
                        v[4 * q + 1] = x.y;
                        v[4 * q + 2] = x.z;
                        v[4 * q + 3] = x.w;
                    }
                    tmemStore32(tmemBase + laneField + kTmemA + c, v);
                }
                tmemStoreWait();
                fenceBefore();
                mbarArrive(aFull);
            }

            RowState st;
            st.rowId = valid ? uint32_t(p.rowBegin + localRow) : 0xffffffffu;
            st.count = 0;
            st.appended = 0;
            st.tau = valid ? p.tau0 : 0;
            st.buf = p.cand + (uint64_t(seg * kSubStreams + sub) * p.rows + (valid ? localRow : 0)) * p.cap;
            int32_t dotThr = int32_t(dotK) - 2 * int32_t(st.tau);      // hamming < tau  <=>  dot > K - 2 tau
            for (uint32_t t = 0; t < tiles; t++, tileIter++) {
                const uint32_t buf = tileIter & 1;
                mbarWait(accFull + buf, (tileIter >> 1) & 1);
                fenceAfter();
                const uint32_t idBase = uint32_t(colBegin) + t * kTileN + sub * kSubCols;
                const uint32_t taddr = tmemBase + buf * kTileN + sub * kSubCols + laneField;
#pragma unroll 1
                for (int c = 0; c < kSubCols; c += 32) {
                    uint32_t v[32];
                    tmemLoad32(taddr + c, v);
                    tmemLoadWait();
                    if (DUMP) {
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const uint32_t id = idBase + c + j;
                                if (id < colEnd)
                                    p.dump[localRow * p.cellCount + id] = uint16_t((int32_t(dotK) - int32_t(v[j])) >> 1);
                            }
                        }
                    } else {
                        int32_t mx = int32_t(v[0]);
#pragma unroll
                        for (int j = 1; j < 32; j++) mx = max(mx, int32_t(v[j]));
                        if (EPI == 0) {
                            // lanes with a passing column diverge; straight-line mask + select tree
                            if (mx > dotThr) {
                                uint32_t mask = 0;
#pragma unroll
                                for (int j = 0; j < 32; j++) mask |= uint32_t(int32_t(v[j]) > dotThr) << j;
                                do {
                                    const int j = __ffs(int(mask)) - 1;
                                    mask &= mask - 1;
                                    const int32_t val = pick32(v, j);
                                    consider(st, uint32_t(int32_t(dotK) - val) >> 1, idBase + c + j, colEnd);
                                } while (mask);
                            }
                            if (__any_sync(0xffffffffu, mx > dotThr)) {
                                warpPruneIfNeeded(st, p.k, p.cap);
                                dotThr = int32_t(dotK) - 2 * int32_t(st.tau);
                            }
                        } else if (__any_sync(0xffffffffu, mx > dotThr)) {
                            // Some row of this warp has a passing column (common with clustered data).
                            // Keep the warp converged: every lane pushes its passing columns into its
                            // private shared-memory ring with predicated stores, then drains the ring.
                            uint32_t n = 0;
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                if (int32_t(v[j]) > dotThr) {
                                    ring[n * kEpiThreads + threadIdx.x] = (v[j] << 5) | uint32_t(j);
                                    n++;
                                }
                            }
                            for (uint32_t i = 0; i < n; i++) {
                                const uint32_t e = ring[i * kEpiThreads + threadIdx.x];
                                const int32_t val = int32_t(e) >> 5;
                                consider(st, uint32_t(int32_t(dotK) - val) >> 1, idBase + c + (e & 31u), colEnd);
                            }
                            warpPruneIfNeeded(st, p.k, p.cap);
                            dotThr = int32_t(dotK) - 2 * int32_t(st.tau);
                        }
                    }
                }
                fenceBefore();
                mbarArrive(accEmpty + buf);
            }
            if (!DUMP && valid) {
                p.candCount[uint64_t(seg * kSubStreams + sub) * p.rows + localRow] = st.count;
                if (p.appendedTotal && st.appended) atomicAdd(p.appendedTotal, (unsigned long long)st.appended);
            }
        }
    }

    fenceBefore();
    __syncthreads();
    if (warp == kEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmemBase) : "memory");
    }
}

// +-1 int8 expansion of the packed signatures: E[n][p] = bit p set ? +1 : -1, p < K; bits at and
// beyond lshCount (zero in the packed words, or beyond them) encode as -1 in every row.
__global__ void encodeKernel(const uint64_t* __restrict__ sig, uint32_t W, uint64_t cellCount, uint32_t K,
                             uint8_t* __restrict__ enc)
{
    const uint32_t groupsPerRow = K / 16;      // 16 bits -> 16 bytes per thread
    const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (idx >= cellCount * groupsPerRow) return;
    const uint64_t row = idx / groupsPerRow;
    const uint32_t g = uint32_t(idx - row * groupsPerRow);
    const uint32_t w = g >> 2;
    uint32_t bits = 0;
    if (w < W) bits = uint32_t(sig[row * W + w] >> (48 - 16 * (g & 3))) & 0xFFFFu;   // MSB-first
    uint32_t out[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t bit = (bits >> (15 - (4 * q + b))) & 1u;
            word |= (bit ? 0x01u : 0xFFu) << (8 * b);
        }
        out[q] = word;
    }
    *reinterpret_cast<uint4*>(enc + row * K + size_t(g) * 16) = make_uint4(out[0], out[1], out[2], out[3]);
}

int runMma(em2_context* ctx, const uint64_t* signatures, uint64_t cellCount, uint64_t lshCount, uint64_t rowBegin,
           uint64_t rowEnd, uint64_t k, int64_t mismatchMax, const float* lut, em2_pair* pairs, uint32_t* usedCount,
           uint16_t* dump, cudaStream_t s)
{
    const uint64_t rows = rowEnd - rowBegin;
    const uint32_t W = uint32_t(wordCount(lshCount));
    const uint32_t K = uint32_t(roundUp(lshCount, kChunkBytes));
    if (K > kMaxPanels * kChunkBytes)
        return fail(ctx, EM2_ERR_INVALID, "EM2_VARIANT_MMA_I8 supports lshCount <= 1024 (use EM2_VARIANT_POPC)");
    if (cellCount > 0x7fffff00ull) return fail(ctx, EM2_ERR_INVALID, "cellCount too large for the MMA variant");

    // 1. encode
    void* enc = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_ENC, cellCount * K, &enc));
    {
        const uint64_t threads = cellCount * (K / 16);
        encodeKernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>(signatures, W, cellCount, K, static_cast<uint8_t*>(enc));
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());
    }

    // 2. plan + scratch
    MmaParams p{};
    const uint32_t panels = K / kChunkBytes;
    p.stages = kStages;
    ScanPlan plan = makeScanPlan(ctx, rows, cellCount, dump ? 1 : k, kTileN, kRowsPerItem, 1);
    if (dump) {
        plan.segments = 1;
        plan.segmentCols = roundUp(cellCount, kTileN);
    }
    void* cand = nullptr;
    void* candCount = nullptr;
    void* counters = nullptr;
    if (!dump) {
        EM2_TRY(reserve(ctx, em2_context::S_CAND, size_t(plan.segments) * kSubStreams * rows * plan.cap * sizeof(uint64_t), &cand));
        EM2_TRY(reserve(ctx, em2_context::S_CANDCOUNT, size_t(plan.segments) * kSubStreams * rows * sizeof(uint32_t), &candCount));
    }
    EM2_TRY(reserve(ctx, em2_context::S_COUNTERS, 64, &counters));
    p.cellCount = cellCount;
    p.rowBegin = rowBegin;
    p.rows = rows;
    p.K = K;
    p.panels = panels;
    p.segments = plan.segments;
    p.segmentCols = plan.segmentCols;
    p.rowBlocks = plan.rowBlocks;
    p.k = uint32_t(k);
    p.cap = plan.cap;
    p.tau0 = mismatchMax < 0 ? 0u : uint32_t(std::min<int64_t>(mismatchMax, int64_t(lshCount)) + 1);
    p.cand = static_cast<uint64_t*>(cand);
    p.candCount = static_cast<uint32_t*>(candCount);
    p.appendedTotal = static_cast<unsigned long long*>(counters) + 1;
    p.dump = dump;

    CUtensorMap mapB;
    EM2_TRY(makeTensorMapU8(ctx, &mapB, enc, cellCount, K, K, kTileN));

    const size_t smem = 1024 + size_t(kStages) * kStageBytes + 512 + kRingBytes;
    const uint32_t items = plan.rowBlocks * plan.segments;
    const unsigned grid = unsigned(std::min<uint32_t>(items, uint32_t(ctx->smCount)));
    static const int epi = [] { const char* e = std::getenv("EM2_MMA_EPI"); return e ? std::atoi(e) : 0; }();
    auto go = [&](auto kernel) -> int {
        EM2_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kernel<<<grid, kThreads, smem, s>>>(mapB, static_cast<const uint8_t*>(enc), p);
        return EM2_OK;
    };
    if (dump) EM2_TRY(go(scanMmaKernel<true, 0>));
    else if (epi == 0) EM2_TRY(go(scanMmaKernel<false, 0>));
    else EM2_TRY(go(scanMmaKernel<false, 1>));
    ctx->stats.kernel_launches++;
    EM2_CUDA(ctx, cudaGetLastError());
    if (dump) return EM2_OK;
    ScanPlan merged = plan;                    // every (segment, sub-stream) buffer is one list to merge
    merged.segments = plan.segments * kSubStreams;
    return launchFinalize(ctx, merged, rows, k, p.cand, p.candCount, lut, pairs, usedCount, s);
}

}  // namespace

int launchScanMma(em2_context* ctx, const uint64_t* signatures, uint64_t cellCount, uint64_t lshCount,
                  uint64_t rowBegin, uint64_t rowEnd, uint64_t k, int64_t mismatchMax, const float* lut,
                  em2_pair* pairs, uint32_t* usedCount, cudaStream_t s)
{
    return runMma(ctx, signatures, cellCount, lshCount, rowBegin, rowEnd, k, mismatchMax, lut, pairs, usedCount,
                  nullptr, s);
}

int launchMismatchBlockMma(em2_context* ctx, const uint64_t* signatures, uint64_t cellCount, uint64_t lshCount,
                           uint64_t rowBegin, uint64_t rowEnd, uint16_t* out, cudaStream_t s)
{
    return runMma(ctx, signatures, cellCount, lshCount, rowBegin, rowEnd, 1, int64_t(lshCount), nullptr, nullptr,
                  nullptr, out, s);
}

}  // namespace em2
