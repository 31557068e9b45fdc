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
// Kernel structure (one persistent CTA per SM, 320 threads, warp specialised):
//   warps 0-7  epilogue : thread = (query row == TMEM lane, column sub-stream).  At the start of an
//                         item the threads write their row's K encoded bytes into TMEM (tcgen05.st): the A
//                         operand is ROW STATIONARY IN TENSOR MEMORY (128 lanes x 256 columns) for the whole
//                         sweep and never touches shared memory again.  Per column tile a thread pulls its 64
//                         accumulator columns into registers (two tcgen05.ld), hands the TMEM buffer straight
//                         back to the MMA warp, and only then selects: group maxima against the row's running
//                         bound (dot > K - 2*lim  <=>  hamming < lim), survivors appended to the row's candidate
//                         buffer (topk.cuh).  The two sub-streams of a row trade bounds through shared memory.
//   warp 8     producer : TMA (cp.async.bulk.tensor, 128B swizzle) streams the B operand -- 128 columns
//                         x 128-byte K-chunks (16 KB) -- through a deep ring (all of shared memory).
//   warp 9     MMA      : one elected thread issues tcgen05.mma.cta_group::1.kind::i8 (A from TMEM, B from
//                         shared memory), M=128 N=128 K=32; accumulators in TMEM, double buffered
//                         (2 x 128 columns), so the epilogue of tile t overlaps the MMAs of tile t+1.
//                         tcgen05.commit releases shared-memory stages and publishes accumulators
//                         through mbarriers.  TMEM map: [0,128) acc0, [128,256) acc1, [256,512) A.
// Work item = 128-row block x all columns, or x one column segment for the tail row blocks that cannot fill
// a wave (ScanPlan, common.cuh); items are dealt round-robin to the persistent CTAs.
#include "common.cuh"
#include "tc05.cuh"
#include "topk.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

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
constexpr uint32_t kShareBytes = kEpiThreads * 4;              // bound exchange between a row's two sub-streams
constexpr int kMaxPanels = 8;         // K <= 1024
constexpr uint32_t kChunkTileBytes = kTileN * kChunkBytes;     // 16 KB: one K-chunk of a column tile
constexpr int kChunksPerStage = 2;                             // a ring stage carries two K-chunks (32 KB)
constexpr uint32_t kStageBytes = kChunksPerStage * kChunkTileBytes;
constexpr int kStages = 5;                                     // 160 KB ring
constexpr uint32_t kTmemA = 256;      // first TMEM column of the A operand
// CTA-pair variant (cta_group::2, M = 256): each CTA streams HALF of every B tile (64 columns), so a K-chunk is
// 8 KB per CTA, a stage 16 KB, and the same 160 KB hold 10 stages.
constexpr int kPairRows = 2 * kRowsPerItem;
constexpr uint32_t kPairChunkBytes = (kTileN / 2) * kChunkBytes;
constexpr uint32_t kPairStageBytes = kChunksPerStage * kPairChunkBytes;
constexpr int kPairStages = 10;
constexpr int kMaxStages = 10;

// Instruction descriptor: kind::i8, A/B signed 8-bit K-major, D s32, M=128, N=256.
constexpr uint32_t kInstrDesc = (2u << 4)                        // c_format = S32
                                | (1u << 7)                      // a_format = signed 8-bit
                                | (1u << 10)                     // b_format = signed 8-bit
                                | (uint32_t(kTileN >> 3) << 17)  // n_dim
                                | (uint32_t(kRowsPerItem >> 4) << 24);   // m_dim
constexpr uint32_t kInstrDescPair = (kInstrDesc & ~(0x1Fu << 24)) | (uint32_t(kPairRows >> 4) << 24);   // M = 256 over two CTAs

struct MmaParams {
    uint64_t cellCount;      // N (columns)
    uint64_t rowBegin, rows; // scanned rows [rowBegin, rowBegin + rows)
    uint32_t K;              // padded bit count == bytes per encoded row
    uint32_t panels;         // K / 128
    uint32_t stages;         // B ring depth
    uint32_t mainBlocks, segments, items;   // work decomposition (ScanPlan, common.cuh)
    uint64_t segmentCols;
    uint32_t k, cap, tau0;
    uint64_t* cand;
    uint32_t* candCount;
    unsigned long long* appendedTotal;
    uint16_t* dump;          // optional: all distances of the scanned rows (tests)
    uint32_t flags;          // debug: bit 0 = no bound sharing between sub-streams
    const uint8_t* encRows;  // encoded signatures of the scanned rows, indexed by scan position (TMEM-resident kernel)
    const uint32_t* rowPerm; // scan position -> cell id of the row (nullptr: rowBegin + position)
};

template <bool DUMP, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
scanMmaKernel(const __grid_constant__ CUtensorMap mapB, const uint8_t* __restrict__ enc, const MmaParams p)
{
    constexpr int kNumStages = PAIR ? kPairStages : kStages;
    constexpr uint32_t kStageSz = PAIR ? kPairStageBytes : kStageBytes;
    constexpr uint32_t kChunkSz = PAIR ? kPairChunkBytes : kChunkTileBytes;
    extern __shared__ uint8_t smemRaw[];
    // carve: [B stages][barriers]; 1024-byte alignment for the 128B swizzle
    uint8_t* smB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + size_t(kNumStages) * kStageSz);
    uint64_t* aFull = bars + 0;      // epilogue threads -> MMA: the A operand of this item is in TMEM
    uint64_t* accFull = bars + 2;    // [2]
    uint64_t* accEmpty = bars + 4;   // [2]
    uint64_t* bFull = bars + 6;      // [stages]
    uint64_t* bEmpty = bFull + kMaxStages;
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bEmpty + kMaxStages);
    uint32_t* tauShare = reinterpret_cast<uint32_t*>(bars) + 128;   // [kSubStreams][kRowsPerItem], after 512 B of barriers

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // PAIR: the two CTAs of a cluster work on one 256-row item; rank 0 is the leader (it owns the barriers the
    // MMA thread waits on and issues the MMAs); every CTA runs its own producer and its own epilogue.
    const uint32_t rank = PAIR ? clusterRank() : 0;
    const uint32_t worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
    const uint32_t workers = PAIR ? (gridDim.x >> 1) : gridDim.x;

    if (threadIdx.x == 0) {
        // PAIR: one arrival per epilogue warp of either CTA; otherwise one per epilogue thread
        mbarInit(aFull, PAIR ? 2 * kEpiWarps : kEpiWarps * 32);
        for (int i = 0; i < 2; i++) {
            mbarInit(accFull + i, 1);
            mbarInit(accEmpty + i, PAIR ? 2 * kEpiWarps : kEpiWarps * 32);
        }
        for (uint32_t i = 0; i < kNumStages; i++) {
            mbarInit(bFull + i, 1);
            mbarInit(bEmpty + i, 1);
        }
        mbarInitFence();
    }
    if (warp == kEpiWarps) {
        if (PAIR) tmemAllocPair(tmemSlot, 512);
        else tmemAlloc(tmemSlot, 512);
    }
    fenceBefore();
    if (PAIR) clusterSync();      // barriers of both CTAs are initialised before anyone signals across
    else __syncthreads();
    fenceAfter();
    const uint32_t tmemBase = *tmemSlot;

    const uint32_t items = p.items;

    if (warp == kEpiWarps) {
        // ===================== TMA producer (B operand; PAIR: this CTA's half of every tile) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t stagesPerTile = (p.panels + kChunksPerStage - 1) / kChunksPerStage;
            for (uint32_t item = worker; item < items; item += workers) {
                const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, p.cellCount);
                const uint64_t colBegin = it.colBegin, colEnd = it.colEnd;
                const uint32_t tiles = uint32_t((colEnd - colBegin + kTileN - 1) / kTileN);
                for (uint32_t t = 0; t < tiles; t++) {
                    const int32_t col0 = int32_t(colBegin + uint64_t(t) * kTileN + (PAIR ? rank * (kTileN / 2) : 0));
                    for (uint32_t j = 0; j < stagesPerTile; j++) {
                        const uint32_t kc0 = j * kChunksPerStage;
                        const uint32_t chunks = min(uint32_t(kChunksPerStage), p.panels - kc0);
                        mbarWait(bEmpty + stage, phase ^ 1);
                        uint8_t* dst = smB + size_t(stage) * kStageSz;
                        if (PAIR) {
                            // the leader's barrier collects the bytes of both halves
                            if (rank == 0) mbarExpectTx(bFull + stage, 2 * chunks * kChunkSz);
                            for (uint32_t c = 0; c < chunks; c++)
                                tmaLoad2dPair(dst + c * kChunkSz, &mapB, bFull + stage, int32_t((kc0 + c) * kChunkBytes), col0);
                        } else {
                            mbarExpectTx(bFull + stage, chunks * kChunkSz);
                            for (uint32_t c = 0; c < chunks; c++)
                                tmaLoad2d(dst + c * kChunkSz, &mapB, bFull + stage, int32_t((kc0 + c) * kChunkBytes), col0);
                        }
                        if (++stage == kNumStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ===================== MMA issuer (PAIR: leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            uint32_t itemIter = 0, tileIter = 0, stage = 0, phase = 0;
            const uint32_t stagesPerTile = (p.panels + kChunksPerStage - 1) / kChunksPerStage;
            for (uint32_t item = worker; item < items; item += workers, itemIter++) {
                const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, p.cellCount);
                const uint64_t colBegin = it.colBegin, colEnd = it.colEnd;
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
                        uint32_t bAddr = smemAddr(smB + size_t(stage) * kStageSz);
                        for (uint32_t c = 0; c < chunks; c++, bAddr += kChunkSz) {
#pragma unroll
                            for (int ks = 0; ks < kChunkBytes / kUmmaK; ks++, aCol += kUmmaK / 4) {
                                if (PAIR) mmaI8TsPair(tmemD, aCol, makeSmemDesc(bAddr + ks * kUmmaK), kInstrDescPair, first);
                                else mmaI8Ts(tmemD, aCol, makeSmemDesc(bAddr + ks * kUmmaK), kInstrDesc, first);
                                first = 1;
                            }
                        }
                        // stage reusable once these MMAs have read it (PAIR: in both CTAs)
                        if (PAIR) commitPair(bEmpty + stage);
                        else commit(bEmpty + stage);
                        if (++stage == kNumStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    // accumulator complete (and, on the item's last tile, every read of the A operand is done)
                    if (PAIR) commitPair(accFull + buf);
                    else commit(accFull + buf);
                }
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
        static_assert(kSubCols == 64 && kSubStreams == 2, "the epilogue below handles two 32-column chunks per thread");
        const uint32_t laneField = uint32_t((warp & 3) * 32) << 16;
        uint32_t tileIter = 0;
        for (uint32_t item = worker; item < items; item += workers) {
            const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, p.cellCount);
            const uint32_t seg = it.segment;
            const uint64_t localRow = PAIR ? uint64_t(it.rowBlock) * kPairRows + rank * kRowsPerItem + rowInItem
                                           : uint64_t(it.rowBlock) * kRowsPerItem + rowInItem;
            const bool valid = localRow < p.rows;
            const uint64_t colBegin = it.colBegin;
            const uint32_t colEnd = uint32_t(it.colEnd);
            const uint32_t tiles = uint32_t((it.colEnd - colBegin + kTileN - 1) / kTileN);

            // A operand: this thread's encoded row -> TMEM lane, columns [kTmemA, kTmemA + K/4).
            // The previous item's MMAs have all completed (its last accFull was waited on below).
            {
                const uint4* src = reinterpret_cast<const uint4*>(p.encRows + (valid ? localRow : 0) * uint64_t(p.K));
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
                if (PAIR) {
                    __syncwarp();
                    if (lane == 0) mbarArriveLeader(aFull);
                } else {
                    mbarArrive(aFull);
                }
            }

            RowState st;
            st.rowId = !valid ? 0xffffffffu : p.rowPerm ? p.rowPerm[localRow] : uint32_t(p.rowBegin + localRow);
            st.count = 0;
            st.appended = 0;
            st.tau = valid ? p.tau0 : 0;
            st.lim = st.tau;
            st.buf = p.cand + (uint64_t(seg * kSubStreams + sub) * p.rows + (valid ? localRow : 0)) * p.cap;
            // The two sub-streams of a row exchange their bounds through shared memory (stale values are only
            // looser).  The slot is re-initialised per item; the barrier keeps a fast warp from reading the
            // previous item's value.  The initial value must be harmless for ANY row: the partner warp may still
            // be finishing the previous item (a different row) when it reads it -- so never the 0 of a padding row.
            tauShare[sub * kRowsPerItem + rowInItem] = p.tau0;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            int32_t dotThr = int32_t(dotK) - 2 * int32_t(st.lim);      // hamming < lim  <=>  dot > K - 2 lim

            // One 32-column chunk held in registers: group maxima first (the common case is "nothing passes"),
            // then only the 8-column groups that contain a passing column are examined, with static indexing.
            auto chunk = [&](const uint32_t (&v)[32], uint32_t id0) {
                int32_t m[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    m[g] = int32_t(v[8 * g]);
#pragma unroll
                    for (int j = 1; j < 8; j++) m[g] = max(m[g], int32_t(v[8 * g + j]));
                }
                const int32_t mx = max(max(m[0], m[1]), max(m[2], m[3]));
                if (mx > dotThr) {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (m[g] > dotThr) {
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                if (int32_t(v[8 * g + j]) > dotThr)
                                    consider(st, uint32_t(int32_t(dotK) - int32_t(v[8 * g + j])) >> 1, id0 + 8 * g + j, colEnd);
                        }
                    }
                }
                if (__any_sync(0xffffffffu, mx > dotThr)) {
                    warpPruneIfNeeded(st, p.k, p.cap);
                    dotThr = int32_t(dotK) - 2 * int32_t(st.lim);
                }
            };

            for (uint32_t t = 0; t < tiles; t++, tileIter++) {
                const uint32_t buf = tileIter & 1;
                mbarWait(accFull + buf, (tileIter >> 1) & 1);
                fenceAfter();
                const uint32_t idBase = uint32_t(colBegin) + t * kTileN + sub * kSubCols;
                const uint32_t taddr = tmemBase + buf * kTileN + sub * kSubCols + laneField;
                // Both chunks into registers, then hand the accumulator back at once: selection work (and the
                // occasional prune) overlaps the MMAs of the next TWO tiles instead of holding a TMEM buffer.
                uint32_t v0[32], v1[32];
                tmemLoad32(taddr, v0);
                tmemLoad32(taddr + 32, v1);
                tmemLoadWait();
                fenceBefore();
                if (PAIR) {
                    __syncwarp();
                    if (lane == 0) mbarArriveLeader(accEmpty + buf);
                } else {
                    mbarArrive(accEmpty + buf);
                }
                if (DUMP) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const uint32_t id = idBase + j;
                            if (id < colEnd) p.dump[localRow * p.cellCount + id] = uint16_t((int32_t(dotK) - int32_t(v0[j])) >> 1);
                            if (id + 32 < colEnd)
                                p.dump[localRow * p.cellCount + id + 32] = uint16_t((int32_t(dotK) - int32_t(v1[j])) >> 1);
                        }
                    }
                } else {
                    const uint32_t other = tauShare[(sub ^ 1) * kRowsPerItem + rowInItem];
                    if (other + 1 < st.lim && !(p.flags & 1)) {       // a tie with the other stream's k-th best can still win on id
                        st.lim = other + 1;
                        dotThr = int32_t(dotK) - 2 * int32_t(st.lim);
                    }
                    chunk(v0, idBase);
                    chunk(v1, idBase + 32);
                    if (valid) tauShare[sub * kRowsPerItem + rowInItem] = st.tau;
                }
            }
            if (!DUMP && valid) {
                p.candCount[uint64_t(seg * kSubStreams + sub) * p.rows + localRow] = st.count;
                if (p.appendedTotal && st.appended) atomicAdd(p.appendedTotal, (unsigned long long)st.appended);
            }
        }
    }

    fenceBefore();
    if (PAIR) {
        clusterSync();        // neither CTA may retire while its partner can still signal into it
        if (warp == kEpiWarps) tmemDeallocPair(tmemBase, 512);
    } else {
        __syncthreads();
        if (warp == kEpiWarps) tmemDealloc(tmemBase, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Any L (used for L > 1024, where the A operand no longer fits in tensor memory): both operands stream.
// Tile = 128 rows x 256 columns, K looped in 128-byte chunks; per chunk the producer loads the row block's
// A chunk (16 KB) and the column tile's B chunk (32 KB) by TMA into a 4-stage ring, the MMA thread issues four
// M=128 N=256 K=32 instructions with both descriptors in shared memory, accumulators double buffered in TMEM
// (2 x 256 columns).  The main loop of a tile is K/1024 times longer than in the kernel above, so the epilogue
// (thread = row x one of two 128-column sub-streams, four 32-column chunks per tile) has time to spare.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSsTileN = 256;
constexpr int kSsStages = 4;
constexpr uint32_t kSsABytes = kRowsPerItem * kChunkBytes;      // 16 KB
constexpr uint32_t kSsBBytes = kSsTileN * kChunkBytes;          // 32 KB
constexpr uint32_t kSsStageBytes = kSsABytes + kSsBBytes;
constexpr uint32_t kInstrDescSs = (kInstrDesc & ~(0x3Fu << 17)) | (uint32_t(kSsTileN >> 3) << 17);

template <bool DUMP>
__global__ void __launch_bounds__(kThreads, 1)
scanMmaSsKernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const MmaParams p)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + size_t(kSsStages) * kSsStageBytes);
    uint64_t* accFull = bars + 0;    // [2]
    uint64_t* accEmpty = bars + 2;   // [2]
    uint64_t* full = bars + 4;       // [stages]
    uint64_t* empty = full + kSsStages;
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(empty + kSsStages);
    uint32_t* tauShare = reinterpret_cast<uint32_t*>(bars) + 64;    // [kSubStreams][kRowsPerItem], after 256 B of barriers

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbarInit(accFull + i, 1);
            mbarInit(accEmpty + i, kEpiWarps * 32);
        }
        for (int i = 0; i < kSsStages; i++) {
            mbarInit(full + i, 1);
            mbarInit(empty + i, 1);
        }
        mbarInitFence();
    }
    if (warp == kEpiWarps) tmemAlloc(tmemSlot, 512);
    fenceBefore();
    __syncthreads();
    fenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    const uint32_t items = p.items;

    if (warp == kEpiWarps) {
        // ===================== TMA producer (A chunk of the row block + B chunk of the column tile) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
                const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, p.cellCount);
                const int32_t rowA = int32_t(uint64_t(it.rowBlock) * kRowsPerItem);      // mapA covers the scanned rows only
                const uint32_t tiles = uint32_t((it.colEnd - it.colBegin + kSsTileN - 1) / kSsTileN);
                for (uint32_t t = 0; t < tiles; t++) {
                    const int32_t col0 = int32_t(it.colBegin + uint64_t(t) * kSsTileN);
                    for (uint32_t kc = 0; kc < p.panels; kc++) {
                        mbarWait(empty + stage, phase ^ 1);
                        mbarExpectTx(full + stage, kSsStageBytes);
                        uint8_t* dst = ring + size_t(stage) * kSsStageBytes;
                        tmaLoad2d(dst, &mapA, full + stage, int32_t(kc * kChunkBytes), rowA);
                        tmaLoad2d(dst + kSsABytes, &mapB, full + stage, int32_t(kc * kChunkBytes), col0);
                        if (++stage == kSsStages) {
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
            uint32_t tileIter = 0, stage = 0, phase = 0;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
                const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, p.cellCount);
                const uint32_t tiles = uint32_t((it.colEnd - it.colBegin + kSsTileN - 1) / kSsTileN);
                for (uint32_t t = 0; t < tiles; t++, tileIter++) {
                    const uint32_t buf = tileIter & 1;
                    mbarWait(accEmpty + buf, ((tileIter >> 1) & 1) ^ 1);
                    fenceAfter();
                    const uint32_t tmemD = tmemBase + buf * kSsTileN;
                    uint32_t accumulate = 0;
                    for (uint32_t kc = 0; kc < p.panels; kc++) {
                        mbarWait(full + stage, phase);
                        fenceAfter();
                        const uint32_t aAddr = smemAddr(ring + size_t(stage) * kSsStageBytes);
                        const uint32_t bAddr = aAddr + kSsABytes;
#pragma unroll
                        for (int ks = 0; ks < kChunkBytes / kUmmaK; ks++) {
                            mmaI8Ss(tmemD, makeSmemDesc(aAddr + ks * kUmmaK), makeSmemDesc(bAddr + ks * kUmmaK), kInstrDescSs, accumulate);
                            accumulate = 1;
                        }
                        commit(empty + stage);
                        if (++stage == kSsStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    commit(accFull + buf);
                }
            }
        }
    } else {
        // ===================== epilogue: thread == (query row == TMEM lane, 128-column sub-stream) =====================
        const uint32_t dotK = p.K;
        const uint32_t rowInItem = threadIdx.x & (kRowsPerItem - 1);
        const uint32_t sub = threadIdx.x / kRowsPerItem;
        constexpr int kSubCols = kSsTileN / kSubStreams;
        const uint32_t laneField = uint32_t((warp & 3) * 32) << 16;
        uint32_t tileIter = 0;
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, p.cellCount);
            const uint32_t seg = it.segment;
            const uint64_t localRow = uint64_t(it.rowBlock) * kRowsPerItem + rowInItem;
            const bool valid = localRow < p.rows;
            const uint64_t colBegin = it.colBegin;
            const uint32_t colEnd = uint32_t(it.colEnd);
            const uint32_t tiles = uint32_t((it.colEnd - colBegin + kSsTileN - 1) / kSsTileN);

            RowState st;
            st.rowId = !valid ? 0xffffffffu : p.rowPerm ? p.rowPerm[localRow] : uint32_t(p.rowBegin + localRow);
            st.count = 0;
            st.appended = 0;
            st.tau = valid ? p.tau0 : 0;
            st.lim = st.tau;
            st.buf = p.cand + (uint64_t(seg * kSubStreams + sub) * p.rows + (valid ? localRow : 0)) * p.cap;
            tauShare[sub * kRowsPerItem + rowInItem] = p.tau0;      // harmless for any row (see scanMmaKernel)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            int32_t dotThr = int32_t(dotK) - 2 * int32_t(st.lim);      // hamming < lim  <=>  dot > K - 2 lim

            auto chunk = [&](const uint32_t (&v)[32], uint32_t id0) {
                int32_t m[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    m[g] = int32_t(v[8 * g]);
#pragma unroll
                    for (int j = 1; j < 8; j++) m[g] = max(m[g], int32_t(v[8 * g + j]));
                }
                const int32_t mx = max(max(m[0], m[1]), max(m[2], m[3]));
                if (mx > dotThr) {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (m[g] > dotThr) {
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                if (int32_t(v[8 * g + j]) > dotThr)
                                    consider(st, uint32_t(int32_t(dotK) - int32_t(v[8 * g + j])) >> 1, id0 + 8 * g + j, colEnd);
                        }
                    }
                }
                if (__any_sync(0xffffffffu, mx > dotThr)) {
                    warpPruneIfNeeded(st, p.k, p.cap);
                    dotThr = int32_t(dotK) - 2 * int32_t(st.lim);
                }
            };

            for (uint32_t t = 0; t < tiles; t++, tileIter++) {
                const uint32_t buf = tileIter & 1;
                mbarWait(accFull + buf, (tileIter >> 1) & 1);
                fenceAfter();
                const uint32_t idBase = uint32_t(colBegin) + t * kSsTileN + sub * kSubCols;
                const uint32_t taddr = tmemBase + buf * kSsTileN + sub * kSubCols + laneField;
                if (!DUMP) {
                    const uint32_t other = tauShare[(sub ^ 1) * kRowsPerItem + rowInItem];
                    if (other + 1 < st.lim && !(p.flags & 1)) {
                        st.lim = other + 1;
                        dotThr = int32_t(dotK) - 2 * int32_t(st.lim);
                    }
                }
#pragma unroll 1
                for (int c = 0; c < kSubCols; c += 32) {
                    // one chunk per round trip (two at a time measured slower: 168 registers and spills)
                    uint32_t v[32];
                    tmemLoad32(taddr + c, v);
                    tmemLoadWait();
                    if (c + 32 == kSubCols) {          // last chunk is in registers: hand the accumulator back
                        fenceBefore();
                        mbarArrive(accEmpty + buf);
                    }
                    if (DUMP) {
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const uint32_t id = idBase + c + j;
                                if (id < colEnd) p.dump[localRow * p.cellCount + id] = uint16_t((int32_t(dotK) - int32_t(v[j])) >> 1);
                            }
                        }
                    } else {
                        chunk(v, idBase + c);
                    }
                }
                if (!DUMP && valid) tauShare[sub * kRowsPerItem + rowInItem] = st.tau;
            }
            if (!DUMP && valid) {
                p.candCount[uint64_t(seg * kSubStreams + sub) * p.rows + localRow] = st.count;
                if (p.appendedTotal && st.appended) atomicAdd(p.appendedTotal, (unsigned long long)st.appended);
            }
        }
    }
    fenceBefore();
    __syncthreads();
    if (warp == kEpiWarps) tmemDealloc(tmemBase, 512);
}

// ---------------------------------------------------------------------------------------------------------
// Symmetric scan (whole-matrix jobs, K <= 1024): every unordered pair of cells is evaluated ONCE.
//
// Cells are taken in scan-position order (the grouped order when row grouping is on; rows AND columns) and cut
// into super blocks of 256 positions.  The CTA that owns row block a (128 rows of super block A = a / 2) visits
// the column tiles C = (A - d) mod S for the offsets d = 0 .. S/2 -- backwards: row blocks are processed in
// increasing order, so the column cells of most tiles have already had their own sweep and carry their final,
// tight bounds.  Of every unordered pair of distinct super blocks exactly one owner sees the pair's tile (for even
// S the offset S/2 is seen by both owners and treated as two one-directional tiles, like the diagonal d = 0;
// tests/test_symmetric_schedule.py restates the rule).  A tile's accumulators then feed BOTH directions:
//   row direction     the thread's own row, exactly as in the kernels above (private bound, private candidate
//                     region), except that candidates do not arrive in id order: the order-independent prune of
//                     topk.cuh is used and every new bound is published to limEx[row position] (atomicMin);
//   column direction  the same 32 accumulators are compared with the bounds of the tile's COLUMN cells
//                     (limEx, staged per tile in shared memory as dot-product thresholds, with one loosest
//                     threshold per 8 columns so that the common case is four extra compares per chunk);
//                     survivors go to a log, a pool of 64-entry chunks of which a thread owns one at a time (a plain
//                     store per survivor, one atomic per 64: an atomic slot per survivor cost a ~700-cycle round
//                     trip each and made the sweep 3x slower on clustered data); afterwards the log is filed into
//                     per-cell inboxes of exactly the needed length (count -> scan -> fill).
//                     For the same reason candidate keys carry scan positions, not cell ids; ids are looked up where
//                     they decide something (ties at a prune, the finalize kernel).
// limEx starts from a sampling pre-pass (the k-th best of every cell against N/32 sample cells, a valid upper
// bound of its final k-th best) and tightens as the cell's own row streams progress; a stale bound is only
// looser.  The finalize kernel merges a cell's row streams and its inbox.  If the log pool runs dry a flag is
// raised and the caller reruns the job with the one-directional kernels: exactness never depends on the bounds
// being tight.
//
// Operand traffic: per tile the one-directional kernel re-reads the row block's A chunks from L2 (48 KB per
// K-chunk and SM, ~19 TB/s at full tensor rate -- sustainable only because all CTAs stream the SAME B tile).
// Here the B tiles of concurrently running CTAs differ (a sliding window of ~74 super blocks), so A is kept
// RESIDENT IN SHARED MEMORY for the whole item (K x 128 B <= 128 KB) and only B streams: 32 KB per K-chunk.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSymMaxStages = 6;
constexpr uint32_t kLogChunk = 64;
constexpr uint32_t kPaceWindow = 256;      // tiles a CTA may run ahead of the slowest one (64 MB of distinct B tiles)
constexpr uint32_t kSymSmallBytes = 192 + 2 * kSsTileN * 2 + 2 * (kSsTileN / 8) * 2 + kEpiThreads * 2;

struct SymParams {
    uint64_t cellCount;            // N: rows == columns, in scan-position order
    uint32_t K, panels, stages;
    uint32_t mainBlocks, segments, items;
    uint64_t segmentCols;          // in virtual columns: offset * 256
    uint32_t superBlocks;          // S = ceil(N / 256)
    uint32_t offsets;              // column tiles per row block: S / 2 + 1
    uint32_t halfOffset;           // S even: the offset visited by both owners (no column direction); else 0
    uint32_t dBegin;               // this launch sweeps the offsets [dBegin, dBegin + offsetsHere)
    uint32_t offsetsHere;
    uint32_t resume;               // 1: segment 0 of every row CONTINUES the row's streams of the previous launch
    uint32_t k, cap;
    uint64_t* cand;
    uint32_t* candCount;
    unsigned long long* appendedTotal;   // both directions
    uint32_t* limEx;               // per position: accept iff mismatch count < limEx
    ulonglong2* colLog;            // column-direction survivors {mismatch << 32 | row cell id, column position}: a pool of
    uint32_t* chunkFill;           //   64-entry chunks; a thread takes a chunk at a time (one atomic per 64 survivors);
    uint32_t* chunkNext;           //   chunkFill[c] = valid entries of chunk c; scattered to the inboxes afterwards
    uint32_t chunkCap;
    uint32_t* overflow;
    uint32_t* progress;            // [grid] pacing of whole-sweep items (nullptr: none)
    const uint32_t* perm;          // position -> cell id (nullptr: identity)
    uint32_t flags;
};

// Out of line on purpose (the call sits in 32 unrolled places of the epilogue): closes the thread's full chunk and
// takes the next one from the pool.  A dry pool hands out the spill chunk at index chunkCap again and again; its
// content is never read and the kernel raises the overflow flag at the end.
static __device__ __noinline__ ulonglong2* nextLogChunk(uint32_t* chunkFill, uint32_t* chunkNext, uint32_t chunkCap, ulonglong2* pool,
                                                        ulonglong2* logNext, uint32_t& logFill)
{
    if (logNext) {
        const uint64_t chunk = uint64_t(logNext - 1 - pool) / kLogChunk;
        if (chunk < chunkCap) chunkFill[chunk] = kLogChunk;
    }
    const uint32_t c = min(atomicAdd(chunkNext, 1u), chunkCap);
    logFill = 0;
    return pool + uint64_t(c) * kLogChunk;
}

__global__ void __launch_bounds__(kThreads, 1)
scanMmaSymKernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const SymParams p)
{
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smA + size_t(p.panels) * kSsABytes;
    uint8_t* small = ring + size_t(p.stages) * kSsBBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(small);
    uint64_t* accFull = bars + 0;    // [2]
    uint64_t* accEmpty = bars + 2;   // [2]
    uint64_t* aFull = bars + 4;      // the item's A operand has landed
    uint64_t* aEmpty = bars + 5;     // every MMA of the item has read it
    uint64_t* full = bars + 6;       // [stages]
    uint64_t* empty = full + kSymMaxStages;
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(empty + kSymMaxStages);
    int16_t* colThr = reinterpret_cast<int16_t*>(small + 192);          // [2][256] dot thresholds of the tile's columns
    int16_t* grpThr = colThr + 2 * kSsTileN;                             // [2][32]  loosest threshold of each 8 columns
    uint16_t* tauShare = reinterpret_cast<uint16_t*>(grpThr + 2 * (kSsTileN / 8));   // [kSubStreams][kRowsPerItem]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbarInit(accFull + i, 1);
            mbarInit(accEmpty + i, kEpiWarps * 32);
        }
        mbarInit(aFull, 1);
        mbarInit(aEmpty, 1);
        for (int i = 0; i < kSymMaxStages; i++) {
            mbarInit(full + i, 1);
            mbarInit(empty + i, 1);
        }
        mbarInitFence();
    }
    if (warp == kEpiWarps) tmemAlloc(tmemSlot, 512);
    fenceBefore();
    __syncthreads();
    fenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    const uint32_t items = p.items;
    const uint64_t virtualCols = uint64_t(p.offsetsHere) * kSsTileN;

    if (warp == kEpiWarps) {
        // ===================== TMA producer: A once per item, B per tile =====================
        // Lane 0 issues the loads; the whole warp takes part in the PACING of whole-sweep items: every CTA publishes how
        // far it is (in tiles) and none runs more than kPaceWindow tiles ahead of the slowest.  The B tiles of
        // neighbouring row blocks are the same blocks one step apart, so in step they are read from DRAM once and from
        // L2 147 times; without pacing a CTA that falls behind starts missing L2, gets slower still, and the sweep
        // settles DRAM-bound (1 M cells: 3.25 TB read from DRAM, L2 hit rate 33 %, tensor pipe 47 %).
        uint32_t stage = 0, phase = 0, itemIter = 0;
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x, itemIter++) {
            const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, virtualCols);
            const uint32_t super = it.rowBlock >> 1;
            const uint32_t d0 = p.dBegin + uint32_t(it.colBegin / kSsTileN);
            const uint32_t d1 = p.dBegin + uint32_t((it.colEnd + kSsTileN - 1) / kSsTileN);
            const bool paced = p.progress != nullptr && item < p.mainBlocks;
            if (lane == 0) {
                if (p.progress && !paced) *reinterpret_cast<volatile uint32_t*>(p.progress + blockIdx.x) = 0xffffffffu;
                mbarWait(aEmpty, (itemIter & 1) ^ 1);
                mbarExpectTx(aFull, p.panels * kSsABytes);
                for (uint32_t kc = 0; kc < p.panels; kc++)
                    tmaLoad2d(smA + size_t(kc) * kSsABytes, &mapA, aFull, int32_t(kc * kChunkBytes), int32_t(it.rowBlock * kRowsPerItem));
            }
            for (uint32_t d = d0; d < d1; d++) {
                if (paced && ((d - d0) & 7u) == 0) {
                    const uint32_t vt = itemIter * p.offsetsHere + (d - d0);
                    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(p.progress + blockIdx.x) = vt;
                    for (int spin = 0; spin < 4000; spin++) {           // bounded: pacing is an optimisation, never a dependency
                        uint32_t slowest = 0xffffffffu;
                        for (uint32_t c = lane; c < gridDim.x; c += 32)
                            slowest = min(slowest, *reinterpret_cast<volatile const uint32_t*>(p.progress + c));
                        slowest = __reduce_min_sync(0xffffffffu, slowest);
                        if (slowest >= vt || slowest + kPaceWindow >= vt) break;
                        __nanosleep(500);
                    }
                }
                if (lane == 0) {
                    const int32_t col0 = int32_t(((super + p.superBlocks - d) % p.superBlocks) * kSsTileN);
                    for (uint32_t kc = 0; kc < p.panels; kc++) {
                        mbarWait(empty + stage, phase ^ 1);
                        mbarExpectTx(full + stage, kSsBBytes);
                        tmaLoad2d(ring + size_t(stage) * kSsBBytes, &mapB, full + stage, int32_t(kc * kChunkBytes), col0);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0 && p.progress) *reinterpret_cast<volatile uint32_t*>(p.progress + blockIdx.x) = 0xffffffffu;
    } else if (warp == kEpiWarps + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t tileIter = 0, stage = 0, phase = 0, itemIter = 0;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x, itemIter++) {
                const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, virtualCols);
                const uint32_t d0 = p.dBegin + uint32_t(it.colBegin / kSsTileN);
                const uint32_t d1 = p.dBegin + uint32_t((it.colEnd + kSsTileN - 1) / kSsTileN);
                mbarWait(aFull, itemIter & 1);
                fenceAfter();
                const uint32_t aBase = smemAddr(smA);
                for (uint32_t d = d0; d < d1; d++, tileIter++) {
                    const uint32_t buf = tileIter & 1;
                    mbarWait(accEmpty + buf, ((tileIter >> 1) & 1) ^ 1);
                    fenceAfter();
                    const uint32_t tmemD = tmemBase + buf * kSsTileN;
                    uint32_t accumulate = 0;
                    for (uint32_t kc = 0; kc < p.panels; kc++) {
                        mbarWait(full + stage, phase);
                        fenceAfter();
                        const uint32_t aAddr = aBase + kc * kSsABytes;
                        const uint32_t bAddr = smemAddr(ring + size_t(stage) * kSsBBytes);
#pragma unroll
                        for (int ks = 0; ks < kChunkBytes / kUmmaK; ks++) {
                            mmaI8Ss(tmemD, makeSmemDesc(aAddr + ks * kUmmaK), makeSmemDesc(bAddr + ks * kUmmaK), kInstrDescSs, accumulate);
                            accumulate = 1;
                        }
                        commit(empty + stage);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    commit(accFull + buf);
                }
                commit(aEmpty);      // fires when the item's last MMAs have read A: the producer may overwrite it
            }
        }
    } else {
        // ===================== epilogue: thread == (row == TMEM lane, 128-column sub-stream) =====================
        const uint32_t dotK = p.K;
        const uint32_t N = uint32_t(p.cellCount);
        const uint32_t rowInItem = threadIdx.x & (kRowsPerItem - 1);
        const uint32_t sub = threadIdx.x / kRowsPerItem;
        constexpr int kSubCols = kSsTileN / kSubStreams;
        const uint32_t laneField = uint32_t((warp & 3) * 32) << 16;
        uint32_t tileIter = 0;
        ulonglong2* logNext = nullptr;           // next free entry of the thread's current log chunk
        uint32_t logFill = kLogChunk;            // entries used in it (no chunk yet)
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const ScanItem it = decodeScanItem(item, p.mainBlocks, p.segments, p.segmentCols, virtualCols);
            const uint32_t seg = it.segment;
            const uint32_t super = it.rowBlock >> 1;
            const uint32_t d0 = p.dBegin + uint32_t(it.colBegin / kSsTileN);
            const uint32_t d1 = p.dBegin + uint32_t((it.colEnd + kSsTileN - 1) / kSsTileN);
            const uint32_t rowPos = it.rowBlock * kRowsPerItem + rowInItem;
            const bool valid = rowPos < N;
            const uint32_t rowCell = !valid ? 0xffffffffu : p.perm ? p.perm[rowPos] : rowPos;
            uint32_t* limPtr = p.limEx + (valid ? rowPos : 0);

            RowState st;
            st.rowId = rowPos;            // self test is on positions
            // A stream that continues the previous launch's region keeps its k best so far: the next prune then yields
            // the k-th best of everything the row has seen.  (A fresh region needs ~2k survivors of the old bound
            // before its first prune tightens anything: measured 250 instead of ~65 row-direction survivors per cell.)
            st.count = (p.resume && seg == 0 && valid) ? p.candCount[uint64_t(sub) * N + rowPos] : 0;
            st.appended = 0;
            st.tau = valid ? __ldcg(limPtr) : 0;
            st.lim = st.tau;
            st.buf = p.cand + (uint64_t(seg * kSubStreams + sub) * N + (valid ? rowPos : 0)) * p.cap;
            tauShare[sub * kRowsPerItem + rowInItem] = 0xffffu;      // harmless for any row (see scanMmaKernel)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            int32_t dotThr = int32_t(dotK) - 2 * int32_t(st.lim);      // mismatch < lim  <=>  dot > K - 2 lim

            for (uint32_t d = d0; d < d1; d++, tileIter++) {
                const uint32_t buf = tileIter & 1;
                const uint32_t colSuper = (super + p.superBlocks - d) % p.superBlocks;
                const bool colDir = d != 0 && d != p.halfOffset;          // CTA-uniform
                const int16_t* thr = colThr + buf * kSsTileN + sub * kSubCols;
                const int16_t* grp = grpThr + buf * (kSsTileN / 8) + sub * (kSubCols / 8);
                if (colDir) {
                    // stage the bounds of this sub-stream's 128 columns (the same 128 threads use them)
                    const uint32_t c = threadIdx.x;                       // == sub * 128 + rowInItem
                    const uint32_t pos = colSuper * kSsTileN + c;
                    const uint32_t lim = pos < N ? __ldcg(p.limEx + pos) : 0u;       // padding columns: nothing passes
                    int32_t t = int32_t(dotK) - 2 * int32_t(lim);
                    colThr[buf * kSsTileN + c] = int16_t(t);
                    t = min(t, __shfl_xor_sync(0xffffffffu, t, 1));
                    t = min(t, __shfl_xor_sync(0xffffffffu, t, 2));
                    t = min(t, __shfl_xor_sync(0xffffffffu, t, 4));
                    if ((lane & 7) == 0) grpThr[buf * (kSsTileN / 8) + (c >> 3)] = int16_t(t);
                    asm volatile("bar.sync %0, %1;" ::"r"(2 + sub), "n"(kRowsPerItem) : "memory");
                }
                mbarWait(accFull + buf, (tileIter >> 1) & 1);
                fenceAfter();
                const uint32_t posBase = colSuper * kSsTileN + sub * kSubCols;
                const uint32_t taddr = tmemBase + buf * kSsTileN + sub * kSubCols + laneField;
                {
                    const uint32_t other = tauShare[(sub ^ 1) * kRowsPerItem + rowInItem];
                    if (other < st.lim && !(p.flags & 1)) {
                        st.lim = other;
                        dotThr = int32_t(dotK) - 2 * int32_t(st.lim);
                    }
                }
#pragma unroll 1
                for (int c = 0; c < kSubCols; c += 32) {
                    uint32_t v[32];
                    tmemLoad32(taddr + c, v);
                    tmemLoadWait();
                    if (c + 32 == kSubCols) {          // last chunk is in registers: hand the accumulator back
                        fenceBefore();
                        mbarArrive(accEmpty + buf);
                    }
                    int32_t m[4];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        m[g] = int32_t(v[8 * g]);
#pragma unroll
                        for (int j = 1; j < 8; j++) m[g] = max(m[g], int32_t(v[8 * g + j]));
                    }
                    const int32_t mx = max(max(m[0], m[1]), max(m[2], m[3]));
                    // ---- row direction
                    if (mx > dotThr) {
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            if (m[g] > dotThr) {
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const int32_t dv = int32_t(v[8 * g + j]);
                                    const uint32_t pos = posBase + c + 8 * g + j;
                                    if (dv > dotThr && pos < N && pos != rowPos) {
                                        const uint32_t ham = uint32_t(int32_t(dotK) - dv) >> 1;
                                        st.buf[st.count++] = (uint64_t(ham) << 32) | pos;
                                        st.appended++;
                                    }
                                }
                            }
                        }
                    }
                    // ---- column direction
                    if (colDir && valid) {
                        const short4 gt = *reinterpret_cast<const short4*>(grp + (c >> 3));
                        const int32_t gtv[4] = {gt.x, gt.y, gt.z, gt.w};
                        if (m[0] > gtv[0] || m[1] > gtv[1] || m[2] > gtv[2] || m[3] > gtv[3]) {
#pragma unroll
                            for (int g = 0; g < 4; g++) {
                                if (m[g] > gtv[g]) {
#pragma unroll
                                    for (int j = 0; j < 8; j++) {
                                        const int32_t dv = int32_t(v[8 * g + j]);
                                        if (dv > int32_t(thr[c + 8 * g + j])) {
                                            const uint32_t ham = uint32_t(int32_t(dotK) - dv) >> 1;
                                            if (logFill == kLogChunk) logNext = nextLogChunk(p.chunkFill, p.chunkNext, p.chunkCap, p.colLog, logNext, logFill);
                                            *logNext++ = make_ulonglong2((uint64_t(ham) << 32) | rowCell, posBase + c + 8 * g + j);
                                            logFill++;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, mx > dotThr)) {
                        warpPruneIfNeededAnyOrder(st, p.k, p.cap, limPtr, p.perm);
                        dotThr = int32_t(dotK) - 2 * int32_t(st.lim);
                    }
                }
                if (valid) tauShare[sub * kRowsPerItem + rowInItem] = uint16_t(min(st.tau, 0xffffu));
            }
            if (valid) {
                p.candCount[uint64_t(seg * kSubStreams + sub) * N + rowPos] = st.count;
                if (p.appendedTotal && st.appended) atomicAdd(p.appendedTotal, (unsigned long long)st.appended);
            }
        }
        if (logNext) {
            const uint64_t chunk = uint64_t(logNext - 1 - p.colLog) / kLogChunk;
            if (chunk < p.chunkCap) p.chunkFill[chunk] = logFill;
            else atomicOr(p.overflow, 1u);     // bit 0: the log pool (entries went to the spill chunk)
        }
    }
    fenceBefore();
    __syncthreads();
    if (warp == kEpiWarps) tmemDealloc(tmemBase, 512);
}

// Sample rows of the encoded matrix (every stride-th scan position) as the column operand of the pre-pass, and for
// every scan position its index in the sample (the pre-pass must not count a cell as its own neighbour).
__global__ void sampleGatherKernel(const uint8_t* __restrict__ enc, uint32_t K, uint64_t cellCount, uint32_t stride,
                                   uint32_t sampleCount, uint8_t* __restrict__ out, uint32_t* __restrict__ selfIndex)
{
    const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    const uint32_t perRow = K / 16;
    if (idx < uint64_t(sampleCount) * perRow) {
        const uint64_t r = idx / perRow, o = idx % perRow;
        reinterpret_cast<uint4*>(out + r * K)[o] = reinterpret_cast<const uint4*>(enc + r * stride * uint64_t(K))[o];
    }
    if (idx < cellCount) selfIndex[idx] = (idx % stride == 0 && idx / stride < sampleCount) ? uint32_t(idx / stride) : 0xffffffffu;
}

// Pre-pass result -> limEx: the k-th smallest mismatch count a cell has against the sample, plus one (exclusive
// bound, ties still accepted); tau0 when the sample holds fewer than k admissible cells.  One warp per cell.
__global__ void __launch_bounds__(128)
sampleBoundKernel(uint64_t cellCount, uint32_t streams, uint32_t cap, uint32_t k, const uint64_t* __restrict__ cand,
                  const uint32_t* __restrict__ candCount, uint32_t tau0, uint32_t* __restrict__ limEx)
{
    const uint64_t row = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (row >= cellCount) return;
    uint32_t n = 0;
    for (uint32_t s = 0; s < streams; s++) n += candCount[uint64_t(s) * cellCount + row];
    uint32_t bound = tau0;
    if (n >= k && tau0 > 0) {
        uint32_t lo = 0, hi = tau0 - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            uint32_t c = 0;
            for (uint32_t s = 0; s < streams; s++) {
                const uint32_t cs = candCount[uint64_t(s) * cellCount + row];
                const uint64_t* src = cand + (uint64_t(s) * cellCount + row) * cap;
                for (uint32_t i = lane; i < cs; i += 32) c += (uint32_t(src[i] >> 32) <= mid);
            }
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= k) hi = mid;
            else lo = mid + 1;
        }
        bound = lo + 1;
    }
    if (lane == 0) limEx[row] = bound;
}

// Files the column-direction log into per-cell inboxes of exactly the needed length (count -> exclusive scan -> fill,
// one warp per log chunk): FILL = false counts the entries per column cell, FILL = true writes them at
// inOffset[cell] + (a running cursor per cell).
template <bool FILL>
__global__ void __launch_bounds__(256)
scatterLogKernel(const uint32_t* __restrict__ chunkNext, uint32_t chunkCap, const ulonglong2* __restrict__ log,
                 const uint32_t* __restrict__ chunkFill, uint32_t* __restrict__ inCount, const uint32_t* __restrict__ inOffset,
                 uint64_t* __restrict__ inbox, unsigned long long* __restrict__ appendedTotal)
{
    const uint32_t chunks = min(*chunkNext, chunkCap);
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < chunks; w += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t n = chunkFill[w];
        if (!FILL && lane == 0 && n) atomicAdd(appendedTotal, (unsigned long long)n);      // statistics
        for (uint32_t i = lane; i < n; i += 32) {
            const ulonglong2 e = log[uint64_t(w) * kLogChunk + i];
            const uint32_t pos = uint32_t(e.y);
            const uint32_t slot = atomicAdd(inCount + pos, 1u);
            if (FILL) inbox[uint64_t(inOffset[pos]) + slot] = e.x;
        }
    }
}

// Merge of a cell's row streams and inbox (symmetric scan), one warp per cell.  Pass 1 finds h, the k-th smallest
// mismatch count among the keys below the cell's final bound: their 16-bit counts are staged in shared memory for the
// bisection when they fit, else (a cell with an unusually long inbox) every bisection step re-reads the regions.
// Pass 2 stages the keys with count <= h -- k plus the ties at h -- which are ranked like in finalizeKernel
// (scan_popc.cu).  Stream keys carry scan positions (translated here), inbox keys cell ids.
constexpr int kSymFinalWarps = 4;

__global__ void __launch_bounds__(kSymFinalWarps * 32)
finalizeSymKernel(uint64_t cellCount, uint32_t streams, uint32_t cap, uint32_t k, const uint64_t* __restrict__ cand,
                  const uint32_t* __restrict__ candCount, const uint64_t* __restrict__ inbox, const uint32_t* __restrict__ inOffset,
                  const uint32_t* __restrict__ limEx, const float* __restrict__ lut, em2_pair* __restrict__ pairs,
                  uint32_t* __restrict__ usedCount, const uint32_t* __restrict__ perm, uint32_t* __restrict__ overflow,
                  uint32_t hamsPerWarp, uint32_t keysPerWarp)
{
    extern __shared__ __align__(16) uint64_t skeys[];      // [warps][keysPerWarp] keys, then [warps][hamsPerWarp] uint16
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t row = uint64_t(blockIdx.x) * kSymFinalWarps + warp;
    if (row >= cellCount) return;
    uint64_t* keys = skeys + size_t(warp) * keysPerWarp;
    uint16_t* hams = reinterpret_cast<uint16_t*>(skeys + size_t(kSymFinalWarps) * keysPerWarp) + size_t(warp) * hamsPerWarp;
    const uint32_t lim = limEx[row];
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t* in = inbox + inOffset[row];
    const uint32_t inboxCount = inOffset[row + 1] - inOffset[row];
    uint32_t total = inboxCount;
    for (uint32_t s = 0; s < streams; s++) total += candCount[uint64_t(s) * cellCount + row];
    const bool staged = total <= hamsPerWarp;
    // ---- pass 1
    uint32_t n = 0;
    auto stageHams = [&](const uint64_t* src, uint32_t c) {
        for (uint32_t base = 0; base < c; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t m = i < c ? uint32_t(src[i] >> 32) : 0xffffffffu;
            const bool keep = m < lim;
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (keep && staged) hams[n + __popc(mask & lt)] = uint16_t(m);
            n += __popc(mask);
        }
    };
    for (uint32_t s = 0; s < streams; s++)
        stageHams(cand + (uint64_t(s) * cellCount + row) * cap, candCount[uint64_t(s) * cellCount + row]);
    stageHams(in, inboxCount);
    __syncwarp();
    auto countAtMost = [&](uint32_t mid) {
        uint32_t c = 0;
        if (staged) {
            for (uint32_t e = lane; e < n; e += 32) c += (hams[e] <= mid);
        } else {
            for (uint32_t s = 0; s < streams; s++) {
                const uint64_t* src = cand + (uint64_t(s) * cellCount + row) * cap;
                const uint32_t cs = candCount[uint64_t(s) * cellCount + row];
                for (uint32_t i = lane; i < cs; i += 32) c += (uint32_t(src[i] >> 32) <= mid);
            }
            for (uint32_t i = lane; i < inboxCount; i += 32) c += (uint32_t(in[i] >> 32) <= mid);
        }
        return __reduce_add_sync(0xffffffffu, c);
    };
    uint32_t h = lim;                         // fewer than k keys below the bound: all of them
    if (n > k) {
        uint32_t lo = 0, hi = lim - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (countAtMost(mid) >= k) hi = mid;
            else lo = mid + 1;
        }
        h = lo;
    }
    // ---- pass 2: the keys with count <= h.  Normally all ties at h are staged and the ranking below picks the
    // smallest ids; a cell with more ties than the staging holds (hundreds of identical cells) first finds the id of
    // the last tie that still fits by bisection over the regions themselves.
    uint32_t idCut = 0xffffffffu;
    if (n > k) {
        const uint32_t less = h ? countAtMost(h - 1) : 0;
        const uint32_t ties = countAtMost(h) - less;
        if (less + ties > keysPerWarp) {
            const uint32_t r = k - less;              // ties that still fit (>= 1 by the choice of h)
            auto tiesUpTo = [&](uint32_t id) {
                uint32_t c = 0;
                for (uint32_t s = 0; s < streams; s++) {
                    const uint64_t* src = cand + (uint64_t(s) * cellCount + row) * cap;
                    const uint32_t cs = candCount[uint64_t(s) * cellCount + row];
                    for (uint32_t i = lane; i < cs; i += 32) {
                        const uint64_t key = src[i];
                        if (uint32_t(key >> 32) == h) c += ((perm ? perm[uint32_t(key)] : uint32_t(key)) <= id);
                    }
                }
                for (uint32_t i = lane; i < inboxCount; i += 32) c += (uint32_t(in[i] >> 32) == h && uint32_t(in[i]) <= id);
                return __reduce_add_sync(0xffffffffu, c);
            };
            uint32_t a = 0, b = 0xffffffffu;
            while (a < b) {
                const uint32_t mid = a + ((b - a) >> 1);
                if (tiesUpTo(mid) >= r) b = mid;
                else a = mid + 1;
            }
            idCut = a;                                // ids are unique: exactly r ties have id <= idCut
        }
    }
    n = 0;
    bool over = false;
    auto stageKeys = [&](const uint64_t* src, uint32_t c, bool positions) {
        for (uint32_t base = 0; base < c && !over; base += 32) {
            const uint32_t i = base + lane;
            uint64_t key = i < c ? src[i] : ~0ull;
            const uint32_t m = uint32_t(key >> 32);
            bool keep = i < c && m < lim && m <= h;
            if (keep && positions && perm) key = (key & 0xffffffff00000000ull) | perm[uint32_t(key)];
            if (keep && m == h && uint32_t(key) > idCut) keep = false;
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (n + __popc(mask) > keysPerWarp) {
                over = true;
                break;
            }
            if (keep) keys[n + __popc(mask & lt)] = key;
            n += __popc(mask);
        }
    };
    for (uint32_t s = 0; s < streams; s++)
        stageKeys(cand + (uint64_t(s) * cellCount + row) * cap, candCount[uint64_t(s) * cellCount + row], true);
    stageKeys(in, inboxCount, false);
    if (over) {
        if (lane == 0) atomicOr(overflow, 4u);      // bit 2: cannot happen (k <= keysPerWarp); kept as a guard
        return;
    }
    __syncwarp();
    const uint32_t used = n < k ? n : k;
    const uint64_t outRow = perm ? uint64_t(perm[row]) : row;
    for (uint32_t e = lane; e < n; e += 32) {
        const uint64_t key = keys[e];
        uint32_t rank = 0;
        for (uint32_t f = 0; f < n; f++) rank += (keys[f] < key);      // keys are unique (ids are)
        if (rank < k) {
            em2_pair pr;
            pr.cell = uint32_t(key);
            pr.similarity = lut[uint32_t(key >> 32)];
            pairs[outRow * k + rank] = pr;
        }
    }
    for (uint32_t i = used + lane; i < k; i += 32) {
        em2_pair z;
        z.cell = 0;
        z.similarity = 0.f;
        pairs[outRow * k + i] = z;
    }
    if (lane == 0) usedCount[outRow] = used;
}

// ---------------------------------------------------------------------------------------------------------
// Row grouping.  A warp of the epilogue serves 32 query rows; when those rows are unrelated, almost every
// 32-column chunk holds a passing column for SOME lane and the selection code runs with two or three lanes
// active (ncu: ~90 % of the chunks, 10 of 32 threads per instruction on clustered data).  Rows that are
// similar to each other pass on the SAME columns, so putting similar rows into the same warp makes most chunks
// miss for the whole warp and the rest hit with most lanes active.  Only the ORDER IN WHICH ROWS ARE SCANNED
// changes: columns stay in cell-id order (the tie-break of topk.cuh needs that), every row still sees every
// column, results are identical.  Grouping = nearest of 256 pivot cells by Hamming distance on the first
// <= 512 bits, then a radix sort of (pivot, cell id).
// ---------------------------------------------------------------------------------------------------------
constexpr int kPivots = 256;
constexpr int kPivotWords = 8;        // 512 bits are plenty to tell clusters apart; halves the assignment pass

__global__ void __launch_bounds__(256)
pivotAssignKernel(const uint64_t* __restrict__ sig, uint32_t W, uint64_t cellCount, uint64_t rowBegin, uint64_t rows,
                  unsigned long long* __restrict__ keys)
{
    __shared__ uint64_t piv[kPivots][kPivotWords];
    const uint32_t wp = W < kPivotWords ? W : kPivotWords;
    for (uint32_t i = threadIdx.x; i < kPivots * wp; i += blockDim.x) {
        const uint32_t pv = i / wp, w = i % wp;
        const uint64_t cell = uint64_t(pv) * cellCount / kPivots;
        piv[pv][w] = sig[cell * W + w];
    }
    __syncthreads();
    const uint64_t q = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (q >= rows) return;
    uint64_t x[kPivotWords];
#pragma unroll
    for (int w = 0; w < kPivotWords; w++) x[w] = uint32_t(w) < wp ? sig[(rowBegin + q) * W + w] : 0;
    uint32_t best = 0xffffffffu, bestPivot = 0;
    for (int pv = 0; pv < kPivots; pv++) {
        uint32_t d = 0;
#pragma unroll
        for (int w = 0; w < kPivotWords; w++)
            if (uint32_t(w) < wp) d += __popcll(x[w] ^ piv[pv][w]);
        if (d < best) {
            best = d;
            bestPivot = pv;
        }
    }
    keys[q] = (uint64_t(bestPivot) << 32) | uint32_t(rowBegin + q);
}

__global__ void keysToPermKernel(const unsigned long long* __restrict__ keys, uint64_t rows, uint32_t* __restrict__ perm)
{
    const uint64_t q = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (q < rows) perm[q] = uint32_t(keys[q]);
}

// +-1 int8 expansion of the packed signatures: E[n][p] = bit p set ? +1 : -1, p < K; bits at and
// beyond lshCount (zero in the packed words, or beyond them) encode as -1 in every row.
__global__ void encodeKernel(const uint64_t* __restrict__ sig, uint32_t W, uint64_t cellCount, uint32_t K,
                             uint8_t* __restrict__ enc, const uint32_t* __restrict__ index)
{
    const uint32_t groupsPerRow = K / 16;      // 16 bits -> 16 bytes per thread
    const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (idx >= cellCount * groupsPerRow) return;
    const uint64_t row = idx / groupsPerRow;
    const uint32_t g = uint32_t(idx - row * groupsPerRow);
    const uint32_t w = g >> 2;
    uint32_t bits = 0;
    const uint64_t srcRow = index ? uint64_t(index[row]) : row;      // output row `row` = signature of cell index[row]
    if (w < W) bits = uint32_t(sig[srcRow * W + w] >> (48 - 16 * (g & 3))) & 0xFFFFu;   // MSB-first
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

// Symmetric scan of the whole matrix; encP = encoded signatures in scan-position order (rows and columns),
// perm = position -> cell id (nullptr: identity).  *overflowed != 0 means the log pool ran dry (bit 0) and NOTHING
// that was written may be used: rerun one-directionally.
int runSymmetric(em2_context* ctx, const uint8_t* encP, const uint32_t* perm, uint64_t cellCount, uint32_t K, uint64_t k,
                 uint32_t tau0, const float* lut, em2_pair* pairs, uint32_t* usedCount, cudaStream_t s, int* overflowed)
{
    const uint64_t N = cellCount;
    const uint32_t panels = K / kChunkBytes;
    // ---- sampling pre-pass: every cell against M ~ N/32 sample cells with the one-directional kernel
    const uint32_t M = uint32_t(std::min<uint64_t>(N, std::max<uint64_t>(256, roundUp(N / 32, kSsTileN))));
    const uint32_t stride = uint32_t(N / M);
    ScanPlan pre = makeScanPlan(ctx, N, M, k, kSsTileN, kRowsPerItem, 1, kSubStreams);
    const uint32_t superBlocks = uint32_t((N + kSsTileN - 1) / kSsTileN);
    const uint32_t offsets = superBlocks / 2 + 1;
    // Two launches: the diagonal tiles first, for ALL row blocks (row direction only).  In grouped order a cell's
    // nearest neighbours sit in its own super block, so after this short launch every cell's published bound is already
    // tight on clustered data (the sample bound is the tight one on unstructured data) -- before any CTA of the long
    // second launch compares the cell with its rows.
    const uint32_t nearOffsets = 1;
    ScanPlan nearPlan = makeScanPlan(ctx, N, uint64_t(nearOffsets) * kSsTileN, k, kSsTileN, kRowsPerItem, 1, kSubStreams);
    ScanPlan plan = nearPlan;
    if (offsets > nearOffsets)
        plan = makeScanPlan(ctx, N, uint64_t(offsets - nearOffsets) * kSsTileN, k, kSsTileN, kRowsPerItem, 1, kSubStreams);
    // small regions here: a cell's bound is published when its region is pruned, and the column direction of other CTAs
    // lives on fresh bounds (with the one-directional kernels' 4k + 32 keys: 90 M instead of 70 M survivors at config 2)
    nearPlan.cap = plan.cap = pre.cap = candidateCapacity(uint32_t(k)) + uint32_t(k) * uint32_t(ctx->candCapExtra);
    // stream pair 0 of a row is shared by the diagonal launch and segment 0 of the second launch (which continues it)
    const uint32_t streams = std::max(nearPlan.segments, plan.segments) * kSubStreams;
    const uint32_t maxSegments = std::max(pre.segments, streams / kSubStreams);

    void *cand = nullptr, *candCount = nullptr, *counters = nullptr, *sample = nullptr, *sym = nullptr, *inbox = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_CAND, size_t(maxSegments) * kSubStreams * N * plan.cap * sizeof(uint64_t), &cand));
    EM2_TRY(reserve(ctx, em2_context::S_CANDCOUNT, size_t(maxSegments) * kSubStreams * N * sizeof(uint32_t), &candCount));
    EM2_TRY(reserve(ctx, em2_context::S_COUNTERS, 64, &counters));
    EM2_TRY(reserve(ctx, em2_context::S_SAMPLE, size_t(M) * K, &sample));
    // [limEx N][inCount N][selfIndex N][inOffset N + 1][overflow 1][pad][progress 1024]
    EM2_TRY(reserve(ctx, em2_context::S_SYM, (4 * N + 8 + 1024) * sizeof(uint32_t), &sym));
    // column-direction log pool: 24 k entries per cell (measured: 1.5-12 k per cell on clustered data); the inboxes are
    // cut from a buffer of the same number of keys (count -> scan -> fill)
    const uint64_t poolEntries = std::min<uint64_t>(0xf0000000ull, 24 * N * k + 2 * uint64_t(ctx->smCount) * kEpiThreads * kLogChunk);
    const uint32_t chunkCap = uint32_t(poolEntries / kLogChunk);
    EM2_TRY(reserve(ctx, em2_context::S_INBOX, size_t(chunkCap) * kLogChunk * sizeof(uint64_t), &inbox));
    void *colLog = nullptr, *chunkFill = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_COLLOG, (size_t(chunkCap) + 1) * kLogChunk * sizeof(ulonglong2), &colLog));   // + spill chunk
    EM2_TRY(reserve(ctx, em2_context::S_COLLOGFILL, (size_t(chunkCap) + 4) * sizeof(uint32_t), &chunkFill));
    uint32_t* chunkNext = static_cast<uint32_t*>(chunkFill) + chunkCap;
    EM2_CUDA(ctx, cudaMemsetAsync(chunkFill, 0, (size_t(chunkCap) + 4) * sizeof(uint32_t), s));
    // debug_flags bit 3: report the candidate counters after every stage (synchronises)
    auto report = [&](const char* what) {
        if (!(ctx->debugFlags & 8)) return;
        unsigned long long v = 0;
        uint32_t chunksUsed = 0;
        cudaStreamSynchronize(s);
        cudaMemcpy(&v, static_cast<unsigned long long*>(counters) + 1, sizeof(v), cudaMemcpyDeviceToHost);
        cudaMemcpy(&chunksUsed, chunkNext, sizeof(chunksUsed), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "[em2 sym] after %s: row-direction candidates %llu, column-direction log chunks %u of %u\n", what, v,
                     chunksUsed, chunkCap);
    };
    uint32_t* limEx = static_cast<uint32_t*>(sym);
    uint32_t* inCount = limEx + N;
    uint32_t* selfIndex = inCount + N;
    uint32_t* inOffset = selfIndex + N;
    uint32_t* overflow = inOffset + N + 1;
    uint32_t* progress = overflow + 4;
    EM2_CUDA(ctx, cudaMemsetAsync(inCount, 0, N * sizeof(uint32_t), s));
    EM2_CUDA(ctx, cudaMemsetAsync(overflow, 0, sizeof(uint32_t), s));
    {
        const uint64_t threads = std::max<uint64_t>(uint64_t(M) * (K / 16), N);
        sampleGatherKernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>(encP, K, N, stride, M, static_cast<uint8_t*>(sample), selfIndex);
        EM2_CUDA(ctx, cudaGetLastError());
    }
    CUtensorMap mapA, mapB, mapS;
    EM2_TRY(makeTensorMapU8(ctx, &mapA, encP, N, K, K, kRowsPerItem));
    EM2_TRY(makeTensorMapU8(ctx, &mapB, encP, N, K, K, kSsTileN));
    EM2_TRY(makeTensorMapU8(ctx, &mapS, sample, M, K, K, kSsTileN));
    {
        MmaParams q{};
        q.cellCount = M;
        q.rowBegin = 0;
        q.rows = N;
        q.K = K;
        q.panels = panels;
        q.mainBlocks = pre.mainBlocks;
        q.segments = pre.segments;
        q.items = pre.items;
        q.segmentCols = pre.segmentCols;
        q.k = uint32_t(k);
        q.cap = pre.cap;
        q.tau0 = tau0;
        q.cand = static_cast<uint64_t*>(cand);
        q.candCount = static_cast<uint32_t*>(candCount);
        q.appendedTotal = static_cast<unsigned long long*>(counters) + 1;
        q.flags = uint32_t(ctx->debugFlags);
        q.rowPerm = selfIndex;
        if (pre.segments > 1)
            EM2_CUDA(ctx, cudaMemsetAsync(candCount, 0, size_t(pre.segments) * kSubStreams * N * sizeof(uint32_t), s));
        const size_t smem = 1024 + size_t(kSsStages) * kSsStageBytes + 256 + kShareBytes;
        EM2_CUDA(ctx, cudaFuncSetAttribute(scanMmaSsKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        scanMmaSsKernel<false><<<unsigned(std::min<uint32_t>(pre.items, uint32_t(ctx->smCount))), kThreads, smem, s>>>(mapA, mapS, q);
        EM2_CUDA(ctx, cudaGetLastError());
        sampleBoundKernel<<<unsigned((N * 32 + 127) / 128), 128, 0, s>>>(N, pre.segments * kSubStreams, pre.cap, uint32_t(k), q.cand,
                                                                         q.candCount, tau0, limEx);
        EM2_CUDA(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 3;
        report("pre-pass");
    }
    // ---- the symmetric sweep
    SymParams p{};
    p.cellCount = N;
    p.K = K;
    p.panels = panels;
    const size_t budget = 227 * 1024 - 1024 - kSymSmallBytes - size_t(panels) * kSsABytes;
    p.stages = uint32_t(std::min<size_t>(kSymMaxStages, budget / kSsBBytes));
    p.superBlocks = superBlocks;
    p.offsets = offsets;
    p.halfOffset = superBlocks % 2 == 0 ? superBlocks / 2 : 0;
    p.k = uint32_t(k);
    p.cap = plan.cap;
    p.cand = static_cast<uint64_t*>(cand);
    p.candCount = static_cast<uint32_t*>(candCount);
    p.appendedTotal = static_cast<unsigned long long*>(counters) + 1;
    p.limEx = limEx;
    p.colLog = static_cast<ulonglong2*>(colLog);
    p.chunkFill = static_cast<uint32_t*>(chunkFill);
    p.chunkNext = chunkNext;
    p.chunkCap = chunkCap;
    p.overflow = overflow;
    p.perm = perm;
    p.flags = uint32_t(ctx->debugFlags);
    EM2_CUDA(ctx, cudaMemsetAsync(candCount, 0, size_t(streams) * N * sizeof(uint32_t), s));   // untouched streams read as empty
    {
        const size_t smem = 1024 + size_t(panels) * kSsABytes + size_t(p.stages) * kSsBBytes + kSymSmallBytes;
        EM2_CUDA(ctx, cudaFuncSetAttribute(scanMmaSymKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        auto sweep = [&](const ScanPlan& pl, uint32_t dBegin, uint32_t count, uint32_t resume) -> int {
            p.mainBlocks = pl.mainBlocks;
            p.segments = pl.segments;
            p.items = pl.items;
            p.segmentCols = pl.segmentCols;
            p.dBegin = dBegin;
            p.offsetsHere = count;
            p.resume = resume;
            // only sweeps long enough for the CTAs to drift apart by more blocks than L2 holds (~490): paced at 200k cells
            // (390 tiles per item) the sweep was 1.5x SLOWER, at 1 M cells (1953 tiles) 1.2x faster
            p.progress = (count >= 768 && !(ctx->debugFlags & 16)) ? progress : nullptr;
            if (p.progress) EM2_CUDA(ctx, cudaMemsetAsync(progress, 0, 1024 * sizeof(uint32_t), s));
            scanMmaSymKernel<<<unsigned(std::min<uint32_t>(pl.items, uint32_t(ctx->smCount))), kThreads, smem, s>>>(mapA, mapB, p);
            EM2_CUDA(ctx, cudaGetLastError());
            ctx->stats.kernel_launches++;
            return EM2_OK;
        };
        EM2_TRY(sweep(nearPlan, 0, nearOffsets, 0));
        report("diagonal launch");
        if (offsets > nearOffsets) EM2_TRY(sweep(plan, nearOffsets, offsets - nearOffsets, 1));
        report("main sweep");
        // inboxes: count the log entries per column cell, exclusive scan, fill
        scatterLogKernel<false><<<unsigned(ctx->smCount) * 8, 256, 0, s>>>(chunkNext, chunkCap, p.colLog, p.chunkFill, inCount, nullptr,
                                                                          nullptr, p.appendedTotal);
        EM2_CUDA(ctx, cudaGetLastError());
        {
            size_t cubBytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, cubBytes, inCount, inOffset, int(N + 1), s);
            void* cubTemp = nullptr;
            EM2_TRY(reserve(ctx, em2_context::S_MISC, cubBytes, &cubTemp));
            // inCount[N] is selfIndex[0]: only the first N + 1 outputs' prefix property matters, inOffset[N] = total
            EM2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cubTemp, cubBytes, inCount, inOffset, int(N + 1), s));
        }
        EM2_CUDA(ctx, cudaMemsetAsync(inCount, 0, N * sizeof(uint32_t), s));
        scatterLogKernel<true><<<unsigned(ctx->smCount) * 8, 256, 0, s>>>(chunkNext, chunkCap, p.colLog, p.chunkFill, inCount, inOffset,
                                                                         static_cast<uint64_t*>(inbox), p.appendedTotal);
        EM2_CUDA(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 2;      // + the scan's own kernels (library code, not counted)
        // staging: 16-bit mismatch counts of everything below the bound (cells with longer lists bisect in place), then
        // the k best keys plus the ties at the k-th place
        const uint32_t hamsPerWarp = uint32_t(roundUp(uint64_t(streams) * plan.cap + 40 * k + 256, 64));
        const uint32_t keysPerWarp = uint32_t(roundUp(2 * k + 256, 64));
        const size_t smemF = size_t(kSymFinalWarps) * (size_t(keysPerWarp) * sizeof(uint64_t) + size_t(hamsPerWarp) * sizeof(uint16_t));
        if (smemF > 200 * 1024) return fail(ctx, EM2_ERR_INVALID, "k too large for the symmetric finalize kernel");
        EM2_CUDA(ctx, cudaFuncSetAttribute(finalizeSymKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemF)));
        finalizeSymKernel<<<unsigned((N + kSymFinalWarps - 1) / kSymFinalWarps), kSymFinalWarps * 32, smemF, s>>>(
            N, streams, plan.cap, uint32_t(k), p.cand, p.candCount, static_cast<const uint64_t*>(inbox), inOffset, limEx, lut, pairs,
            usedCount, perm, overflow, hamsPerWarp, keysPerWarp);
        EM2_CUDA(ctx, cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    // the overflow flag decides whether the result stands
    uint32_t* flagHost = nullptr;
    {
        void* pin = nullptr;
        EM2_TRY(reservePinned(ctx, 2, 64, &pin));
        flagHost = static_cast<uint32_t*>(pin);
    }
    EM2_CUDA(ctx, cudaMemcpyAsync(flagHost, overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    EM2_CUDA(ctx, cudaStreamSynchronize(s));
    *overflowed = int(*flagHost);
    return EM2_OK;
}

int runMma(em2_context* ctx, const uint64_t* signatures, uint64_t cellCount, uint64_t lshCount, uint64_t rowBegin,
           uint64_t rowEnd, uint64_t k, int64_t mismatchMax, const float* lut, em2_pair* pairs, uint32_t* usedCount,
           uint16_t* dump, cudaStream_t s)
{
    const uint64_t rows = rowEnd - rowBegin;
    const uint32_t W = uint32_t(wordCount(lshCount));
    const uint32_t K = uint32_t(roundUp(lshCount, kChunkBytes));
    // Which tcgen05 kernel: above 1024 bits the A operand does not fit in tensor memory, so both operands stream
    // (scanMmaSsKernel, N = 256 instructions).  Measured at 512..1024 bits the streamed kernel is also the faster one
    // (3500 vs 2480 TOP/s on iid signatures at L = 1024: N = 128 instructions with A in TMEM top out at 2971), so it
    // is the default from 512 bits; "mma_kernel" = 1 / 2 forces the TMEM-resident / the streamed form.
    const bool streamed = K > kMaxPanels * kChunkBytes || ctx->mmaKernel == 2 || (ctx->mmaKernel == 0 && K >= 512);
    if (cellCount > 0x7fffff00ull) return fail(ctx, EM2_ERR_INVALID, "cellCount too large for the MMA variant");

    // 1. encode
    void* enc = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_ENC, cellCount * K, &enc));
    {
        const uint64_t threads = cellCount * (K / 16);
        encodeKernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>(signatures, W, cellCount, K, static_cast<uint8_t*>(enc), nullptr);
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());
    }

    // 1b. the scanned rows in grouped order (see pivotAssignKernel) and their encoded signatures by scan position
    const bool grouped = !dump && (ctx->rowGrouping == 2 || (ctx->rowGrouping == 0 && rows >= 8192));
    const uint8_t* encRows = static_cast<const uint8_t*>(enc) + rowBegin * uint64_t(K);
    const uint32_t* rowPerm = nullptr;
    if (grouped) {
        size_t cubBytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, cubBytes, static_cast<const unsigned long long*>(nullptr),
                                       static_cast<unsigned long long*>(nullptr), int(rows), 0, 40, s);
        const size_t keyBytes = roundUp(rows * sizeof(unsigned long long), 256);
        void* scratch = nullptr;
        EM2_TRY(reserve(ctx, em2_context::S_ROWPERM, 2 * keyBytes + roundUp(rows * 4, 256) + cubBytes, &scratch));
        auto* keysIn = static_cast<unsigned long long*>(scratch);
        auto* keysOut = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(scratch) + keyBytes);
        auto* perm = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(scratch) + 2 * keyBytes);
        void* cubTemp = static_cast<uint8_t*>(scratch) + 2 * keyBytes + roundUp(rows * 4, 256);
        pivotAssignKernel<<<unsigned((rows + 255) / 256), 256, 0, s>>>(signatures, W, cellCount, rowBegin, rows, keysIn);
        EM2_CUDA(ctx, cudaGetLastError());
        EM2_CUDA(ctx, cub::DeviceRadixSort::SortKeys(cubTemp, cubBytes, keysIn, keysOut, int(rows), 0, 40, s));
        keysToPermKernel<<<unsigned((rows + 255) / 256), 256, 0, s>>>(keysOut, rows, perm);
        EM2_CUDA(ctx, cudaGetLastError());
        void* er = nullptr;
        EM2_TRY(reserve(ctx, em2_context::S_ENCROWS, rows * uint64_t(K), &er));
        const uint64_t threads = rows * (K / 16);
        encodeKernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>(signatures, W, rows, K, static_cast<uint8_t*>(er), perm);
        EM2_CUDA(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 4;      // + the radix sort's own passes (library code, not counted)
        encRows = static_cast<const uint8_t*>(er);
        rowPerm = perm;
    }

    // 1c. whole-matrix jobs, on request: every unordered pair once (scanMmaSymKernel); falls through to the
    //     one-directional kernels if a capacity ran out (nothing of the symmetric attempt is kept)
    const uint32_t tau0 = mismatchMax < 0 ? 0u : uint32_t(std::min<int64_t>(mismatchMax, int64_t(lshCount)) + 1);
    ctx->stats.scan_symmetric = 0;
    {
        const uint32_t capSym = scanCandidateCapacity(uint32_t(k), uint32_t(ctx->candCapExtra));
        const bool eligible = streamed && !dump && K <= kMaxPanels * kChunkBytes && rowBegin == 0 && rows == cellCount &&
                              capSym <= 32 * kPruneRegsPerLane && cellCount >= 1024 && tau0 > 0;
        // "scan_symmetric" = 2 asks for the symmetric kernel whenever it is eligible.  AUTO takes it for 400k..2M cells:
        // there the one-directional sweep sits at the power cap and the paced symmetric sweep is 1.3x faster (1 M clustered
        // cells: 570 vs 756 ms); around 100k densely clustered cells it is 5-15 % slower -- the selection epilogue, not
        // the MMA pipe, is what the sweep waits for -- and above 2 M cells its logs outgrow the memory.
        bool automatic = ctx->scanSymmetric == 0 && cellCount >= 400000 && cellCount <= 2000000;
        if (automatic) {
            // its scratch (log pool 16 B + inbox 8 B per entry, 24 k entries per cell; candidate regions; sample) must fit
            // beside what is already allocated -- otherwise stay with the one-directional kernels instead of failing
            const size_t want = size_t(cellCount) * (24 * k * 24 + 8 * (2 * k + 32) * 8 + 64) + (size_t(1) << 30);
            size_t have = ctx->scratch[em2_context::S_COLLOG].bytes + ctx->scratch[em2_context::S_INBOX].bytes +
                          ctx->scratch[em2_context::S_CAND].bytes;
            size_t freeBytes = 0, totalBytes = 0;
            if (cudaMemGetInfo(&freeBytes, &totalBytes) != cudaSuccess) freeBytes = 0;
            automatic = want <= have + freeBytes / 10 * 9;
        }
        if (eligible && (ctx->scanSymmetric == 2 || automatic)) {
            int overflowed = 0;
            EM2_TRY(runSymmetric(ctx, encRows, rowPerm, cellCount, K, k, tau0, lut, pairs, usedCount, s, &overflowed));
            ctx->stats.scan_symmetric = overflowed ? 2 + 16 * overflowed : 1;      // 2 + 16 * (which capacity ran out)
            if (!overflowed) return EM2_OK;
        }
    }

    // 2. plan + scratch
    MmaParams p{};
    const uint32_t panels = K / kChunkBytes;
    p.stages = kStages;
    const bool pair = ctx->mmaCtaPair != 0 && !streamed;
    ScanPlan plan = streamed ? makeScanPlan(ctx, rows, cellCount, dump ? 1 : k, kSsTileN, kRowsPerItem, 1, kSubStreams) :
                    pair ? makeScanPlan(ctx, rows, cellCount, dump ? 1 : k, kTileN, kPairRows, 1, kSubStreams,
                                        uint32_t(ctx->smCount / 2))
                         : makeScanPlan(ctx, rows, cellCount, dump ? 1 : k, kTileN, kRowsPerItem, 1, kSubStreams);
    if (dump) {
        plan.mainBlocks = plan.rowBlocks;
        plan.segments = 1;
        plan.segmentCols = roundUp(cellCount, kSsTileN);
        plan.items = plan.rowBlocks;
    }
    void* cand = nullptr;
    void* candCount = nullptr;
    void* counters = nullptr;
    if (!dump) {
        EM2_TRY(reserve(ctx, em2_context::S_CAND, size_t(plan.segments) * kSubStreams * rows * plan.cap * sizeof(uint64_t), &cand));
        EM2_TRY(reserve(ctx, em2_context::S_CANDCOUNT, size_t(plan.segments) * kSubStreams * rows * sizeof(uint32_t), &candCount));
        if (plan.segments > 1)   // streams a main row block never touches must read as empty
            EM2_CUDA(ctx, cudaMemsetAsync(candCount, 0, size_t(plan.segments) * kSubStreams * rows * sizeof(uint32_t), s));
    }
    EM2_TRY(reserve(ctx, em2_context::S_COUNTERS, 64, &counters));
    p.cellCount = cellCount;
    p.rowBegin = rowBegin;
    p.rows = rows;
    p.K = K;
    p.panels = panels;
    p.mainBlocks = plan.mainBlocks;
    p.segments = plan.segments;
    p.items = plan.items;
    p.segmentCols = plan.segmentCols;
    p.k = uint32_t(k);
    p.cap = plan.cap;
    p.tau0 = tau0;
    p.cand = static_cast<uint64_t*>(cand);
    p.candCount = static_cast<uint32_t*>(candCount);
    p.appendedTotal = static_cast<unsigned long long*>(counters) + 1;
    p.dump = dump;
    p.flags = uint32_t(ctx->debugFlags);
    p.encRows = encRows;
    p.rowPerm = rowPerm;

    if (streamed) {
        CUtensorMap mapA, mapB;
        EM2_TRY(makeTensorMapU8(ctx, &mapA, encRows, rows, K, K, kRowsPerItem));
        EM2_TRY(makeTensorMapU8(ctx, &mapB, enc, cellCount, K, K, kSsTileN));
        const size_t smem = 1024 + size_t(kSsStages) * kSsStageBytes + 256 + kShareBytes;
        const unsigned grid = unsigned(std::min<uint32_t>(plan.items, uint32_t(ctx->smCount)));
        if (dump) {
            EM2_CUDA(ctx, cudaFuncSetAttribute(scanMmaSsKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            scanMmaSsKernel<true><<<grid, kThreads, smem, s>>>(mapA, mapB, p);
        } else {
            EM2_CUDA(ctx, cudaFuncSetAttribute(scanMmaSsKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            scanMmaSsKernel<false><<<grid, kThreads, smem, s>>>(mapA, mapB, p);
        }
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());
        if (dump) return EM2_OK;
        ScanPlan merged = plan;
        merged.segments = plan.segments * kSubStreams;
        return launchFinalize(ctx, merged, rows, k, p.cand, p.candCount, lut, pairs, usedCount, s, rowPerm, rowBegin);
    }

    CUtensorMap mapB;
    EM2_TRY(makeTensorMapU8(ctx, &mapB, enc, cellCount, K, K, pair ? kTileN / 2 : kTileN));

    const size_t smem = 1024 + size_t(kStages) * kStageBytes + 512 + kShareBytes;
    static_assert(size_t(kStages) * kStageBytes == size_t(kPairStages) * kPairStageBytes, "both variants use the same ring bytes");
    const uint32_t items = plan.items;
    auto go = [&](auto kernel) -> int {
        EM2_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        if (pair) {
            // clusters of two CTAs (one TPC); as many pairs as the device can keep resident
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(2 * unsigned(ctx->smCount / 2));
            int clusters = 0;
            EM2_CUDA(ctx, cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg));
            if (clusters < 1) return fail(ctx, EM2_ERR_CUDA, "no CTA pair of the MMA scan kernel fits on this device");
            cfg.gridDim = dim3(2 * std::min<unsigned>(unsigned(clusters), std::min<unsigned>(items, unsigned(ctx->smCount / 2))));
        } else {
            cfg.gridDim = dim3(std::min<uint32_t>(items, uint32_t(ctx->smCount)));
        }
        EM2_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, mapB, static_cast<const uint8_t*>(enc), p));
        return EM2_OK;
    };
    if (pair) {
        if (dump) EM2_TRY(go(scanMmaKernel<true, true>));
        else EM2_TRY(go(scanMmaKernel<false, true>));
    } else {
        if (dump) EM2_TRY(go(scanMmaKernel<true, false>));
        else EM2_TRY(go(scanMmaKernel<false, false>));
    }
    ctx->stats.kernel_launches++;
    EM2_CUDA(ctx, cudaGetLastError());
    if (dump) return EM2_OK;
    ScanPlan merged = plan;                    // every (segment, sub-stream) buffer is one list to merge
    merged.segments = plan.segments * kSubStreams;
    return launchFinalize(ctx, merged, rows, k, p.cand, p.candCount, lut, pairs, usedCount, s, rowPerm, rowBegin);
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
