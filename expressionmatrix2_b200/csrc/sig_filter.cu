// Signature construction, tensor-core FILTER path (sm_100a): the sign of every projection is decided by an
// exact-integer tcgen05 GEMM with a rigorous error bound; the few projections the bound cannot decide are
// recomputed with the reference's own FP64 operation sequence.  The resulting signature bits are identical
// to Lsh::computeCellLshSignatures (reference src/Lsh.cpp:160-207) -- this path only changes HOW MUCH FP64
// work is needed, never the answer.
//
// Why: the FP64 kernel (signatures.cu) needs one 8-byte hyperplane element per multiply-add and is bound by
// the L1 data pipe.  Expression counts are small non-negative integers (UMI counts), so
//     s_i = sum_g c_g U[g][i] - mean * sumU_i
// can be bracketed as follows.  Quantise every hyperplane column to 22-bit fixed point,
//     U[g][i] = scale_i * q[g][i] + eps,   |eps| <= scale_i / 2,   q = (128 d2 + d1) * 128 + d0,  d* int8,
// expand the cell's counts to a dense uint8 row, and let the tensor cores compute the three EXACT integer
// sums  S_j = sum_g c_g d_j[g][i]  (kind::i8, s32 accumulators).  Then
//     s~_i = scale_i * ((128 S_2 + S_1) * 128 + S_0) - mean * sumU_i
// differs from the value the reference computes by at most
//     B_i = sum1 * scale_i / 2                                   (quantisation, sum1 = sum_g c_g)
//         + (nnz + 16) * 2^-52 * (sum1 * max_g|U[g][i]| + |mean * sumU_i|)   (FP64 rounding of both sides)
// so whenever |s~_i| > B_i the reference's sign is the sign of s~_i.  Otherwise (a few per 100,000
// projections at 30k genes, 5 % density; 0.35 % with two digits, which is why there are three) the pair
// (cell, i) goes to a list and `fixupKernel` evaluates the
// reference's sequence  s = (-mean) * sumU_i;  s = s + double(c) * U[g][i]  in stored gene order with
// __dmul_rn/__dadd_rn.  Cells with a count that is not an integer in [0, 255] (or an enormous sum) are not
// eligible; they are listed and computed entirely by the FP64 kernel.  If the uncertain list overflows, the
// FP64 kernel recomputes every cell (device-side predicate, no host round trip).
//
// GEMM: persistent CTA per SM, warp specialised -- warps 0-3 epilogue (thread = cell = TMEM lane), warp 4
// TMA producer, warp 5 MMA issuer.  Tile = 128 cells x 128 hyperplanes x 3 digits (UMMA M=128, one N=256
// and one N=128 instruction per 32 genes), K = genes in 128-byte chunks, both operands by TMA (128B swizzle)
// through a 3-stage ring of 64 KB stages, accumulators in 384 TMEM columns (the epilogue is ~1 % of a tile's
// 235-chunk main loop, so it is not double buffered).
// PAIR form (default): the two CTAs of a cluster (the two SMs of a TPC) take 256 cells x 128 hyperplanes as ONE
// cta_group::2 tile (M = 256): each CTA streams its own 128 cells of A and HALF of every B tile (digit 0 or digit 1 of
// the N=256 instruction, 64 of digit 2's 128 rows for the N=128 one) -- 40 KB per K chunk and SM instead of 64, five
// stages, and per MMA 8 / 6 KB of shared-memory operand reads per SM instead of 12 / 8.  With both operands in shared
// memory and one CTA per tile the pipe tops out at 0.81 (N=256) / 0.64 (N=128) of its peak (tools/mma_peak.cu); the
// single-CTA kernel sat at that ceiling (ncu: tensor pipe 73 % at 1 M cells).
#include "common.cuh"
#include "tc05.cuh"

#include <algorithm>

namespace em2 {

namespace {

using namespace tc05;

constexpr int kFM = 128;              // cells per tile (UMMA M)
constexpr int kFHyper = 128;          // hyperplanes per tile
constexpr int kFDigits = 3;
constexpr int kFN = kFDigits * kFHyper;   // accumulator columns: digit d of hyperplane j at column d*128 + j (d = 0 highest)
constexpr int kFChunk = 128;          // genes (K bytes) per pipeline stage
constexpr int kFStages = 3;
constexpr uint32_t kFABytes = kFM * kFChunk;
constexpr uint32_t kFBBytes = kFN * kFChunk;
constexpr uint32_t kFStageBytes = kFABytes + kFBBytes;     // 64 KB
constexpr int kFThreads = 192;
constexpr int kFPairStages = 5;
constexpr uint32_t kFPairBBytes = (kFN / 2) * kFChunk;                 // this CTA's half of the B rows: 128 + 64
constexpr uint32_t kFPairStageBytes = kFABytes + kFPairBBytes;          // 40 KB
constexpr double kQuantMax = 2080768.;   // |q| <= 127 * 128 * 128
constexpr double kNearZeroEps = 1e-12;
constexpr double kMaxSum1 = 1.6e7;    // keeps |sum c * qh| <= 127 * sum1 below 2^31

// Per hyperplane column: sum over genes in ascending order (src/Lsh.cpp:137-144), max |U|, and the
// constants of the filter bound.  The sum is a G-long dependent chain of FP64 adds per column, so the
// kernel is organised around keeping that chain fed: a block owns 32 columns, all 8 warps stream 128-gene
// tiles into a double-buffered shared-memory panel with cp.async, warp 0 runs the 32 chains from there.
constexpr int kCsRows = 128;
constexpr int kCsStages = 4;          // 4 x 32 KB panels in flight per block: the loads stay ahead of the add chain

__device__ __forceinline__ void cpAsync8(void* smemDst, const void* gmemSrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smemAddr(smemDst)), "l"(gmemSrc) : "memory");
}

__global__ void __launch_bounds__(256)
columnStatsKernel(uint64_t geneCount, const double* __restrict__ U, uint64_t ld, uint32_t lshCount, uint32_t cols,
                  double* __restrict__ sumU, double* __restrict__ scale, double* __restrict__ e1,
                  double* __restrict__ e2)
{
    extern __shared__ __align__(16) double panel[];      // [kCsStages][kCsRows][32]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * 32 + tx;
    const bool active = i < lshCount;
    const uint32_t tiles = uint32_t((geneCount + kCsRows - 1) / kCsRows);
    auto issue = [&](uint32_t t) {
        if (t < tiles && active) {
            double* dst = panel + size_t(t % kCsStages) * kCsRows * 32;
            for (int r = ty; r < kCsRows; r += 8) {
                const uint64_t g = uint64_t(t) * kCsRows + r;
                if (g < geneCount) cpAsync8(dst + r * 32 + tx, U + g * ld + i);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");      // one group per tile slot, possibly empty
    };
    double s = 0., m = 0.;
    for (uint32_t t = 0; t < kCsStages - 1; t++) issue(t);
    for (uint32_t t = 0; t < tiles; t++) {
        issue(t + kCsStages - 1);        // refills the panel consumed in the previous iteration
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kCsStages - 1) : "memory");
        __syncthreads();
        if (ty == 0 && active) {
            const double* src = panel + size_t(t % kCsStages) * kCsRows * 32 + tx;
            const int rows = int(min(uint64_t(kCsRows), geneCount - uint64_t(t) * kCsRows));
            if (rows == kCsRows) {
#pragma unroll 16
                for (int r = 0; r < kCsRows; r++) {
                    const double v = src[r * 32];
                    s = __dadd_rn(s, v);
                    m = fmax(m, fabs(v));
                }
            } else {
                for (int r = 0; r < rows; r++) {
                    const double v = src[r * 32];
                    s = __dadd_rn(s, v);
                    m = fmax(m, fabs(v));
                }
            }
        }
        __syncthreads();
    }
    if (ty != 0 || i >= cols) return;
    sumU[i] = s;
    if (scale) {
        const double sc = m > 0. ? m / kQuantMax : 1.;
        scale[i] = sc;
        e1[i] = 0.5 * sc * (1. + 1e-6);
        e2[i] = 2.220446049250313e-16 * (m + fabs(s) / double(geneCount));
    }
}

// Quantise + transpose the hyperplanes into the GEMM's B operand: int8 [nBlocks*384][Gpad], row
// nb*384 + d*128 + j = digit d (0 = highest) of hyperplane nb*128 + j, genes contiguous.  Block = 128 genes x 32 hyperplanes.  The buffer is zeroed beforehand (pads stay 0).
__global__ void __launch_bounds__(256)
quantizeKernel(uint64_t geneCount, const double* __restrict__ U, uint64_t ld, uint32_t lshCount,
               const double* __restrict__ scale, uint64_t gPad, int8_t* __restrict__ Uq)
{
    __shared__ int8_t sDig[kFDigits][32][132];
    const uint64_t g0 = uint64_t(blockIdx.x) * 128;
    const uint32_t i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t i = i0 + tx;
    const double inv = (i < lshCount) ? 1. / scale[i] : 0.;
    const double sc = (i < lshCount) ? scale[i] : 1.;
    for (int r = ty; r < 128; r += 8) {
        const uint64_t g = g0 + r;
        int q = 0;
        if (g < geneCount && i < lshCount) {
            const double u = __ldg(U + g * ld + i);
            // nearest integer to u/scale: the product with the reciprocal can be off by one ulp, which
            // moves the quantisation error by at most 1e-12 * scale -- inside e1's safety factor.
            q = __double2int_rn(u * inv);
            // keep the invariant |u - sc*q| <= sc/2 (1 + 1e-6) under every rounding
            if (fabs(u - sc * double(q)) > 0.5 * sc * (1. + 5e-7)) q = __double2int_rn(u / sc);
        }
        const int d2 = (q + 8192) >> 14;        // floor: q - 16384 d2 in [-8192, 8191], |d2| <= 127
        const int rem = q - 16384 * d2;
        const int d1 = (rem + 64) >> 7;         // in [-64, 64]
        sDig[0][tx][r] = int8_t(d2);
        sDig[1][tx][r] = int8_t(d1);
        sDig[2][tx][r] = int8_t(rem - 128 * d1);   // in [-64, 63]
    }
    __syncthreads();
    // write: warp w handles hyperplanes w, w+8, ...; lane writes 4 consecutive genes
    for (int c = ty; c < 32; c += 8) {
        const uint32_t ic = i0 + c;
        if (ic >= lshCount) continue;
        const uint32_t nb = ic / kFHyper, j = ic % kFHyper;
#pragma unroll
        for (int d = 0; d < kFDigits; d++) {
            int8_t* row = Uq + (uint64_t(nb) * kFN + d * kFHyper + j) * gPad + g0;
            *reinterpret_cast<uint32_t*>(row + 4 * tx) = *reinterpret_cast<const uint32_t*>(&sDig[d][c][4 * tx]);
        }
    }
}

// Dense uint8 expansion of a chunk of cells (the GEMM's A operand) + eligibility.  One warp per cell.
__global__ void __launch_bounds__(256)
densifyKernel(uint64_t chunkBegin, uint32_t chunkCells, uint64_t geneCount, uint64_t gPad,
              const uint64_t* __restrict__ toc, const em2_count* __restrict__ counts,
              float maxCount, uint8_t* __restrict__ dense,
              uint8_t* __restrict__ flags, uint32_t* __restrict__ fallbackList, uint32_t* __restrict__ fallbackCount)
{
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= chunkCells) return;
    const uint64_t cell = chunkBegin + w;
    uint8_t* row = dense + uint64_t(w) * gPad;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t o = uint64_t(lane) * 16; o < gPad; o += 512) *reinterpret_cast<uint4*>(row + o) = z;
    __syncwarp();
    bool ok = true;
    double rowSum = 0.;      // the eligibility bound on the cell's sum, computed here so that this kernel does not
                             // depend on the per-cell sums kernel (it runs ahead on its own stream)
    const uint64_t end = toc[cell + 1];
    for (uint64_t e = toc[cell] + lane; e < end; e += 32) {
        const em2_count p = counts[e];
        const float c = p.count;
        ok = ok && (c >= 0.f) && (c <= maxCount) && (c == truncf(c)) && (p.gene < geneCount);
        rowSum += double(fminf(fmaxf(c, 0.f), 65536.f));
        if (p.gene < geneCount) row[p.gene] = uint8_t(min(255u, __float2uint_rz(fmaxf(c, 0.f))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rowSum += __shfl_xor_sync(0xffffffffu, rowSum, o);
    ok = __all_sync(0xffffffffu, ok) && rowSum < kMaxSum1;
    if (lane == 0) {
        flags[w] = ok ? 1 : 0;
        if (!ok) fallbackList[atomicAdd(fallbackCount, 1u)] = uint32_t(cell);
    }
}

// Same result, one CTA per cell with the row built in SHARED memory: the scatter of ~5 % non-zero bytes lands in
// shared memory instead of read-modify-writing 32-byte sectors of a freshly zeroed row in L2, and the row leaves
// the SM once, as coalesced 16-byte stores (the copy loop re-zeroes the buffer for the next cell).  Measured on
// config 2 (100k cells x 30k genes): 1.70 -> see DESIGN.md 4.1; the warp-per-cell kernel above stays for gene
// counts whose row does not fit (gPad > kDenseSmemMax).
constexpr uint32_t kDenseSmemMax = 96 * 1024;

__global__ void __launch_bounds__(256)
densifySmemKernel(uint64_t chunkBegin, uint32_t chunkCells, uint64_t geneCount, uint64_t gPad,
                  const uint64_t* __restrict__ toc, const em2_count* __restrict__ counts,
                  float maxCount, uint8_t* __restrict__ dense,
                  uint8_t* __restrict__ flags, uint32_t* __restrict__ fallbackList, uint32_t* __restrict__ fallbackCount)
{
    extern __shared__ __align__(16) uint8_t srow[];       // gPad bytes
    __shared__ double warpSum[8];
    __shared__ int warpOk[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t o = uint64_t(threadIdx.x) * 16; o < gPad; o += 256 * 16) *reinterpret_cast<uint4*>(srow + o) = z;
    __syncthreads();
    for (uint32_t w = blockIdx.x; w < chunkCells; w += gridDim.x) {
        const uint64_t cell = chunkBegin + w;
        bool ok = true;
        double rowSum = 0.;
        const uint64_t end = toc[cell + 1];
        for (uint64_t e = toc[cell] + threadIdx.x; e < end; e += 256) {
            const em2_count p = counts[e];
            const float c = p.count;
            ok = ok && (c >= 0.f) && (c <= maxCount) && (c == truncf(c)) && (p.gene < geneCount);
            rowSum += double(fminf(fmaxf(c, 0.f), 65536.f));
            if (p.gene < geneCount) srow[p.gene] = uint8_t(min(255u, __float2uint_rz(fmaxf(c, 0.f))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rowSum += __shfl_xor_sync(0xffffffffu, rowSum, o);
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) {
            warpSum[warp] = rowSum;
            warpOk[warp] = ok ? 1 : 0;
        }
        __syncthreads();
        uint8_t* row = dense + uint64_t(w) * gPad;
        for (uint64_t o = uint64_t(threadIdx.x) * 16; o < gPad; o += 256 * 16) {
            *reinterpret_cast<uint4*>(row + o) = *reinterpret_cast<const uint4*>(srow + o);
            *reinterpret_cast<uint4*>(srow + o) = z;
        }
        if (threadIdx.x == 0) {
            double sum = 0.;
            bool all = true;
            for (int i = 0; i < 8; i++) {      // fixed order: the sum only feeds the eligibility test
                sum += warpSum[i];
                all = all && warpOk[i];
            }
            all = all && sum < kMaxSum1;
            flags[w] = all ? 1 : 0;
            if (!all) fallbackList[atomicAdd(fallbackCount, 1u)] = uint32_t(cell);
        }
        __syncthreads();
    }
}

struct FilterParams {
    uint64_t chunkBegin;
    uint32_t chunkCells;
    uint64_t geneCount;
    uint32_t kChunks;          // ceil(G / 128)
    uint32_t mBlocks, nBlocks;
    uint32_t lshCount, wordsPerCell;
    uint32_t idesc256, idesc128;
    const uint64_t* toc;
    const double* sum1;
    const uint8_t* flags;
    const double *sumU, *scale, *e1, *e2;
    uint64_t* signatures;
    uint64_t* uncertain;       // (cell << 32 | hyperplane)
    uint32_t* uncertainCount;
    uint32_t uncertainCap;
};

// PAIR: mapB64 is mapB with a 64-row box (this CTA's half of digit 2); the leader (cluster rank 0) owns the barriers the
// MMA issuer waits on (full, accEmpty) and issues for the pair; empty / accFull are signalled in both CTAs (commitPair).
template <bool PAIR>
__global__ void __launch_bounds__(kFThreads, 1)
sigFilterKernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapB64, const FilterParams p)
{
    constexpr int kNumStages = PAIR ? kFPairStages : kFStages;
    constexpr uint32_t kStageSz = PAIR ? kFPairStageBytes : kFStageBytes;
    extern __shared__ uint8_t smemRaw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + size_t(kNumStages) * kStageSz);
    uint64_t* full = bars;                     // [stages]  TMA -> MMA
    uint64_t* empty = bars + kNumStages;       // [stages]  MMA -> TMA
    uint64_t* accFull = bars + 2 * kNumStages; // MMA -> epilogue
    uint64_t* accEmpty = accFull + 1;          // epilogue -> MMA
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(accEmpty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? clusterRank() : 0;
    const uint32_t worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
    const uint32_t workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumStages; i++) {
            mbarInit(full + i, 1);
            mbarInit(empty + i, 1);
        }
        mbarInit(accFull, 1);
        mbarInit(accEmpty, PAIR ? 8 : 128);      // PAIR: one arrival per epilogue warp of either CTA
        mbarInitFence();
    }
    if (warp == 4) {
        if (PAIR) tmemAllocPair(tmemSlot, 512);
        else tmemAlloc(tmemSlot, 512);
    }
    fenceBefore();
    if (PAIR) clusterSync();      // barriers of both CTAs are initialised before anyone signals across
    else __syncthreads();
    fenceAfter();
    const uint32_t tmemBase = *tmemSlot;
    // PAIR: an item is 256 cells (m-blocks 2 q and 2 q + 1; an odd last block reads zeros and writes nothing)
    const uint32_t mItems = PAIR ? (p.mBlocks + 1) / 2 : p.mBlocks;
    const uint32_t items = mItems * p.nBlocks;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            prefetchMap(&mapA);
            prefetchMap(&mapB);
            uint32_t stage = 0, phase = 0;
            if (PAIR) prefetchMap(&mapB64);
            for (uint32_t item = worker; item < items; item += workers) {
                const int32_t rowA = int32_t(((item / p.nBlocks) * (PAIR ? 2 : 1) + rank) * kFM);
                const int32_t rowB = int32_t((item % p.nBlocks) * kFN);
                for (uint32_t kc = 0; kc < p.kChunks; kc++) {
                    mbarWait(empty + stage, phase ^ 1);
                    uint8_t* dst = ring + size_t(stage) * kStageSz;
                    if (PAIR) {
                        // this CTA's cells, its digit (0 or 1) of the N=256 instruction, its 64 rows of digit 2;
                        // the leader's barrier collects both CTAs' bytes
                        if (rank == 0) mbarExpectTx(full + stage, 2 * kFPairStageBytes);
                        tmaLoad2dPair(dst, &mapA, full + stage, int32_t(kc * kFChunk), rowA);
                        tmaLoad2dPair(dst + kFABytes, &mapB, full + stage, int32_t(kc * kFChunk), rowB + int32_t(rank) * kFHyper);
                        tmaLoad2dPair(dst + kFABytes + kFHyper * kFChunk, &mapB64, full + stage, int32_t(kc * kFChunk),
                                      rowB + 2 * kFHyper + int32_t(rank) * (kFHyper / 2));
                    } else {
                        mbarExpectTx(full + stage, kFStageBytes);
                        tmaLoad2d(dst, &mapA, full + stage, int32_t(kc * kFChunk), rowA);
#pragma unroll
                        for (int d = 0; d < kFDigits; d++)
                            tmaLoad2d(dst + kFABytes + d * kFHyper * kFChunk, &mapB, full + stage, int32_t(kc * kFChunk),
                                      rowB + d * kFHyper);
                    }
                    if (++stage == kNumStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues (tc05.cuh, electOne) =====================
        uint32_t stage = 0, phase = 0, tileIter = 0;
        const uint32_t ringBase = smemAddr(ring);
        for (uint32_t item = worker; item < items && rank == 0; item += workers, tileIter++) {
            mbarWait(accEmpty, (tileIter & 1) ^ 1);
            fenceAfter();
            const uint32_t tmemD = tmemBase;
            for (uint32_t kc = 0; kc < p.kChunks; kc++) {
                mbarWait(full + stage, phase);
                fenceAfter();
                const uint32_t aAddr = ringBase + stage * kStageSz;
                const uint64_t descA = makeSmemDesc(aAddr);
                const uint64_t descB0 = makeSmemDesc(aAddr + kFABytes);
                // second instruction's B rows: digit 2 (PAIR: this CTA's 64 of them, behind its 128 rows of digit 0 / 1)
                const uint64_t descB1 = makeSmemDesc(aAddr + kFABytes + (PAIR ? kFHyper : 2 * kFHyper) * kFChunk);
                if (electOne()) {
#pragma unroll
                    for (int ks = 0; ks < kFChunk / 32; ks++) {      // + 32 bytes along K = + 2 in the descriptor's address field
                        const uint32_t accumulate = (kc | uint32_t(ks)) != 0 ? 1u : 0u;
                        if (PAIR) {
                            mmaI8SsPair(tmemD, descA + uint64_t(2 * ks), descB0 + uint64_t(2 * ks), p.idesc256, accumulate);
                            mmaI8SsPair(tmemD + 256, descA + uint64_t(2 * ks), descB1 + uint64_t(2 * ks), p.idesc128, accumulate);
                        } else {
                            mmaI8Ss(tmemD, descA + uint64_t(2 * ks), descB0 + uint64_t(2 * ks), p.idesc256, accumulate);
                            mmaI8Ss(tmemD + 256, descA + uint64_t(2 * ks), descB1 + uint64_t(2 * ks), p.idesc128, accumulate);
                        }
                    }
                    if (PAIR) {      // in both CTAs: the stage is free, the accumulator halves are complete
                        commitPair(empty + stage);
                        if (kc + 1 == p.kChunks) commitPair(accFull);
                    } else {
                        commit(empty + stage);
                        if (kc + 1 == p.kChunks) commit(accFull);
                    }
                }
                __syncwarp();
                if (++stage == kNumStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: thread == cell == TMEM lane =====================
        const uint32_t laneField = uint32_t(warp * 32) << 16;
        uint32_t tileIter = 0;
        for (uint32_t item = worker; item < items; item += workers, tileIter++) {
            const uint32_t mb = (item / p.nBlocks) * (PAIR ? 2 : 1) + rank, nb = item % p.nBlocks;
            const uint32_t wLocal = mb * kFM + threadIdx.x;
            const bool valid = wLocal < p.chunkCells && p.flags[wLocal] != 0;
            const uint64_t cell = p.chunkBegin + (wLocal < p.chunkCells ? wLocal : 0);
            const double s1 = p.sum1[cell];
            const double mean = __ddiv_rn(s1, double(p.geneCount));           // src/Lsh.cpp:167-168
            const double terms = double(p.toc[cell + 1] - p.toc[cell] + 16);
            mbarWait(accFull, tileIter & 1);
            fenceAfter();
            const uint32_t taddr = tmemBase + laneField;
            uint32_t masks[4];
#pragma unroll 1
            for (int q = 0; q < 4; q++) {
                uint32_t d2[32], d1[32], d0[32];
                tmemLoad32(taddr + q * 32, d2);
                tmemLoad32(taddr + kFHyper + q * 32, d1);
                tmemLoad32(taddr + 2 * kFHyper + q * 32, d0);
                tmemLoadWait();
                uint32_t mask = 0;
                const uint32_t iBase = nb * kFHyper + q * 32;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t i = iBase + j;
                    const double acc = fma(fma(double(int32_t(d2[j])), 128., double(int32_t(d1[j]))), 128.,
                                           double(int32_t(d0[j])));                  // exact integer < 2^53
                    const double sh = __dsub_rn(__dmul_rn(acc, __ldg(p.scale + i)), __dmul_rn(mean, __ldg(p.sumU + i)));
                    const double bound = s1 * (__ldg(p.e1 + i) + terms * __ldg(p.e2 + i));
                    if (fabs(sh) > bound) {
                        mask |= uint32_t(sh > 0.) << (31 - j);
                    } else if (valid && i < p.lshCount) {
                        const uint32_t slot = atomicAdd(p.uncertainCount, 1u);
                        if (slot < p.uncertainCap) p.uncertain[slot] = (cell << 32) | i;
                    }
                }
                masks[q] = mask;
            }
            fenceBefore();
            if (PAIR) {
                __syncwarp();
                if (lane == 0) mbarArriveLeader(accEmpty);
            } else {
                mbarArrive(accEmpty);
            }
            if (valid) {
                const uint32_t w0 = nb * 2;
                uint64_t* out = p.signatures + cell * p.wordsPerCell;
                if (w0 < p.wordsPerCell) out[w0] = (uint64_t(masks[0]) << 32) | masks[1];
                if (w0 + 1 < p.wordsPerCell) out[w0 + 1] = (uint64_t(masks[2]) << 32) | masks[3];
            }
        }
    }
    fenceBefore();
    if (PAIR) {
        clusterSync();        // neither CTA may retire while its partner can still signal into it
        if (warp == 4) tmemDeallocPair(tmemBase, 512);
    } else {
        __syncthreads();
        if (warp == 4) tmemDealloc(tmemBase, 512);
    }
}

// Exact FP64 evaluation of the listed (cell, hyperplane) projections: the reference's operation sequence
// (src/Lsh.cpp:180-198).  One warp per entry: lanes fetch 32 stored counts and their hyperplane elements and
// form the 32 products in parallel; the additions run in stored order.
__global__ void __launch_bounds__(256)
fixupKernel(uint64_t geneCount, const uint64_t* __restrict__ toc, const em2_count* __restrict__ counts,
            const double* __restrict__ sum1, const double* __restrict__ sum2, const double* __restrict__ U, uint64_t ld,
            const double* __restrict__ sumU, uint32_t wordsPerCell, const uint64_t* __restrict__ entries,
            const uint32_t* __restrict__ entryCount, uint32_t cap, uint64_t* __restrict__ signatures,
            unsigned long long* __restrict__ nearZero)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(*entryCount, cap);
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += warps) {
        const uint64_t entry = entries[e];
        const uint64_t cell = entry >> 32;
        const uint32_t i = uint32_t(entry);
        const double mean = __ddiv_rn(sum1[cell], double(geneCount));
        double s = __dmul_rn(-mean, sumU[i]);
        const uint64_t end = toc[cell + 1];
        for (uint64_t base = toc[cell]; base < end; base += 32) {
            double prod = 0.;
            if (base + lane < end) {
                const em2_count c = counts[base + lane];
                prod = __dmul_rn(double(c.count), __ldg(U + uint64_t(c.gene) * ld + i));
            }
            const int nv = int(min(uint64_t(32), end - base));
            for (int j = 0; j < nv; j++) s = __dadd_rn(s, __shfl_sync(0xffffffffu, prod, j));
        }
        if (lane == 0) {
            if (s > 0.) atomicOr(reinterpret_cast<unsigned long long*>(signatures + cell * wordsPerCell + (i >> 6)),
                                 1ull << (63 - (i & 63)));
            if (nearZero && sum2 &&
                fabs(s) < kNearZeroEps * (sqrt(sum2[cell]) + fabs(__dmul_rn(mean, sumU[i]))))
                atomicAdd(nearZero, 1ull);
        }
    }
}

__global__ void recordFilterCounters(const uint32_t* uncertainCount, const uint32_t* fallbackCount, uint32_t chunkCells,
                                     unsigned long long* counters)
{
    if (counters) {
        atomicAdd(counters + 2, (unsigned long long)(chunkCells - *fallbackCount));
        atomicAdd(counters + 3, (unsigned long long)(*uncertainCount));
    }
}

}  // namespace

int launchColumnStats(em2_context* ctx, uint64_t geneCount, const double* U, uint64_t ld, uint64_t lshCount,
                      uint64_t cols, double* sumU, double* scale, double* e1, double* e2, cudaStream_t s)
{
    const size_t smem = size_t(kCsStages) * kCsRows * 32 * sizeof(double);
    EM2_CUDA(ctx, cudaFuncSetAttribute(columnStatsKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    columnStatsKernel<<<unsigned((cols + 31) / 32), 256, smem, s>>>(geneCount, U, ld, uint32_t(lshCount), uint32_t(cols),
                                                                    sumU, scale, e1, e2);
    ctx->stats.kernel_launches++;
    EM2_CUDA(ctx, cudaGetLastError());
    return EM2_OK;
}

// Hyperplane-side preparation of the filter path (once per job): column constants + quantised operand.
int prepareSignaturesFiltered(em2_context* ctx, SignaturePlan& pl, uint64_t cellCountHint, cudaStream_t s)
{
    const uint64_t geneCount = pl.geneCount, lshCount = pl.lshCount;
    pl.gPad = roundUp(geneCount, kFChunk);
    pl.nBlocks = uint32_t((lshCount + kFHyper - 1) / kFHyper);
    const uint64_t Lpad = uint64_t(pl.nBlocks) * kFHyper;
    if (geneCount > 0x7fffff00ull) return fail(ctx, EM2_ERR_INVALID, "geneCount too large");

    void* stats = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_SUMU, 4 * Lpad * sizeof(double), &stats));
    pl.sumU = static_cast<double*>(stats);
    pl.scale = pl.sumU + Lpad;
    pl.e1 = pl.scale + Lpad;
    pl.e2 = pl.e1 + Lpad;
    const size_t uqBytes = size_t(pl.nBlocks) * kFN * pl.gPad;
    EM2_TRY(reserve(ctx, em2_context::S_UQ, uqBytes, &pl.uq));
    // The column constants are G-long dependent FP64 add chains on 32 SMs (0.6 ms at 30k genes) and, like the
    // quantisation, depend on the hyperplanes only: they run on a side stream while the main stream computes the
    // per-cell sums and the dense counts; the GEMM launch waits for them (evPrep).
    cudaStream_t a = ctx->auxStream;
    EM2_CUDA(ctx, cudaEventRecord(ctx->evFork, s));
    EM2_CUDA(ctx, cudaStreamWaitEvent(a, ctx->evFork, 0));
    EM2_TRY(launchColumnStats(ctx, geneCount, pl.U, pl.ld, lshCount, Lpad, pl.sumU, pl.scale, pl.e1, pl.e2, a));
    EM2_CUDA(ctx, cudaMemsetAsync(pl.uq, 0, uqBytes, a));
    {
        const dim3 grid(unsigned(pl.gPad / 128), unsigned((lshCount + 31) / 32));
        quantizeKernel<<<grid, 256, 0, a>>>(geneCount, pl.U, pl.ld, uint32_t(lshCount), pl.scale, pl.gPad,
                                            static_cast<int8_t*>(pl.uq));
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());
    }
    EM2_CUDA(ctx, cudaEventRecord(ctx->evPrep, a));
    pl.prepOnAux = true;

    // scratch of the cell chunks: two slots (dense operand, flags, lists) so that the dense expansion of chunk
    // c+1 (HBM-write bound, on a side stream) overlaps the GEMM of chunk c (tensor bound)
    const uint64_t denseBudget = 3ull << 30;                 // per slot
    const uint64_t hint = std::max<uint64_t>(cellCountHint, 1);
    // Measured (100k cells x 30k genes): 1 / 2 / 4 / 8 chunks per call take 7.6 / 7.9 / 8.1 / 9.3 ms -- the GEMM loses
    // more to wave quantisation and the extra small launches than the overlap of the next expansion wins, so a call
    // is cut only where the 3 GB slot demands it; the two slots still overlap consecutive calls (the blocking API's
    // PCIe chunks) and let the expansion start while the hyperplane-side preparation runs.
    const uint64_t parts = ctx->filterParts ? uint64_t(ctx->filterParts) : 1;
    pl.chunkMax = std::max<uint64_t>(kFM, denseBudget / pl.gPad / kFM * kFM);
    pl.chunkMax = std::min<uint64_t>(pl.chunkMax, roundUp((hint + parts - 1) / parts, kFM));
    EM2_TRY(reserve(ctx, em2_context::S_DENSE, 2 * pl.chunkMax * pl.gPad, &pl.dense));
    pl.uncertainCap = ctx->filterUncertainCap ? ctx->filterUncertainCap :
        uint32_t(std::min<uint64_t>(std::max<uint64_t>(1u << 20, pl.chunkMax * lshCount / 32), 1u << 28));
    // layout of one slot of S_FLAGS: [counters: 2 x u32 (+pad to 16)] [flags: chunkMax bytes] [fallback list: chunkMax u32] [uncertain: cap u64]
    pl.offFlags = 16;
    pl.offFallback = roundUp(pl.offFlags + pl.chunkMax, 16);
    pl.offUncertain = roundUp(pl.offFallback + 4 * pl.chunkMax, 16);
    pl.slotBytes = roundUp(pl.offUncertain + size_t(pl.uncertainCap) * 8, 256);
    EM2_TRY(reserve(ctx, em2_context::S_FLAGS, 2 * pl.slotBytes, &pl.lists));
    return EM2_OK;
}

// Signatures of cells [cellBegin, cellEnd) through the filter path.
int launchSignaturesFiltered(em2_context* ctx, const SignaturePlan& pl, const uint64_t* toc, const em2_count* counts,
                             const double* sum1, const double* sum2, uint64_t cellBegin, uint64_t cellEnd,
                             uint64_t* signatures, uint64_t* nearZero, cudaStream_t s)
{
    const uint64_t geneCount = pl.geneCount, lshCount = pl.lshCount, gPad = pl.gPad, chunkMax = pl.chunkMax;
    const uint64_t W = wordCount(lshCount);
    const uint32_t nBlocks = pl.nBlocks, uncertainCap = pl.uncertainCap;
    const double *U = pl.U, *Upadded = pl.Upadded, *sumU = pl.sumU, *scale = pl.scale, *e1 = pl.e1, *e2 = pl.e2;
    const uint64_t ld = pl.ld, ldPadded = pl.ldPadded;
    void* uq = pl.uq;
    const uint64_t cellCount = cellEnd;

    const bool unsignedCounts = ctx->filterCountsSigned == 0;
    const bool pair = ctx->filterCtaPair != 1 && ctx->smCount >= 2;
    const size_t smem = 1024 + (pair ? size_t(kFPairStages) * kFPairStageBytes : size_t(kFStages) * kFStageBytes) + 256;
    EM2_CUDA(ctx, cudaFuncSetAttribute(sigFilterKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    EM2_CUDA(ctx, cudaFuncSetAttribute(sigFilterKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    CUtensorMap mapB, mapB64;
    EM2_TRY(makeTensorMapU8(ctx, &mapB, uq, uint64_t(nBlocks) * kFN, gPad, gPad, 128));
    EM2_TRY(makeTensorMapU8(ctx, &mapB64, uq, uint64_t(nBlocks) * kFN, gPad, gPad, 64));

    // the dense expansion runs ahead on its own stream; it only needs the CSR, which everything enqueued on `s` so far provides
    cudaStream_t d = ctx->auxStream2;
    EM2_CUDA(ctx, cudaEventRecord(ctx->evFork2, s));
    EM2_CUDA(ctx, cudaStreamWaitEvent(d, ctx->evFork2, 0));

    for (uint64_t begin = cellBegin; begin < cellEnd; begin += chunkMax) {
        const uint32_t chunkCells = uint32_t(std::min<uint64_t>(chunkMax, cellEnd - begin));
        const int slot = int(ctx->filterChunkSeq++ & 1);
        uint8_t* dense = static_cast<uint8_t*>(pl.dense) + size_t(slot) * chunkMax * gPad;
        uint8_t* base = static_cast<uint8_t*>(pl.lists) + size_t(slot) * pl.slotBytes;
        uint32_t* uncertainCount = reinterpret_cast<uint32_t*>(base);
        uint32_t* fallbackCount = uncertainCount + 1;
        uint8_t* dFlags = base + pl.offFlags;
        uint32_t* fallbackList = reinterpret_cast<uint32_t*>(base + pl.offFallback);
        uint64_t* uncertain = reinterpret_cast<uint64_t*>(base + pl.offUncertain);

        // side stream: wait until the slot's previous user (GEMM + fix-ups) is done, then expand
        if (ctx->slotUsed[slot]) EM2_CUDA(ctx, cudaStreamWaitEvent(d, ctx->evGemm[slot], 0));
        EM2_CUDA(ctx, cudaMemsetAsync(base, 0, 16, d));
        if (gPad <= kDenseSmemMax && ctx->denseWarpKernel == 0) {
            EM2_CUDA(ctx, cudaFuncSetAttribute(densifySmemKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gPad)));
            const unsigned ctasPerSm = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(8, (200 * 1024) / (gPad + 1024))));
            densifySmemKernel<<<std::min<unsigned>(chunkCells, unsigned(ctx->smCount) * ctasPerSm), 256, gPad, d>>>(
                begin, chunkCells, geneCount, gPad, toc, counts, unsignedCounts ? 255.f : 127.f, dense, dFlags, fallbackList,
                fallbackCount);
        } else {
            densifyKernel<<<(chunkCells + 7) / 8, 256, 0, d>>>(begin, chunkCells, geneCount, gPad, toc, counts,
                                                               unsignedCounts ? 255.f : 127.f, dense, dFlags, fallbackList,
                                                               fallbackCount);
        }
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());
        EM2_CUDA(ctx, cudaEventRecord(ctx->evDense[slot], d));
        EM2_CUDA(ctx, cudaStreamWaitEvent(s, ctx->evDense[slot], 0));

        FilterParams p{};
        p.chunkBegin = begin;
        p.chunkCells = chunkCells;
        p.geneCount = geneCount;
        p.kChunks = uint32_t(gPad / kFChunk);
        p.mBlocks = (chunkCells + kFM - 1) / kFM;
        p.nBlocks = nBlocks;
        p.lshCount = uint32_t(lshCount);
        p.wordsPerCell = uint32_t(W);
        p.idesc256 = instrDescI8(!unsignedCounts, true, pair ? 2 * kFM : kFM, 256);
        p.idesc128 = instrDescI8(!unsignedCounts, true, pair ? 2 * kFM : kFM, 128);
        p.toc = toc;
        p.sum1 = sum1;
        p.flags = dFlags;
        p.sumU = sumU;
        p.scale = scale;
        p.e1 = e1;
        p.e2 = e2;
        p.signatures = signatures;
        p.uncertain = uncertain;
        p.uncertainCount = uncertainCount;
        p.uncertainCap = uncertainCap;
        CUtensorMap mapA;
        EM2_TRY(makeTensorMapU8(ctx, &mapA, dense, chunkCells, gPad, gPad, kFM));
        if (pl.prepOnAux) EM2_CUDA(ctx, cudaStreamWaitEvent(s, ctx->evPrep, 0));
        if (pair) {
            const uint32_t items = ((p.mBlocks + 1) / 2) * p.nBlocks;
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(kFThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;      // the two CTAs of a pair: one TPC
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(2 * std::min<uint32_t>(items, uint32_t(ctx->smCount / 2)));
            EM2_CUDA(ctx, cudaLaunchKernelEx(&cfg, sigFilterKernel<true>, mapA, mapB, mapB64, p));
        } else {
            const uint32_t items = p.mBlocks * p.nBlocks;
            const unsigned grid = std::min<uint32_t>(items, uint32_t(ctx->smCount));
            sigFilterKernel<false><<<grid, kFThreads, smem, s>>>(mapA, mapB, mapB64, p);
        }
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());

        fixupKernel<<<ctx->smCount * 8, 256, 0, s>>>(geneCount, toc, counts, sum1, sum2, U, ld, sumU, uint32_t(W), uncertain,
                                                     uncertainCount, uncertainCap, signatures,
                                                     reinterpret_cast<unsigned long long*>(nearZero));
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());

        // cells the filter could not take (non-integer / large counts): the FP64 kernel on the list;
        // and, should the uncertain list have overflowed, the FP64 kernel on the whole chunk.
        EM2_TRY(launchSignaturesFp64(ctx, cellCount, geneCount, toc, counts, sum1, sum2, Upadded, ldPadded, sumU, lshCount,
                                     signatures, nearZero, fallbackList, fallbackCount, chunkCells, nullptr, 0, begin,
                                     chunkCells, s));
        EM2_TRY(launchSignaturesFp64(ctx, cellCount, geneCount, toc, counts, sum1, sum2, Upadded, ldPadded, sumU, lshCount,
                                     signatures, nearZero, nullptr, nullptr, 0, uncertainCount, uncertainCap, begin,
                                     chunkCells, s));
        recordFilterCounters<<<1, 1, 0, s>>>(uncertainCount, fallbackCount, chunkCells,
                                             reinterpret_cast<unsigned long long*>(ctx->scratch[em2_context::S_COUNTERS].ptr));
        ctx->stats.kernel_launches++;
        EM2_CUDA(ctx, cudaGetLastError());
        EM2_CUDA(ctx, cudaEventRecord(ctx->evGemm[slot], s));
        ctx->slotUsed[slot] = true;
    }
    return EM2_OK;
}

}  // namespace em2
