// Signature construction on sm_100a.
//
// Replaces, bit for bit, the reference's CPU loops
//   ExpressionMatrixSubset::computeSums         (reference src/ExpressionMatrixSubset.cpp:47-58)
//   lshVectorsSums                              (reference src/Lsh.cpp:137-144)
//   Lsh::computeCellLshSignatures               (reference src/Lsh.cpp:160-207)
// Bit-exactness rule: every projection s_i is the SAME sequence of IEEE double operations the
// reference executes (its build is -O3 -msse4.2: separate mulsd/addsd, no FMA):
//     s_i = (-mean) * sumU_i;   for each stored (gene,count) of the cell, in stored order:
//     s_i = s_i + double(count) * U[gene][i]
// so the kernels use __dmul_rn/__dadd_rn (never contracted) and keep one accumulator per
// (cell, hyperplane) with no cross-thread reduction.  Parallelism is over cells and hyperplanes.
//
// Data movement (v1): one warp owns (cell, 128-hyperplane slice).  The cell's CSR entries are staged
// through shared memory 32 at a time (coalesced 256 B reads); for every entry the warp reads the 1 KB
// slice row U[gene][slice] with two 512 B LDG.128 requests.  Blocks are ordered slice-major so that
// the G x 128 x 8 B slice (30 MB at G = 30k) stays L2 resident while all cells pass over it.
#include "common.cuh"

#include <algorithm>

namespace em2 {

namespace {

constexpr int kSigWarps = 8;          // warps (cells) per block
constexpr int kSliceCols = 128;       // hyperplanes per warp slice
constexpr double kNearZeroEps = 1e-12;

// One thread per cell, sequential like the reference (order matters for the rounding of sum1).
__global__ void cellSumsKernel(uint64_t cellBegin, uint64_t cellCount, const uint64_t* __restrict__ toc,
                               const em2_count* __restrict__ counts, double* __restrict__ sum1,
                               double* __restrict__ sum2)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= cellCount) return;
    const uint64_t c = cellBegin + i;
    double s1 = 0., s2 = 0.;
    const uint64_t e = toc[c + 1];
    for (uint64_t j = toc[c]; j < e; j++) {
        const float x = counts[j].count;
        s1 = __dadd_rn(s1, double(x));                 // sum.sum1 += count
        s2 = __dadd_rn(s2, double(__fmul_rn(x, x)));   // sum.sum2 += count*count (float product)
    }
    sum1[c] = s1;
    if (sum2) sum2[c] = s2;
}

__device__ __forceinline__ uint64_t spreadBits(uint32_t v)
{
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

__device__ __forceinline__ double2 ldU(const double* p)
{
    // read-only, 16-byte vector load
    return __ldg(reinterpret_cast<const double2*>(p));
}

// Work unit = (128-hyperplane slice, group of kSigWarps cells), slice-major; blocks take units grid-stride
// (the plain launch has one block per unit).  U must have `ld` >= slices*128 columns readable (zero padded
// beyond lshCount), ld even, base 16-byte aligned.  Cells are [rangeBegin, rangeBegin + rangeCells), or, when
// `cellList` is given, the first *cellListCount entries of that list.  With `runIfCount` the whole launch is
// a no-op unless *runIfCount > runIfCap (device-side predicate of the filter path's overflow rescue).
template <bool PLAIN>      // PLAIN: one block per unit of a cell range, no list, no predicate (the common launch)
__global__ void __launch_bounds__(kSigWarps * 32, 4)
signatureKernel(uint64_t rangeBegin, uint64_t rangeCells, uint64_t geneCount, const uint64_t* __restrict__ toc,
                const em2_count* __restrict__ counts, const double* __restrict__ sum1,
                const double* __restrict__ sum2, const double* __restrict__ U, uint64_t ld,
                const double* __restrict__ sumU, uint32_t lshCount, uint32_t wordsPerCell, uint32_t slices,
                uint64_t* __restrict__ signatures, unsigned long long* __restrict__ nearZero,
                const uint32_t* __restrict__ cellList, const uint32_t* __restrict__ cellListCount,
                const uint32_t* __restrict__ runIfCount, uint32_t runIfCap)
{
    __shared__ uint32_t sGene[kSigWarps][32];
    __shared__ double sCount[kSigWarps][32];

    if (!PLAIN && runIfCount != nullptr && *runIfCount <= runIfCap) return;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint64_t nCells = (!PLAIN && cellList) ? min(uint64_t(*cellListCount), rangeCells) : rangeCells;
    const uint32_t cellBlocks = uint32_t((nCells + kSigWarps - 1) / kSigWarps);
    const uint32_t units = PLAIN ? blockIdx.x + 1 : cellBlocks * slices;
  for (uint32_t unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const uint32_t slice = unit / cellBlocks;
    const uint64_t idx = uint64_t(unit % cellBlocks) * kSigWarps + warp;
    if (idx >= nCells) continue;
    const uint64_t cell = (!PLAIN && cellList) ? uint64_t(cellList[idx]) : rangeBegin + idx;

    const uint32_t colBase = slice * kSliceCols;
    const uint32_t c0 = colBase + 2 * lane;        // hyperplanes c0, c0+1
    const uint32_t c1 = c0 + 64;                   // hyperplanes c1, c1+1

    const double mean = __ddiv_rn(sum1[cell], double(geneCount));      // src/Lsh.cpp:167-168
    const double negMean = -mean;
    const double2 su0 = *reinterpret_cast<const double2*>(sumU + c0);
    const double2 su1 = *reinterpret_cast<const double2*>(sumU + c1);
    double a0 = __dmul_rn(negMean, su0.x);                               // src/Lsh.cpp:180-182
    double a1 = __dmul_rn(negMean, su0.y);
    double a2 = __dmul_rn(negMean, su1.x);
    double a3 = __dmul_rn(negMean, su1.y);

    const double* Uslice = U + c0;
    const uint64_t begin = toc[cell];
    const uint64_t end = toc[cell + 1];
    for (uint64_t base = begin; base < end; base += 32) {
        // Stage up to 32 entries of the CSR row (coalesced), padded to a multiple of 4 with
        // (gene 0, count 0): adding 0*U leaves every accumulator unchanged.
        const uint64_t left = end - base;
        const int n = left < 32 ? int(left) : 32;
        uint32_t g = 0;
        double c = 0.;
        if (lane < n) {
            const em2_count p = counts[base + lane];
            g = p.gene;
            c = double(p.count);                                         // src/Lsh.cpp:190
        }
        sGene[warp][lane] = g;
        sCount[warp][lane] = c;
        __syncwarp();
        const int n4 = (n + 3) & ~3;
        for (int j = 0; j < n4; j += 4) {
            double2 u[4][2];
            double cnt[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double* row = Uslice + uint64_t(sGene[warp][j + q]) * ld;
                cnt[q] = sCount[warp][j + q];
                u[q][0] = ldU(row);
                u[q][1] = ldU(row + 64);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {                                // src/Lsh.cpp:194-197, in order
                a0 = __dadd_rn(a0, __dmul_rn(cnt[q], u[q][0].x));
                a1 = __dadd_rn(a1, __dmul_rn(cnt[q], u[q][0].y));
                a2 = __dadd_rn(a2, __dmul_rn(cnt[q], u[q][1].x));
                a3 = __dadd_rn(a3, __dmul_rn(cnt[q], u[q][1].y));
            }
        }
        __syncwarp();
    }

    // Sign bits -> two 64-bit words, first hyperplane in the most significant bit
    // (src/Lsh.cpp:201-206, src/BitSet.hpp:48-62).  Columns >= lshCount never set a bit.
    const bool p0 = (c0 < lshCount) && (a0 > 0.);
    const bool p1 = (c0 + 1 < lshCount) && (a1 > 0.);
    const bool p2 = (c1 < lshCount) && (a2 > 0.);
    const bool p3 = (c1 + 1 < lshCount) && (a3 > 0.);
    const uint32_t b0 = __ballot_sync(0xffffffffu, p0);
    const uint32_t b1 = __ballot_sync(0xffffffffu, p1);
    const uint32_t b2 = __ballot_sync(0xffffffffu, p2);
    const uint32_t b3 = __ballot_sync(0xffffffffu, p3);

    if (nearZero != nullptr && sum2 != nullptr) {
        const double scale = sqrt(sum2[cell]);
        int nz = 0;
        nz += (c0 < lshCount) && (fabs(a0) < kNearZeroEps * (scale + fabs(__dmul_rn(mean, su0.x))));
        nz += (c0 + 1 < lshCount) && (fabs(a1) < kNearZeroEps * (scale + fabs(__dmul_rn(mean, su0.y))));
        nz += (c1 < lshCount) && (fabs(a2) < kNearZeroEps * (scale + fabs(__dmul_rn(mean, su1.x))));
        nz += (c1 + 1 < lshCount) && (fabs(a3) < kNearZeroEps * (scale + fabs(__dmul_rn(mean, su1.y))));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
        if (lane == 0 && nz) atomicAdd(nearZero, (unsigned long long)nz);
    }

    if (lane == 0) {
        const uint32_t w0 = slice * 2;
        if (w0 < wordsPerCell)
            signatures[cell * wordsPerCell + w0] = __brevll(spreadBits(b0) | (spreadBits(b1) << 1));
        if (w0 + 1 < wordsPerCell)
            signatures[cell * wordsPerCell + w0 + 1] = __brevll(spreadBits(b2) | (spreadBits(b3) << 1));
    }
    __syncwarp();
  }
}

}  // namespace

int launchCellSums(em2_context* ctx, uint64_t cellCount, const uint64_t* toc, const em2_count* counts,
                   double* sum1, double* sum2, cudaStream_t s, uint64_t cellBegin)
{
    if (cellCount == 0) return EM2_OK;
    const int threads = 128;
    const unsigned blocks = unsigned((cellCount + threads - 1) / threads);
    cellSumsKernel<<<blocks, threads, 0, s>>>(cellBegin, cellCount, toc, counts, sum1, sum2);
    ctx->stats.kernel_launches++;
    EM2_CUDA(ctx, cudaGetLastError());
    return EM2_OK;
}

int launchSignaturesFp64(em2_context* ctx, uint64_t cellCount, uint64_t geneCount, const uint64_t* toc,
                         const em2_count* counts, const double* sum1, const double* sum2, const double* Uk,
                         uint64_t ldk, const double* sumU, uint64_t lshCount, uint64_t* signatures, uint64_t* nearZero,
                         const uint32_t* cellList, const uint32_t* cellListCount, uint64_t maxListed,
                         const uint32_t* runIfCount, uint32_t runIfCap, uint64_t rangeBegin, uint64_t rangeCells,
                         cudaStream_t s)
{
    (void)cellCount;
    const uint64_t W = wordCount(lshCount);
    const uint32_t slices = uint32_t(roundUp(lshCount, kSliceCols) / kSliceCols);
    const uint64_t cells = cellList ? maxListed : rangeCells;
    if (cells == 0) return EM2_OK;
    uint64_t blocks = (cells + kSigWarps - 1) / kSigWarps * slices;
    // device-sized work (a list, or a predicate that is normally false): a persistent grid is enough
    if (cellList || runIfCount) blocks = std::min<uint64_t>(blocks, uint64_t(ctx->smCount) * 8);
    if (blocks > 0x7fffffffull) return fail(ctx, EM2_ERR_INVALID, "too many cells for one signature launch");
    if (cellList || runIfCount)
        signatureKernel<false><<<unsigned(blocks), kSigWarps * 32, 0, s>>>(
            rangeBegin, cells, geneCount, toc, counts, sum1, sum2, Uk, ldk, sumU, uint32_t(lshCount), uint32_t(W), slices,
            signatures, reinterpret_cast<unsigned long long*>(nearZero), cellList, cellListCount, runIfCount, runIfCap);
    else
        signatureKernel<true><<<unsigned(blocks), kSigWarps * 32, 0, s>>>(
            rangeBegin, cells, geneCount, toc, counts, sum1, sum2, Uk, ldk, sumU, uint32_t(lshCount), uint32_t(W), slices,
            signatures, reinterpret_cast<unsigned long long*>(nearZero), nullptr, nullptr, nullptr, 0);
    ctx->stats.kernel_launches++;
    EM2_CUDA(ctx, cudaGetLastError());
    return EM2_OK;
}

int prepareSignatures(em2_context* ctx, uint64_t cellCount, uint64_t geneCount, const double* U, uint64_t ld,
                      uint64_t lshCount, uint64_t nnzHint, SignaturePlan* plan, cudaStream_t s)
{
    if (lshCount == 0 || lshCount > 65535) return fail(ctx, EM2_ERR_INVALID, "lshCount must be in [1, 65535]");
    if (geneCount == 0) return fail(ctx, EM2_ERR_INVALID, "geneCount must be positive");
    const uint64_t Lpad = roundUp(lshCount, kSliceCols);
    SignaturePlan& pl = *plan;
    pl = SignaturePlan();
    pl.geneCount = geneCount;
    pl.lshCount = lshCount;
    pl.U = U;
    pl.ld = ld;

    // The FP64 kernel wants `Lpad` readable, zero padded columns with an even pitch and 16-byte alignment.
    pl.Upadded = U;
    pl.ldPadded = ld;
    if (ld < Lpad || (ld & 1) || (reinterpret_cast<uintptr_t>(U) & 15)) {
        void* p = nullptr;
        EM2_TRY(reserve(ctx, em2_context::S_UPAD, geneCount * Lpad * sizeof(double), &p));
        EM2_CUDA(ctx, cudaMemsetAsync(p, 0, geneCount * Lpad * sizeof(double), s));
        EM2_CUDA(ctx, cudaMemcpy2DAsync(p, Lpad * sizeof(double), U, ld * sizeof(double), lshCount * sizeof(double),
                                        geneCount, cudaMemcpyDeviceToDevice, s));
        pl.Upadded = static_cast<const double*>(p);
        pl.ldPadded = Lpad;
    }

    // Path choice (both are bit-identical): the tensor-core filter does G*3L int8 MACs per cell
    // at ~1.6e15/s, the FP64 kernel nnz*L at ~2.9e12/s, so the filter wins above ~0.5 % density; it also
    // needs enough cells to fill the GPU with 128 x 128 tiles.
    if (ctx->signatureMode == 2) pl.filter = true;
    else if (ctx->signatureMode == 0 && nnzHint != 0 && cellCount != 0) {
        const double density = double(nnzHint) / (double(cellCount) * double(geneCount));
        pl.filter = density >= 0.012 && cellCount >= 4096 && geneCount >= 1024 && lshCount >= 128;
    }
    if (pl.filter) return prepareSignaturesFiltered(ctx, pl, cellCount, s);

    void* sumU = nullptr;
    EM2_TRY(reserve(ctx, em2_context::S_SUMU, Lpad * sizeof(double), &sumU));
    pl.sumU = static_cast<double*>(sumU);
    return launchColumnStats(ctx, geneCount, pl.Upadded, pl.ldPadded, Lpad, Lpad, pl.sumU, nullptr, nullptr, nullptr, s);
}

int launchSignaturesRange(em2_context* ctx, const SignaturePlan& pl, const uint64_t* toc, const em2_count* counts,
                          const double* sum1, const double* sum2, uint64_t cellBegin, uint64_t cellEnd,
                          uint64_t* signatures, uint64_t* nearZero, cudaStream_t s)
{
    if (cellEnd <= cellBegin) return EM2_OK;
    if (pl.filter)
        return launchSignaturesFiltered(ctx, pl, toc, counts, sum1, sum2, cellBegin, cellEnd, signatures, nearZero, s);
    return launchSignaturesFp64(ctx, cellEnd, pl.geneCount, toc, counts, sum1, sum2, pl.Upadded, pl.ldPadded, pl.sumU,
                                pl.lshCount, signatures, nearZero, nullptr, nullptr, 0, nullptr, 0, cellBegin,
                                cellEnd - cellBegin, s);
}

int launchSignatures(em2_context* ctx, uint64_t cellCount, uint64_t geneCount, const uint64_t* toc,
                     const em2_count* counts, const double* sum1, const double* sum2, const double* U,
                     uint64_t ld, uint64_t lshCount, uint64_t nnzHint, uint64_t* signatures, uint64_t* nearZero,
                     cudaStream_t s)
{
    if (cellCount == 0) return EM2_OK;
    SignaturePlan pl;
    EM2_TRY(prepareSignatures(ctx, cellCount, geneCount, U, ld, lshCount, nnzHint, &pl, s));
    return launchSignaturesRange(ctx, pl, toc, counts, sum1, sum2, 0, cellCount, signatures, nearZero, s);
}

}  // namespace em2
