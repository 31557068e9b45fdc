#include "Lsh.hpp"

#include <fstream>
#include <iostream>

#include "ExpressionMatrixSubset.hpp"
#include "Gpu.hpp"
#include "SimilarPairs.hpp"

using namespace ChanZuckerberg::ExpressionMatrix2;

Lsh::Lsh(const std::string& name, ExpressionMatrixSubset& subset, size_t lshCount, uint32_t seed)
{
    if (lshCount == 0) throw std::runtime_error("lshCount must be positive.");
    info.createNew(name + "-Info");
    info->lshCount = lshCount;
    info->cellCount = subset.cellCount();
    signatureWordCount = (lshCount - 1) / 64 + 1;

    const size_t geneCount = subset.geneCount();
    const size_t cellCount = subset.cellCount();

    // hyperplanes [gene][lshVector]: host, threads (em2_generate_lsh_vectors <- Lsh.cpp:68-113)
    std::vector<double> lshVectors(geneCount * lshCount);
    if (em2_generate_lsh_vectors(geneCount, lshCount, seed, lshVectors.data()) != EM2_OK)
        throw std::runtime_error("Could not generate LSH vectors.");

    // signatures and per-cell sums on the GPU, written into the mapped file
    signatures.createNew(name + "-Signatures", cellCount * signatureWordCount);
    std::vector<double> sum1(cellCount), sum2(cellCount);
    Gpu& gpu = Gpu::instance();
    static_assert(sizeof(std::pair<GeneId, float>) == sizeof(em2_count), "pair<GeneId,float> must be 8 bytes");
    gpu.check(em2_compute_signatures(gpu.context(), cellCount, geneCount, subset.toc(),
                                     reinterpret_cast<const em2_count*>(subset.data()), lshVectors.data(), lshCount,
                                     signatures.begin(), sum1.data(), sum2.data()),
              "em2_compute_signatures");
    nearZeroProjections = gpu.stats().near_zero_projections;
    subset.sums.resize(cellCount);
    for (size_t c = 0; c < cellCount; c++) {
        subset.sums[c].sum1 = sum1[c];
        subset.sums[c].sum2 = sum2[c];
    }
    computeSimilarityTable();
}

Lsh::Lsh(const std::string& name)
{
    info.accessExistingReadOnly(name + "-Info");
    signatures.accessExistingReadOnly(name + "-Signatures");
    signatureWordCount = (lshCount() - 1) / 64 + 1;
    if (signatures.size() != size_t(cellCount()) * signatureWordCount)
        throw std::runtime_error("Lsh object " + name + " has a signature file of inconsistent length.");
    computeSimilarityTable();
}

void Lsh::remove()
{
    signatures.remove();
    info.remove();
}

void Lsh::computeSimilarityTable()
{
    similarityTable.resize(lshCount() + 1);
    em2_similarity_table(lshCount(), similarityTable.data());
}

size_t Lsh::computeMismatchCount(CellId c0, CellId c1) { return countMismatches(getSignature(c0), getSignature(c1)); }

double Lsh::computeCellSimilarity(CellId c0, CellId c1) { return similarityTable[computeMismatchCount(c0, c1)]; }

size_t Lsh::computeMismatchCountThresholdFromSimilarityThreshold(double similarityThreshold) const
{
    for (size_t m = 0; m < similarityTable.size(); m++)
        if (similarityTable[m] < similarityThreshold) return m - 1;
    throw std::runtime_error("No mismatch count has similarity below the requested threshold.");
}

void Lsh::writeSignatureStatistics(const std::string& csvFileName)
{
    std::ofstream csv(csvFileName);
    writeSignatureStatistics(csv);
}

void Lsh::writeSignatureStatistics(std::ostream& csv)
{
    csv << "Bit,Set,Unset,Total\n";
    for (size_t i = 0; i < lshCount(); i++) {
        size_t set = 0;
        for (CellId c = 0; c < cellCount(); c++) set += getSignature(c).get(i);
        csv << i << "," << set << "," << cellCount() - set << "," << cellCount() << "\n";
    }
}

void Lsh::findSimilarPairs7(SimilarPairs& similarPairs, size_t k, double similarityThreshold,
                            const std::vector<int>& lshSliceLengths, CellId maxCheck, size_t log2BucketCount)
{
    if (similarPairs.cellCount() != cellCount()) throw std::runtime_error("SimilarPairs and Lsh disagree on the cell count.");
    if (similarPairs.k() != k) throw std::runtime_error("SimilarPairs was created with a different k.");
    const size_t n = cellCount();
    std::vector<uint32_t> used(n);
    std::vector<int32_t> slices(lshSliceLengths.begin(), lshSliceLengths.end());
    Gpu& gpu = Gpu::instance();
    gpu.check(em2_find_similar_pairs7(gpu.context(), signatures.begin(), n, lshCount(), k, similarityThreshold, slices.data(),
                                      slices.size(), maxCheck, log2BucketCount,
                                      reinterpret_cast<em2_pair*>(similarPairs.begin(0)), used.data()),
              "em2_find_similar_pairs7");
    similarPairs.setUsedCounts(used);
}

void Lsh::findSimilarPairs(SimilarPairs& similarPairs, size_t k, double similarityThreshold, int variant)
{
    if (similarPairs.cellCount() != cellCount()) throw std::runtime_error("SimilarPairs and Lsh disagree on the cell count.");
    if (similarPairs.k() != k) throw std::runtime_error("SimilarPairs was created with a different k.");
    const size_t n = cellCount();
    std::vector<uint32_t> used(n);
    static_assert(sizeof(SimilarPairs::Pair) == sizeof(em2_pair), "pair<CellId,float> must be 8 bytes");
    Gpu& gpu = Gpu::instance();
    gpu.check(em2_find_similar_pairs(gpu.context(), signatures.begin(), n, lshCount(), 0, n, k, similarityThreshold,
                                     variant, reinterpret_cast<em2_pair*>(similarPairs.begin(0)), used.data()),
              "em2_find_similar_pairs");
    similarPairs.setUsedCounts(used);
}
