// class Lsh -- LSH signatures of the cells of an ExpressionMatrixSubset, computed on the GPU.
// Same public interface and the same files (<name>-Info, <name>-Signatures) as the reference's Lsh
// (reference src/Lsh.hpp:35-142, src/Lsh.cpp:18-64); the OpenCL members (Lsh.hpp:145-266) are replaced by
// findSimilarPairs(), which runs the whole scan + top-k on the device.
#pragma once
#include <iosfwd>
#include <string>
#include <vector>

#include "Ids.hpp"
#include "MemoryMapped.hpp"

namespace ChanZuckerberg {
namespace ExpressionMatrix2 {

class ExpressionMatrixSubset;
class SimilarPairs;

// View of one cell signature: MSB-first bit order (reference src/BitSet.hpp:29-62).
class BitSetPointer {
public:
    uint64_t* begin;
    uint64_t* end;
    BitSetPointer(uint64_t* b = nullptr, uint64_t wordCount = 0) : begin(b), end(b + wordCount) {}
    uint64_t wordCount() const { return uint64_t(end - begin); }
    bool get(uint64_t bit) const { return (begin[bit >> 6] >> (63ULL - (bit & 63ULL))) & 1ULL; }
};
inline uint64_t countMismatches(const BitSetPointer& x, const BitSetPointer& y)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i < x.wordCount(); i++) n += uint64_t(__builtin_popcountll(x.begin[i] ^ y.begin[i]));
    return n;
}

class Lsh {
public:
    // Create: hyperplanes from `seed`, signatures of every cell of the subset on the GPU, stored on disk.
    Lsh(const std::string& name, ExpressionMatrixSubset&, size_t lshCount, uint32_t seed);
    // Access an existing Lsh object.
    explicit Lsh(const std::string& name);
    void remove();

    // Accessors on single pairs (results inspection; the bulk path is findSimilarPairs).
    double computeCellSimilarity(CellId localCellId0, CellId localCellId1);
    size_t computeMismatchCount(CellId localCellId0, CellId localCellId1);
    BitSetPointer getSignature(CellId cellId)
    {
        return BitSetPointer(signatures.begin() + size_t(cellId) * signatureWordCount, signatureWordCount);
    }
    void writeSignatureStatistics(const std::string& csvFileName);
    void writeSignatureStatistics(std::ostream&);

    CellId cellCount() const { return CellId(info->cellCount); }
    size_t lshCount() const { return info->lshCount; }
    size_t wordCount() const { return signatureWordCount; }
    size_t computeMismatchCountThresholdFromSimilarityThreshold(double similarityThreshold) const;
    double getSimilarity(size_t mismatchCount) const { return similarityTable[mismatchCount]; }

    // All-pairs Hamming scan + per-cell top-k on the GPU, written straight into the mapped SimilarPairs
    // rows (already in SimilarPairs::sort() order).  variant: em2_variant.
    void findSimilarPairs(SimilarPairs&, size_t k, double similarityThreshold, int variant = 0);
    // Bucketed search with the candidate order and lists of the reference's findSimilarPairs7
    // (reference src/ExpressionMatrixLsh.cpp:507-827), on the GPU.
    void findSimilarPairs7(SimilarPairs&, size_t k, double similarityThreshold, const std::vector<int>& lshSliceLengths,
                           CellId maxCheck, size_t log2BucketCount);

    // Number of projections that fell inside the epsilon band around zero (reported, expected 0).
    uint64_t nearZeroProjections = 0;

private:
    size_t signatureWordCount = 0;
    MemoryMapped::Vector<uint64_t> signatures;
    std::vector<double> similarityTable;
    void computeSimilarityTable();
    struct Info {
        size_t cellCount;
        size_t lshCount;
    };
    MemoryMapped::Object<Info> info;
};

}  // namespace ExpressionMatrix2
}  // namespace ChanZuckerberg
