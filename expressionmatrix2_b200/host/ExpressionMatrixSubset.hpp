// The (gene set x cell set) view of the expression counts that the LSH path works on -- host side.
// Replaces reference src/ExpressionMatrixSubset.{hpp,cpp}: same constructor arguments, same public members
// (geneCount, cellCount, cellExpressionCounts, sums, totalExpressionCounts, remove).  Differences by design:
//   * when the sets are "all genes x all cells" the global arrays are used in place (no 21 GB temp copy,
//     SURVEY.md 8f row 1); otherwise the re-indexed CSR is built in ONE pass into a temp mapped file;
//   * the per-cell sums are produced by the GPU (cellSumsKernel) when the signatures are computed and
//     stored back into `sums`, bit-identical to ExpressionMatrixSubset::computeSums (.cpp:47-58).
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "GeneSet.hpp"
#include "Ids.hpp"
#include "MemoryMapped.hpp"

namespace ChanZuckerberg {
namespace ExpressionMatrix2 {

class ExpressionMatrixSubset {
public:
    using CellExpressionCounts = MemoryMapped::VectorOfVectors<std::pair<GeneId, float>, uint64_t>;
    ExpressionMatrixSubset(const std::string& name, const GeneSet& geneSet, const CellSet& cellSet,
                           const CellExpressionCounts& globalExpressionCounts);
    ~ExpressionMatrixSubset();
    ExpressionMatrixSubset(const ExpressionMatrixSubset&) = delete;

    const GeneSet& geneSet;
    const CellSet& cellSet;
    GeneId geneCount() const { return geneSet.size(); }
    CellId cellCount() const { return CellId(cellSet.size()); }
    size_t totalExpressionCounts() const { return size_t(toc()[cellCount()]); }

    // CSR arrays in the C-ABI's layout (local gene ids, ascending per cell).
    const uint64_t* toc() const { return inPlace_ ? global_.tocBegin() : local_.tocBegin(); }
    const std::pair<GeneId, float>* data() const { return inPlace_ ? global_.dataBegin() : local_.dataBegin(); }

    struct Sum {
        double sum1 = 0.;
        double sum2 = 0.;
    };
    std::vector<Sum> sums;      // filled by Lsh (from the GPU) or by computeSums()
    void computeSums();         // host restatement for callers that need sums without signatures

    // Exact similarity of two cells of the subset (Pearson; reference .cpp:83-133), for spot checks.
    double computeCellSimilarity(CellId localCellId0, CellId localCellId1);

    void remove();

private:
    const CellExpressionCounts& global_;
    CellExpressionCounts local_;
    bool inPlace_ = false;
};

}  // namespace ExpressionMatrix2
}  // namespace ChanZuckerberg
