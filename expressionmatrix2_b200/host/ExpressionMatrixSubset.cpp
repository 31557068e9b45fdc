#include "ExpressionMatrixSubset.hpp"

#include <cmath>

using namespace ChanZuckerberg::ExpressionMatrix2;

ExpressionMatrixSubset::ExpressionMatrixSubset(const std::string& name, const GeneSet& geneSetArg,
                                               const CellSet& cellSetArg, const CellExpressionCounts& global)
    : geneSet(geneSetArg), cellSet(cellSetArg), global_(global)
{
    if (!std::is_sorted(geneSet.begin(), geneSet.end())) throw std::runtime_error("Gene set is not sorted.");
    if (!std::is_sorted(cellSet.begin(), cellSet.end())) throw std::runtime_error("Cell set is not sorted.");

    // all cells x all genes: local ids == global ids, use the global arrays where they lie
    bool allCells = cellSet.size() == global.size();
    for (size_t i = 0; allCells && i < cellSet.size(); i += std::max<size_t>(1, cellSet.size() / 64))
        allCells = cellSet[i] == CellId(i);
    if (allCells && cellSet.size() > 0) allCells = cellSet[cellSet.size() - 1] == CellId(cellSet.size() - 1);
    bool allGenes = geneSet.isIdentity();
    if (allCells && allGenes) {
        // every stored gene id must be inside the gene set for the identity shortcut to hold
        GeneId maxGene = 0;
        const auto* d = global.dataBegin();
        for (size_t i = 0; i < global.totalSize(); i++) maxGene = std::max(maxGene, d[i].first);
        allGenes = global.totalSize() == 0 || maxGene < geneSet.size();
    }
    inPlace_ = allCells && allGenes;
    if (inPlace_) return;

    // one pass: count, then fill (no per-element push_back / remap as in the reference)
    local_.createNew(name);
    std::vector<std::pair<GeneId, float>> row;
    for (CellId local = 0; local != cellSet.size(); local++) {
        const CellId globalCell = cellSet[local];
        row.clear();
        for (const auto* p = global.begin(globalCell); p != global.end(globalCell); ++p) {
            const GeneId l = geneSet.getLocalGeneId(p->first);
            if (l != invalidGeneId) row.push_back(std::make_pair(l, p->second));
        }
        local_.appendVector(row.begin(), row.end());
    }
}

ExpressionMatrixSubset::~ExpressionMatrixSubset()
{
    try {
        remove();
    } catch (...) {
    }
}

void ExpressionMatrixSubset::remove()
{
    if (!inPlace_ && local_.isOpen()) local_.remove();
}

void ExpressionMatrixSubset::computeSums()
{
    sums.assign(cellCount(), Sum());
    const uint64_t* t = toc();
    const auto* d = data();
    for (CellId c = 0; c < cellCount(); c++) {
        Sum& s = sums[c];
        for (uint64_t j = t[c]; j < t[c + 1]; j++) {
            const float x = d[j].second;
            s.sum1 += x;
            const float xx = x * x;
            s.sum2 += xx;
        }
    }
}

double ExpressionMatrixSubset::computeCellSimilarity(CellId c0, CellId c1)
{
    if (sums.size() != cellCount()) computeSums();
    const uint64_t* t = toc();
    const auto* d = data();
    uint64_t i0 = t[c0], e0 = t[c0 + 1], i1 = t[c1], e1 = t[c1 + 1];
    double dot = 0.;
    while (i0 != e0 && i1 != e1) {
        if (d[i0].first < d[i1].first) ++i0;
        else if (d[i1].first < d[i0].first) ++i1;
        else {
            const float prod = d[i0].second * d[i1].second;
            dot += prod;
            ++i0;
            ++i1;
        }
    }
    const double n = double(geneCount());
    const Sum& a = sums[c0];
    const Sum& b = sums[c1];
    return (n * dot - a.sum1 * b.sum1) / std::sqrt((n * a.sum2 - a.sum1 * a.sum1) * (n * b.sum2 - b.sum1 * b.sum1));
}
