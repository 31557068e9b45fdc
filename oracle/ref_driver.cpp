// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// C-ABI driver around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src by oracle/Makefile into oracle/_ref/libem2ref.so).  It lets the Python
// tests and bench.py's cpu_baseline leg run the reference's own
//   ExpressionMatrixSubset (src/ExpressionMatrixSubset.cpp:9-58),
//   Lsh                    (src/Lsh.cpp:18-274),
//   SimilarPairs           (src/SimilarPairs.cpp:11-42, 369-405),
//   keepBest               (src/heap.hpp:116-126)
// on arrays handed over from numpy.
//
// Two things cannot be compiled from the reference and are RESTATED here, each next to the lines
// it follows:
//   * the pair loop of ExpressionMatrix::findSimilarPairs4 (src/ExpressionMatrixLsh.cpp:199-286);
//     the member function itself needs ExpressionMatrix.hpp -> Boost.Graph/Asio, which are absent.
//   * the deterministic top-k of the reference's OpenCL host path
//     (src/ExpressionMatrixLshGpu.cpp:132-157), with the `similarity > threshold` filter of
//     src/ExpressionMatrixLsh.cpp:244.
// Both restatements operate on the real reference Lsh / SimilarPairs objects.

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <map>
#include <regex>
#include <random>
#include <cmath>
#include <unistd.h>

// Lsh::lshVectors is private (src/Lsh.hpp:113); the oracle needs to read it to hand the very
// same hyperplanes to the GPU path.  Access specifiers do not change layout with GCC.
#include <boost/lexical_cast.hpp>
#include <boost/random/mersenne_twister.hpp>
#include <boost/random/normal_distribution.hpp>
#include <boost/random/variate_generator.hpp>
#define private public
#include "Lsh.hpp"
#include "GeneSet.hpp"
#include "CellSets.hpp"
#include "ExpressionMatrixSubset.hpp"
#include "SimilarPairs.hpp"
#undef private
#include "heap.hpp"
#include "orderPairs.hpp"
#include "filesystem.hpp"
#include "MurmurHash2.hpp"
#include "CellGraph.hpp"
#include <boost/graph/iteration_macros.hpp>

using namespace ChanZuckerberg;
using namespace ExpressionMatrix2;

namespace {

thread_local std::string g_error;

struct Handle {
    std::string dir;
    GeneSet geneSet;
    CellSet cellSet;
    ExpressionMatrixSubset::CellExpressionCounts global;
    std::unique_ptr<ExpressionMatrixSubset> subset;
    std::unique_ptr<Lsh> lsh;
    bool ownsLshFiles = false;
    double signatureSeconds = 0.;   // the region the reference itself times (Lsh.cpp:160,209)
    double lshCtorSeconds = 0.;
    std::string log;
};

// Capture what the reference prints to cout while `f` runs.
template <class F> std::string captureCout(F&& f)
{
    std::ostringstream capture;
    std::streambuf* old = std::cout.rdbuf(capture.rdbuf());
    try {
        f();
    } catch (...) {
        std::cout.rdbuf(old);
        throw;
    }
    std::cout.rdbuf(old);
    return capture.str();
}

double parseSignatureSeconds(const std::string& log)
{
    // "Computation of LSH cell signatures took <t>s."  (Lsh.cpp:218)
    const std::string key = "Computation of LSH cell signatures took ";
    const size_t p = log.find(key);
    if (p == std::string::npos) return -1.;
    return std::atof(log.c_str() + p + key.size());
}

void makeSets(Handle& h, uint64_t cellCount, uint64_t geneCount)
{
    // GeneSet "AllGenes" and CellSet "AllCells" files, as ExpressionMatrix creates them
    // (src/ExpressionMatrix.cpp:56-153); SimilarPairs re-opens them by name.
    h.geneSet.createNew(h.dir + "/GeneSet-AllGenes");
    for (GeneId g = 0; g < GeneId(geneCount); g++) h.geneSet.addGene(g);
    h.geneSet.forceSorted();
    h.cellSet.createNew(h.dir + "/CellSet-AllCells", cellCount);
    for (CellId c = 0; c < CellId(cellCount); c++) h.cellSet[c] = c;
}

template <class F> int guarded(F&& f)
{
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown exception";
        return 1;
    }
}

}  // namespace

extern "C" {

const char* em2ref_last_error() { return g_error.c_str(); }

// Build the reference objects from a CSR matrix and run the reference Lsh constructor
// (hyperplane generation + signatures + similarity table).
int em2ref_open_csr(const char* dir, uint64_t cellCount, uint64_t geneCount, const uint64_t* toc,
                    const uint32_t* geneIds, const float* counts, uint64_t lshCount, uint32_t seed,
                    void** out)
{
    return guarded([&] {
        std::unique_ptr<Handle> h(new Handle);
        h->dir = dir;
        makeSets(*h, cellCount, geneCount);
        h->global.createNew(h->dir + "/CellExpressionCounts");
        std::vector<std::pair<GeneId, float>> row;
        for (uint64_t c = 0; c < cellCount; c++) {
            row.clear();
            for (uint64_t j = toc[c]; j < toc[c + 1]; j++) row.push_back(std::make_pair(geneIds[j], counts[j]));
            h->global.appendVector(row.begin(), row.end());
        }
        h->log = captureCout([&] {
            h->subset.reset(new ExpressionMatrixSubset(h->dir + "/tmp-ExpressionMatrixSubset-oracle", h->geneSet,
                                                       h->cellSet, h->global));
            const auto t0 = std::chrono::steady_clock::now();
            h->lsh.reset(new Lsh(h->dir + "/tmp-Lsh", *h->subset, lshCount, seed));
            const auto t1 = std::chrono::steady_clock::now();
            h->lshCtorSeconds = 1e-9 * double(std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
        });
        h->signatureSeconds = parseSignatureSeconds(h->log);
        h->ownsLshFiles = true;
        *out = h.release();
    });
}

// Open the reference Lsh on signatures supplied by the caller (written in the reference's own
// file format through its own containers), for scan-only runs.
int em2ref_open_signatures(const char* dir, uint64_t cellCount, uint64_t lshCount, const uint64_t* signatures,
                           void** out)
{
    return guarded([&] {
        std::unique_ptr<Handle> h(new Handle);
        h->dir = dir;
        makeSets(*h, cellCount, 1);
        const uint64_t wordCount = (lshCount - 1) / 64 + 1;
        {
            MemoryMapped::Object<Lsh::Info> info;
            info.createNew(h->dir + "/tmp-Lsh-Info");
            info->cellCount = cellCount;
            info->lshCount = lshCount;
            MemoryMapped::Vector<uint64_t> sig;
            sig.createNew(h->dir + "/tmp-Lsh-Signatures", cellCount * wordCount);
            std::memcpy(sig.begin(), signatures, cellCount * wordCount * sizeof(uint64_t));
        }
        h->lsh.reset(new Lsh(h->dir + "/tmp-Lsh"));
        h->ownsLshFiles = true;
        *out = h.release();
    });
}

// The reference's own ExpressionMatrixSubset constructor (src/ExpressionMatrixSubset.cpp:9-42) on an arbitrary
// gene set / cell set: returns the local CSR it builds and its per-cell sums.  outToc has cellSetSize + 1
// entries; outGenes/outCounts must hold at least the global nnz; *outNnz receives the number of entries kept.
int em2ref_subset(const char* dir, uint64_t cellCount, uint64_t geneCount, const uint64_t* toc, const uint32_t* geneIds,
                  const float* counts, uint64_t geneSetSize, const uint32_t* geneSet, uint64_t cellSetSize,
                  const uint32_t* cellSet, uint64_t* outToc, uint32_t* outGenes, float* outCounts, uint64_t* outNnz,
                  double* sum1, double* sum2)
{
    return guarded([&] {
        (void)geneCount;
        const std::string d = dir;
        GeneSet gs;
        gs.createNew(d + "/GeneSet-subset");
        for (uint64_t i = 0; i < geneSetSize; i++) gs.addGene(geneSet[i]);
        gs.forceSorted();
        CellSet cs;
        cs.createNew(d + "/CellSet-subset", cellSetSize);
        for (uint64_t i = 0; i < cellSetSize; i++) cs[i] = cellSet[i];
        ExpressionMatrixSubset::CellExpressionCounts global;
        global.createNew(d + "/CellExpressionCounts-subset");
        std::vector<std::pair<GeneId, float>> row;
        for (uint64_t c = 0; c < cellCount; c++) {
            row.clear();
            for (uint64_t j = toc[c]; j < toc[c + 1]; j++) row.push_back(std::make_pair(geneIds[j], counts[j]));
            global.appendVector(row.begin(), row.end());
        }
        {
            ExpressionMatrixSubset subset(d + "/tmp-ExpressionMatrixSubset-subset", gs, cs, global);
            uint64_t n = 0;
            outToc[0] = 0;
            for (uint64_t c = 0; c < cellSetSize; c++) {
                for (const auto& p : subset.cellExpressionCounts[c]) {
                    outGenes[n] = p.first;
                    outCounts[n] = p.second;
                    n++;
                }
                outToc[c + 1] = n;
                if (sum1) sum1[c] = subset.sums[c].sum1;
                if (sum2) sum2[c] = subset.sums[c].sum2;
            }
            *outNnz = n;
        }   // the destructor removes the subset's temp files
        global.remove();
        gs.remove();
        cs.remove();
    });
}

int em2ref_close(void* handle)
{
    return guarded([&] {
        std::unique_ptr<Handle> h(static_cast<Handle*>(handle));
        if (h->lsh) {
            captureCout([&] { h->lsh->remove(); });
            h->lsh.reset();
        }
        if (h->subset) h->subset.reset();   // destructor removes its temp files
        if (h->global.size() > 0 || !h->global.empty()) {
            try { h->global.remove(); } catch (...) {}
        }
        h->geneSet.remove();
        h->cellSet.remove();
    });
}

uint64_t em2ref_word_count(void* handle) { return static_cast<Handle*>(handle)->lsh->wordCount(); }
double em2ref_signature_seconds(void* handle) { return static_cast<Handle*>(handle)->signatureSeconds; }
double em2ref_lsh_ctor_seconds(void* handle) { return static_cast<Handle*>(handle)->lshCtorSeconds; }

int em2ref_get_signatures(void* handle, uint64_t* out)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        const uint64_t n = uint64_t(h.lsh->cellCount()) * h.lsh->wordCount();
        std::memcpy(out, h.lsh->signatures.begin(), n * sizeof(uint64_t));
    });
}

// Hyperplanes as the reference holds them: [gene][lshVector] (Lsh.hpp:105-113), row-major copy.
int em2ref_get_lsh_vectors(void* handle, double* out)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        const size_t L = h.lsh->lshCount();
        for (size_t g = 0; g < h.lsh->lshVectors.size(); g++)
            std::memcpy(out + g * L, h.lsh->lshVectors[g].data(), L * sizeof(double));
    });
}

int em2ref_get_sums(void* handle, double* sum1, double* sum2)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        for (size_t c = 0; c < h.subset->sums.size(); c++) {
            sum1[c] = h.subset->sums[c].sum1;
            sum2[c] = h.subset->sums[c].sum2;
        }
    });
}

int em2ref_get_similarity_table(void* handle, double* out)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        for (size_t m = 0; m <= h.lsh->lshCount(); m++) out[m] = h.lsh->getSimilarity(m);
    });
}

// Lsh::computeMismatchCount for a list of pairs (Lsh.cpp:266-274 -> BitSet.hpp:277-288).
int em2ref_mismatch_counts(void* handle, uint64_t pairCount, const uint32_t* cell0, const uint32_t* cell1,
                           uint32_t* out)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        for (uint64_t i = 0; i < pairCount; i++) out[i] = uint32_t(h.lsh->computeMismatchCount(cell0[i], cell1[i]));
    });
}

// All mismatch counts of one cell against every cell (full row), same reference call.
int em2ref_mismatch_row(void* handle, uint32_t cell0, uint32_t* out)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        for (CellId c = 0; c < h.lsh->cellCount(); c++) out[c] = uint32_t(h.lsh->computeMismatchCount(cell0, c));
    });
}

// Exact similarity of the reference (ExpressionMatrixSubset.cpp:83-133) for a list of pairs.
int em2ref_exact_similarity(void* handle, uint64_t pairCount, const uint32_t* cell0, const uint32_t* cell1,
                            double* out)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        if (!h.subset) throw std::runtime_error("handle was opened without an expression matrix");
        for (uint64_t i = 0; i < pairCount; i++) out[i] = h.subset->computeCellSimilarity(cell0[i], cell1[i]);
    });
}

// RESTATEMENT of the pair loop of ExpressionMatrix::findSimilarPairs4
// (src/ExpressionMatrixLsh.cpp:199-286) over the real reference Lsh, keepBest and SimilarPairs.
// rowBegin/rowEnd bound the outer `begin0` loop (full job: 0, cellCount) so that a bounded sample
// of the same loop can be timed at sizes where the whole job would take hours.
// Outputs: the -Pairs payload [cellCount*k] and usedCount[cellCount] after copy()+sort()
// (only when the whole job was run), the seconds spent in the region the reference times
// (ExpressionMatrixLsh.cpp:217,270) and the number of pairs visited.
int em2ref_find_similar_pairs4_loop(void* handle, uint64_t k, double similarityThreshold, uint32_t rowBegin,
                                    uint32_t rowEnd, uint32_t* pairIds, float* pairSimilarities, uint32_t* usedCount,
                                    double* loopSeconds, uint64_t* pairsVisited)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        Lsh& lsh = *h.lsh;
        const CellId cellCount = lsh.cellCount();
        rowEnd = std::min<uint32_t>(rowEnd, cellCount);

        // :199-207
        vector<vector<pair<CellId, float>>> tmp(cellCount);
        const size_t tmpStore = 2 * k;
        for (auto& v : tmp) v.reserve(tmpStore);
        vector<float> cellThreshold(cellCount, float(similarityThreshold));

        // :216-263
        const auto t0 = std::chrono::steady_clock::now();
        const CellId blockSize = 64;
        size_t pairCount = 0;
        for (CellId begin0 = rowBegin; begin0 < rowEnd; begin0 += blockSize) {
            const CellId end0 = min(begin0 + blockSize, rowEnd);
            for (CellId begin1 = 0; begin1 <= begin0; begin1 += blockSize) {
                const CellId end1 = min(begin1 + blockSize, end0);
                for (CellId cell0 = begin0; cell0 != end0; ++cell0) {
                    auto& tmp0 = tmp[cell0];
                    for (CellId cell1 = begin1; cell1 != end1 && cell1 < cell0; ++cell1) {
                        auto& tmp1 = tmp[cell1];
                        ++pairCount;
                        const double similarity = lsh.computeCellSimilarity(cell0, cell1);
                        if (similarity > similarityThreshold) {
                            if (similarity > cellThreshold[cell0]) {
                                tmp0.push_back(make_pair(cell1, similarity));
                                if (tmp0.size() == tmpStore) {
                                    keepBest(tmp0, k, OrderPairsBySecondGreater<pair<CellId, float>>());
                                    cellThreshold[cell0] = tmp0.back().second;
                                }
                            }
                            if (similarity > cellThreshold[cell1]) {
                                tmp1.push_back(make_pair(cell0, similarity));
                                if (tmp1.size() == tmpStore) {
                                    keepBest(tmp1, k, OrderPairsBySecondGreater<pair<CellId, float>>());
                                    cellThreshold[cell1] = tmp1.back().second;
                                }
                            }
                        }
                    }
                }
            }
        }
        // :265-269
        for (auto& tmp0 : tmp) {
            if (tmp0.size() > k) keepBest(tmp0, k, OrderPairsBySecondGreater<pair<CellId, float>>());
        }
        const auto t1 = std::chrono::steady_clock::now();
        if (loopSeconds)
            *loopSeconds = 1e-9 * double(std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
        if (pairsVisited) *pairsVisited = pairCount;

        // :277-285 -- real SimilarPairs: constructor, copy, sort; then read the mapped payload back.
        if (pairIds && pairSimilarities && usedCount) {
            SimilarPairs similarPairs(h.dir, "oracle-fsp4", "AllGenes", "AllCells", k);
            similarPairs.copy(tmp);
            similarPairs.sort();
            for (CellId c = 0; c < cellCount; c++) {
                usedCount[c] = uint32_t(similarPairs.size(c));
                const SimilarPairs::Pair* p = similarPairs.begin(c);
                for (size_t i = 0; i < k; i++) {
                    pairIds[size_t(c) * k + i] = i < usedCount[c] ? p[i].first : 0;
                    pairSimilarities[size_t(c) * k + i] = i < usedCount[c] ? p[i].second : 0.f;
                }
            }
            similarPairs.remove();
        }
    });
}

// RESTATEMENT of the reference's deterministic top-k (src/ExpressionMatrixLshGpu.cpp:132-157):
// per cell, candidates = all other cells passing the filter, keepBest(k, less<pair<mismatch, id>>),
// sort, addUnsymmetricNoCheck in that order.  The filter is findSimilarPairs4's
// `similarity > similarityThreshold` (src/ExpressionMatrixLsh.cpp:244) applied to the table value.
// Runs rows [rowBegin,rowEnd) against all cells, so it can also be sampled.
int em2ref_topk_deterministic(void* handle, uint64_t k, double similarityThreshold, uint32_t rowBegin,
                              uint32_t rowEnd, uint32_t* pairIds, float* pairSimilarities, uint32_t* usedCount,
                              double* seconds)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        Lsh& lsh = *h.lsh;
        const CellId cellCount = lsh.cellCount();
        rowEnd = std::min<uint32_t>(rowEnd, cellCount);
        SimilarPairs similarPairs(h.dir, "oracle-topk", "AllGenes", "AllCells", k);
        vector<pair<uint16_t, CellId>> neighbors;
        const auto t0 = std::chrono::steady_clock::now();
        for (CellId cellId0 = rowBegin; cellId0 < rowEnd; cellId0++) {
            neighbors.clear();
            for (CellId cellId1 = 0; cellId1 != cellCount; cellId1++) {
                if (cellId1 == cellId0) continue;
                const size_t mismatchCount = lsh.computeMismatchCount(cellId0, cellId1);
                if (lsh.getSimilarity(mismatchCount) > similarityThreshold)
                    neighbors.push_back(make_pair(uint16_t(mismatchCount), cellId1));
            }
            keepBest(neighbors, k, std::less<pair<uint16_t, CellId>>());
            sort(neighbors.begin(), neighbors.end());
            for (const auto& neighbor : neighbors)
                similarPairs.addUnsymmetricNoCheck(cellId0, neighbor.second, lsh.getSimilarity(neighbor.first));
        }
        const auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = 1e-9 * double(std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
        for (CellId c = rowBegin; c < rowEnd; c++) {
            const size_t r = size_t(c - rowBegin);
            usedCount[r] = uint32_t(similarPairs.size(c));
            const SimilarPairs::Pair* p = similarPairs.begin(c);
            for (size_t i = 0; i < k; i++) {
                pairIds[r * k + i] = i < usedCount[r] ? p[i].first : 0;
                pairSimilarities[r * k + i] = i < usedCount[r] ? p[i].second : 0.f;
            }
        }
        similarPairs.remove();
    });
}

// Write a SimilarPairs object with the reference's own class (for file-format parity tests):
// rows are given already ordered; uses addUnsymmetricNoCheck like the reference GPU path.
int em2ref_write_similar_pairs(const char* dir, const char* name, uint64_t cellCount, uint64_t geneCount,
                               uint64_t k, const uint32_t* pairIds, const float* sims, const uint32_t* usedCount)
{
    return guarded([&] {
        Handle h;
        h.dir = dir;
        makeSets(h, cellCount, geneCount);
        SimilarPairs sp(dir, name, "AllGenes", "AllCells", k);
        for (CellId c = 0; c < CellId(cellCount); c++)
            for (uint32_t i = 0; i < usedCount[c]; i++)
                sp.addUnsymmetricNoCheck(c, pairIds[size_t(c) * k + i], sims[size_t(c) * k + i]);
    });
}

// Read a SimilarPairs object back with the reference's own class (validates the hashes,
// SimilarPairs.cpp:47-83).  Returns k through *kOut; arrays may be null for a size query.
int em2ref_read_similar_pairs(const char* dir, const char* name, uint64_t* kOut, uint64_t* cellCountOut,
                              uint32_t* pairIds, float* sims, uint32_t* usedCount)
{
    return guarded([&] {
        SimilarPairs sp(dir, name, true);
        *kOut = sp.k();
        *cellCountOut = sp.cellCount();
        if (!pairIds) return;
        for (CellId c = 0; c < sp.cellCount(); c++) {
            usedCount[c] = uint32_t(sp.size(c));
            const SimilarPairs::Pair* p = sp.begin(c);
            for (size_t i = 0; i < sp.k(); i++) {
                pairIds[size_t(c) * sp.k() + i] = p[i].first;
                sims[size_t(c) * sp.k() + i] = p[i].second;
            }
        }
    });
}

// The reference's OWN CellGraph constructor (src/CellGraph.cpp:33-117, compiled unmodified against the Boost.Graph
// stand-in of boost_shim/) on a SimilarPairs object written by em2ref_write_similar_pairs: the graph's cell set is
// `cellSet` (sorted cell ids); edges come back in the graph's edge order (= insertion order) as vertex indices
// (positions in cellSet) plus the stored similarity.  *edgeCount receives the number of edges; the arrays may be
// null for a size query.
int em2ref_cell_graph_edges(const char* dir, const char* similarPairsName, uint64_t cellSetSize, const uint32_t* cellSet,
                            double similarityThreshold, uint64_t maxConnectivity, uint64_t capacity, uint32_t* vertex0,
                            uint32_t* vertex1, float* similarity, uint64_t* edgeCount)
{
    return guarded([&] {
        MemoryMapped::Vector<CellId> cells;
        const std::string name = std::string(dir) + "/tmp-CellGraphCellSet";
        cells.createNew(name, cellSetSize);
        for (uint64_t i = 0; i < cellSetSize; i++) cells[i] = cellSet[i];
        {
            CellGraph graph(cells, dir, similarPairsName, similarityThreshold, size_t(maxConnectivity));
            std::map<CellId, uint32_t> indexOf;
            for (uint64_t i = 0; i < cellSetSize; i++) indexOf[cellSet[i]] = uint32_t(i);
            uint64_t n = 0;
            BGL_FORALL_EDGES(e, graph, CellGraph) {
                if (vertex0 && n < capacity) {
                    vertex0[n] = indexOf[graph[source(e, graph)].cellId];
                    vertex1[n] = indexOf[graph[target(e, graph)].cellId];
                    similarity[n] = graph[e].similarity;
                }
                n++;
            }
            *edgeCount = n;
        }
        cells.remove();
    });
}

// keepBest known-answer vectors of the reference's own unit test (src/heap.cpp:21-38) are driven
// from Python through this.
int em2ref_keep_best_less(uint64_t n, const int64_t* in, uint64_t k, int64_t* out, uint64_t* outCount)
{
    return guarded([&] {
        vector<int64_t> v(in, in + n);
        keepBest(v, k, std::less<int64_t>());
        *outCount = v.size();
        std::copy(v.begin(), v.end(), out);
    });
}

// Raw generator outputs, to pin the C oracle's MT19937 + normal sampler against the shimmed
// reference types (boost::mt19937 == std::mt19937).
int em2ref_normal_stream(uint32_t seed, uint64_t n, double* out)
{
    return guarded([&] {
        boost::mt19937 eng(seed);
        boost::normal_distribution<> dist;
        boost::variate_generator<boost::mt19937, boost::normal_distribution<>> gen(eng, dist);
        for (uint64_t i = 0; i < n; i++) out[i] = gen();
    });
}

uint64_t em2ref_murmur64a(const void* p, int len, uint64_t seed) { return MurmurHash64A(p, len, seed); }

// ExpressionMatrix::findSimilarPairs7 (ExpressionMatrixLsh.cpp:507-687) and its bucket assignment
// (findSimilarPairs7AssignCellsToBuckets, :707-827).  ExpressionMatrix itself cannot be compiled here (HTTP server,
// HDF5, Boost.Graph ...), so the two member functions are restated statement by statement over the reference's OWN
// Lsh object (getSignature, computeMismatchCount, getSimilarity, computeMismatchCountThresholdFromSimilarityThreshold),
// BitSet (getBits, the cellMap), MurmurHash64A and keepBest; what SimilarPairs::addUnsymmetricNoCheck would store
// (cellId1, float(similarity), in this order) is returned row by row.
int em2ref_find_similar_pairs7(void* handle, uint64_t k, double similarityThreshold, const int32_t* sliceLengths,
                               uint64_t sliceLengthCount, uint32_t maxCheck, uint64_t log2BucketCount, uint32_t* outIds,
                               float* outSims, uint32_t* outUsed)
{
    return guarded([&] {
        Handle& h = *static_cast<Handle*>(handle);
        Lsh& lsh = *h.lsh;
        const CellId cellCount = lsh.cellCount();
        const size_t lshBitCount = lsh.lshCount();
        const vector<int> lshSliceLengths(sliceLengths, sliceLengths + sliceLengthCount);
        for (size_t i = 1; i < sliceLengthCount; i++)
            if (lshSliceLengths[i] >= lshSliceLengths[i - 1]) throw runtime_error("The slice lengths are not in decreasing order.");
        for (size_t i = 0; i < sliceLengthCount; i++)
            if (lshSliceLengths[i] > 64) throw runtime_error("Each slice length can be at most 64 bits.");

        // ---- findSimilarPairs7AssignCellsToBuckets
        vector<vector<vector<vector<CellId>>>> table4(sliceLengthCount);
        vector<vector<vector<size_t>>> sliceBits3(sliceLengthCount);
        const uint64_t bucketCount = (1ULL << log2BucketCount);
        const uint64_t bucketMask = bucketCount - 1ULL;
        for (size_t sliceLengthId = 0; sliceLengthId < sliceLengthCount; sliceLengthId++) {
            const size_t sliceLength = lshSliceLengths[sliceLengthId];
            const size_t sliceCount = lshBitCount / sliceLength;
            table4[sliceLengthId].resize(sliceCount);
            sliceBits3[sliceLengthId].resize(sliceCount);
            const uint64_t tableSize = std::min(uint64_t(1ULL << sliceLength), bucketCount);
            for (size_t sliceId = 0; sliceId < sliceCount; sliceId++) {
                table4[sliceLengthId][sliceId].resize(tableSize);
                auto& sliceBits1 = sliceBits3[sliceLengthId][sliceId];
                sliceBits1.resize(sliceLength);
                size_t bitPosition = sliceId * sliceLength;
                for (size_t bitId = 0; bitId < sliceLength; bitId++, ++bitPosition) sliceBits1[bitId] = bitPosition;
            }
        }
        for (CellId cellId = 0; cellId < cellCount; cellId++) {
            const BitSetPointer signature = lsh.getSignature(cellId);
            for (size_t sliceLengthId = 0; sliceLengthId < sliceLengthCount; sliceLengthId++) {
                const size_t sliceLength = lshSliceLengths[sliceLengthId];
                const size_t sliceCount = lshBitCount / sliceLength;
                for (size_t sliceId = 0; sliceId < sliceCount; sliceId++) {
                    const uint64_t signatureSlice = signature.getBits(sliceBits3[sliceLengthId][sliceId]);
                    const uint64_t bucketId = (sliceLength < log2BucketCount) ? signatureSlice
                                                                              : (MurmurHash64A(&signatureSlice, 8, 231) & bucketMask);
                    table4[sliceLengthId][sliceId].at(bucketId).push_back(cellId);
                }
            }
        }

        // ---- the search loop
        BitSet cellMap(cellCount);
        vector<CellId> candidateNeighbors;
        vector<pair<uint32_t, CellId>> neighbors;
        const size_t mismatchCountThreshold = lsh.computeMismatchCountThresholdFromSimilarityThreshold(similarityThreshold);
        for (CellId cellId0 = 0; cellId0 < cellCount; cellId0++) {
            const BitSetPointer signature = lsh.getSignature(cellId0);
            for (size_t sliceLengthId = 0; sliceLengthId < sliceLengthCount; sliceLengthId++) {
                const size_t sliceLength = lshSliceLengths[sliceLengthId];
                const size_t sliceCount = lshBitCount / sliceLength;
                for (size_t sliceId = 0; sliceId < sliceCount; sliceId++) {
                    const uint64_t signatureSlice = signature.getBits(sliceBits3[sliceLengthId][sliceId]);
                    const uint64_t bucketId = (sliceLength < log2BucketCount) ? signatureSlice
                                                                              : (MurmurHash64A(&signatureSlice, 8, 231) & bucketMask);
                    const auto& table1 = table4[sliceLengthId][sliceId].at(bucketId);
                    for (const CellId cellId1 : table1) {
                        if (cellId1 == cellId0) continue;
                        if (cellMap.get(cellId1)) continue;
                        cellMap.set(cellId1);
                        candidateNeighbors.push_back(cellId1);
                        const uint32_t mismatchCount = uint32_t(lsh.computeMismatchCount(cellId0, cellId1));
                        if (mismatchCount < mismatchCountThreshold) neighbors.push_back(make_pair(mismatchCount, cellId1));
                        if (candidateNeighbors.size() == maxCheck) break;
                    }
                    if (candidateNeighbors.size() == maxCheck) break;
                }
                if (candidateNeighbors.size() == maxCheck) break;
            }
            keepBest(neighbors, k, std::less<pair<uint32_t, CellId>>());
            sort(neighbors.begin(), neighbors.end());
            uint32_t used = 0;
            for (const auto& neighbor : neighbors) {
                outIds[uint64_t(cellId0) * k + used] = neighbor.second;
                outSims[uint64_t(cellId0) * k + used] = float(lsh.getSimilarity(neighbor.first));
                used++;
            }
            outUsed[cellId0] = used;
            for (const CellId cellId1 : candidateNeighbors) cellMap.clear(cellId1);
            candidateNeighbors.clear();
            neighbors.clear();
        }
    });
}

// SignatureGraph vertices and edges.  SignatureGraph itself needs Boost.Graph (absent here), so the two loops of
// ExpressionMatrix::createSignatureGraph (ExpressionMatrixSignatureGraph.cpp:69-75, 111-125) and
// SignatureGraph::createEdges (SignatureGraph.cpp:23-48) are restated -- over the reference's OWN BitSetPointer /
// BitSet (BitSet.hpp): the std::map order is its operator<, the bit tests are its get() / set().
// Vertices are numbered in map order (boost add_vertex on a vecS graph hands out 0, 1, 2, ...).
int em2ref_signature_graph(const uint64_t* signatures, uint64_t cellCount, uint64_t lshCount, uint64_t minCellCount,
                           uint32_t* cellOrder, uint64_t* vertexOffsets, uint64_t* vertexCount, uint32_t* edgeV0,
                           uint32_t* edgeV1, uint64_t edgeCapacity, uint64_t* edgeCount)
{
    return guarded([&] {
        const uint64_t W = (lshCount - 1) / 64 + 1;
        std::map<BitSetPointer, vector<CellId>> signatureMap;
        for (CellId cellId = 0; cellId < cellCount; cellId++)
            signatureMap[BitSetPointer(const_cast<uint64_t*>(signatures) + uint64_t(cellId) * W, W)].push_back(cellId);
        std::map<BitSetPointer, uint32_t> vertexMap;
        vector<BitSetPointer> vertexSignature;
        uint64_t cells = 0;
        vertexOffsets[0] = 0;
        for (const auto& p : signatureMap) {
            if (p.second.size() < minCellCount) continue;
            vertexMap.insert(std::make_pair(p.first, uint32_t(vertexSignature.size())));
            vertexSignature.push_back(p.first);
            for (const CellId c : p.second) cellOrder[cells++] = c;
            vertexOffsets[vertexSignature.size()] = cells;
        }
        *vertexCount = vertexSignature.size();
        uint64_t edges = 0;
        BitSet signature1(lshCount);
        for (uint32_t v0 = 0; v0 < vertexSignature.size(); v0++) {
            const BitSetPointer signature0 = vertexSignature[v0];
            for (size_t bit = 0; bit != lshCount; bit++) {
                if (signature0.get(bit) == 0) {
                    std::copy(signature0.begin, signature0.end, signature1.begin);
                    signature1.set(bit);
                    const auto it1 = vertexMap.find(signature1);
                    if (it1 != vertexMap.end()) {
                        if (edges < edgeCapacity) {
                            edgeV0[edges] = v0;
                            edgeV1[edges] = it1->second;
                        }
                        edges++;
                    }
                }
            }
        }
        *edgeCount = edges;
    });
}

}  // extern "C"
