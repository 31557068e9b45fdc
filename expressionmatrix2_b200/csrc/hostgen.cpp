// Host-side O(G*L) / O(L) pieces of the path that stay on the CPU by design:
//   em2_generate_lsh_vectors  <- Lsh::generateLshVectors      (reference src/Lsh.cpp:68-113)
//   em2_similarity_table      <- Lsh::computeSimilarityTable  (reference src/Lsh.cpp:229-249)
//   em2_mismatch_max          <- the `similarity > similarityThreshold` filter of
//                                findSimilarPairs4            (reference src/ExpressionMatrixLsh.cpp:244)
//
// The reference draws its hyperplane components from boost::mt19937 + boost::normal_distribution<>.
// Boost is an un-vendored dependency without a pinned version; this implements the Boost <= 1.55
// recipe (Box-Muller over uniform_01, see oracle/boost_shim/boost/random/normal_distribution.hpp).
// The Mersenne-Twister stream is inherently sequential; the transcendental work and the column
// normalisation are spread over host threads without changing a single rounding: every normal is the
// same libm expression of the same two uniforms, and every column norm is summed in gene order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

#include "../../include/em2b200.h"

namespace {

unsigned hostThreads()
{
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    return std::min(n, 32u);
}

template <class F> void parallelFor(uint64_t n, uint64_t grain, F&& f)
{
    const unsigned T = unsigned(std::min<uint64_t>(hostThreads(), std::max<uint64_t>(1, n / std::max<uint64_t>(grain, 1))));
    if (T <= 1) {
        f(uint64_t(0), n);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t chunk = (n + T - 1) / T;
    for (unsigned t = 0; t < T; t++) {
        const uint64_t b = std::min<uint64_t>(n, t * chunk), e = std::min<uint64_t>(n, b + chunk);
        if (b < e) th.emplace_back([=, &f] { f(b, e); });
    }
    for (auto& x : th) x.join();
}

const double kPi = 3.141592653589793238462643383279502884;

}  // namespace

extern "C" {

int em2_generate_lsh_vectors(uint64_t geneCount, uint64_t lshCount, uint32_t seed, double* U)
{
    if (!U || geneCount == 0 || lshCount == 0) return EM2_ERR_INVALID;
    const uint64_t total = geneCount * lshCount;
    const uint64_t pairs = (total + 1) / 2;

    // 1. the sequential part: raw 32-bit outputs, two per Box-Muller pair
    std::vector<uint32_t> raw(2 * pairs);
    {
        std::mt19937 engine(seed);                       // == boost::mt19937
        for (auto& r : raw) r = uint32_t(engine());
    }

    // 2. normals, in draw order: value 2p = rho*cos(2 pi r1), value 2p+1 = rho*sin(2 pi r1)
    const double factor = 1.0 / (4294967295.0 + 1.0);    // uniform_01 over a 32-bit engine
    const double twoPi = 2.0 * 3.14159265358979323846264338327950288;
    parallelFor(pairs, 1 << 14, [&](uint64_t b, uint64_t e) {
        for (uint64_t p = b; p < e; p++) {
            const double r1 = double(raw[2 * p]) * factor;
            const double r2 = double(raw[2 * p + 1]) * factor;
            const double rho = std::sqrt(-2.0 * std::log(1.0 - r2));
            U[2 * p] = rho * std::cos(twoPi * r1) * 1.0 + 0.0;
            if (2 * p + 1 < total) U[2 * p + 1] = rho * std::sin(twoPi * r1) * 1.0 + 0.0;
        }
    });

    // 3. scale every hyperplane (column) to unit norm; sums run over genes in ascending order
    parallelFor(lshCount, 16, [&](uint64_t b, uint64_t e) {
        std::vector<double> norm(e - b, 0.);
        for (uint64_t g = 0; g < geneCount; g++) {
            const double* row = U + g * lshCount;
            for (uint64_t i = b; i < e; i++) norm[i - b] += row[i] * row[i];
        }
        for (auto& f : norm) f = 1. / std::sqrt(f);
        for (uint64_t g = 0; g < geneCount; g++) {
            double* row = U + g * lshCount;
            for (uint64_t i = b; i < e; i++) row[i] *= norm[i - b];
        }
    });
    return EM2_OK;
}

int em2_similarity_table(uint64_t lshCount, double* table)
{
    if (!table || lshCount == 0) return EM2_ERR_INVALID;
    for (uint64_t m = 0; m <= lshCount; m++) table[m] = std::cos(double(m) * kPi / double(lshCount));
    return EM2_OK;
}

int64_t em2_mismatch_max(uint64_t lshCount, double similarityThreshold)
{
    int64_t best = -1;
    for (uint64_t m = 0; m <= lshCount; m++) {
        if (std::cos(double(m) * kPi / double(lshCount)) > similarityThreshold) best = int64_t(m);
        else break;
    }
    return best;
}

}  // extern "C"
