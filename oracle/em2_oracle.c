/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's LSH cell-similarity hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (expressionmatrix2_b200/) never
 * links, imports or executes it.
 *
 * Parity pin: every function below is checked in tests/test_oracle_vs_reference.py against the
 * reference's OWN sources compiled unmodified (oracle/_ref/libem2ref.so, see oracle/Makefile) and
 * against the golden vectors under tests/golden/ that were generated from that build.  The one
 * unpinned item is the normal sampler behind the hyperplanes: Boost is an absent, un-pinned
 * dependency of the reference (see boost_shim/boost/random/normal_distribution.hpp); the
 * hyperplanes are therefore an INPUT shared by oracle and product.
 *
 * Build: gcc -std=c11 -O3 -msse4.2 -ffp-contract=off (no FMA contraction, like the reference's
 * -O3 -msse4.2 build, src/CMakeLists.txt:56-69).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* MT19937 (boost::mt19937 == std::mt19937, the engine of src/Lsh.cpp:75-79).                   */
typedef struct { uint32_t s[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t* m, uint32_t seed)
{
    m->s[0] = seed;
    for (int i = 1; i < 624; i++) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(mt19937_t* m)
{
    if (m->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (m->s[i] & 0x80000000u) | (m->s[(i + 1) % 624] & 0x7fffffffu);
            m->s[i] = m->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->s[m->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* boost::normal_distribution<> as of Boost <= 1.55 (Box-Muller with one cached value) over
 * boost::uniform_01 of a 32-bit engine; see boost_shim/boost/random/normal_distribution.hpp. */
typedef struct { mt19937_t eng; double r1, r2, rho; int valid; } normal_gen_t;

static double uniform01(mt19937_t* e)
{
    const double factor = 1.0 / (4294967295.0 + 1.0);
    for (;;) {
        const double r = (double)mt_next(e) * factor;
        if (r < 1.0) return r;
    }
}
static double normal_next(normal_gen_t* g)
{
    const double twoPi = 2.0 * 3.14159265358979323846264338327950288;
    if (!g->valid) {
        g->r1 = uniform01(&g->eng);
        g->r2 = uniform01(&g->eng);
        g->rho = sqrt(-2.0 * log(1.0 - g->r2));
        g->valid = 1;
    } else {
        g->valid = 0;
    }
    return g->rho * (g->valid ? cos(twoPi * g->r1) : sin(twoPi * g->r1)) * 1.0 + 0.0;
}

void em2o_normal_stream(uint32_t seed, uint64_t n, double* out)
{
    normal_gen_t g;
    mt_seed(&g.eng, seed);
    g.valid = 0;
    for (uint64_t i = 0; i < n; i++) out[i] = normal_next(&g);
}

/* Lsh::generateLshVectors, src/Lsh.cpp:68-113.  U is [gene][lshVector], row-major. */
void em2o_generate_lsh_vectors(uint64_t geneCount, uint64_t lshCount, uint32_t seed, double* U)
{
    normal_gen_t g;
    mt_seed(&g.eng, seed);
    g.valid = 0;
    double* norm = (double*)calloc(lshCount, sizeof(double));
    for (uint64_t gene = 0; gene < geneCount; gene++) {          /* :90-100, gene outer, vector inner */
        for (uint64_t i = 0; i < lshCount; i++) {
            const double x = normal_next(&g);
            U[gene * lshCount + i] = x;
            norm[i] += x * x;
        }
    }
    for (uint64_t i = 0; i < lshCount; i++) norm[i] = 1. / sqrt(norm[i]);   /* :103-105 */
    for (uint64_t gene = 0; gene < geneCount; gene++)                        /* :106-110 */
        for (uint64_t i = 0; i < lshCount; i++) U[gene * lshCount + i] *= norm[i];
    free(norm);
}

/* ExpressionMatrixSubset::computeSums, src/ExpressionMatrixSubset.cpp:47-58:
 * sum1 += count (float widened), sum2 += count*count (product in float, then widened). */
void em2o_cell_sums(uint64_t cellCount, const uint64_t* toc, const float* counts, double* sum1, double* sum2)
{
    for (uint64_t c = 0; c < cellCount; c++) {
        double s1 = 0., s2 = 0.;
        for (uint64_t j = toc[c]; j < toc[c + 1]; j++) {
            const float x = counts[j];
            s1 += x;
            const float xx = x * x;
            s2 += xx;
        }
        sum1[c] = s1;
        if (sum2) sum2[c] = s2;
    }
}

/* Lsh::computeCellLshSignatures, src/Lsh.cpp:118-224.
 * signatures: cellCount * W words, bit p of a cell at word p>>6, bit 63-(p&63) (BitSet.hpp:48-62).
 * nearZero (optional): number of projections with |s| < eps * (|mean*sumU| + sum|count*U|);
 * scalarOut (optional): the cellCount*lshCount projections themselves. */
void em2o_signatures(uint64_t cellCount, uint64_t geneCount, const uint64_t* toc, const uint32_t* geneIds,
                     const float* counts, const double* sum1, const double* U, uint64_t lshCount,
                     uint64_t* signatures, double eps, uint64_t* nearZero, double* scalarOut)
{
    const uint64_t W = (lshCount - 1) / 64 + 1;                      /* :127 */
    double* sumU = (double*)calloc(lshCount, sizeof(double));
    for (uint64_t g = 0; g < geneCount; g++)                         /* :137-144 */
        for (uint64_t i = 0; i < lshCount; i++) sumU[i] += U[g * lshCount + i];
    memset(signatures, 0, cellCount * W * sizeof(uint64_t));         /* :148, createNew zero-fills */
    double* s = (double*)malloc(lshCount * sizeof(double));
    double* mag = (double*)malloc(lshCount * sizeof(double));
    uint64_t nz = 0;
    for (uint64_t c = 0; c < cellCount; c++) {
        const double mean = sum1[c] / (double)geneCount;             /* :167-168 */
        for (uint64_t i = 0; i < lshCount; i++) {                    /* :180-182 */
            s[i] = -mean * sumU[i];
            mag[i] = fabs(s[i]);
        }
        for (uint64_t j = toc[c]; j < toc[c + 1]; j++) {             /* :188-198 */
            const double count = (double)counts[j];
            const double* v = U + (uint64_t)geneIds[j] * lshCount;
            for (uint64_t i = 0; i < lshCount; i++) {
                const double p = count * v[i];
                s[i] += p;
                mag[i] += fabs(p);
            }
        }
        for (uint64_t i = 0; i < lshCount; i++) {                    /* :201-206 */
            if (s[i] > 0.) signatures[c * W + (i >> 6)] |= 1ULL << (63 - (i & 63));
            if (fabs(s[i]) < eps * mag[i]) nz++;
        }
        if (scalarOut) memcpy(scalarOut + c * lshCount, s, lshCount * sizeof(double));
    }
    if (nearZero) *nearZero = nz;
    free(mag);
    free(s);
    free(sumU);
}

/* Lsh::computeSimilarityTable, src/Lsh.cpp:229-249. table has lshCount+1 entries. */
void em2o_similarity_table(uint64_t lshCount, double* table)
{
    const double pi = 3.141592653589793238462643383279502884;
    for (uint64_t m = 0; m <= lshCount; m++) {
        const double angle = (double)m * pi / (double)lshCount;
        table[m] = cos(angle);
    }
}

/* Largest mismatch count m whose table value passes findSimilarPairs4's strict filter
 * `similarity > similarityThreshold` (src/ExpressionMatrixLsh.cpp:244); -1 if none. */
int64_t em2o_mismatch_max(uint64_t lshCount, double similarityThreshold)
{
    const double pi = 3.141592653589793238462643383279502884;
    int64_t best = -1;
    for (uint64_t m = 0; m <= lshCount; m++) {
        if (cos((double)m * pi / (double)lshCount) > similarityThreshold) best = (int64_t)m;
        else break;                                  /* the table is non-increasing */
    }
    return best;
}

/* countMismatches, src/BitSet.hpp:277-288. */
static inline uint32_t mismatches(const uint64_t* x, const uint64_t* y, uint64_t W)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i < W; i++) n += (uint64_t)__builtin_popcountll(x[i] ^ y[i]);
    return (uint32_t)n;
}

void em2o_mismatch_counts(const uint64_t* sig, uint64_t W, uint64_t pairCount, const uint32_t* c0,
                          const uint32_t* c1, uint32_t* out)
{
    for (uint64_t i = 0; i < pairCount; i++) out[i] = mismatches(sig + (uint64_t)c0[i] * W, sig + (uint64_t)c1[i] * W, W);
}

void em2o_mismatch_row(const uint64_t* sig, uint64_t W, uint64_t cellCount, uint32_t cell0, uint32_t* out)
{
    for (uint64_t c = 0; c < cellCount; c++) out[c] = mismatches(sig + (uint64_t)cell0 * W, sig + c * W, W);
}

/* Checksum over all unordered pairs: sum of m and sum of m*m (mod 2^64), for bit-exactness checks
 * at sizes where storing every distance is impractical. */
void em2o_mismatch_checksum(const uint64_t* sig, uint64_t W, uint64_t cellCount, uint64_t* sumM, uint64_t* sumM2)
{
    uint64_t a = 0, b = 0;
    for (uint64_t i = 1; i < cellCount; i++)
        for (uint64_t j = 0; j < i; j++) {
            const uint64_t m = mismatches(sig + i * W, sig + j * W, W);
            a += m;
            b += m * m;
        }
    *sumM = a;
    *sumM2 = b;
}

/* Deterministic top-k: the reference's own GPU-host semantics (src/ExpressionMatrixLshGpu.cpp:132-157)
 * with findSimilarPairs4's filter (src/ExpressionMatrixLsh.cpp:244):
 * for each cell in [rowBegin,rowEnd): candidates = every other cell with table[m] > threshold,
 * ordered by (mismatch asc, cellId asc); keep the first k; similarity = float(table[m])
 * (SimilarPairs.cpp:290-302 stores float(similarity)).
 * Identical to the order SimilarPairs::sort() produces (orderPairs.hpp:44-52) because the float
 * table is strictly decreasing for lshCount <= 8192.
 * Outputs are indexed from rowBegin: ids/sims [(rowEnd-rowBegin)*k], used [rowEnd-rowBegin]. */
typedef struct { uint32_t m, id; } cand_t;
static int cand_less(const void* a, const void* b)
{
    const cand_t* x = (const cand_t*)a;
    const cand_t* y = (const cand_t*)b;
    if (x->m != y->m) return x->m < y->m ? -1 : 1;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    return 0;
}
double em2o_topk(const uint64_t* sig, uint64_t cellCount, uint64_t lshCount, uint64_t k, double similarityThreshold,
                 uint64_t rowBegin, uint64_t rowEnd, uint32_t* ids, float* sims, uint32_t* used)
{
    const uint64_t W = (lshCount - 1) / 64 + 1;
    double* table = (double*)malloc((lshCount + 1) * sizeof(double));
    em2o_similarity_table(lshCount, table);
    const int64_t mmax = em2o_mismatch_max(lshCount, similarityThreshold);
    cand_t* cand = (cand_t*)malloc((cellCount + 1) * sizeof(cand_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint64_t r = rowBegin; r < rowEnd; r++) {
        uint64_t n = 0;
        for (uint64_t c = 0; c < cellCount; c++) {
            if (c == r) continue;
            const uint32_t m = mismatches(sig + r * W, sig + c * W, W);
            if ((int64_t)m <= mmax) { cand[n].m = m; cand[n].id = (uint32_t)c; n++; }
        }
        qsort(cand, n, sizeof(cand_t), cand_less);
        const uint64_t keep = n < k ? n : k;
        const uint64_t o = (r - rowBegin) * k;
        for (uint64_t i = 0; i < k; i++) {
            ids[o + i] = i < keep ? cand[i].id : 0;
            sims[o + i] = i < keep ? (float)table[cand[i].m] : 0.f;
        }
        used[r - rowBegin] = (uint32_t)keep;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(cand);
    free(table);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* The bare pair loop of findSimilarPairs4 (src/ExpressionMatrixLsh.cpp:218-263) WITHOUT the
 * candidate bookkeeping: mismatch + table lookup + threshold test for rows [rowBegin,rowEnd) x all
 * cell1 < cell0, 64x64 blocked like the reference.  Used only to time a "port" CPU baseline when
 * oracle/_ref is not available.  Returns seconds; *pairs = pairs visited, *passed = pairs over threshold. */
double em2o_pair_loop(const uint64_t* sig, uint64_t cellCount, uint64_t lshCount, double similarityThreshold,
                      uint64_t rowBegin, uint64_t rowEnd, uint64_t* pairs, uint64_t* passed)
{
    const uint64_t W = (lshCount - 1) / 64 + 1;
    double* table = (double*)malloc((lshCount + 1) * sizeof(double));
    em2o_similarity_table(lshCount, table);
    uint64_t n = 0, p = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const uint64_t B = 64;
    if (rowEnd > cellCount) rowEnd = cellCount;
    for (uint64_t b0 = rowBegin; b0 < rowEnd; b0 += B) {
        const uint64_t e0 = b0 + B < rowEnd ? b0 + B : rowEnd;
        for (uint64_t b1 = 0; b1 <= b0; b1 += B) {
            const uint64_t e1 = b1 + B < e0 ? b1 + B : e0;
            for (uint64_t c0 = b0; c0 != e0; ++c0)
                for (uint64_t c1 = b1; c1 != e1 && c1 < c0; ++c1) {
                    ++n;
                    const double s = table[mismatches(sig + c0 * W, sig + c1 * W, W)];
                    if (s > similarityThreshold) ++p;
                }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(table);
    *pairs = n;
    *passed = p;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ExpressionMatrixSubset::computeCellSimilarity, src/ExpressionMatrixSubset.cpp:83-133
 * (Pearson correlation over all genes; dot by sorted merge, product in float, sum in double). */
double em2o_exact_similarity(uint64_t geneCount, const uint64_t* toc, const uint32_t* geneIds, const float* counts,
                             const double* sum1, const double* sum2, uint32_t c0, uint32_t c1)
{
    uint64_t i0 = toc[c0], e0 = toc[c0 + 1], i1 = toc[c1], e1 = toc[c1 + 1];
    double dot = 0.;
    while (i0 != e0 && i1 != e1) {
        const uint32_t g0 = geneIds[i0], g1 = geneIds[i1];
        if (g0 < g1) ++i0;
        else if (g1 < g0) ++i1;
        else {
            const float prod = counts[i0] * counts[i1];
            dot += prod;
            ++i0;
            ++i1;
        }
    }
    const double n = (double)geneCount;
    const double num = n * dot - sum1[c0] * sum1[c1];
    const double den = sqrt((n * sum2[c0] - sum1[c0] * sum1[c0]) * (n * sum2[c1] - sum1[c1] * sum1[c1]));
    return num / den;
}

void em2o_exact_similarities(uint64_t geneCount, const uint64_t* toc, const uint32_t* geneIds, const float* counts,
                             const double* sum1, const double* sum2, uint64_t pairCount, const uint32_t* c0,
                             const uint32_t* c1, double* out)
{
    for (uint64_t i = 0; i < pairCount; i++)
        out[i] = em2o_exact_similarity(geneCount, toc, geneIds, counts, sum1, sum2, c0[i], c1[i]);
}

/* All exact similarities of rows [rowBegin,rowEnd) against every cell, as a dense block
 * out[(r-rowBegin)*cellCount + c]; the diagonal is computed like any other pair. */
void em2o_exact_rows(uint64_t cellCount, uint64_t geneCount, const uint64_t* toc, const uint32_t* geneIds,
                     const float* counts, const double* sum1, const double* sum2, uint64_t rowBegin,
                     uint64_t rowEnd, double* out)
{
    for (uint64_t r = rowBegin; r < rowEnd; r++)
        for (uint64_t c = 0; c < cellCount; c++)
            out[(r - rowBegin) * cellCount + c] =
                em2o_exact_similarity(geneCount, toc, geneIds, counts, sum1, sum2, (uint32_t)r, (uint32_t)c);
}

/* MurmurHash64A (Austin Appleby, public domain), the hash of MemoryMapped::Vector::hash
 * (src/MemoryMappedVector.hpp:715-723, seed 231). */
uint64_t em2o_murmur64a(const void* key, int len, uint64_t seed)
{
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    const int r = 47;
    uint64_t h = seed ^ ((uint64_t)len * m);
    const unsigned char* data = (const unsigned char*)key;
    const unsigned char* end = data + (size_t)(len / 8) * 8;
    while (data != end) {
        uint64_t k;
        memcpy(&k, data, 8);
        data += 8;
        k *= m;
        k ^= k >> r;
        k *= m;
        h ^= k;
        h *= m;
    }
    switch (len & 7) {
    case 7: h ^= (uint64_t)data[6] << 48; /* fallthrough */
    case 6: h ^= (uint64_t)data[5] << 40; /* fallthrough */
    case 5: h ^= (uint64_t)data[4] << 32; /* fallthrough */
    case 4: h ^= (uint64_t)data[3] << 24; /* fallthrough */
    case 3: h ^= (uint64_t)data[2] << 16; /* fallthrough */
    case 2: h ^= (uint64_t)data[1] << 8;  /* fallthrough */
    case 1: h ^= (uint64_t)data[0];
            h *= m;
    }
    h ^= h >> r;
    h *= m;
    h ^= h >> r;
    return h;
}
