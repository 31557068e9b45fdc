"""Parity at BASELINE.json's own configurations (-m gpu).

config 1 (10k cells x 20k genes, 5 %, 1024 bits, top-50) runs WHOLE against the oracle: every signature word and every
list.  config 2 (100k x 30k) runs whole on the GPU and is compared on a sample: all signature words of 4096 cells and
the complete lists of 96 rows spread over the matrix.  The metric's 1M-cell configuration is covered through
size-independent properties in tests/test_gpu_parity.py::test_one_million_cells_sampled_rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import expressionmatrix2_b200 as em2  # noqa: E402
from expressionmatrix2_b200 import synthetic  # noqa: E402


def _equal_lists(got, want):
    return (np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0]) and
            np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)))


def test_config1_whole_job_against_the_oracle(engine, oracle):
    N, G, m, L, k, thr = 10_000, 20_000, 1000, 1024, 50, 0.2
    toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345)
    U = em2.generate_lsh_vectors(G, L, 231)
    assert np.array_equal(U, oracle.generate_lsh_vectors(G, L, 231))
    s1, s2 = oracle.cell_sums(toc, counts)
    want_sig, _ = oracle.signatures(toc, genes, counts, s1, U)
    want = oracle.topk(want_sig, L, k, thr)[:3]
    ids, sims, used, sig = engine.lsh_similar_pairs(toc, counts, U, k, thr, gene_ids=genes, want_signatures=True)
    assert engine.stats()["near_zero_projections"] == 0
    assert np.array_equal(sig, want_sig)                      # all 10k x 16 words
    assert _equal_lists((ids, sims, used), want)              # all 10k lists, ids in order, 0 ULP similarities
    # both scan variants and the symmetric kernel give the same lists on the same signatures
    for variant in (em2.VARIANT_POPC, em2.VARIANT_MMA_I8):
        assert _equal_lists(engine.find_similar_pairs(sig, L, k, thr, variant=variant), want)
    engine.set_option("scan_symmetric", 2)
    try:
        assert _equal_lists(engine.find_similar_pairs(sig, L, k, thr, variant=em2.VARIANT_MMA_I8), want)
        assert engine.stats()["scan_symmetric"] == 1
    finally:
        engine.set_option("scan_symmetric", 0)
    # the literal findSimilarPairs4 loop never beats the deterministic selection (dominance)
    if oracle.have_ref():
        with oracle.Reference.from_signatures(want_sig, L) as ref:
            r = ref.find_similar_pairs4_loop(k, thr, 0, N)
        lit_sims, lit_used = r["sims"], r["used"]
        assert np.all(lit_used <= used)
        for c in range(0, N, 97):
            n = int(lit_used[c])
            assert np.all(sims[c, :n] >= lit_sims[c, :n])


def test_config2_sampled_against_the_oracle(engine, oracle):
    N, G, m, L, k, thr = 100_000, 30_000, 1500, 1024, 50, 0.2
    toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345)
    U = em2.generate_lsh_vectors(G, L, 231)
    ids, sims, used, sig = engine.lsh_similar_pairs(toc, counts, U, k, thr, gene_ids=genes, want_signatures=True)
    # signatures: every word of 4096 cells (4 runs of 1024 spread over the matrix)
    for b in (0, 33_000, 66_000, N - 1024):
        e = b + 1024
        lt = (toc[b:e + 1] - toc[b]).astype(np.uint64)
        lo, hi = int(toc[b]), int(toc[e])
        s1, _ = oracle.cell_sums(lt, counts[lo:hi])
        want_sig, _ = oracle.signatures(lt, genes[lo:hi], counts[lo:hi], s1, U)
        assert np.array_equal(sig[b:e], want_sig)
    # lists: 96 complete rows against the oracle's selection over all 100k columns
    for b in (0, 12_345, 50_000, 77_777, 99_000, N - 16):
        wi, ws, wu, _ = oracle.topk(sig, L, k, thr, b, b + 16)
        assert _equal_lists((ids[b:b + 16], sims[b:b + 16], used[b:b + 16]), (wi, ws, wu))
    # size-independent properties of all 100k lists
    assert np.all(used <= k)
    table = em2.similarity_table(L).astype(np.float32)
    rows = np.arange(N)[:, None]
    valid = np.arange(k)[None, :] < used[:, None]
    assert not np.any((ids == rows) & valid)                                  # no self pairs
    assert np.all(np.diff(sims, axis=1)[valid[:, 1:]] <= 0)                   # similarity descending
    assert np.all(sims[valid] > np.float32(thr))
    assert np.all(np.isin(sims[valid], table))                                # every value is a table entry
