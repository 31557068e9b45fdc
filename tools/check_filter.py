"""GPU check of the tensor-core signature filter path against the oracle + timing vs the FP64 kernel."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
import oracle
oracle.build()
eng = em2.Engine(0)
res = {}
def case(name, N, G, dens, L, tweak=None, opts=()):
    toc, genes, counts = synthetic.gen_expression_matrix(N, G, dens, seed=N + L, mode="clustered", clusters=7)
    if tweak: tweak(counts)
    U = em2.generate_lsh_vectors(G, L, 231)
    s1, _ = oracle.cell_sums(toc, counts)
    want, _ = oracle.signatures(toc, genes, counts, s1, U)
    out = {}
    for mode in (1, 2):
        eng.set_option("signature_mode", mode)
        for k, v in opts: eng.set_option(k, v)
        sig = eng.compute_signatures(toc, counts, U, gene_ids=genes)
        st = eng.stats()
        diff = int(np.count_nonzero(sig != want))
        bits = int(sum(bin(int(x)).count("1") for x in (sig ^ want).ravel()[:200000]))
        out[mode] = dict(word_diffs=diff, bit_diffs_sampled=bits, sig_ms=st["signatures_ms"], filter_cells=st["filter_cells"],
                         uncertain=st["filter_uncertain"], near_zero=st["near_zero_projections"], launches=st["kernel_launches"])
    for k, v in opts: eng.set_option(k, 0)
    res[name] = out
    print(name, json.dumps(out), flush=True)

case("small_1024", 1500, 700, 0.05, 1024)
case("ragged_200", 900, 333, 0.07, 200)
case("one_bit", 700, 300, 0.05, 1)
def big_counts(c): c[::97] = 300.0; c[5::1013] = 2.5
case("ineligible", 1200, 640, 0.05, 256, tweak=big_counts)
case("signed", 1500, 700, 0.05, 512, opts=(("filter_counts_signed", 1),))
case("overflow", 1500, 700, 0.05, 512, opts=(("filter_uncertain_cap", 16),))
# timing at bench shape (fast generator), 20k and 100k cells
for N in (20000, 100000):
    G, m, L = 30000, 1500, 1024
    toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345)
    U = em2.generate_lsh_vectors(G, L, 231)
    sigs = {}
    for mode in (1, 2, 2):
        eng.set_option("signature_mode", mode)
        sig = eng.compute_signatures(toc, counts, U, gene_ids=genes)
        st = eng.stats()
        sigs[mode] = sig
        print(f"N={N} mode={mode} sig_ms={st['signatures_ms']:.3f} sums_ms={st['sums_ms']:.3f} uncertain={st['filter_uncertain']} "
              f"filter_cells={st['filter_cells']} launches={st['kernel_launches']}", flush=True)
    print(f"N={N} filter == fp64: {bool(np.array_equal(sigs[1], sigs[2]))}  word diffs {int(np.count_nonzero(sigs[1] != sigs[2]))}", flush=True)
    s1, _ = oracle.cell_sums(toc[:257], counts[:int(toc[256])])
    want, _ = oracle.signatures(toc[:257], genes[:int(toc[256])], counts[:int(toc[256])], s1, U)
    print(f"N={N} first 256 cells == oracle: {bool(np.array_equal(sigs[2][:256], want))}", flush=True)
eng.close()
