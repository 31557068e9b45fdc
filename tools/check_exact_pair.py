"""Exact path: CTA-pair form vs single-CTA form (results must be identical) + timing."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
eng = em2.Engine(0)
for N in (3000, 50000):
    G, m, k, thr = (1500, 100, 20, 0.1) if N == 3000 else (20000, 1000, 50, 0.2)
    toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=7)
    res = {}
    for pair in (0, 1):
        eng.set_option("exact_cta_pair", pair)
        for it in range(2):
            ids, sims, used = eng.exact_similar_pairs(toc, counts, G, k, thr, gene_ids=genes)
            st = eng.stats()
        res[pair] = (ids, sims, used)
        print(f"N={N} pair={pair} scan_ms={st['scan_ms']:.2f}", flush=True)
    same = all(np.array_equal(a, b) for a, b in zip(res[0], res[1]))
    print(f"N={N} identical: {same}", flush=True)
eng.close()
