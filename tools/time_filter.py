"""Per-kernel timing driver for the signature stage at the bench shape (run under ncu for a launch list)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 2, 1]
G, m, L = 30000, 1500, 1024
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345)
U = em2.generate_lsh_vectors(G, L, 231)
with em2.Engine(0) as eng:
    for mode in modes:
        eng.set_option("signature_mode", mode)
        eng.compute_signatures(toc, counts, U, gene_ids=genes)
        st = eng.stats()
        print(f"N={N} mode={mode} sig_ms={st['signatures_ms']:.3f} uncertain={st['filter_uncertain']}", flush=True)
