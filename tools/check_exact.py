"""GPU check + timing of the exact path (config 5 shape: 50k x 20k @ 5 %)."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
import oracle
oracle.build()
eng = em2.Engine(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
G, m, k, thr = 20000, 1000, 50, 0.2
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=7)
for it in range(2):
    t0 = time.time()
    ids, sims, used = eng.exact_similar_pairs(toc, counts, G, k, thr, gene_ids=genes)
    st = eng.stats()
    print(f"N={N} exact: wall {time.time()-t0:.3f}s  scan_ms={st['scan_ms']:.2f} h2d={st['h2d_ms']:.2f} d2h={st['d2h_ms']:.2f} "
          f"launches={st['kernel_launches']} mean used={used.mean():.2f}", flush=True)
rows = [0, 1, N // 2, N - 1]
for r in rows:
    wi, ws, wu, _ = oracle.exact_topk(G, toc, genes, counts, k, thr, r, r + 1)
    ok = np.array_equal(wi[0], ids[r]) and np.array_equal(ws[0].view(np.uint32), sims[r].view(np.uint32)) and wu[0] == used[r]
    print(f"row {r}: equal to oracle = {ok} used={used[r]}", flush=True)
pairs = N * (N - 1) / 2
print(json.dumps(dict(N=N, G=G, scan_ms=st["scan_ms"], pairs_per_s=pairs / (st["scan_ms"] * 1e-3),
                      int8_tops_executed=2.0 * N * N * G / (st["scan_ms"] * 1e-3) / 1e12)))
eng.close()
