"""Candidate-region capacity sweep of the one-directional MMA scan on BASELINE config 2's signatures."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, G, m, L, k, thr = 100000, 30000, 1500, 1024, 50, 0.2
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345)
U = em2.generate_lsh_vectors(G, L, 231)
eng = em2.Engine(0)
sig = eng.compute_signatures(toc, counts, U, gene_ids=genes)
d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
mm = em2.mismatch_max(L, thr)
s = torch.cuda.current_stream().cuda_stream
pairs = torch.zeros((N, k, 2), dtype=torch.int32, device="cuda")
used = torch.zeros(N, dtype=torch.int32, device="cuda")
ref = None
for extra in (0, 1, 2, 3):
    eng.set_option("cand_cap_extra", extra)
    ts = []
    for r in range(4):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        eng.scan_topk_device(d_sig, N, L, 0, N, k, mm, lut, pairs, used, variant=2, stream=s)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    cur = pairs.cpu().numpy().copy()
    ref = cur if ref is None else ref
    print(json.dumps(dict(cand_cap_extra=extra, cap=(2 + extra) * k + 32, ms=min(ts[1:]), equal=bool(np.array_equal(cur, ref)))), flush=True)
