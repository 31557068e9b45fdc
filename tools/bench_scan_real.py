"""Time the scan on the bench workload's real signatures (heavy-hit regime): N cells from c2."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, 30000, 1500, seed=12345)
U = em2.generate_lsh_vectors(30000, 1024, 231)
eng = em2.Engine(0)
sig = eng.compute_signatures(toc, counts, U, gene_ids=genes)
L, k = 1024, 50
d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
pairs = torch.zeros((N, k, 2), dtype=torch.int32, device="cuda")
used = torch.zeros(N, dtype=torch.int32, device="cuda")
mm = em2.mismatch_max(L, 0.2)
s = torch.cuda.current_stream().cuda_stream
ts = []
for r in range(4):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    eng.scan_topk_device(d_sig, N, L, 0, N, k, mm, lut, pairs, used, variant=variant, stream=s)
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(json.dumps(dict(real=True, N=N, variant=variant, ms=min(ts[1:]))))
