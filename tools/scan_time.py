"""Time the scan kernel alone: python tools/scan_time.py N L [variant] [k] [clusters]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, L = int(sys.argv[1]), int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
k = int(sys.argv[4]) if len(sys.argv) > 4 else 50
clusters = int(sys.argv[5]) if len(sys.argv) > 5 else 500
sig = synthetic.gen_signatures(N, L, seed=1, clusters=clusters)
eng = em2.Engine(0)
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
t = min(ts[1:])
print(json.dumps(dict(N=N, L=L, variant=variant, ms=t, ordered_pairs_per_s=N * N / (t * 1e-3),
                      used_mean=float(used.float().mean()))))
