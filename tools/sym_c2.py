"""BASELINE config 2 data through the symmetric / one-directional scan: python tools/sym_c2.py [N] [clusters]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
clusters = int(sys.argv[2]) if len(sys.argv) > 2 else 64
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["one_directional", "symmetric"]
G, m, L, k, thr = 30000, 1500, 1024, 50, 0.2
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345, clusters=clusters)
U = em2.generate_lsh_vectors(G, L, 231)
eng = em2.Engine(0)
sig = eng.compute_signatures(toc, counts, U, gene_ids=genes)
d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
mm = em2.mismatch_max(L, thr)
s = torch.cuda.current_stream().cuda_stream
out, res = {}, {}
for name, opt in (("one_directional", 1), ("symmetric", 2)):
    if name not in modes:
        continue
    eng.set_option("scan_symmetric", opt)
    pairs = torch.zeros((N, k, 2), dtype=torch.int32, device="cuda")
    used = torch.zeros(N, dtype=torch.int32, device="cuda")
    ts = []
    for r in range(4):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        eng.scan_topk_device(d_sig, N, L, 0, N, k, mm, lut, pairs, used, variant=2, stream=s)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    sym = eng.stats()["scan_symmetric"]
    eng.set_option("debug_flags", 8)
    eng.find_similar_pairs(sig, L, k, thr, variant=2)
    eng.set_option("debug_flags", 0)
    out[name] = dict(ms=min(ts[1:]), appended=eng.stats()["candidates_appended"], sym=sym)
    res[name] = (pairs.cpu().numpy().copy(), used.cpu().numpy().copy())
out["equal"] = len(res) < 2 or bool(np.array_equal(res["one_directional"][0], res["symmetric"][0]) and
                    np.array_equal(res["one_directional"][1], res["symmetric"][1]))
out["used_mean"] = float(list(res.values())[-1][1].mean())
print(json.dumps(out))
