"""Symmetric vs one-directional MMA scan: python tools/sym_time.py N L [k] [clusters]  (clusters 0 = iid bits)"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, L = int(sys.argv[1]), int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 50
clusters = int(sys.argv[4]) if len(sys.argv) > 4 else 500
modes = sys.argv[5].split(",") if len(sys.argv) > 5 else ["one_directional", "symmetric"]
sig = synthetic.gen_signatures(N, L, seed=1, clusters=clusters) if clusters else synthetic.gen_signatures(N, L, seed=1)
eng = em2.Engine(0)
d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
mm = em2.mismatch_max(L, 0.2)
s = torch.cuda.current_stream().cuda_stream
out = {}
res = {}
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
    st = eng.stats()
    eng.set_option("debug_flags", 8)
    got = eng.find_similar_pairs(sig, L, k, 0.2, variant=2)
    eng.set_option("debug_flags", 0)
    out[name] = dict(ms=min(ts[1:]), appended=eng.stats()["candidates_appended"], sym=st["scan_symmetric"])
    res[name] = (pairs.cpu().numpy().copy(), used.cpu().numpy().copy())
out["equal"] = len(res) < 2 or bool(np.array_equal(res["one_directional"][0], res["symmetric"][0]) and
                    np.array_equal(res["one_directional"][1], res["symmetric"][1]))
out["N"], out["L"], out["k"], out["clusters"] = N, L, k, clusters
print(json.dumps(out))
