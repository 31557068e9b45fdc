"""Device-API timing of sums + signatures at the bench shape under option knobs: python tools/sig_device_time.py cfg..."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, G, m, L = 100000, 30000, 1500, 1024
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, m, seed=12345)
U = em2.generate_lsh_vectors(G, L, 231)
eng = em2.Engine(0)
dev = torch.device("cuda", 0)
d_toc = torch.from_numpy(toc.view(np.int64)).to(dev)
d_counts = torch.from_numpy(em2.to_pairs(genes, counts).view(np.int64)).to(dev)
d_U = torch.from_numpy(U).to(dev)
d_s1 = torch.empty(N, dtype=torch.float64, device=dev); d_s2 = torch.empty(N, dtype=torch.float64, device=dev)
d_sig = torch.zeros((N, 16), dtype=torch.int64, device=dev)
d_nz = torch.zeros(8, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
ref = None
for cfg in sys.argv[1:] or ["base"]:
    opts = {} if cfg == "base" else dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in cfg.split(","))
    for o, v in opts.items(): eng.set_option(o, v)
    ts = []
    for r in range(5):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        eng.cell_sums_device(N, d_toc, d_counts, d_s1, d_s2, stream=s)
        eng.signatures_device(N, G, d_toc, d_counts, d_s1, d_s2, d_U, L, L, d_sig, d_nz, stream=s, nnz=int(toc[-1]))
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out = d_sig.cpu().numpy()
    if ref is None: ref = out.copy()
    print(json.dumps(dict(cfg=cfg, ms=round(min(ts[1:]), 3), all=[round(t, 2) for t in ts], same=bool(np.array_equal(out, ref)))), flush=True)
    for o in opts: eng.set_option(o, 0)
eng.close()
