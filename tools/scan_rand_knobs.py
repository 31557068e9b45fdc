"""Scan timing on synthetic signatures under option knobs: python tools/scan_rand_knobs.py N L clusters cfg..."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, L, cl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sig = synthetic.gen_signatures(N, L, seed=1, clusters=cl)
eng = em2.Engine(0)
k = 50
d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
pairs = torch.zeros((N, k, 2), dtype=torch.int32, device="cuda")
used = torch.zeros(N, dtype=torch.int32, device="cuda")
mm = em2.mismatch_max(L, 0.2)
s = torch.cuda.current_stream().cuda_stream
ref = None
for cfg in sys.argv[4:] or ["base"]:
    opts = {} if cfg == "base" else dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in cfg.split(","))
    variant = opts.pop("variant", 2)
    for o, v in opts.items():
        eng.set_option(o, v)
    ts = []
    for r in range(4):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        eng.scan_topk_device(d_sig, N, L, 0, N, k, mm, lut, pairs, used, variant=variant, stream=s)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out = pairs.cpu().numpy()
    if ref is None:
        ref = out.copy()
    t = min(ts[1:])
    K = (L + 127) // 128 * 128
    print(json.dumps(dict(N=N, L=L, clusters=cl, cfg=cfg, ms=round(t, 3), tops=round(N * N * 2.0 * K / (t * 1e-3) / 1e12), same_as_first=bool(np.array_equal(out, ref)))), flush=True)
    for o in opts:
        eng.set_option(o, 0)
eng.close()
