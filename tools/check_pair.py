"""GPU check of the CTA-pair (cta_group::2) MMA scan: distances, lists, timing."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
import oracle
oracle.build()
eng = em2.Engine(0)
s = torch.cuda.current_stream().cuda_stream
stage = sys.argv[1] if len(sys.argv) > 1 else "all"
if stage in ("all", "dist"):
    for L in (1024, 512, 200):
        N = 3001
        sig = synthetic.gen_signatures(N, L, seed=L, clusters=11)
        d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
        for pair in (0, 1):
            eng.set_option("mma_cta_pair", pair)
            out = torch.zeros((700, N), dtype=torch.int16, device="cuda")
            eng.mismatch_block_device(d_sig, N, L, 1000, 1700, out, variant=em2.VARIANT_MMA_I8, stream=s)
            torch.cuda.synchronize()
            got = out.cpu().numpy().view(np.uint16)
            ok = all(np.array_equal(got[r].astype(np.uint32), oracle.mismatch_row(sig, 1000 + r)) for r in (0, 127, 128, 255, 256, 511, 699))
            print(f"L={L} pair={pair} distances ok={ok}", flush=True)
if stage in ("all", "lists"):
    for (N, L, k, thr, cl) in ((2049, 1024, 50, 0.2, 13), (5000, 512, 20, -1.0, 0), (777, 1024, 10, 0.2, 5)):
        sig = synthetic.gen_signatures(N, L, seed=N, clusters=cl)
        want = oracle.topk(sig, L, k, thr)[:3]
        for pair in (0, 1):
            eng.set_option("mma_cta_pair", pair)
            ids, sims, used = eng.find_similar_pairs(sig, L, k, thr, variant=em2.VARIANT_MMA_I8)
            ok = np.array_equal(ids, want[0]) and np.array_equal(sims.view(np.uint32), want[1].view(np.uint32)) and np.array_equal(used, want[2])
            print(f"N={N} L={L} k={k} pair={pair} lists ok={ok}", flush=True)
if stage in ("all", "time"):
    for (N, cl) in ((100000, 0), (200000, 500)):
        L, k = 1024, 50
        sig = synthetic.gen_signatures(N, L, seed=1, clusters=cl)
        d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
        lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
        pairs = torch.zeros((N, k, 2), dtype=torch.int32, device="cuda")
        used = torch.zeros(N, dtype=torch.int32, device="cuda")
        mm = em2.mismatch_max(L, 0.2)
        res = {}
        for pair in (0, 1):
            eng.set_option("mma_cta_pair", pair)
            ts = []
            for r in range(4):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record()
                eng.scan_topk_device(d_sig, N, L, 0, N, k, mm, lut, pairs, used, variant=em2.VARIANT_MMA_I8, stream=s)
                b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            res[pair] = (min(ts[1:]), pairs.cpu().numpy().copy(), used.cpu().numpy().copy())
            print(f"N={N} clusters={cl} pair={pair} ms={min(ts[1:]):.3f} ordered pairs/s={N*N/(min(ts[1:])*1e-3):.4g} TOP/s={N*N*2048/(min(ts[1:])*1e-3)/1e12:.0f}", flush=True)
        print("identical results:", bool(np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])), flush=True)
eng.close()
