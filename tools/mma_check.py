"""Quick validation + timing of the tcgen05 variant against the POPC variant (both on the GPU)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic

eng = em2.Engine(0)
ok = True
for N, L, k, thr, clusters in ((300, 1024, 10, 0.2, 5), (1000, 1024, 50, -1.0, 7), (5000, 256, 50, 0.2, 20), (4099, 1000, 20, 0.2, 9),
                               (20000, 1024, 50, 0.2, 50), (3000, 64, 50, -1.0, 0)):
    sig = synthetic.gen_signatures(N, L, seed=N, clusters=clusters)
    d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
    # Hamming block, bit-exact
    R = min(N, 200)
    a = torch.zeros((R, N), dtype=torch.int16, device="cuda")
    b = torch.zeros((R, N), dtype=torch.int16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    eng.mismatch_block_device(d_sig, N, L, 3, 3 + R if 3 + R <= N else N, a[: (R if 3 + R <= N else N - 3)], variant=em2.VARIANT_POPC, stream=s)
    eng.mismatch_block_device(d_sig, N, L, 3, 3 + R if 3 + R <= N else N, b[: (R if 3 + R <= N else N - 3)], variant=em2.VARIANT_MMA_I8, stream=s)
    torch.cuda.synchronize()
    same_h = bool(torch.equal(a, b))
    p = eng.find_similar_pairs(sig, L, k, thr, variant=em2.VARIANT_POPC)
    m = eng.find_similar_pairs(sig, L, k, thr, variant=em2.VARIANT_MMA_I8)
    same = all(np.array_equal(x, y) for x, y in zip(p, m))
    print(f"N={N} L={L} k={k} thr={thr}: hamming_equal={same_h} lists_equal={same} used={m[2].mean():.1f}", flush=True)
    if not same_h:
        d = (a != b).nonzero()
        print("  first diffs", d[:5].tolist(), a[d[0][0], d[0][1]].item(), b[d[0][0], d[0][1]].item())
    ok &= same and same_h
print("ALL OK" if ok else "MISMATCH", flush=True)

def time_scan(N, L, variant, k=50, clusters=500):
    sig = synthetic.gen_signatures(N, L, seed=1, clusters=clusters)
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
    return dict(N=N, L=L, variant=variant, ms=t, ordered_pairs_per_s=N * N / (t * 1e-3), tops=N * N * 2 * L / (t * 1e-3) / 1e12)

for N, L in ((100_000, 1024), (100_000, 256), (400_000, 1024)):
    for v in (em2.VARIANT_MMA_I8,):
        print(time_scan(N, L, v), flush=True)
