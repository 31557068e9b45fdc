"""First-contact GPU script: microbenchmarks + timing sweep of the scan variants (not the bench)."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic

out = {}
mb = os.path.join(ROOT, "expressionmatrix2_b200", "build", "microbench")
if os.path.exists(mb):
    out["microbench"] = json.loads(subprocess.check_output([mb]).decode())
    print(out["microbench"], flush=True)

def time_scan(N, L, k=50, thr=0.2, variant=em2.VARIANT_POPC, reps=2, clusters=500):
    sig = synthetic.gen_signatures(N, L, seed=1, clusters=clusters)
    eng = em2.Engine(0)
    d_sig = torch.from_numpy(sig.view(np.int64)).cuda()
    lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).cuda()
    pairs = torch.zeros((N, k, 2), dtype=torch.int32, device="cuda")
    used = torch.zeros(N, dtype=torch.int32, device="cuda")
    mm = em2.mismatch_max(L, thr)
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for r in range(reps + 1):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        eng.scan_topk_device(d_sig, N, L, 0, N, k, mm, lut, pairs, used, variant=variant, stream=s)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    eng.close()
    t = min(ts[1:])
    return dict(N=N, L=L, ms=t, ordered_pairs_per_s=N * N / (t * 1e-3), used_mean=float(used.float().mean()))

for csa in (0, 1, 2):
    os.environ["EM2_POPC_CSA"] = str(csa)
    # the env var is read once per process -> run in a subprocess
    code = ("import sys; sys.path.insert(0, %r); import tools.quick_gpu as q" % ROOT)
for N, L in ((100_000, 1024), (100_000, 256), (50_000, 4096)):
    r = time_scan(N, L)
    print(r, flush=True)
    out[f"scan_{N}_{L}"] = r

# full pipeline at a reduced C2 (20k x 30k) to time the signature kernel
N, G, nnzc, L = 20000, 30000, 1500, 1024
toc, genes, counts = synthetic.gen_expression_matrix_fast(N, G, nnzc, seed=3)
U = em2.generate_lsh_vectors(G, L, 231)
eng = em2.Engine(0)
for _ in range(2):
    ids, sims, used = eng.lsh_similar_pairs(toc, counts, U, 50, 0.2, gene_ids=genes)
    st = eng.stats()
print(st, flush=True)
nnz = len(genes)
st["sig_mac_per_s"] = nnz * L / (st["signatures_ms"] * 1e-3)
out["pipeline_20k"] = st
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "quick_gpu.json"), "w"), indent=1)
