"""Wall time of the bucketed LSH search (em2_find_similar_pairs7) and its overlap with the all-pairs lists:
python tools/bucketed_time.py N L"""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, L = int(sys.argv[1]), int(sys.argv[2])
k, thr, slices, max_check, log2b = 50, 0.2, [22, 20, 18], 1000, 20
sig = synthetic.gen_signatures(N, L, seed=1, clusters=max(1, N // 2000))
eng = em2.Engine(0)
out = {}
for it in range(2):
    t0 = time.time()
    ids, sims, used = eng.find_similar_pairs7(sig, L, k, thr, slices, max_check, log2b)
    out["bucketed_wall_s"] = time.time() - t0
    out["bucketed_stats_total_ms"] = eng.stats()["total_ms"]
    out["kernel_launches"] = eng.stats()["kernel_launches"]
t0 = time.time()
aids, asims, aused = eng.find_similar_pairs(sig, L, k, thr)
out["all_pairs_wall_s"] = time.time() - t0
rows = np.random.default_rng(0).integers(0, N, 2000)
hit = [len(np.intersect1d(ids[r, :used[r]], aids[r, :aused[r]])) / max(1, aused[r]) for r in rows]
out.update(N=N, L=L, slices=slices, max_check=max_check, log2_bucket_count=log2b, used_mean=float(used.mean()),
           overlap_with_all_pairs_lists=float(np.mean(hit)))
print(json.dumps(out))
