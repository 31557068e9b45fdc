import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
import oracle
oracle.build()
N, L, cl, k, thr = 30000, 512, 100, 50, 0.2
sig = synthetic.gen_signatures(N, L, seed=1, clusters=cl)
eng = em2.Engine(0)
res = {}
for name, variant, opts in (("popc", 1, {}), ("ts", 2, {}), ("ss", 2, {"mma_streamed": 1})):
    for o, v in opts.items(): eng.set_option(o, v)
    res[name] = eng.find_similar_pairs(sig, L, k, thr, variant=variant)
    for o in opts: eng.set_option(o, 0)
for a in ("ts", "ss"):
    d = np.nonzero((res[a][0] != res["popc"][0]).any(1) | (res[a][2] != res["popc"][2]))[0]
    print(a, "rows differing from popc:", len(d), d[:10])
    for r in d[:3]:
        wi, ws, wu, _ = oracle.topk(sig, L, k, thr, int(r), int(r) + 1)
        m = oracle.mismatch_row(sig, int(r))
        print(" row", r, "oracle used", wu[0], "popc used", res["popc"][2][r], a, "used", res[a][2][r])
        print("  oracle==popc", np.array_equal(wi[0], res["popc"][0][r]), " oracle==", a, np.array_equal(wi[0], res[a][0][r]))
        bad = np.nonzero(wi[0] != res[a][0][r])[0]
        j = bad[0] if len(bad) else 0
        print("  first diff slot", j, "oracle ids", wi[0][j:j+6], "ham", m[wi[0][j:j+6]], a, "ids", res[a][0][r][j:j+6], "ham", m[res[a][0][r][j:j+6]])
        print("  row block", r // 128, "lane", r % 128, "kth ham", m[wi[0][wu[0]-1]] if wu[0] else None)
eng.close()
