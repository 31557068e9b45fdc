import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, L, cl, k, thr = 30000, 512, 100, 50, 0.2
sig = synthetic.gen_signatures(N, L, seed=1, clusters=cl)
eng = em2.Engine(0)
ref = eng.find_similar_pairs(sig, L, k, thr, variant=1)
for cfg in sys.argv[1:]:
    opts = {} if cfg == "base" else dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in cfg.split(","))
    for o, v in opts.items(): eng.set_option(o, v)
    for rep in range(3):
        got = eng.find_similar_pairs(sig, L, k, thr, variant=2)
        d = np.nonzero((got[0] != ref[0]).any(1) | (got[2] != ref[2]))[0]
        info = []
        for r in d[:4]:
            j = np.nonzero(got[0][r] != ref[0][r])[0]
            info.append((int(r), int(j[0]) if len(j) else -1, int(ref[0][r][j[0]]) if len(j) else -1, int(got[0][r][j[0]]) if len(j) else -1))
        print(cfg, "rep", rep, "differing rows:", len(d), info, flush=True)
    for o in opts: eng.set_option(o, 0)
eng.close()
