import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import expressionmatrix2_b200 as em2
from expressionmatrix2_b200 import synthetic
N, L, cl, k, thr = 30000, 512, 100, 50, 0.2
sig = synthetic.gen_signatures(N, L, seed=1, clusters=cl)
eng = em2.Engine(0)
eng.set_option("debug_flags", 4)
ids, sims, used = eng.find_similar_pairs(sig, L, k, thr, variant=2)
print("slot 48,49:", ids[11070][48:50], sims[11070][48:50])
eng.close()
