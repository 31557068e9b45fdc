"""Golden vectors for the section-8f rows that have no plain-C oracle: the bucketed LSH search (findSimilarPairs7) and the
SignatureGraph construction, both produced by the reference's own classes through oracle/_ref
(ref_driver.cpp: em2ref_find_similar_pairs7, em2ref_signature_graph).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_next_rows.py

Writes tests/golden/next_bucketed.npz and tests/golden/next_siggraph.npz (a few hundred KB)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from expressionmatrix2_b200 import synthetic  # noqa: E402

oracle.build()
assert oracle.have_ref(), "the reference build (oracle/_ref) is required"
here = os.path.dirname(os.path.abspath(__file__))

# ---- bucketed search: two parameter sets on one signature set (hashed + direct buckets; a cut-off inside buckets)
sig = synthetic.gen_signatures(1500, 192, seed=21, clusters=12)
out = dict(signatures=sig, lsh_count=192)
cases = [(12, 0.3, [20, 9], 120, 10), (4, 0.25, [33, 7], 9, 5)]
out["cases"] = len(cases)
with oracle.Reference.from_signatures(sig, 192) as ref:
    for i, (k, thr, slices, max_check, log2b) in enumerate(cases):
        ids, sims, used = ref.find_similar_pairs7(k, thr, slices, max_check, log2b)
        out.update({f"c{i}_k": k, f"c{i}_thr": thr, f"c{i}_slices": np.array(slices, np.int32), f"c{i}_max_check": max_check,
                    f"c{i}_log2b": log2b, f"c{i}_ids": ids, f"c{i}_sims": sims, f"c{i}_used": used})
np.savez_compressed(os.path.join(here, "next_bucketed.npz"), **out)

# ---- signature graph: 14-bit signatures of 4000 cells, minimum vertex size 2
sig = synthetic.gen_signatures(4000, 14, seed=22, clusters=8, flip_fraction=0.15)
order, offsets, edges = oracle.ref_signature_graph(sig, 14, 2)
np.savez_compressed(os.path.join(here, "next_siggraph.npz"), signatures=sig, lsh_count=14, min_cell_count=2, cell_order=order,
                    vertex_offsets=offsets, edges=edges)
print("bucketed:", [int(out[f"c{i}_used"].sum()) for i in range(len(cases))], "list entries;",
      "signature graph:", len(offsets) - 1, "vertices,", len(edges), "edges")
