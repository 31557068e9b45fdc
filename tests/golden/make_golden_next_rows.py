"""Golden vectors for the section-8f rows that have no plain-C oracle: the bucketed LSH search (findSimilarPairs7), the
SignatureGraph construction and the CellGraph edge construction, all produced by the reference's own classes through
oracle/_ref (ref_driver.cpp: em2ref_find_similar_pairs7, em2ref_signature_graph, em2ref_cell_graph_edges -- the last one
is the reference's CellGraph constructor itself, src/CellGraph.cpp compiled unmodified).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_next_rows.py

Writes tests/golden/next_bucketed.npz, next_siggraph.npz and next_cellgraph.npz (a few hundred KB)."""
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
# ---- cell graph: SimilarPairs rows of 2500 cells (deterministic top-k of the reference), three (cell set, threshold,
# maxConnectivity) cases incl. cells outside the graph's cell set, a threshold equal to a stored float, and
# maxConnectivity 0 (never stops the reference's loop)
sig = synthetic.gen_signatures(2500, 256, seed=23, clusters=20)
sig[100:140] = sig[100]
with oracle.Reference.from_signatures(sig, 256) as ref:
    ids, sims, used = ref.topk_deterministic(24, 0.1)[:3]
rng = np.random.default_rng(6)
cg = dict(ids=ids, sims=sims, used=used)
cg_cases = [(0.2, 10, 0.0), (float(np.float32(sims[7, 3])), 5, 0.3), (-1.0, 0, 0.4)]
cg["cases"] = len(cg_cases)
for i, (thr, max_conn, drop) in enumerate(cg_cases):
    cell_set = np.flatnonzero(rng.random(len(used)) >= drop).astype(np.uint32)
    v0, v1, ss = oracle.ref_cell_graph_edges(ids, sims, used, cell_set, thr, max_conn)
    cg.update({f"c{i}_thr": thr, f"c{i}_max_conn": max_conn, f"c{i}_cell_set": cell_set, f"c{i}_v0": v0, f"c{i}_v1": v1, f"c{i}_sim": ss})
np.savez_compressed(os.path.join(here, "next_cellgraph.npz"), **cg)
print("cell graph:", [len(cg[f"c{i}_v0"]) for i in range(len(cg_cases))], "edges")

print("bucketed:", [int(out[f"c{i}_used"].sum()) for i in range(len(cases))], "list entries;",
      "signature graph:", len(offsets) - 1, "vertices,", len(edges), "edges")
