"""Generate tests/golden/*.npz from the reference's OWN sources (oracle/_ref/libem2ref.so).

Run in the authoring container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
The reference ships no golden vectors for the LSH path, so these fixtures are
outputs of the reference build itself on small seeded inputs; they travel to the GPU box, where
/root/reference does not exist.  Inputs are stored too, so nothing has to be regenerated bit-exactly.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from expressionmatrix2_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, N, G, density, L, mode, clusters, [(k, threshold)]
    ("lsh_n300_g200_l64", 300, 200, 0.10, 64, "clustered", 6, [(5, 0.2), (8, -1.0)]),
    ("lsh_n257_g150_l100", 257, 150, 0.08, 100, "iid", 0, [(10, 0.2), (3, -1.0)]),
    ("lsh_n400_g300_l256", 400, 300, 0.05, 256, "clustered", 8, [(10, 0.2), (20, -1.0), (4, 0.6)]),
    ("lsh_n500_g400_l1024", 500, 400, 0.05, 1024, "clustered", 10, [(50, 0.2), (16, -1.0)]),
    ("lsh_n200_g120_l2048", 200, 120, 0.10, 2048, "clustered", 4, [(10, 0.2)]),
]


def main():
    assert O.have_ref(), "build oracle/_ref first"
    for name, N, G, dens, L, mode, clusters, combos in CASES:
        toc, genes, counts = synthetic.gen_expression_matrix(N, G, dens, seed=sum(map(ord, name)), mode=mode, clusters=max(clusters, 1))
        data = dict(toc=toc, genes=genes, counts=counts, gene_count=G, lsh_count=L, seed=231)
        with O.Reference.from_csr(toc, genes, counts, G, L, 231) as R:
            U = R.lsh_vectors()
            data["U_head"] = U[:3, :8].copy()
            data["U_checksum"] = np.array([U.sum(), np.abs(U).sum(), (U * np.arange(1, G + 1)[:, None]).sum()])
            data["signatures"] = R.signatures()
            s1, s2 = R.sums()
            data["sum1"], data["sum2"] = s1, s2
            data["table"] = R.similarity_table()
            rng = np.random.default_rng(5)
            c0 = rng.integers(0, N, 2000).astype(np.uint32)
            c1 = rng.integers(0, N, 2000).astype(np.uint32)
            data["pair_c0"], data["pair_c1"] = c0, c1
            data["pair_mismatch"] = R.mismatch_counts(c0, c1)
            data["pair_exact"] = R.exact_similarity(c0[:300], c1[:300])
            data["row0_mismatch"] = R.mismatch_row(0)
            for i, (k, thr) in enumerate(combos):
                ids, sims, used, _ = R.topk_deterministic(k, thr)
                lit = R.find_similar_pairs4_loop(k, thr)
                data[f"combo{i}_k"] = k
                data[f"combo{i}_thr"] = thr
                data[f"combo{i}_ids"], data[f"combo{i}_sims"], data[f"combo{i}_used"] = ids, sims, used
                data[f"combo{i}_lit_ids"], data[f"combo{i}_lit_sims"], data[f"combo{i}_lit_used"] = (
                    lit["ids"], lit["sims"], lit["used"])
            data["combos"] = len(combos)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
        print(name, "nnz", len(genes), "sig", data["signatures"].shape)
    # generator / hash known answers
    np.savez_compressed(os.path.join(OUT, "generator.npz"), normal_231=O.ref_normal_stream(231, 4096),
                        normal_7=O.ref_normal_stream(7, 1001),
                        murmur_inputs=np.frombuffer(b"ExpressionMatrix2 LSH hot path golden", np.uint8),
                        murmur=np.array([O.ref_murmur64a(b"ExpressionMatrix2 LSH hot path golden"[:n])
                                         for n in range(0, 38)], np.uint64))


if __name__ == "__main__":
    main()
