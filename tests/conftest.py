import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "lsh_*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def golden_bucketed_cases():
    """(signatures, L, k, thr, slices, max_check, log2b, ids, sims, used) of tests/golden/next_bucketed.npz: lists the
    reference's own classes produce for findSimilarPairs7 (tests/golden/make_golden_next_rows.py)."""
    g = load_golden("next_bucketed")
    for i in range(int(g["cases"])):
        yield (g["signatures"], int(g["lsh_count"]), int(g[f"c{i}_k"]), float(g[f"c{i}_thr"]), g[f"c{i}_slices"].tolist(),
               int(g[f"c{i}_max_check"]), int(g[f"c{i}_log2b"]), g[f"c{i}_ids"], g[f"c{i}_sims"], g[f"c{i}_used"])


def golden_cellgraph_cases():
    """(ids, sims, used, cell_set, thr, max_conn, v0, v1, sim) of tests/golden/next_cellgraph.npz: edges built by the
    reference's own CellGraph constructor (tests/golden/make_golden_next_rows.py)."""
    g = load_golden("next_cellgraph")
    for i in range(int(g["cases"])):
        yield (g["ids"], g["sims"], g["used"], g[f"c{i}_cell_set"], float(g[f"c{i}_thr"]), int(g[f"c{i}_max_conn"]),
               g[f"c{i}_v0"], g[f"c{i}_v1"], g[f"c{i}_sim"])


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine():
    """One em2 context on cuda:0 for the whole GPU session (fails loudly without the library/GPU)."""
    import expressionmatrix2_b200 as em2
    eng = em2.Engine(0)
    yield eng
    eng.close()
