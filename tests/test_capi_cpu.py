"""CPU tests of the product library: it loads, exports every symbol include/em2b200.h declares, its
host-side helpers agree with the oracle, and it FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_cases, load_golden


@pytest.fixture(scope="module")
def em2():
    import expressionmatrix2_b200 as m
    from expressionmatrix2_b200 import build
    build.build()
    return m


def test_library_exports_every_declared_symbol(em2):
    header = open(os.path.join(ROOT, "include", "em2b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(em2_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 17
    L = ctypes.CDLL(em2.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert em2.lib().em2_abi_version() == 2


def test_struct_layouts_match_reference_types(em2):
    assert em2.PAIR_DTYPE.itemsize == 8 and em2.SIMPAIR_DTYPE.itemsize == 8   # pair<uint32,float>
    assert ctypes.sizeof(em2.Stats) == 8 * 8 + 7 * 8 + 8 + 8 + 2 * 8 + 8 + 8


@pytest.mark.parametrize("name", golden_cases())
def test_host_generator_matches_golden(em2, name):
    g = load_golden(name)
    G, L = int(g["gene_count"]), int(g["lsh_count"])
    U = em2.generate_lsh_vectors(G, L, int(g["seed"]))
    assert np.array_equal(U[:3, :8], g["U_head"])
    chk = np.array([U.sum(), np.abs(U).sum(), (U * np.arange(1, G + 1)[:, None]).sum()])
    assert np.array_equal(chk, g["U_checksum"])
    assert np.array_equal(em2.similarity_table(L), g["table"])


def test_host_helpers_match_oracle(em2, oracle):
    for G, L, seed in ((37, 65, 1), (300, 128, 231), (11, 1, 5), (1, 1024, 9)):
        assert np.array_equal(em2.generate_lsh_vectors(G, L, seed), oracle.generate_lsh_vectors(G, L, seed))
    for L in (1, 63, 64, 1000, 4096):
        assert np.array_equal(em2.similarity_table(L), oracle.similarity_table(L))
        for thr in (-1.0, -0.5, 0.0, 0.2, 0.9999, 1.0, 2.0):
            assert em2.mismatch_max(L, thr) == oracle.mismatch_max(L, thr)


def test_no_cpu_fallback(em2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(em2.Em2Error) as e:
        em2.Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "expressionmatrix2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "em2_oracle" not in text and "libem2ref" not in text, f


def test_dist_partition_rule(em2):
    """em2_dist_partition (no GPU needed): rank r owns [r S, (r + 1) S) with S a multiple of 256 when there is more than
    one rank; the Python mirror used by bench.py / the gloo tests (parallel.Partition) follows the same rule."""
    from expressionmatrix2_b200.parallel import Partition
    for N in (0, 1, 255, 256, 257, 1000, 100_000, 1_000_000, 1_300_000):
        for P in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(P):
                b, e, sh = em2.dist_partition(N, P, r)
                assert b == covered and b <= e <= N and e - b <= sh
                assert P == 1 or sh % 256 == 0
                part = Partition(N, P, r)
                assert (part.row_begin, part.row_end, part.shard) == (b, e, sh)
                covered = e
            assert covered == N


def test_multi_gpu_entry_points_fail_loudly_without_a_gpu(em2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(em2.Em2Error) as e:
        em2.MultiEngine()
    assert "no CPU fallback" in str(e.value)
