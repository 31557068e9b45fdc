"""The C++ host layer (expressionmatrix2_b200/host) behind the reference's API names, through the pybind11
module `ExpressionMatrix2`.  CPU tests: the memory-mapped file formats are interchangeable with the
reference's own classes (both directions, hashes included).  GPU tests: findSimilarPairs4 /
computeLshSignatures on gene/cell subsets against the oracle."""
import os

import numpy as np
import pytest

from expressionmatrix2_b200 import synthetic


@pytest.fixture(scope="module")
def M():
    from expressionmatrix2_b200 import hostmodule
    hostmodule.build()
    return hostmodule.load()


def _make_matrix(M, path, N=200, G=90, seed=3):
    toc, genes, counts = synthetic.gen_expression_matrix(N, G, 0.08, seed=seed, mode="clustered", clusters=5)
    e = M.ExpressionMatrix(str(path))
    e.addGenes(G)
    e.addCells(toc, genes, counts)
    assert e.cellCount() == N and e.geneCount() == G
    return e, (toc, genes, counts)


def test_module_has_reference_api(M):
    e = M.ExpressionMatrix
    for name in ("findSimilarPairs4", "findSimilarPairs0", "findSimilarPairs7", "computeLshSignatures", "geneCount", "cellCount"):
        assert hasattr(e, name)
    doc = e.findSimilarPairs4.__doc__
    for token in ("geneSetName: str = 'AllGenes'", "cellSetName: str = 'AllCells'", "similarPairsName: str",
                  "k: typing.SupportsInt = 100", "similarityThreshold: typing.SupportsFloat = 0.2",
                  "lshCount: typing.SupportsInt = 1024", "seed: typing.SupportsInt = 231"):
        assert token.split(":")[0] in doc and token.split("=")[-1].strip() in doc


def test_similar_pairs_files_are_read_by_the_reference(M, oracle, tmp_path):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libem2ref.so not present")
    e, _ = _make_matrix(M, tmp_path / "data")
    N, k = 200, 6
    rng = np.random.default_rng(0)
    ids = rng.integers(0, N, (N, k)).astype(np.uint32)
    sims = np.sort(rng.random((N, k)).astype(np.float32), axis=1)[:, ::-1].copy()
    used = rng.integers(0, k + 1, N).astype(np.uint32)
    M.writeSimilarPairs(str(tmp_path / "data"), "mine", "AllGenes", "AllCells", ids, sims, used)
    rid, rsim, rused = oracle.ref_read_similar_pairs(str(tmp_path / "data"), "mine")   # reference's own reader
    assert np.array_equal(rused, used)
    for c in range(N):
        assert np.array_equal(rid[c, :used[c]], ids[c, :used[c]])
        assert np.array_equal(rsim[c, :used[c]], sims[c, :used[c]])
    # header bytes: 256-byte header, page-rounded size, magic number of MemoryMapped::Vector
    raw = np.fromfile(tmp_path / "data" / "SimilarPairs-mine-Pairs", np.uint64, 7)
    size = os.path.getsize(tmp_path / "data" / "SimilarPairs-mine-Pairs")
    assert raw[0] == 256 and raw[1] == 8 and raw[2] == N * k and raw[4] == size and size % 4096 == 0
    assert raw[6] == 0xa3756fd4b5d8bcc1
    info = np.fromfile(tmp_path / "data" / "SimilarPairs-mine-Info", np.uint64, 7)
    assert info[6] == 0xb7756f4515d8bc94 and info[1] == 536


def test_reference_files_are_read_by_the_host_layer(M, oracle, tmp_path):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libem2ref.so not present")
    N, G, k = 77, 31, 4
    rng = np.random.default_rng(1)
    ids = rng.integers(0, N, (N, k)).astype(np.uint32)
    sims = rng.random((N, k)).astype(np.float32)
    used = rng.integers(0, k + 1, N).astype(np.uint32)
    d = tmp_path / "refdata"
    d.mkdir()
    oracle.ref_write_similar_pairs(str(d), "theirs", G, ids, sims, used)    # reference's own writer + sets
    mid, msim, mused = M.readSimilarPairs(str(d), "theirs")                 # validates both hashes
    assert np.array_equal(mused, used)
    for c in range(N):
        assert np.array_equal(mid[c, :used[c]], ids[c, :used[c]])
        assert np.array_equal(msim[c, :used[c]], sims[c, :used[c]])


def test_lookup_errors_match_the_reference(M, tmp_path):
    e, _ = _make_matrix(M, tmp_path / "data", N=20, G=10)
    with pytest.raises(RuntimeError, match="Gene set Nope does not exist."):
        e.findSimilarPairs4(geneSetName="Nope", similarPairsName="x")
    with pytest.raises(RuntimeError, match="Cell set Nope does not exist."):
        e.findSimilarPairs4(cellSetName="Nope", similarPairsName="x")
    e.createCellSet("Empty", [])
    with pytest.raises(RuntimeError, match="Cell set Empty is empty."):
        e.findSimilarPairs4(cellSetName="Empty", similarPairsName="x")


def test_device_limits_are_checked_before_anything_is_written(M, tmp_path):
    """k <= 1024 and lshCount <= 65535 are limits of the device path the reference does not have: the host layer refuses
    them BEFORE it creates the SimilarPairs object, and a call that fails on the device removes the object it created
    (here: no GPU) -- no empty but valid-looking SimilarPairs-<name> set stays in the data directory."""
    import glob
    e, _ = _make_matrix(M, tmp_path / "data", N=20, G=10)
    with pytest.raises(RuntimeError, match="k must be in"):
        e.findSimilarPairs4(similarPairsName="TooWide", k=5000)
    with pytest.raises(RuntimeError, match="lshCount must be in"):
        e.findSimilarPairs4(similarPairsName="TooLong", lshCount=70000)
    with pytest.raises(RuntimeError, match="k must be in"):
        e.findSimilarPairs0(similarPairsName="TooWide0", k=0)
    import torch
    if not torch.cuda.is_available():      # the device call itself fails: the object it created must be gone again
        with pytest.raises(RuntimeError, match="GPU"):
            e.findSimilarPairs4(similarPairsName="NoDevice", k=5)
    assert glob.glob(str(tmp_path / "data" / "SimilarPairs-*")) == []


def test_reopen_existing_directory(M, tmp_path):
    e, _ = _make_matrix(M, tmp_path / "data", N=30, G=12)
    e.createGeneSet("Some", [1, 5, 7])
    del e
    e2 = M.ExpressionMatrix(str(tmp_path / "data"))
    assert e2.cellCount() == 30 and e2.geneCount() == 12


@pytest.mark.gpu
def test_find_similar_pairs4_end_to_end(M, oracle, tmp_path):
    N, G, L, k, thr = 1500, 600, 1024, 20, 0.2
    e, (toc, genes, counts) = _make_matrix(M, tmp_path / "data", N=N, G=G, seed=8)
    e.findSimilarPairs4(similarPairsName="Lsh", k=k, similarityThreshold=thr, lshCount=L, seed=231)
    ids, sims, used = e.getSimilarPairs("Lsh")
    U = oracle.generate_lsh_vectors(G, L, 231)
    s1, _ = oracle.cell_sums(toc, counts)
    sig, _ = oracle.signatures(toc, genes, counts, s1, U)
    wi, ws, wu, _ = oracle.topk(sig, L, k, thr)
    assert np.array_equal(used, wu) and np.array_equal(ids, wi)
    assert np.array_equal(sims.view(np.uint32), ws.view(np.uint32))
    assert not os.path.exists(tmp_path / "data" / "tmp-Lsh-Lsh-Signatures")       # temporaries removed
    if oracle.have_ref():                                                        # and the reference reads it
        rid, rsim, rused = oracle.ref_read_similar_pairs(str(tmp_path / "data"), "Lsh")
        assert np.array_equal(rused, wu) and np.array_equal(rid * (np.arange(k)[None, :] < wu[:, None]), wi)


@pytest.mark.gpu
def test_subsets_and_persistent_signatures(M, oracle, tmp_path):
    N, G, L = 900, 400, 256
    e, (toc, genes, counts) = _make_matrix(M, tmp_path / "data", N=N, G=G, seed=9)
    gene_ids = np.arange(0, G, 3, dtype=np.uint32)            # every third gene
    cell_ids = np.arange(5, N, 2, dtype=np.uint32)            # odd subset of cells
    e.createGeneSet("Thirds", gene_ids.tolist())
    e.createCellSet("Odd", cell_ids.tolist())
    e.computeLshSignatures(geneSetName="Thirds", cellSetName="Odd", lshName="S", lshCount=L, seed=7)
    sig = e.getLshSignatures("S")
    # oracle on the re-indexed subset
    local = -np.ones(G, np.int64)
    local[gene_ids] = np.arange(len(gene_ids))
    rows_g, rows_c, stoc = [], [], [0]
    for c in cell_ids:
        g = genes[int(toc[c]):int(toc[c + 1])]
        x = counts[int(toc[c]):int(toc[c + 1])]
        keep = local[g] >= 0
        rows_g.append(local[g][keep].astype(np.uint32))
        rows_c.append(x[keep])
        stoc.append(stoc[-1] + int(keep.sum()))
    sg, sc, stoc = np.concatenate(rows_g), np.concatenate(rows_c), np.array(stoc, np.uint64)
    U = oracle.generate_lsh_vectors(len(gene_ids), L, 7)
    s1, _ = oracle.cell_sums(stoc, sc)
    want, _ = oracle.signatures(stoc, sg, sc, s1, U)
    assert np.array_equal(sig, want)
    e.findSimilarPairs4(geneSetName="Thirds", cellSetName="Odd", similarPairsName="Sub", k=9, similarityThreshold=0.1,
                        lshCount=L, seed=7)
    ids, sims, used = e.getSimilarPairs("Sub")
    wi, ws, wu, _ = oracle.topk(want, L, 9, 0.1)
    assert np.array_equal(used, wu) and np.array_equal(ids, wi) and np.array_equal(sims, ws)


@pytest.mark.gpu
def test_find_similar_pairs7_on_persistent_signatures(M, oracle, tmp_path):
    """computeLshSignatures -> findSimilarPairs7 with the reference's argument names: the stored lists must equal the
    reference loops run over the reference's own Lsh object on the same signatures."""
    if not oracle.have_ref():
        pytest.skip("reference build (oracle/_ref) not present")
    N, G, L, k, thr = 1200, 500, 256, 15, 0.3
    e, _ = _make_matrix(M, tmp_path / "data", N=N, G=G, seed=11)
    e.computeLshSignatures(lshName="S", lshCount=L, seed=231)
    e.findSimilarPairs7(lshName="S", similarPairsName="Bucketed", k=k, similarityThreshold=thr, lshSliceLengths=[14, 10, 6],
                        maxCheck=150, log2BucketCount=9)
    ids, sims, used = e.getSimilarPairs("Bucketed")
    sig = e.getLshSignatures("S")
    with oracle.Reference.from_signatures(sig, L) as ref:
        wi, ws, wu = ref.find_similar_pairs7(k, thr, [14, 10, 6], 150, 9)
    assert np.array_equal(used, wu) and np.array_equal(ids, wi)
    assert np.array_equal(sims.view(np.uint32), ws.view(np.uint32))
    with pytest.raises(RuntimeError):
        e.findSimilarPairs7(lshName="S", similarPairsName="Bad", k=k, similarityThreshold=thr, lshSliceLengths=[6, 10],
                            maxCheck=150, log2BucketCount=9)
