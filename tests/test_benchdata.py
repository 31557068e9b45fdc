"""bench.py's counter-based synthetic generator (benchdata.py): any range of cells is reproducible anywhere -- the B200
arm generates its cells on the device per rank, the reference arm and the CPU baseline regenerate ranges on the host --
so the bytes must not depend on how the range is cut, and rows must have the layout the reference stores
(src/ExpressionMatrix.cpp:265-277: distinct gene ids ascending per cell)."""
import numpy as np
import torch

import benchdata as bd


def test_counts_are_range_independent_and_well_formed():
    G, m = 3000, 150
    whole = bd.gen_counts(0, 500, G, m, clusters=7)
    assert whole.dtype == torch.int64 and whole.numel() == 500 * m
    parts = torch.cat([bd.gen_counts(0, 123, G, m, clusters=7), bd.gen_counts(123, 124, G, m, clusters=7),
                       bd.gen_counts(124, 500, G, m, clusters=7)])
    assert torch.equal(whole, parts)
    genes, counts = bd.counts_to_numpy(whole)
    genes = genes.reshape(500, m).astype(np.int64)
    assert genes.min() >= 0 and genes.max() < G
    assert np.all(np.diff(genes, axis=1) > 0)                       # distinct and ascending
    assert np.all(counts >= 2) and np.all(counts == np.round(counts)) and counts.max() < 256      # integer UMI-like counts
    assert np.array_equal(bd.toc_of(500, m), np.arange(501, dtype=np.uint64) * m)
    other = bd.gen_counts(0, 500, G, m, seed=999, clusters=7)
    assert not torch.equal(whole, other)


def test_cells_of_a_cluster_share_genes():
    G, m, n, clusters = 3000, 150, 400, 5
    genes, _ = bd.counts_to_numpy(bd.gen_counts(0, n, G, m, clusters=clusters))
    genes = genes.reshape(n, m)
    member = bd.cluster_of(torch.arange(n), 12345, clusters).numpy()
    same, diff = [], []
    for i in range(0, 60):
        for j in range(i + 1, 60):
            shared = np.intersect1d(genes[i], genes[j]).size
            (same if member[i] == member[j] else diff).append(shared)
    assert np.mean(same) > 3 * np.mean(diff)                         # planted neighbour structure


def test_signatures_are_range_independent_with_zero_pad_bits():
    for L in (100, 1024):
        W = (L - 1) // 64 + 1
        whole = bd.gen_signatures(0, 300, L, clusters=6)
        assert whole.shape == (300, W)
        parts = torch.cat([bd.gen_signatures(0, 77, L, clusters=6), bd.gen_signatures(77, 300, L, clusters=6)])
        assert torch.equal(whole, parts)
        sig = whole.numpy().view(np.uint64)
        pad = W * 64 - L
        if pad:
            assert np.all(sig[:, -1] & np.uint64((1 << pad) - 1) == 0)      # MSB-first: the pad bits are the low bits of the last word
        bits = np.unpackbits(sig.view(np.uint8), axis=1)
        member = bd.cluster_of(torch.arange(300), 1000, 6).numpy()
        i, j = np.flatnonzero(member == member[0])[:2]
        k = np.flatnonzero(member != member[0])[0]
        assert (bits[i] != bits[j]).sum() < (bits[i] != bits[k]).sum()     # cluster mates are nearer than strangers
    iid = bd.gen_signatures(0, 50, 256, clusters=0).numpy().view(np.uint64)
    frac = np.unpackbits(iid.view(np.uint8)).mean()
    assert 0.45 < frac < 0.55
