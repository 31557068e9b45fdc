"""Synthetic sparse cell x gene expression matrices and synthetic signatures.

The reference ships no generator; shapes follow BASELINE.json's configs.
Rows are in the layout the reference stores (src/ExpressionMatrix.cpp:265-277): per cell, distinct
gene ids in ascending order with float counts; counts are integer valued like 10x UMI counts
(src/ExpressionMatrixHdf5.cpp:172).
"""
from __future__ import annotations

import numpy as np

PAIR_DTYPE = np.dtype([("gene", "<u4"), ("count", "<f4")])  # pair<GeneId,float>, 8 bytes


def gen_expression_matrix(cell_count: int, gene_count: int, density: float, seed: int = 12345,
                          mode: str = "clustered", clusters: int = 64):
    """Return (toc uint64[N+1], gene_ids uint32[nnz], counts float32[nnz]).

    mode "iid": every gene equally likely in every cell (all similarities near 0: worst case for ties).
    mode "clustered": cells belong to `clusters` groups, each with its own gene-propensity profile, so
    that real neighbour structure exists.
    """
    rng = np.random.default_rng(seed)
    nnz_per_cell = rng.binomial(gene_count, density, size=cell_count).astype(np.int64)
    nnz_per_cell = np.clip(nnz_per_cell, 1, gene_count)
    toc = np.zeros(cell_count + 1, np.uint64)
    toc[1:] = np.cumsum(nnz_per_cell).astype(np.uint64)
    nnz = int(toc[-1])
    gene_ids = np.empty(nnz, np.uint32)
    if mode == "iid":
        # Sample distinct genes per cell by taking the nnz smallest of G random keys, in blocks.
        block = max(1, min(cell_count, (1 << 24) // max(gene_count, 1)))
        for b in range(0, cell_count, block):
            e = min(cell_count, b + block)
            keys = rng.random((e - b, gene_count), dtype=np.float32)
            for i in range(b, e):
                n = int(nnz_per_cell[i])
                idx = np.argpartition(keys[i - b], n - 1)[:n]
                idx.sort()
                gene_ids[int(toc[i]):int(toc[i + 1])] = idx
    elif mode == "clustered":
        # Each cluster prefers a random 4*density fraction of the genes (8x more likely than the rest).
        member = rng.integers(0, clusters, size=cell_count)
        hot = max(1, int(gene_count * min(0.5, 4 * density)))
        cluster_hot = [rng.choice(gene_count, hot, replace=False) for _ in range(clusters)]
        for i in range(cell_count):
            n = int(nnz_per_cell[i])
            n_hot = min(hot, rng.binomial(n, 0.7))
            a = rng.choice(cluster_hot[member[i]], n_hot, replace=False)
            # fill the rest uniformly, resolve collisions by oversampling + unique
            need = n - n_hot
            got = np.unique(a)
            while len(got) < n:
                extra = rng.integers(0, gene_count, size=2 * (n - len(got)) + 8)
                got = np.unique(np.concatenate([got, extra]))
            if len(got) > n:
                # keep all hot genes, drop random others
                drop = rng.choice(len(got), len(got) - n, replace=False)
                got = np.delete(got, drop)
            gene_ids[int(toc[i]):int(toc[i + 1])] = got
            del need
    else:
        raise ValueError("unknown mode " + mode)
    counts = (1 + rng.geometric(0.3, size=nnz)).astype(np.float32)
    return toc, gene_ids, counts


def gen_expression_matrix_fast(cell_count: int, gene_count: int, nnz_per_cell: int, seed: int = 12345,
                               clusters: int = 64):
    """Large-shape generator (bench): fixed nnz per cell, vectorised.

    Every cell draws `nnz_per_cell` gene ids as sorted(unique(cluster_offset + stride pattern)); built
    so that generation is O(nnz) numpy work without per-cell Python loops.  Rows are strictly ascending
    and duplicate free.
    """
    rng = np.random.default_rng(seed)
    n, m = cell_count, nnz_per_cell
    # Stratified sampling: one gene from each of m equal strata of [0, G) -> ascending, distinct.
    edges = np.linspace(0, gene_count, m + 1).astype(np.int64)
    width = np.diff(edges)
    assert width.min() >= 1, "nnz_per_cell must not exceed gene_count"
    member = rng.integers(0, clusters, size=n)
    # Cluster profile: a preferred offset inside every stratum; cells use it with probability 0.6.
    profile = (rng.random((clusters, m)) * width[None, :]).astype(np.int64)
    gene_ids = np.empty((n, m), np.uint32)
    chunk = max(1, (1 << 22) // m)
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        own = (rng.random((e - b, m)) * width[None, :]).astype(np.int64)
        use = rng.random((e - b, m)) < 0.6
        off = np.where(use, profile[member[b:e]], own)
        gene_ids[b:e] = (edges[None, :-1] + off).astype(np.uint32)
    counts = (1 + rng.geometric(0.3, size=n * m)).astype(np.float32)
    toc = (np.arange(n + 1, dtype=np.uint64) * np.uint64(m))
    return toc, gene_ids.reshape(-1), counts


def to_pairs(gene_ids: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Interleave into the reference's on-disk element type pair<GeneId,float> (8 bytes, AoS)."""
    out = np.empty(len(gene_ids), PAIR_DTYPE)
    out["gene"] = gene_ids
    out["count"] = counts
    return out


def gen_signatures(cell_count: int, lsh_count: int, seed: int = 12345, clusters: int = 0,
                   flip_fraction: float = 0.12, centre_seed: int | None = None) -> np.ndarray:
    """Signatures-only synthetic input for scan-only runs: uint64 [N, W], MSB-first bit order.

    clusters == 0: iid random bits (all Hamming distances near L/2).
    clusters > 0 : each cell = its cluster centre with each bit flipped with probability
    `flip_fraction` (planted neighbours).  centre_seed: draw the cluster centres from their own generator, so
    that shards generated with different `seed`s (one per rank) share the same clusters."""
    rng = np.random.default_rng(seed)
    W = (lsh_count - 1) // 64 + 1
    if clusters <= 0:
        sig = rng.integers(0, 1 << 63, size=(cell_count, W), dtype=np.uint64)
        sig ^= rng.integers(0, 2, size=(cell_count, W), dtype=np.uint64) << np.uint64(63)
    else:
        crng = rng if centre_seed is None else np.random.default_rng(centre_seed)
        centres = crng.integers(0, 1 << 63, size=(clusters, W), dtype=np.uint64)
        centres ^= crng.integers(0, 2, size=(clusters, W), dtype=np.uint64) << np.uint64(63)
        member = rng.integers(0, clusters, size=cell_count)
        sig = centres[member].copy()
        # flip mask with P(bit)=flip_fraction, built from AND/OR of uniform words (p = 1/8 = 0.125 approx)
        p = flip_fraction
        mask = np.zeros((cell_count, W), np.uint64)
        # binary expansion of p to 6 bits
        bits = [(int(p * 64) >> i) & 1 for i in range(6)]
        for b in bits:  # LSB first: mask = b ? (r | mask) : (r & mask)
            r = rng.integers(0, 1 << 63, size=(cell_count, W), dtype=np.uint64)
            r ^= rng.integers(0, 2, size=(cell_count, W), dtype=np.uint64) << np.uint64(63)
            mask = (r | mask) if b else (r & mask)
        sig ^= mask
    pad = W * 64 - lsh_count
    if pad:
        sig[:, -1] &= np.uint64(~((1 << pad) - 1) & 0xFFFFFFFFFFFFFFFF)
    return sig
