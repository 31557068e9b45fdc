"""Synthetic inputs of bench.py -- counter-based, so that any range of cells can be produced anywhere (on a GPU by
the B200 arm, on the host by the reference arm) and is the same bytes everywhere.

This module belongs to the bench harness, not to the product: it imports neither `expressionmatrix2_b200` nor
`oracle`, so `bench.py --impl reference` never touches the product library through it.

Shape of the data (the reference ships no generator): every cell stores `nnz_per_cell` distinct
genes in ascending order -- one gene from each of `nnz_per_cell` equal strata of [0, gene_count) -- with integer-valued
float counts >= 2 (1 + a geometric(0.3) variate, like 10x UMI counts).  Cells belong to `clusters` groups; a cell takes
its cluster's preferred gene of a stratum with probability 0.6, its own random gene otherwise, which plants real
neighbour structure.  Every random decision is a splitmix64 hash of (seed, cell, stratum): no generator state.
"""
from __future__ import annotations

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    """Two's-complement view of a 64-bit constant (torch has no uint64 arithmetic)."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_GOLD = _i64(0x9E3779B97F4A7C15)
_MUL1 = _i64(0xBF58476D1CE4E5B9)
_MUL2 = _i64(0x94D049BB133111EB)


def _shr(z: torch.Tensor, s: int) -> torch.Tensor:
    """Logical right shift of int64 bit patterns."""
    return (z >> s) & ((1 << (64 - s)) - 1)


def mix64(z: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 bit patterns (wrapping arithmetic; identical on CPU and CUDA)."""
    z = z + _GOLD
    z = (z ^ _shr(z, 30)) * _MUL1
    z = (z ^ _shr(z, 27)) * _MUL2
    return z ^ _shr(z, 31)


def _geometric_thresholds(p: float = 0.3, bits: int = 16, kmax: int = 40) -> list[int]:
    """u < T[0] -> 1, u < T[1] -> 2, ... for a `bits`-bit uniform u: integer CDF of the geometric distribution."""
    return [int((1.0 - (1.0 - p) ** k) * (1 << bits)) for k in range(1, kmax + 1)]


def cluster_of(cells: torch.Tensor, seed: int, clusters: int) -> torch.Tensor:
    return _shr(mix64(cells ^ _i64(seed * 0x2545F4914F6CDD1D + 17)), 33) % clusters


def gen_counts(cell_begin: int, cell_end: int, gene_count: int, nnz_per_cell: int, seed: int = 12345, clusters: int = 64,
               device: str | torch.device = "cpu") -> torch.Tensor:
    """Stored counts of cells [cell_begin, cell_end) as int64 [cells * nnz_per_cell]: every element is the bit pattern of
    one pair<GeneId,float> (gene id in the low word, float count in the high word -- the reference's 8-byte AoS element,
    src/ExpressionMatrixSubset.hpp:37).  toc is implicit: cell c starts at (c - cell_begin) * nnz_per_cell."""
    m = nnz_per_cell
    assert gene_count >= m, "nnz_per_cell must not exceed gene_count"
    dev = torch.device(device)
    j = torch.arange(m, dtype=torch.int64, device=dev)
    edges = (torch.arange(m + 1, dtype=torch.int64, device=dev) * gene_count) // m
    width = edges[1:] - edges[:-1]
    thresholds = torch.tensor(_geometric_thresholds(), dtype=torch.int64, device=dev)
    cells = torch.arange(cell_begin, cell_end, dtype=torch.int64, device=dev)
    member = cluster_of(cells, seed, clusters)
    s0 = _i64(seed * 0xD6E8FEB86659FD93)
    h = mix64((cells[:, None] * m + j[None, :]) ^ s0)                       # per (cell, stratum)
    hp = mix64((member[:, None] * m + j[None, :]) ^ _i64(s0 + 0x5851F42D4C957F2D))   # per (cluster, stratum)
    own = (_shr(h, 40) * width[None, :]) >> 24                              # 24-bit uniform -> [0, width)
    pref = (_shr(hp, 40) * width[None, :]) >> 24
    use = (_shr(h, 24) & 0xFFFF) < int(0.6 * 65536)
    gene = edges[None, :-1] + torch.where(use, pref, own)
    u = h & 0xFFFF
    count = 2 + torch.searchsorted(thresholds, u.reshape(-1), right=True).reshape(u.shape)
    bits = count.to(torch.float32).view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return (gene | (bits << 32)).reshape(-1)


def counts_to_numpy(packed: torch.Tensor):
    """(gene_ids uint32[nnz], counts float32[nnz]) of a gen_counts() result."""
    a = packed.cpu().numpy().view(np.uint32).reshape(-1, 2)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]).view(np.float32)


def toc_of(cells: int, nnz_per_cell: int) -> np.ndarray:
    return np.arange(cells + 1, dtype=np.uint64) * np.uint64(nnz_per_cell)


def gen_signatures(cell_begin: int, cell_end: int, lsh_count: int, seed: int = 1000, clusters: int = 500,
                   flip_sixty_fourths: int = 7, device: str | torch.device = "cpu") -> torch.Tensor:
    """Signatures-only synthetic input (config 4): int64 bit patterns [cells, W], MSB-first like the reference's BitSet.
    Cell = its cluster's centre with every bit flipped with probability flip_sixty_fourths / 64 (7/64 = 11 %)."""
    dev = torch.device(device)
    W = (lsh_count - 1) // 64 + 1
    cells = torch.arange(cell_begin, cell_end, dtype=torch.int64, device=dev)
    w = torch.arange(W, dtype=torch.int64, device=dev)
    member = cluster_of(cells, seed, clusters) if clusters > 0 else cells
    centre = mix64((member[:, None] * W + w[None, :]) ^ _i64(seed * 0xA24BAED4963EE407 + 5))
    if clusters <= 0:
        sig = centre
    else:
        # flip mask with P(bit) = p/64 from the binary expansion of p over six independent uniform words
        mask = torch.zeros_like(centre)
        base = (cells[:, None] * W + w[None, :]) * 8
        for b in range(6):      # LSB first: mask = bit ? (r | mask) : (r & mask)
            r = mix64((base + b) ^ _i64(seed * 0x9FB21C651E98DF25 + 11))
            mask = (r | mask) if (flip_sixty_fourths >> b) & 1 else (r & mask)
        sig = centre ^ mask
    pad = W * 64 - lsh_count
    if pad:
        sig[:, -1] &= _i64(~((1 << pad) - 1))
    return sig
