"""Row-block partition of the LSH job over the GPUs of one box (one process per GPU).

The path shards by cells: rank p owns the contiguous block of cells
[p*S, (p+1)*S) with S = ceil(N/P); it builds the signatures of its block, ONE all-gather over
NCCL/NVLink makes every signature visible everywhere, and each rank then scans its rows against all
N columns.  Neighbour lists never cross GPUs: every rank returns (or writes into the mmapped
SimilarPairs file range of) its own rows.  There is no second collective.

The host logic here is backend agnostic (NCCL on GPUs; gloo in the CPU tests, where the two compute
stages are injected), and contains no arithmetic of the path itself.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Partition:
    cell_count: int
    world_size: int
    rank: int

    @property
    def shard(self) -> int:
        """Rows per rank (the last rank may own fewer real cells; shards are padded to this size).  The rule of
        em2_dist_partition (csrc/multi.cu): whole super blocks of 256 cells per rank when there is more than one."""
        if self.world_size <= 1:
            return self.cell_count
        per = (self.cell_count + self.world_size - 1) // self.world_size
        return (per + 255) // 256 * 256

    @property
    def row_begin(self) -> int:
        return min(self.cell_count, self.rank * self.shard)

    @property
    def row_end(self) -> int:
        return min(self.cell_count, (self.rank + 1) * self.shard)

    @property
    def rows(self) -> int:
        return self.row_end - self.row_begin

    def slice_csr(self, toc: np.ndarray, *arrays):
        """This rank's rows of a CSR matrix, re-based to start at 0."""
        b, e = int(toc[self.row_begin]), int(toc[self.row_end])
        local_toc = (toc[self.row_begin:self.row_end + 1] - toc[self.row_begin]).astype(np.uint64)
        return (local_toc,) + tuple(a[b:e] for a in arrays)


def all_gather_signatures(local_sig, partition: Partition, group=None):
    """local_sig: torch int64 tensor [rows, W] on this rank's device.  Returns [shard*P, W]; only the
    first cell_count rows are real (padding rows of the last shard are zero)."""
    import torch
    import torch.distributed as dist

    W = local_sig.shape[1]
    shard = partition.shard
    if partition.world_size == 1:
        return local_sig
    if local_sig.shape[0] != shard:
        pad = torch.zeros((shard, W), dtype=local_sig.dtype, device=local_sig.device)
        pad[: local_sig.shape[0]] = local_sig
        local_sig = pad
    out = torch.empty((shard * partition.world_size, W), dtype=local_sig.dtype, device=local_sig.device)
    dist.all_gather_into_tensor(out, local_sig.contiguous(), group=group)
    return out


def run_sharded(partition: Partition, toc, gene_ids, counts, signatures_fn, scan_fn, group=None, device="cpu"):
    """Generic driver used by the gloo tests and by bench.py's multi-GPU path.

    signatures_fn(local_toc, local_genes, local_counts) -> uint64 ndarray / int64 tensor [rows, W]
    scan_fn(all_signatures [>=N, W], row_begin, row_end) -> (ids, sims, used) for this rank's rows
    Returns this rank's (ids, sims, used)."""
    import torch

    local = partition.slice_csr(toc, gene_ids, counts)
    sig = signatures_fn(*local)
    if isinstance(sig, np.ndarray):
        sig = torch.from_numpy(np.ascontiguousarray(sig).view(np.int64)).to(device)
    full = all_gather_signatures(sig, partition, group)[: partition.cell_count]
    return scan_fn(full, partition.row_begin, partition.row_end)


def gather_lists_to_rank0(partition: Partition, ids, sims, used, group=None):
    """Host-side collection of the per-rank lists (numpy) on rank 0, in cell order."""
    import torch.distributed as dist

    if partition.world_size == 1:
        return ids, sims, used
    objs = [None] * partition.world_size if partition.rank == 0 else None
    dist.gather_object((ids, sims, used), objs, dst=0, group=group)
    if partition.rank != 0:
        return None
    return tuple(np.concatenate([o[i] for o in objs], axis=0) for i in range(3))
