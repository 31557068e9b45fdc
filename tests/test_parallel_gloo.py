"""Host-side logic of the multi-GPU partition on CPU: world_size 2, gloo backend.  The two compute stages
are injected (the oracle stands in for the CUDA kernels here, tests may use it); what is tested is the
row-block partition, the padded all-gather of signature shards and the per-rank list assembly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from expressionmatrix2_b200 import synthetic
from expressionmatrix2_b200.parallel import Partition, gather_lists_to_rank0, run_sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, G, L, k, thr, out_path):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    toc, genes, counts = synthetic.gen_expression_matrix(N, G, 0.06, seed=5, mode="clustered", clusters=6)
    U = oracle.generate_lsh_vectors(G, L, 231)
    part = Partition(N, world, rank)

    def signatures_fn(ltoc, lgenes, lcounts):
        s1, _ = oracle.cell_sums(ltoc, lcounts)
        return oracle.signatures(ltoc, lgenes, lcounts, s1, U)[0]

    def scan_fn(full, b, e):
        sig = full.numpy().view(np.uint64)
        assert sig.shape[0] == N
        return oracle.topk(sig, L, k, thr, b, e)[:3]

    ids, sims, used = run_sharded(part, toc, genes, counts, signatures_fn, scan_fn)
    assert ids.shape[0] == part.rows
    res = gather_lists_to_rank0(part, ids, sims, used)
    if rank == 0:
        np.savez(out_path, ids=res[0], sims=res[1], used=res[2])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [301, 128])      # ragged (last shard shorter) and even
def test_sharded_job_equals_single_process(tmp_path, oracle, N):
    G, L, k, thr = 150, 128, 7, 0.2
    port = _free_port()
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, port, N, G, L, k, thr, out), nprocs=2, join=True)
    got = np.load(out)
    toc, genes, counts = synthetic.gen_expression_matrix(N, G, 0.06, seed=5, mode="clustered", clusters=6)
    U = oracle.generate_lsh_vectors(G, L, 231)
    s1, _ = oracle.cell_sums(toc, counts)
    sig, _ = oracle.signatures(toc, genes, counts, s1, U)
    ids, sims, used, _ = oracle.topk(sig, L, k, thr)
    assert np.array_equal(got["ids"], ids) and np.array_equal(got["sims"], sims) and np.array_equal(got["used"], used)


def test_partition_covers_all_rows():
    for N in (1, 7, 100, 1001):
        for P in (1, 2, 3, 8):
            rows = []
            for r in range(P):
                p = Partition(N, P, r)
                assert 0 <= p.row_begin <= p.row_end <= N and p.rows <= p.shard
                rows += list(range(p.row_begin, p.row_end))
            assert rows == list(range(N))


# ---------------------------------------------------------------------------------------------------------------
# The multi-GPU SYMMETRIC scan's data flow (csrc/scan_mma.cu, runSymmetric / launchScanSymDist) restated on the host:
# every rank owns the super blocks of its row block; near window (row direction only, both owners visit a pair), far
# sweep (offsets w+1 .. S/2, both directions; the offset S/2 of an even S row-only), column-direction candidates
# travel to the owner of their column cell, every rank merges its own rows' candidates with what it received.
# The oracle's distances stand in for the tcgen05 tiles; bounds only prune, so without them the merged top-k must equal
# the oracle's lists exactly -- which checks partition, schedule, ownership and exchange under a real process group.
# ---------------------------------------------------------------------------------------------------------------
def _sym_worker(rank, world, port, N, L, k, thr, w, out_path):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = Partition(N, world, rank)
    sig_all = synthetic.gen_signatures(N, L, seed=N + L, clusters=4)
    # signature all-gather (padded shards), as the product does
    local = torch.zeros((part.shard, sig_all.shape[1]), dtype=torch.int64)
    local[: part.rows] = torch.from_numpy(sig_all[part.row_begin:part.row_end].view(np.int64).copy())
    gathered = torch.empty((part.shard * world, sig_all.shape[1]), dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, local)
    sig = gathered.numpy().view(np.uint64)[:N]
    assert np.array_equal(sig, sig_all)
    mm = oracle.mismatch_max(L, thr)
    S = (N + 255) // 256
    w = min(w, (S - 1) // 2)
    half = S // 2 if S % 2 == 0 else 0
    own_supers = range(part.row_begin // 256, (part.row_end + 255) // 256)
    mine = {c: [] for c in range(part.row_begin, part.row_end)}        # row-direction candidates of my cells
    outbox = [[] for _ in range(world)]                                 # column-direction candidates by owner rank

    def cells_of(sb):
        return range(sb * 256, min(N, sb * 256 + 256))

    for A in own_supers:
        rows = [c for c in cells_of(A) if part.row_begin <= c < part.row_end]
        dist_rows = {c: oracle.mismatch_row(sig, c) for c in rows}
        offsets = [(d, False) for d in range(-w, w + 1)] + [(d, d != half) for d in range(w + 1, S // 2 + 1)]
        for d, col_dir in offsets:
            C = (A - d) % S
            for c in rows:
                for j in cells_of(C):
                    m = int(dist_rows[c][j])
                    if j == c or m > mm:
                        continue
                    mine[c].append((m, j))
                    if col_dir:
                        outbox[min(world - 1, j // part.shard)].append((j, m, c))
    everything = [None] * world      # gloo has no all_to_all: every rank gathers every outbox and keeps its share
    dist.all_gather_object(everything, outbox)
    for src in range(world):
        for (j, m, c) in everything[src][rank]:
            assert part.row_begin <= j < part.row_end
            mine[j].append((m, c))
    table = oracle.similarity_table(L).astype(np.float32)
    ids = np.zeros((part.rows, k), np.uint32)
    sims = np.zeros((part.rows, k), np.float32)
    used = np.zeros(part.rows, np.uint32)
    for c in range(part.row_begin, part.row_end):
        cand = sorted(mine[c])
        assert len(cand) == len(set(cand)), "a pair was produced twice"
        best = cand[:k]
        used[c - part.row_begin] = len(best)
        for i, (m, j) in enumerate(best):
            ids[c - part.row_begin, i] = j
            sims[c - part.row_begin, i] = table[m]
    res = gather_lists_to_rank0(part, ids, sims, used)
    if rank == 0:
        np.savez(out_path, ids=res[0], sims=res[1], used=res[2])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,w", [(1100, 1), (1536, 0), (700, 16)])      # 5 / 6 / 3 super blocks; rank 1 owns 332 / 768 / 188 cells
def test_symmetric_scan_data_flow_across_two_ranks(tmp_path, oracle, N, w):
    L, k, thr = 128, 9, 0.2
    port = _free_port()
    out = str(tmp_path / "sym.npz")
    mp.spawn(_sym_worker, args=(2, port, N, L, k, thr, w, out), nprocs=2, join=True)
    got = np.load(out)
    sig = synthetic.gen_signatures(N, L, seed=N + L, clusters=4)
    ids, sims, used, _ = oracle.topk(sig, L, k, thr)
    assert np.array_equal(got["used"], used) and np.array_equal(got["ids"], ids)
    assert np.array_equal(got["sims"].view(np.uint32), sims.view(np.uint32))
