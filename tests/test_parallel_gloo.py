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
