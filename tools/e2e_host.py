"""API-level end-to-end time of the reference's own call: ExpressionMatrix::findSimilarPairs4 of the C++ host layer
(expressionmatrix2_b200/host, pybind11 module ExpressionMatrix2) on a DATA DIRECTORY -- memory-mapped
CellExpressionCounts in, SimilarPairs-<name>-{Info,Pairs,CellInfo} files out -- with everything the reference's call
contains inside the timed region: hyperplane generation (Lsh::generateLshVectors), creation of the SimilarPairs files,
the device job on all visible GPUs straight from / into the mapped (pageable) files through the library's pinned bounce
buffers, and the msync of the result files when the objects close.  Ingest (addCells) is outside, as in the reference.

    python tools/e2e_host.py --workload c2 [--dir /dev/shm/em2-e2e] [--repeat 3]
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import benchdata as bd  # noqa: E402

SHAPES = {
    "c1": dict(cells=10_000, genes=20_000, nnz_per_cell=1000, clusters=64),
    "c2": dict(cells=100_000, genes=30_000, nnz_per_cell=1500, clusters=64),
    "m1": dict(cells=1_000_000, genes=30_000, nnz_per_cell=1500, clusters=512),
    "m1s": dict(cells=60_000, genes=30_000, nnz_per_cell=1500, clusters=32),
    "c3": dict(cells=1_300_000, genes=28_000, nnz_per_cell=2000, clusters=64),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=sorted(SHAPES))
    ap.add_argument("--dir", default=None)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--out", default=None, help="write the JSON line here (the C++ layer prints its progress to stdout)")
    args = ap.parse_args()
    w = SHAPES[args.workload]
    base = args.dir or ("/dev/shm/em2-e2e" if os.path.isdir("/dev/shm") else "/tmp/em2-e2e")
    shutil.rmtree(base, ignore_errors=True)
    from expressionmatrix2_b200 import hostmodule
    hostmodule.build()
    M = hostmodule.load()
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    N, G, m = w["cells"], w["genes"], w["nnz_per_cell"]
    t0 = time.time()
    e = M.ExpressionMatrix(base)
    e.addGenes(G)
    step = 50_000
    for b in range(0, N, step):          # ingest in slabs: the generator runs on the GPU, addCells appends to the mapped file
        n = min(N, b + step) - b
        genes, counts = bd.counts_to_numpy(bd.gen_counts(b, b + n, G, m, clusters=w["clusters"], device=dev))
        e.addCells(bd.toc_of(n, m), genes, counts)
    ingest_s = time.time() - t0
    assert e.cellCount() == N
    runs = []
    for r in range(args.repeat):
        name = f"bench{r}"
        t0 = time.perf_counter()
        e.findSimilarPairs4(similarPairsName=name, k=args.k, similarityThreshold=0.2, lshCount=1024, seed=231)
        dt = time.perf_counter() - t0
        runs.append(dict(seconds=dt, device_signature_ms=e.lastSignatureMs, device_scan_ms=e.lastScanMs))
    ids, sims, used = M.readSimilarPairs(base, "bench0")
    out = dict(tool="tools/e2e_host.py", api="ExpressionMatrix.findSimilarPairs4 (C++ host layer, data directory)", workload=args.workload,
               cells=N, genes=G, nnz_per_cell=m, k=args.k, lsh=1024, gpu=M.gpuName(), directory=base, ingest_seconds_untimed=ingest_s,
               runs=runs, best_seconds=min(x["seconds"] for x in runs), cell_pairs_per_s=N * (N - 1) / 2 / min(x["seconds"] for x in runs),
               mean_neighbours_stored=float(np.asarray(used).mean()),
               includes="hyperplane generation, SimilarPairs file creation, H2D from the mapped counts file and D2H into the mapped "
                        "Pairs file through pinned bounce buffers, msync on close")
    if args.out:
        with open(args.out, "w") as f:
            f.write(json.dumps(out) + "\n")
    else:
        sys.stdout.flush()
        print(json.dumps(out), flush=True)
    shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    main()
