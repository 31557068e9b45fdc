#!/usr/bin/env python
"""bench.py -- LSH cell-similarity hot path on B200 (BASELINE.json metric: cell-pairs/sec).

A step = one pass of the hot path over the synthetic batch: per-cell sums -> LSH signatures ->
(all-gather when N>1) -> all-pairs Hamming scan with fused top-k -> SimilarPairs payload.
Workload (default "c2" = BASELINE.json configs[1]): 100k cells x 30k genes, 5% density, L=1024, k=50,
similarityThreshold 0.2, synthetic clustered counts, hyperplanes from seed 231.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's own CPU code (oracle/_ref) on a bounded sample

`value`   : unordered cell pairs N(N-1)/2 per second of the whole job, inputs resident in HBM,
            CUDA-event timed, max over ranks.
`e2e`     : same metric through the reference-facing C-ABI call on HOST buffers (H2D + D2H inside).
`roofline`: the dominant kernel (the Hamming scan) against its governing pipe, measured live.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: cells, genes, nnz/cell, L, k, threshold
    "c1": dict(cells=10_000, genes=20_000, nnz_per_cell=1000, lsh=1024, k=50, thr=0.2,
               note="BASELINE configs[0]: 10k x 20k, 5% density"),
    "c2": dict(cells=100_000, genes=30_000, nnz_per_cell=1500, lsh=1024, k=50, thr=0.2,
               note="BASELINE configs[1]: 100k x 30k, 5% density"),
    "c2s": dict(cells=20_000, genes=30_000, nnz_per_cell=1500, lsh=1024, k=50, thr=0.2,
                note="reduced c2 for quick checks (NOT a bench line)"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    samples=len(sm))


def make_workload(w, seed=12345):
    from expressionmatrix2_b200 import synthetic
    import expressionmatrix2_b200 as em2
    toc, genes, counts = synthetic.gen_expression_matrix_fast(w["cells"], w["genes"], w["nnz_per_cell"], seed=seed)
    U = em2.generate_lsh_vectors(w["genes"], w["lsh"], 231)
    return toc, genes, counts, U


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), bounded sample per step
# ------------------------------------------------------------------------------------------------
def cpu_sample(w, toc, genes, counts, signatures, sig_cells=1024, loop_rows=2048):
    """Times the reference's two instrumented regions on a bounded sample of the workload:
    Lsh::computeCellLshSignatures on the first `sig_cells` cells (its own timer, Lsh.cpp:160,209) and the
    findSimilarPairs4 pair loop (ExpressionMatrixLsh.cpp:217,270) for the last `loop_rows` cells against
    all earlier cells.  Returns the extrapolated whole-job figures."""
    import oracle
    N, L, k, thr = w["cells"], w["lsh"], w["k"], w["thr"]
    kind = "reference" if oracle.have_ref() else "port"
    sig_cells = min(sig_cells, N)
    loop_rows = min(loop_rows, N)
    t0 = time.time()
    e = int(toc[sig_cells])
    if kind == "reference":
        with oracle.Reference.from_csr(toc[: sig_cells + 1], genes[:e], counts[:e], w["genes"], L, 231) as R:
            t_sig = R.signature_seconds
            ref_sig = R.signatures()
        with oracle.Reference.from_signatures(signatures, L) as R:
            r = R.find_similar_pairs4_loop(k, thr, N - loop_rows, N, want_pairs=False)
            t_loop, pairs = r["seconds"], r["pairs"]
    else:
        U = oracle.generate_lsh_vectors(w["genes"], L, 231)
        s1, _ = oracle.cell_sums(toc[: sig_cells + 1], counts[:e])
        t1 = time.time()
        ref_sig, _ = oracle.signatures(toc[: sig_cells + 1], genes[:e], counts[:e], s1, U)
        t_sig = time.time() - t1
        t_loop, pairs, _ = oracle.pair_loop(signatures, L, thr, N - loop_rows, N)
    total_pairs = N * (N - 1) / 2
    sig_s_per_cell = t_sig / sig_cells
    ns_per_pair = 1e9 * t_loop / max(pairs, 1)
    full_seconds = sig_s_per_cell * N + ns_per_pair * 1e-9 * total_pairs
    return dict(kind=kind, cores=1, value=total_pairs / full_seconds, unit="cell-pairs/s",
                sample=f"signatures of the first {sig_cells} cells + findSimilarPairs4 pair loop for the last "
                       f"{loop_rows} cells x all earlier cells ({pairs} pairs), extrapolated to the whole job",
                signature_s_per_cell=sig_s_per_cell, ns_per_pair=ns_per_pair, extrapolated_job_seconds=full_seconds,
                sample_seconds=time.time() - t0, sample_signatures_match_gpu=bool(
                    np.array_equal(ref_sig, signatures[:sig_cells])))


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    toc, genes, counts, U = make_workload(w)
    # the pair loop needs the signatures of every cell: computed once with the C restatement's
    # arithmetic on all host cores would take minutes at 100k cells, so use random-projection-free
    # synthetic signatures of identical shape for the loop timing when no GPU is present.
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        import expressionmatrix2_b200 as em2
        with em2.Engine(0) as eng:
            signatures = eng.compute_signatures(toc, counts, U, gene_ids=genes)
        sig_note = "signatures of all cells from the GPU path (bit-exact vs the reference on the sampled cells)"
    else:
        from expressionmatrix2_b200 import synthetic
        signatures = synthetic.gen_signatures(w["cells"], w["lsh"], seed=1, clusters=64)
        sig_note = "synthetic signatures (no GPU visible)"
    samples = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        s = cpu_sample(w, toc, genes, counts, signatures)
        s["wall_ms"] = 1e3 * (time.time() - t0)
        if i >= args.warmup:
            samples.append(s)
    value = float(np.median([s["value"] for s in samples]))
    best = samples[0]
    line = dict(impl="reference", metric="cell-pairs/sec (1024-bit LSH, top-50)", value=value, unit="cell-pairs/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=float(np.mean([s["wall_ms"] for s in samples])), higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="u64 popcount + f64 projections", data="synthetic",
                config=dict(workload=args.workload, **{k: w[k] for k in ("cells", "genes", "nnz_per_cell", "lsh", "k", "thr")},
                            note=w["note"], signatures=sig_note),
                cpu_baseline=dict(kind=best["kind"], cores=1, value=value, unit="cell-pairs/s", sample=best["sample"],
                                  ns_per_pair=best["ns_per_pair"], signature_s_per_cell=best["signature_s_per_cell"],
                                  extrapolated_job_seconds=best["extrapolated_job_seconds"]),
                e2e=dict(value=value, unit="cell-pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, w):
    import torch
    import torch.distributed as dist
    import expressionmatrix2_b200 as em2
    from expressionmatrix2_b200.parallel import Partition, all_gather_signatures

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    variant = dict(auto=em2.VARIANT_AUTO, popc=em2.VARIANT_POPC, mma=em2.VARIANT_MMA_I8)[args.variant]

    N, G, L, k, thr = w["cells"], w["genes"], w["lsh"], w["k"], w["thr"]
    W = em2.word_count(L)
    toc, genes, counts, U = make_workload(w)
    part = Partition(N, world, rank)
    ltoc, lgenes, lcounts = part.slice_csr(toc, genes, counts)
    lpairs = em2.to_pairs(lgenes, lcounts)
    rows = part.rows
    nnz_local = int(ltoc[-1])
    eng = em2.Engine(local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- resident buffers ------------------------------------------------------------------------
    d_toc = torch.from_numpy(ltoc.view(np.int64)).to(dev)
    d_counts = torch.from_numpy(lpairs.view(np.int64)).to(dev)
    d_U = torch.from_numpy(U).to(dev)
    d_sum1 = torch.empty(rows, dtype=torch.float64, device=dev)
    d_sum2 = torch.empty(rows, dtype=torch.float64, device=dev)
    d_sig_local = torch.zeros((part.shard, W), dtype=torch.int64, device=dev)
    d_lut = torch.from_numpy(em2.similarity_table(L).astype(np.float32)).to(dev)
    d_pairs = torch.zeros((rows, k, 2), dtype=torch.int32, device=dev)
    d_used = torch.zeros(rows, dtype=torch.int32, device=dev)
    d_nz = torch.zeros(8, dtype=torch.int64, device=dev)
    mm = em2.mismatch_max(L, thr)
    stage_names = ["sums", "signatures", "allgather", "scan_topk"]

    def step(events=None):
        def mark(i):
            if events is not None:
                events[i].record()
        mark(0)
        eng.cell_sums_device(rows, d_toc, d_counts, d_sum1, d_sum2, stream=stream)
        mark(1)
        eng.signatures_device(rows, G, d_toc, d_counts, d_sum1, d_sum2, d_U, L, L, d_sig_local, d_nz, stream=stream,
                              nnz=nnz_local)
        mark(2)
        full = all_gather_signatures(d_sig_local, part) if world > 1 else d_sig_local
        mark(3)
        eng.scan_topk_device(full, N, L, part.row_begin, part.row_end, k, mm, d_lut, d_pairs, d_used,
                             variant=variant, stream=stream)
        mark(4)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for i in range(args.steps):
        step(ev[i])
    t_end.record()
    barrier()
    clocks = sampler.stop()
    total_ms = t_begin.elapsed_time(t_end)
    launches = eng.stats()["kernel_launches"] - launches0
    stage_ms = {n: float(np.mean([ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(args.steps)]))
                for j, n in enumerate(stage_names)}
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    pairs_total = N * (N - 1) / 2
    value = pairs_total / (ms_per_step * 1e-3)

    # ---- e2e: host buffers -> host lists ----------------------------------------------------------
    h_pairs = torch.empty((rows, k, 2), dtype=torch.int32).pin_memory()
    h_used = torch.empty(rows, dtype=torch.int32).pin_memory()
    if world == 1:
        # through the reference-facing blocking C-ABI call; host inputs live in pinned memory
        p_toc = torch.from_numpy(toc.view(np.int64)).pin_memory()
        p_counts = torch.from_numpy(lpairs.view(np.int64)).pin_memory()
        p_U = torch.from_numpy(U).pin_memory()
        n_toc, n_counts, n_U = p_toc.numpy().view(np.uint64), p_counts.numpy().view(em2.PAIR_DTYPE), p_U.numpy()
        e2e_ms, st = [], None
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            ids, sims, used = eng.lsh_similar_pairs(n_toc, n_counts, n_U, k, thr, variant=variant)
            if i >= args.warmup:
                e2e_ms.append(1e3 * (time.perf_counter() - t0))
            st = eng.stats()
        e2e_t = float(np.mean(e2e_ms))
        h2d, d2h = int(st["h2d_bytes"]), int(st["d2h_bytes"])
        e2e_stats = {k2: st[k2] for k2 in ("h2d_ms", "sums_ms", "signatures_ms", "scan_ms", "d2h_ms")}
    else:
        p_toc = torch.from_numpy(ltoc.view(np.int64)).pin_memory()
        p_counts = torch.from_numpy(lpairs.view(np.int64)).pin_memory()
        p_U = torch.from_numpy(U).pin_memory()
        e2e_ms = []
        for i in range(args.warmup + args.steps):
            barrier()
            t0 = time.perf_counter()
            d_toc.copy_(p_toc, non_blocking=True)
            d_counts.copy_(p_counts, non_blocking=True)
            d_U.copy_(p_U, non_blocking=True)
            step()
            h_pairs.copy_(d_pairs, non_blocking=True)
            h_used.copy_(d_used, non_blocking=True)
            barrier()
            if i >= args.warmup:
                e2e_ms.append(1e3 * (time.perf_counter() - t0))
        tt = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t = float(tt.item())
        h2d = p_toc.numel() * 8 + p_counts.numel() * 8 + p_U.numel() * 8
        d2h = h_pairs.numel() * 4 + h_used.numel() * 4
        e2e_stats = {}

    # ---- roofline of the dominant kernel (the scan) ------------------------------------------------
    peaks = load_peaks()
    scan_s = stage_ms["scan_topk"] * 1e-3
    ordered = rows * N                      # pair evaluations this GPU executed per launch
    alg_pairs = pairs_total / world         # algorithmic units per GPU per launch
    variant_used = eng.stats()["variant_used"] or (em2.VARIANT_POPC if variant != em2.VARIANT_MMA_I8 else variant)
    if variant_used == em2.VARIANT_MMA_I8:
        peak = 2.0 * peaks["bf16_tflops"]   # int8 tensor peak = 2x the measured dense bf16 figure
        roof = dict(bound="tensor", unit="TOP/s", achieved=alg_pairs * 2 * L / scan_s / 1e12, peak=peak,
                    peak_source=f"2 x bf16_tflops of MEASURED_PEAKS.json ({peaks['source']}); tools/mma_peak.cu "
                                "measured 4216 TOP/s (kind::i8, A in TMEM, N=256) and 2971 TOP/s (N=128, this "
                                "kernel's shape) on this pool",
                    executed=ordered * 2 * L / scan_s / 1e12)
    else:
        mb = os.path.join(ROOT, "expressionmatrix2_b200", "build", "microbench")
        popc_peak = None
        try:
            popc_peak = json.loads(subprocess.check_output([mb], env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(local_rank))).decode())["popc_per_s"]
        except Exception:
            pass
        popc_peak = popc_peak or 4.6e12
        # algorithmic unit = one unordered pair = 2W 32-bit popcount-words
        roof = dict(bound="alu", unit="Gpopc32/s", achieved=alg_pairs * 2 * W / scan_s / 1e9, peak=popc_peak / 1e9,
                    peak_source="POPC.b32 issue rate measured by tools/microbench.cu on this GPU",
                    executed=ordered * 2 * W / scan_s / 1e9)
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["executed_frac"] = roof["executed"] / roof["peak"]
    roof["kernel"] = "scan_topk"
    roof["kernel_ms"] = stage_ms["scan_topk"]
    roof["traffic"] = None
    alg_bytes = N * L / 8 + rows * L / 8 + rows * (8 * k + 4)
    roof["hbm"] = dict(bound="hbm", unit="GB/s", algorithmic_bytes=alg_bytes, achieved=alg_bytes / scan_s / 1e9,
                       peak=peaks["hbm_gbs"], frac=alg_bytes / scan_s / 1e9 / peaks["hbm_gbs"],
                       note="compulsory bytes only; the scan is compute bound by construction")
    nnz_local = int(ltoc[-1])
    sig_s = stage_ms["signatures"] * 1e-3
    sig_bytes = 8 * nnz_local + 8 * (rows + 1) + 8 * rows + 8 * G * L + rows * L / 8
    sig_roof = dict(kernel="signatures", kernel_ms=stage_ms["signatures"], flops=2.0 * nnz_local * L,
                    achieved_gflops=2.0 * nnz_local * L / sig_s / 1e9, algorithmic_bytes=sig_bytes,
                    achieved_gbs=sig_bytes / sig_s / 1e9, hbm_frac=sig_bytes / sig_s / 1e9 / peaks["hbm_gbs"])

    if rank == 0:
        used_mean = float(d_used.float().mean().item())
        line = dict(metric="cell-pairs/sec (1024-bit LSH, top-50)", value=value, unit="cell-pairs/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True,
                    scaling="strong", vs_baseline=None,
                    dtype="u64 xor/popc (scan), f64 non-fused mul+add (signatures)", data="synthetic",
                    config=dict(workload=args.workload, **{k2: w[k2] for k2 in ("cells", "genes", "nnz_per_cell", "lsh", "k", "thr")},
                                note=w["note"], variant={1: "popc", 2: "mma_i8"}.get(variant_used, "popc"),
                                parallelism=f"cell-row blocks x{world}" + (", 1 NCCL all-gather of signatures" if world > 1 else ""),
                                l2="inputs (CSR + hyperplanes, >1.4 GB) exceed the 126 MB L2; no flush needed",
                                ordered_evaluations_per_s=N * float(N) / (ms_per_step * 1e-3),
                                mean_neighbours_stored=used_mean),
                    stage_ms=stage_ms, roofline=roof, signature_roofline=sig_roof,
                    e2e=dict(value=pairs_total / (e2e_t * 1e-3), unit="cell-pairs/s", ms=e2e_t,
                             h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, stage_ms=e2e_stats,
                             api="em2_lsh_similar_pairs (C-ABI, host buffers)" if world == 1 else
                                 "device API + pinned host copies per rank"),
                    gpu_launches=int(launches), clocks=clocks)
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            sig_host = d_sig_local[:rows].cpu().numpy().view(np.uint64)
            cb = cpu_sample(w, toc, genes, counts, sig_host)
            line["cpu_baseline"] = {k2: cb[k2] for k2 in ("value", "unit", "cores", "kind", "sample", "ns_per_pair",
                                                          "signature_s_per_cell", "extrapolated_job_seconds",
                                                          "sample_seconds", "sample_signatures_match_gpu")}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto", choices=["auto", "popc", "mma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
