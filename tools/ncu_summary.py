"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [regex-of-extra-metrics]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = [
    r"^Kernel Name$", r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$",
    r"^launch__registers_per_thread$", r"^launch__waves_per_multiprocessor$", r"^launch__occupancy_limit_(registers|shared_mem|warps)$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_(alu|fma|fmaheavy|xu|lsu|fp64|uniform|tc|tma)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_tensor.*cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_(alu|fma|fp64|xu|shared)_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_bytes\.sum$", r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__t_sector_hit_rate\.pct$", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
    r"^smsp__thread_inst_executed_per_inst_executed\.ratio$",
]
if extra:
    WANT.append(extra)
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if any(re.search(w, h) for w in WANT):
            if "issue_stalled" in h:
                try:
                    if float(v) < 0.05:
                        continue
                except ValueError:
                    pass
            print(f"{h:95s} {u:12s} {v}")
