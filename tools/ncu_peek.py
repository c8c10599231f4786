#!/usr/bin/env python
"""Print the handful of ncu --set full metrics the kernel work keeps coming back to, one column per report."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
cols = []
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    for row in r[2:]:
        cols.append((rep.split("/")[-1] + ":" + row[r[0].index("Kernel Name")][:28], dict(zip(r[0], row)), dict(zip(r[0], r[1]))))
stall = sorted({k for _, d, _ in cols for k in d if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k})
for k in KEYS + stall:
    print("%-95s" % k[-95:], "  ".join("%14s" % d.get(k, "-")[:14] for _, d, _ in cols), cols[0][2].get(k, ""))
print(" ".join(c[0] for c in cols))
