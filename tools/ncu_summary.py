#!/usr/bin/env python
"""Turn `ncu --set full` reports (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into the tracked artefacts under
profiles/: a CSV of the columns DESIGN.md quotes (one row per kernel launch) and the traffic JSON bench.py reads for
`roofline.traffic` / `issue_roofline`.

    python tools/ncu_summary.py profiles/r02_ncu_full_main_kernels_640Msamples.csv profiles/r02_traffic.json rep1.ncu-rep [rep2 ...]
"""
import csv
import io
import json
import subprocess
import sys

COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def main():
    csv_out, json_out, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    table, kernels = [], {}
    units_row = None
    for rep in reps:
        hdr, units, rows = rows_of(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        cols = [c for c in COLS if c in ix]
        units_row = ["", *[units[ix[c]] for c in cols]]
        for row in rows:
            name = row[ix["Kernel Name"]]
            table.append((cols, [name] + [row[ix[c]] for c in cols]))
            short = name.replace("void ", "").split("<")[0].split("(")[0]
            rd = float(row[ix["dram__bytes_read.sum"]]) * UNIT_SCALE.get(units[ix["dram__bytes_read.sum"]], 1.0)
            wr = float(row[ix["dram__bytes_write.sum"]]) * UNIT_SCALE.get(units[ix["dram__bytes_write.sum"]], 1.0)
            kernels[short] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                              "duration_ms_under_ncu": float(row[ix["gpu__time_duration.sum"]]),
                              "inst_executed": float(row[ix["smsp__inst_executed.sum"]]),
                              "issue_active_pct": float(row[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                              "lsu_data_pipe_pct": float(row[ix["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]]),
                              "fma_pipe_cycles_pct": float(row[ix["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active" in ix else None,
                              "source": rep.split("/")[-1]}
    with open(csv_out, "w", newline="") as f:
        w = csv.writer(f)
        cols = table[0][0]
        w.writerow(["Kernel Name"] + cols)
        w.writerow(units_row)
        short = lambda name: name.replace("void ", "").split("<")[0].split("(")[0]
        last = {short(r[0]): i for i, (c, r) in enumerate(table)}   # a kernel re-captured by a later report: its newest row
        for i, (c, r) in enumerate(table):
            if last[short(r[0])] == i:
                w.writerow(r)
    with open(json_out, "w") as f:
        json.dump({"workload": "tools/profile_step.py --frames 100000 (bench workload: 640 M samples, QPSK 512/200/128)",
                   "source": "ncu --set full --clock-control none, one launch per kernel; " + ", ".join(r.split("/")[-1] for r in reps),
                   "kernels": kernels}, f, indent=1)
    tot = sum(k["traffic_bytes"] for k in kernels.values())
    print("kernels:", ", ".join(kernels), "| DRAM traffic per step %.2f GB" % (tot / 1e9))


if __name__ == "__main__":
    main()
