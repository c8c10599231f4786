#!/usr/bin/env python
"""Where a kernel's warp-stall samples sit, by SASS opcode: the source page (SASS view) of an `ncu --set full
--import-source on` report folded per opcode -- executed warp instructions, share of the stall samples, and the two
stall reasons that dominate each opcode.  Usage: tools/ncu_hotspots.py report.ncu-rep [kernel-regex ...] > summary.txt
(reads the report here; nothing runs on a GPU)."""
import collections
import csv
import io
import re
import subprocess
import sys

DEFAULT = ["tx_warp_kernel", "chan_filter_warp_kernel", "metric_chunk_kernel", "detect_seg_kernel", "acq_warp_kernel",
           "sink_kernel"]


def opcode(sass):
    """'@!P0 FFMA2.RN R4, ...' -> 'FFMA2' ; 'F2F.F64.F32 R2, R5' -> 'F2F.F64.F32' (conversions keep their types)."""
    t = sass.strip().split()
    if t and t[0].startswith("@"):
        t = t[1:]
    if not t:
        return "?"
    op = t[0].rstrip(";")
    if op.startswith(("F2F", "I2F", "F2I", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL")):
        return ".".join(op.split(".")[:3]) if op.startswith(("F2F", "I2F", "F2I")) else ".".join(op.split(".")[:2])
    return op.split(".")[0]


def fold(report, kern):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hdr_i is None:
        return None
    name = rows[hdr_i - 1][1] if hdr_i and len(rows[hdr_i - 1]) > 1 else kern
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ops = collections.defaultdict(lambda: {"inst": 0, "samples": 0, "stalls": collections.Counter()})
    tot_inst = tot_samp = 0
    reasons = collections.Counter()
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break                                   # a second launch of the same kernel: one is enough
            continue
        o = ops[opcode(r[col["Source"]])]
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["Warp Stall Sampling (All Samples)"]] or 0))
        o["inst"] += inst
        o["samples"] += samp
        tot_inst += inst
        tot_samp += samp
        for s in stall_cols:
            v = int(float(r[col[s]] or 0))
            if v:
                o["stalls"][s[6:]] += v
                reasons[s[6:]] += v
    return name, ops, tot_inst, tot_samp, reasons


def main():
    report = sys.argv[1]
    kerns = sys.argv[2:] or DEFAULT
    print("# %s -- stall samples by SASS opcode (tools/ncu_hotspots.py; `ncu --set full --import-source on`, one launch)" %
          report.split("/")[-1])
    for k in kerns:
        res = fold(report, k)
        if res is None:
            print("\n## %s: not in the report" % k)
            continue
        name, ops, ti, ts, reasons = res
        print("\n## %s" % re.sub(r"\s+", " ", name))
        print("warp instructions %d, stall samples %d; by reason: %s" %
              (ti, ts, ", ".join("%s %.0f%%" % (r, 100.0 * v / max(1, sum(reasons.values()))) for r, v in reasons.most_common(6))))
        print("%-14s %14s %7s %9s   %s" % ("opcode", "warp inst", "inst %", "samples %", "top stall reasons of the opcode"))
        for op, o in sorted(ops.items(), key=lambda kv: -kv[1]["samples"])[:18]:
            top = ", ".join("%s %.0f%%" % (r, 100.0 * v / max(1, sum(o["stalls"].values()))) for r, v in o["stalls"].most_common(2))
            print("%-14s %14d %6.1f%% %8.1f%%   %s" % (op, o["inst"], 100.0 * o["inst"] / max(1, ti),
                                                      100.0 * o["samples"] / max(1, ts), top))


if __name__ == "__main__":
    main()
