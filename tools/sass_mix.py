#!/usr/bin/env python
"""Instruction mix of the hot kernels in libofdm_b200.so (cuobjdump -sass): static counts per kernel of the packed
fp32 instructions (FFMA2 / FADD2 / FMUL2), their scalar counterparts, FP64, conversions, shared/global memory and
TMA / mbarrier instructions.  Usage: python tools/sass_mix.py [lib] > profiles/rNN_sass_mix.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ofdm_uhd_b200", "libofdm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
CLASSES = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL", "F2F", "MUFU", "LDS", "STS", "LDG",
           "STG", "LDGSTS", "SHFL", "BAR", "UTMALDG", "UBLKCP", "SYNCS", "IMAD", "IADD3", "LOP3", "MOV", "PRMT", "SEL",
           "FSETP", "ISETP", "BRA"]
kern = None
mix = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        mix[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        mix[kern]["total"] += 1
        for c in CLASSES:
            if op == c or op.startswith(c + "."):
                mix[kern][c] += 1
                break
print("static SASS instruction mix per kernel (%s)" % os.path.basename(lib))
print("%-58s %6s %6s %6s %6s | %6s %6s %6s | %5s %5s %5s | %5s %5s %5s %5s %6s" % (
    "kernel", "total", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "F2F", "LDS", "STS", "LDG", "STG", "LDGSTS"))
for k, c in mix.items():
    if c["total"] < 200:
        continue
    print("%-58s %6d %6d %6d %6d | %6d %6d %6d | %5d %5d %5d | %5d %5d %5d %5d %6d" % (
        k[:58], c["total"], c["FFMA2"], c["FADD2"], c["FMUL2"], c["FFMA"], c["FADD"], c["FMUL"], c["DFMA"], c["DADD"], c["F2F"],
        c["LDS"], c["STS"], c["LDG"], c["STG"], c["LDGSTS"]))
tot = collections.Counter()
for c in mix.values():
    tot.update(c)
print("library totals: " + ", ".join("%s %d" % (k, tot[k]) for k in ["FFMA2", "FADD2", "FMUL2", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS"]))
