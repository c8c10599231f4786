#!/usr/bin/env python
"""BASELINE configs[3]: 1024-pt Blackman-Harris energy detection over a 100 MS/s synthetic wideband capture
(100 M samples = 1 s, 97 656 frames, dwell 12 frames): throughput of ofdm_sense and the HBM roofline fraction."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ofdm_uhd_b200.engine import SenseEngine

N, n = 1024, 100_000_000
nfr = n // N
se = SenseEngine(N)
g = torch.Generator(device="cuda").manual_seed(4)
x = torch.view_as_complex(torch.randn((nfr * N, 2), device="cuda", generator=g) * (5e-6 / 1024 / 2) ** 0.5)
out = torch.empty((nfr // 12, N), dtype=torch.float32, device="cuda")
for _ in range(3):
    se.maxhold(x, 0, 12, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
a.record()
for _ in range(reps):
    se.maxhold(x, 0, 12, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
peak = 6551.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
gbs = 8.0 * nfr * N / (ms * 1e-3) / 1e9
avg, free, hx = se.decide(out[:10], 1e-3)
print(json.dumps({"workload": "sensing 1024-pt BH, 100 M samples, dwell 12", "ms": ms, "Msamples_s": nfr * N / ms / 1e3,
                  "achieved_GBs": gbs, "peak_GBs": peak, "frac": gbs / peak, "hex_len": len(hx)}))
