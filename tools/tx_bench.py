#!/usr/bin/env python
"""Time the transmit pass alone (make_packets + tx kernel) on the bench layout; OFDM_TX_OLD=1 selects the 3-pass plan."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ofdm_uhd_b200.engine import OfdmEngine

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N, occ, cp, mod = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]) if len(sys.argv) > 5 else (512, 200, 128, "qpsk")
psize = int(sys.argv[6]) if len(sys.argv) > 6 else 402
eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=1, max_pkt_bytes=psize + 16)
rng = np.random.default_rng(1)
body = torch.from_numpy(rng.integers(0, 256, size=F * psize, dtype=np.uint8)).cuda()
plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
x = torch.zeros(plan.n_samples, dtype=torch.complex64, device="cuda")
for _ in range(2):
    eng.tx_run(plan, body, out=x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.tx_run(plan, body, out=x)
b.record(); torch.cuda.synchronize()
print("tx %d frames, %.1f M samples: %.3f ms (OFDM_TX_OLD=%s)" % (F, plan.n_samples / 1e6, a.elapsed_time(b) / 5, os.environ.get("OFDM_TX_OLD", "")))
