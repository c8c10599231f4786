#!/usr/bin/env python
"""Where the end-to-end step (host payloads in, host verdicts + payload bytes out) spends its time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ofdm_uhd_b200.engine import OfdmEngine

F, psize = 100000, 402
eng = OfdmEngine(512, 200, 128, "qpsk", 0.25, pad_seed=1, max_pkt_bytes=416)
rng = np.random.default_rng(1)
h_pay = torch.from_numpy(rng.integers(0, 256, size=F * psize, dtype=np.uint8)).pin_memory()
plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
lead = 2 * eng.L
n = plan.n_samples + 2 * lead
x = torch.zeros(n, dtype=torch.complex64, device="cuda")
xs = x[lead:lead + plan.n_samples]
xc = torch.empty_like(x)
bufs = eng.rx_alloc(n, max_frames=F + 1024)
sync = torch.cuda.synchronize
def T(f, name, reps=3):
    f(); sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    sync()
    print("%-28s %8.2f ms" % (name, (time.perf_counter() - t0) / reps * 1e3))
    return r
dp = T(lambda: h_pay.to("cuda", non_blocking=True), "H2D payloads (40 MB)")
T(lambda: eng.tx_run(plan, dp, out=xs), "tx_run")
T(lambda: eng.channel(x, cfo=0.2, sigma=0.01, seed=3, out=xc), "channel kernel")
T(lambda: eng.demodulate_async(xc, bufs), "demodulate_async")
T(lambda: eng.collect(bufs, want_packets=False, want_payload=True), "collect (payload rows)")
T(lambda: eng.collect(bufs, want_packets=False, want_payload=False), "collect (verdicts only)")
def full():
    d = h_pay.to("cuda", non_blocking=True)
    eng.tx_run(plan, d, out=xs)
    eng.channel(x, cfo=0.2, sigma=0.01, seed=3, out=xc)
    eng.demodulate_async(xc, bufs)
    return eng.collect(bufs, want_packets=False, want_payload=True)
T(full, "full e2e step")
