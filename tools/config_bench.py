#!/usr/bin/env python
"""Throughput of mod + demod on the other BASELINE layouts (not bench lines; a health check of the N >= 1024 paths).

usage: config_bench.py N occ cp mod frames [payload_bytes [snr_db [cfo]]]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ofdm_uhd_b200.engine import OfdmEngine

N, occ, cp, mod, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
psize = int(sys.argv[6]) if len(sys.argv) > 6 else 402
snr_db = float(sys.argv[7]) if len(sys.argv) > 7 else 30.0
eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=1, max_pkt_bytes=psize + 16)
rng = np.random.default_rng(1)
body = torch.from_numpy(rng.integers(0, 256, size=F * psize, dtype=np.uint8)).cuda()
plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
lead = 2 * eng.L
n = plan.n_samples + 2 * lead
x = torch.zeros(n, dtype=torch.complex64, device="cuda")
xs = x[lead:lead + plan.n_samples]
eng.tx_run(plan, body, out=xs)
p = float((xs[:1 << 20].abs() ** 2).mean())
cfo = float(sys.argv[8]) if len(sys.argv) > 8 else 0.27
xc = eng.channel(x, cfo=cfo, sigma=(p / 10 ** (snr_db / 10) / 2) ** 0.5, seed=3)
bufs = eng.rx_alloc(n, max_frames=F + 1024)
st = eng._stream()
io = bufs["io"]
y = eng.ws_view(bufs, 0, n)
stages = {
    "tx": lambda: eng.tx_run(plan, body, out=xs),
    "filter": lambda: eng.L_.ofdm_rx_chan_filter(eng.h, eng._p(xc), n, eng._p(y), st),
    "metric": lambda: eng.L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 0, st),
    "detect": lambda: eng.L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 1, st),
    "gather": lambda: eng.L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 2, st),
    "plan": lambda: eng.L_.ofdm_rx_plan(eng.h, n, C.byref(io), st),
    "acq": lambda: eng.L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 3, st),
    "sink": lambda: eng.L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 4, st),
    "finish": lambda: eng.L_.ofdm_rx_finish(eng.h, C.byref(io), st),
}
eng.demodulate_async(xc, bufs)
res = eng.collect(bufs, want_packets=False)
out = {}
for name, fn in stages.items():
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record(); torch.cuda.synchronize()
    out[name] = a.elapsed_time(b) / 5
tot = sum(out.values())
print("%d/%d/%d %s: %d frames, %.1f M samples, crc ok %d/%d; ms %s; total %.2f ms = %.1f Gsamples/s" % (
    N, occ, cp, mod, F, plan.n_samples / 1e6, int(res.counters[2]), F, {k: round(v, 3) for k, v in out.items()}, tot,
    plan.n_samples / tot / 1e6))
