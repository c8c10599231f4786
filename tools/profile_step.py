#!/usr/bin/env python
"""One warm-up step and one measured step of the bench workload (for ncu): 512/200/128, QPSK by default."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20000)
    ap.add_argument("--mod", default="qpsk")
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    F, psize = a.frames, 402
    eng = OfdmEngine(512, 200, 128, a.mod, 0.25, pad_seed=1, max_pkt_bytes=416)
    rng = np.random.default_rng(1)
    body = torch.from_numpy(rng.integers(0, 256, size=F * psize, dtype=np.uint8)).cuda()
    plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
    lead = 2 * eng.L
    n = plan.n_samples + 2 * lead
    x = torch.zeros(n, dtype=torch.complex64, device="cuda")
    xs = x[lead:lead + plan.n_samples]
    eng.tx_run(plan, body, out=xs)
    p = float((xs[:1 << 20].abs() ** 2).mean())
    xc = eng.channel(x, cfo=0.27, sigma=(p / 100 / 2) ** 0.5, seed=3)
    bufs = eng.rx_alloc(n, max_frames=F + 1024)
    for _ in range(1 + a.steps):
        eng.tx_run(plan, body, out=xs)
        eng.demodulate_async(xc, bufs)
    res = eng.collect(bufs, want_packets=False)
    print("frames", res.n_frames, "counters", res.counters.tolist())


if __name__ == "__main__":
    main()
