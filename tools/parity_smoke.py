#!/usr/bin/env python
"""A few small mod + demod cases (narrow / wide / 1024 / 4096-point layouts) against the oracle: transmit samples
(max abs error) and the receiver's packet list.  Seconds on a GPU box; the quick look after a kernel change."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from oracle import ofdm_oracle as o
from ofdm_uhd_b200.engine import OfdmEngine
from helpers import payloads
for (N, occ, cp, mod) in [(512, 200, 128, "qpsk"), (1024, 400, 256, "qam16"), (512, 400, 64, "qpsk"), (4096, 3200, 512, "qam256")]:
    lay = o.Layout(N, occ, cp, mod)
    rng = np.random.default_rng(3)
    pay = payloads(rng, 6)
    eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=3)
    body = np.frombuffer(b"".join(pay), dtype=np.uint8).copy()
    off = np.concatenate([[0], np.cumsum([len(p) for p in pay])]).astype(np.int64)
    plan = eng.tx_plan(off)
    x = eng.tx_run(plan, torch.from_numpy(body).cuda())
    torch.cuda.synchronize()
    err = float(np.max(np.abs(x.cpu().numpy() - o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, 0.25, seed=3))))
    xo = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, 0.25, seed=3)
    L = N + cp
    cap = o.channel(np.concatenate([np.zeros(700, np.complex64), xo, np.zeros(3 * L, np.complex64)]), 30.0, 0.2, N, seed=4,
                    sig_power=float(np.mean(np.abs(xo) ** 2)))
    got = eng.demodulate(torch.from_numpy(cap).cuda())
    ref = o.rx_demodulate(cap, lay)
    print("oracle ok", sum(1 for g, _ in ref.packets if g), "equal", got.packets == ref.packets)
    print(N, occ, mod, "ok", sum(1 for g, _ in got.packets if g), "of", len(pay), "tx max err", err)
