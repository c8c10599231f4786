"""CPU baseline driver: times the oracle's loopback (mod + demod) on this machine's host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and `--impl reference`).  Uses the C port
(oracle/ofdm_oracle_c.c -> oracle/libofdm_oracle.so, all host threads) when it is built, else the NumPy
oracle on one core.  The reference's own implementation (GNU Radio 3.6 C++ blocks under Python 2) cannot run
here, so the kind is always "port".
"""
from __future__ import annotations

import os
import struct
import time

import numpy as np

from . import ofdm_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))


def _payloads(n, size, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    return [struct.pack("!HH", i & 0xFFFF, 0) + bytes(rng.integers(0, 256, size - 4, dtype=np.uint8)) for i in range(n)]


def time_loopback_numpy(mod: str, frames: int, snr: float):
    lay = o.Layout(512, 200, 128, mod)
    pay = _payloads(frames, 402, 1)
    t0 = time.perf_counter()
    pkts = [o.make_packet(p, 1, 1, False) for p in pay]
    x = o.tx_modulate(pkts, lay, 0.25, seed=3)
    t_mod = time.perf_counter() - t0
    lead = np.zeros(2 * lay.sym_len, dtype=np.complex64)
    xc = o.channel(np.concatenate([lead, x, lead]), snr, 0.2, 512, seed=5, sig_power=float(np.mean(np.abs(x) ** 2)))
    t0 = time.perf_counter()
    r = o.rx_demodulate(xc, lay)
    t_dem = time.perf_counter() - t0
    ok = sum(1 for g, _ in r.packets if g)
    return len(x), t_mod + t_dem, ok


def time_loopback(mod: str = "qpsk", frames: int = 0, snr: float = 20.0, threads=None, cfos=None):
    try:
        from . import c_port
        if c_port.available():
            return c_port.time_loopback(mod, frames, snr, threads, cfos)
    except ImportError:
        pass
    frames = frames or 150
    n, secs, ok = time_loopback_numpy(mod, frames, snr)
    return {"value": n / secs / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port", "ms": secs * 1e3,
            "sample": "%d frames (%d samples) of the bench workload through oracle/ofdm_oracle.py (NumPy, 1 thread); "
                      "%d/%d packets ok" % (frames, n, ok, frames)}
