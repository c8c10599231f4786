#!/usr/bin/env python
"""Accuracy of the channel-filter kernel against the oracle's float64 FIR (relative L2 and max error), per overlap-save size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import ofdm_oracle as o
from ofdm_uhd_b200.engine import OfdmEngine

rng = np.random.default_rng(5)
n = 200000
x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
lay = o.Layout(512, 200, 128, "qpsk")
ref64 = np.convolve(x.astype(np.complex128), o.chan_filter_taps(lay).astype(np.float64))[:n]
eng = OfdmEngine(512, 200, 128, "qpsk", 0.25)
xt = torch.from_numpy(x).cuda()
y = torch.empty_like(xt)
eng.L_.ofdm_rx_chan_filter(eng.h, eng._p(xt), n, eng._p(y), eng._stream())
torch.cuda.synchronize()
g = y.cpu().numpy().astype(np.complex128)
print("NOS=%d rel L2 %.3e  max abs %.3e (|y| rms %.3f)" % (eng.nos, np.linalg.norm(g - ref64) / np.linalg.norm(ref64), np.abs(g - ref64).max(), np.sqrt(np.mean(np.abs(ref64) ** 2))))
