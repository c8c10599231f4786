#!/usr/bin/env python
"""Time the channel-filter stage alone on a 640 M-sample capture (kernel variants via OFDM_FILTER_MODE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ofdm_uhd_b200.engine import OfdmEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 640_000_000
eng = OfdmEngine(512, 200, 128, "qpsk", 0.25)
x = torch.randn(n, 2, device="cuda").view(torch.float32).view(-1)
x = torch.view_as_complex(x.view(n, 2))
y = torch.empty_like(x)
st = eng._stream()
for _ in range(2):
    eng.L_.ofdm_rx_chan_filter(eng.h, eng._p(x), n, eng._p(y), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.L_.ofdm_rx_chan_filter(eng.h, eng._p(x), n, eng._p(y), st)
b.record(); torch.cuda.synchronize()
print("OFDM_FILTER_MODE=%s  %.3f ms" % (os.environ.get("OFDM_FILTER_MODE", "0"), a.elapsed_time(b) / 5))
