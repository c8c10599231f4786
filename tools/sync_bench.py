#!/usr/bin/env python
"""Time ofdm_rx_sync alone on the bench capture shape (noise + frames are irrelevant for timing: any finite stream)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ofdm_uhd_b200.engine import OfdmEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 640_000_000
eng = OfdmEngine(512, 200, 128, "qpsk", 0.25)
y = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
bufs = eng.rx_alloc(n, max_frames=200000)
st = eng._stream()
for _ in range(2):
    eng.L_.ofdm_rx_sync(eng.h, eng._p(y), n, C.byref(bufs["io"]), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.L_.ofdm_rx_sync(eng.h, eng._p(y), n, C.byref(bufs["io"]), st)
b.record(); torch.cuda.synchronize()
print("OFDM_SYNC_SPLIT=%s OFDM_MC_MB=%s  %.3f ms" % (os.environ.get("OFDM_SYNC_SPLIT", "-"), os.environ.get("OFDM_MC_MB", "-"), a.elapsed_time(b) / 5))
