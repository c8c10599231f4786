#!/usr/bin/env python
"""Where the time of ofdm_demod.feed_stream goes: cProfile of a chunked feed of one capture (bench.py's feed_stream leg).

usage: stream_profile.py [frames [buffer_samples [batch_samples]]]"""
import cProfile, os, pstats, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ofdm_uhd_b200 import receive_path as rp_mod, transmit_path as tp_mod
from ofdm_uhd_b200 import channel_model

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
buf = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 25
opt = types.SimpleNamespace(modulation="qpsk", fft_length=512, occupied_tones=200, cp_length=128, snr=20.0, verbose=False,
                            log=False, tx_amplitude=0.25, samples_per_symbol=1)
tx = tp_mod.transmit_path(opt, device=0)
eng = tx.ofdm_tx._engine
rng = np.random.default_rng(5)
body = rng.integers(0, 256, size=(frames, 402), dtype=np.uint8)
plan = eng.tx_plan(np.arange(frames + 1, dtype=np.int64) * 402)
xs = eng.tx_run(plan, torch.from_numpy(body.reshape(-1)).to(eng.dev))
cap = channel_model.channel_model(eng, noise_voltage=0.0108, frequency_offset=0.2, seed=4, lead_in=1280, tail=2560).process(xs)
got = [0]
rx = rp_mod.receive_path(lambda ok, p: None, opt, device=0, max_pkt_bytes=416)
rx.set_batch_callback(lambda oks, blob, off: got.__setitem__(0, got[0] + int(np.count_nonzero(oks))))
for timed in (False, True):
    got[0] = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rx.feed(cap)
    torch.cuda.synchronize(); dt_whole = time.perf_counter() - t0
print("whole-stream feed: %.2f ms, %.1f Gsamples/s, ok %d" % (dt_whole * 1e3, cap.numel() / dt_whole / 1e9, got[0]))
rx3 = rp_mod.receive_path(lambda ok, p: None, opt, device=0, max_pkt_bytes=416)
rx3.ofdm_rx.stream_batch_samples = batch
rx3.set_batch_callback(lambda oks, blob, off: got.__setitem__(0, got[0] + int(np.count_nonzero(oks))))


chunks = [cap[a:a + buf] for a in range(0, cap.numel(), buf)]      # the source's buffers, ready made


def run():
    for c in chunks:
        rx3.feed_stream(c)
    rx3.flush_stream()
    torch.cuda.synchronize()


for timed in (False, True, "profile"):
    got[0] = 0
    rx3.ofdm_rx.reset_stream()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if timed == "profile":
        pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
    else:
        run()
    dt = time.perf_counter() - t0
    print("%s-sample buffers, batch %d: %.2f ms, %.1f Gsamples/s (%.2f of whole-stream), ok %d" %
          (buf, batch, dt * 1e3, cap.numel() / dt / 1e9, dt_whole / dt, got[0]))
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
