#!/usr/bin/env python
"""The secondary-user loop of secondary_tx.py / secondary_rx.py without radios (py3, B200 modules):

  sense the band (sensor + sense_loop: windowed FFT -> max-hold -> 10-dwell mean -> threshold -> carrier-map hex,
  secondary_tx.py:146-266)  ->  hop decision (busy bins around the operating frequency, quietest 17-bin band, :268-300)
  ->  announce the new frequency on the 920 MHz rendezvous channel (synchronization, :54-73)  ->  the receiver's
  rx_callback state machine follows (secondary_rx.py:51-85)  ->  the file is sent on the new frequency
  (run_transmiter, :345-381) and written by the receiver.

A synthetic wideband capture stands for the USRP source; "retuning" selects which of the two simulated channels the
receiver listens to.  Every numeric stage (sensing, decision, modem) runs in the CUDA kernels.
"""
import io
import math
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "ofdm_uhd_b200"))

import transmit_path                                                          # noqa: E402
import receive_path                                                           # noqa: E402
import channel_model                                                          # noqa: E402
import sensing                                                                # noqa: E402
import rendezvous as rv                                                       # noqa: E402


def wideband_capture(fft_size, n_frames, centre, samp_rate, primary_hz, primary_bw_hz, rng, noise=1.5e-4):
    """Complex white noise plus a primary user: a band of +20 dB around primary_hz."""
    X = (rng.standard_normal((n_frames, fft_size)) + 1j * rng.standard_normal((n_frames, fft_size))) * noise * math.sqrt(fft_size / 2.0)
    res = samp_rate / fft_size
    lo = int(round((primary_hz - primary_bw_hz / 2 - (centre - samp_rate / 2)) / res))
    hi = int(round((primary_hz + primary_bw_hz / 2 - (centre - samp_rate / 2)) / res))
    X[:, max(lo, 0):min(hi, fft_size)] *= 10.0
    return np.fft.ifft(np.fft.ifftshift(X, axes=1), axis=1).reshape(-1).astype(np.complex64)


def main(seed=1, source=None, verbose=True):
    import torch
    rng = np.random.default_rng(seed)
    opts = SimpleNamespace(modulation="qpsk", fft_length=512, occupied_tones=200, cp_length=128, snr=30, verbose=False,
                           log=False, tx_amplitude=0.25, samples_per_symbol=2,
                           fft_size=2048, decim=4, tune_delay=0.0, dwell_delay=1e-3, sense_bins=128)
    Frequency = 905 * 10 ** 6                                                 # secondary_tx.py:417
    centre = 905e6
    # ---- sensing: a primary user sits on the operating frequency ------------------------------------------
    tb1 = sensing.sensor(opts)
    dw = tb1.dwell_delay
    cap = wideband_capture(opts.fft_size, 10 * dw, centre, tb1.samp_rate, Frequency, 1.0e6, rng)
    mh = tb1.dwell_vectors(torch.from_numpy(cap).cuda())
    avg_d, free_d, hx_d = tb1.engine.decide_device(mh[:10], 0.001)
    hexa_thr = bytes(hx_d.cpu().numpy()).decode("ascii")
    busy, new_freq = sensing.hop_decision(tb1, avg_d, free_d, Frequency, centre)
    if verbose:
        print("carrier map %s...%s, busy bins around %d Hz: %d -> %s" % (hexa_thr[:8], hexa_thr[-8:], Frequency, busy,
                                                                        "hop to %d Hz" % new_freq if new_freq else "stay"))
    assert new_freq is not None, "the primary user must be detected"
    # ---- the two simulated channels: 920 MHz rendezvous and the data channel ---------------------------------
    tuned = []
    sink = io.BytesIO()
    state = rv.secondary_receiver(set_center_freq=tuned.append, sink=sink)
    rx = receive_path.receive_path(state.rx_callback, opts)

    def air(freq_of_tx):
        """A transmit path whose samples reach the receiver only while it listens on freq_of_tx."""
        tx = transmit_path.transmit_path(opts, pad_seed=seed)
        chan = channel_model.channel_model(tx.ofdm_tx._engine, noise_voltage=0.004, frequency_offset=0.15, seed=seed,
                                           lead_in=1280, tail=2560)
        tx.connect(chan)
        chan.connect(lambda smp: rx.feed(smp) if state.freq == freq_of_tx else None)
        return tx

    # ---- rendezvous: transmitter_control puts the transmitter on 920 MHz while sync == 1 ----------------------
    tx_sync = air(rv.next_tx_frequency(1, new_freq))
    n_sync = rv.synchronization(tx_sync.send_pkt, new_freq, hexa_thr[:4] if False else "FE7F")
    tx_sync.send_pkt(eof=True)
    rx.wait(timeout=60)
    assert state.sync == 0 and state.freq == new_freq, "the receiver must have followed the announcement"
    # ---- data on the new frequency -------------------------------------------------------------------------
    data = source if source is not None else bytes(rng.integers(0, 256, 8000, dtype=np.uint8))
    tx_data = air(rv.next_tx_frequency(0, new_freq))
    n_bytes = rv.run_transmitter(tx_data.send_pkt, data, 204)
    tx_data.send_pkt(eof=True)
    rx.wait(timeout=60)
    got = sink.getvalue()
    if verbose:
        print("%d sync packets, receiver tuned %s, %d payload bytes sent, %d file bytes written, n_right %d / n_rcvd %d"
              % (n_sync, tuned, n_bytes, len(got), state.n_right, state.n_rcvd))
    return new_freq, tuned, data, got, state


if __name__ == "__main__":
    main()
