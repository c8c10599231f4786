"""Shared helpers of the test-suite (CPU side)."""
import struct

import numpy as np

from oracle import ofdm_oracle as o


def payloads(rng, k, size=398):
    return [struct.pack("!HH", i & 0xFFFF, 0) + bytes(rng.integers(0, 256, size, dtype=np.uint8)) for i in range(k)]


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def loopback_capture(lay, pay, snr, cfo, seed, amp=0.25, pad_seed=77, lead=None, tail=None):
    """oracle TX -> lead-in / tail of noise -> AWGN + CFO.  Returns (tx samples, noisy capture)."""
    pk = [o.make_packet(p, 1, 1, False) for p in pay]
    x = o.tx_modulate(pk, lay, amp, seed=pad_seed)
    lead = lay.fft_length + 37 if lead is None else lead
    tail = 4 * lay.sym_len if tail is None else tail
    xin = np.concatenate([np.zeros(lead, np.complex64), x, np.zeros(tail, np.complex64)])
    xc = o.channel(xin, snr, cfo, lay.fft_length, seed=seed, sig_power=float(np.mean(np.abs(x) ** 2)))
    return x, xc


def plan_closed_form(trig, n, N, L, timeout=1000):
    """Python mirror of plan_kernel (ofdm_uhd_b200/csrc/rx_front.cu): the sampler in closed form per trigger.
    Returns (first_ok, frame_start[], n_data[])."""
    trig = [int(t) for t in trig]
    K = len(trig)
    timeout = timeout + 1          # `if (d_timeout-- == 0)`: the FRAME state emits timeout + 1 data vectors
    first_ok = K
    for k in range(K):
        if trig[k] >= N:
            first_ok = k
            break
    F = K - first_ok
    for k in range(first_ok, K):
        t = trig[k]
        if k == first_ok:
            c = (t - N) // (L + 1)
            ok = c * (L + 1) + L + N < n
        else:
            tp = trig[k - 1]
            mm = max(1, -(-(t - tp - 1) // L))
            if mm <= timeout:
                ok = tp + 1 + mm * L < n
            else:
                q0 = tp + 1 + timeout * L
                c = (t - q0) // (L + 1)
                ok = q0 + c * (L + 1) + L < n
        if not ok:
            F = k - first_ok
            break
    starts, ndata = [], []
    for f in range(max(F, 0)):
        k = first_ok + f
        t = trig[k]
        J = timeout
        if k + 1 < K:
            mm = max(1, -(-(trig[k + 1] - t - 1) // L))
            J = min(J, mm - 1)
        room = (n - 2 - t) // L if n - 2 - t >= 0 else 0
        J = max(0, min(J, room))
        starts.append(t - N + 1)
        ndata.append(J)
    return first_ok, np.array(starts, dtype=np.int64), np.array(ndata, dtype=np.int64)
