"""Self-consistency of the CPU oracle: mapper state machine vs closed form, peak detector state machine vs
vectorised form, sampler call-by-call vs closed form, loopback recovery for every modulation/layout, the
reference quirks of SURVEY.md Appendix C."""
import struct

import numpy as np
import pytest

from oracle import ofdm_oracle as o
from helpers import payloads, loopback_capture, plan_closed_form


@pytest.mark.parametrize("mod", ["bpsk", "qpsk", "8psk", "qam16", "qam64", "qam256"])
def test_mapper_closed_form_equals_state_machine(mod):
    lay = o.Layout(512, 200, 128, mod)
    rng = np.random.default_rng(1)
    for n in [0, 1, 2, 5, 33, 223, 411, 74, 75, 297]:
        pkt = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        a = o.mapper_sequential(pkt, lay, 3, 7)
        b = o.mapper_indices(pkt, lay, 3, 7)
        assert a.shape == b.shape and np.array_equal(a, b), (mod, n)


def test_mapper_extra_all_pad_symbol():
    # 8PSK, 223 bytes: 594 groups fill 3 symbols exactly and 2 stray bits remain -> a 4th, all-pad symbol (A.4)
    lay = o.Layout(512, 200, 128, "8psk")
    assert lay.n_data_syms(223) == 4
    assert o.mapper_sequential(bytes(223), lay, 0, 0).shape[0] == 4


def test_packet_roundtrip_and_limits():
    rng = np.random.default_rng(2)
    for n in (0, 1, 15, 402, 4091):
        p = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        for pad in (False, True):
            if pad and n == 4091:
                continue                                      # body would outgrow the 4096-byte mask (C.7)
            pkt = o.make_packet(p, 1, 1, pad)
            assert len(pkt) == (((n + 9) + 15) // 16 * 16 if pad else n + 9)
            L = struct.unpack("!H", pkt[:2])[0] & 0xFFF
            assert o.unmake_packet(pkt[4:4 + L]) == (True, p)
    assert o.unmake_packet(b"\x01\x02") == (False, b"")
    with pytest.raises(ValueError):
        o.make_packet(bytes(4093))


def test_peak_detector_vectorised_equals_sequential():
    rng = np.random.default_rng(3)
    lay = o.Layout(512, 200, 128, "qpsk")
    for snr in (40, 15, 8):
        _, xc = loopback_capture(lay, payloads(rng, 5), snr, 0.2, seed=int(snr))
        y = o.chan_filter(xc, o.chan_filter_taps(lay))
        mf, _, _ = o.sync_pn_metric(y, 512, 128)
        assert np.array_equal(o.peak_detect(mf), o.peak_detect_sequential(mf))
    # synthetic stress: random walks hit the "new maximum keeps the run alive" branch
    for s in range(20):
        r = np.random.default_rng(100 + s)
        mf = (np.cumsum(r.standard_normal(4000)) * 0.02 - 0.5).astype(np.float32)
        assert np.array_equal(o.peak_detect(mf), o.peak_detect_sequential(mf))


def test_plan_closed_form_equals_sampler_calls():
    r = np.random.default_rng(4)
    N, L = 64, 80
    for case in range(300):
        n = int(r.integers(200, 6000))
        k = int(r.integers(0, 14))
        trig = np.unique(r.integers(0, n, size=k))
        if case % 3 == 0 and len(trig) > 1:                   # clusters of close triggers
            trig = np.unique(np.concatenate([trig, trig[:3] + r.integers(1, 5, size=len(trig[:3]))]))
            trig = trig[trig < n]
        to = 5 if case % 2 else 1000                          # small timeout exercises the NO_SIG path
        vs, vf, ft, nd = o.sampler_sim(trig, n, N, L, to)
        first_ok, st, ndc = plan_closed_form(trig, n, N, L, to)
        assert np.array_equal(st, trig[ft] - N + 1 if len(ft) else np.zeros(0, np.int64)), (case, trig, n)
        assert np.array_equal(ndc, nd), (case, trig, n, nd, ndc)


def test_sampler_emits_timeout_plus_one_data_vectors():
    # upstream `if (d_timeout-- == 0) d_state = STATE_NO_SIG` is a post-decrement: a lone trigger followed by more
    # than 1001 symbols of signal-free stream yields exactly 1001 data vectors
    N, L = 64, 80
    trig = np.array([100], dtype=np.int64)
    vs, vf, ft, nd = o.sampler_sim(trig, 100 + 1200 * L, N, L)
    assert list(nd) == [1001] and len(vs) == 1002 and vf[0] == 1 and not vf[1:].any()
    first_ok, st, ndc = plan_closed_form(trig, 100 + 1200 * L, N, L)
    assert list(ndc) == [1001] and list(st) == [100 - N + 1]
    # a second trigger caught by the first NO_SIG call: 1001 data vectors, then the new frame
    trig = np.array([100, 100 + 1001 * L + 7], dtype=np.int64)
    vs, vf, ft, nd = o.sampler_sim(trig, 100 + 1300 * L, N, L)
    first_ok, st, ndc = plan_closed_form(trig, 100 + 1300 * L, N, L)
    assert nd[0] == 1001 and np.array_equal(ndc, nd) and np.array_equal(st, trig[ft] - N + 1)


CASES = [(512, 200, 128, "bpsk", 40, 0.0), (512, 200, 128, "qpsk", 20, 0.3), (512, 200, 128, "8psk", 30, 0.2),
         (512, 200, 128, "qam16", 25, -0.4), (1024, 400, 256, "qam64", 30, 1.3), (1024, 800, 256, "qam64", 30, 0.3),
         (4096, 3200, 512, "qam256", 38, 0.3)]


@pytest.mark.parametrize("N,occ,cp,mod,snr,cfo", CASES)
def test_loopback_recovers_payloads(N, occ, cp, mod, snr, cfo):
    rng = np.random.default_rng(5)
    lay = o.Layout(N, occ, cp, mod)
    pay = payloads(rng, 8)
    _, xc = loopback_capture(lay, pay, snr, cfo, seed=9)
    r = o.rx_demodulate(xc, lay)
    good = [p for ok, p in r.packets if ok]
    assert len(good) >= 6                                     # the first frame after a CFO step may be lost (C.2)
    for p in good:
        assert p == pay[struct.unpack("!H", p[:2])[0]]


def test_zero_gap_poisons_the_detector():
    # C.1: an all-zero span gives 0/0 = NaN in the metric; the IIR average never recovers
    rng = np.random.default_rng(6)
    lay = o.Layout(512, 200, 128, "bpsk")
    x = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in payloads(rng, 2)], lay, 0.25, seed=1)
    burst = o.channel(x, 40, 0.0, 512, seed=1)
    cap = np.concatenate([burst, np.zeros(3000, np.complex64), burst, np.zeros(1500, np.complex64)])
    r = o.rx_demodulate(cap, lay, keep=True)
    assert np.isnan(r.mf).any()
    first_nan = int(np.flatnonzero(np.isnan(r.mf))[0])
    assert len(r.trig) >= 1 and (r.trig < first_nan).all()


def test_sense_chain():
    r = np.random.default_rng(7)
    N = 256
    x = ((r.standard_normal(40 * N) + 1j * r.standard_normal(40 * N)) * 1e-3).astype(np.complex64)
    x += (0.02 * np.exp(2j * np.pi * 0.2 * np.arange(40 * N))).astype(np.complex64)
    mh = o.sense_maxhold(x, N, 1, 3)
    assert mh.shape == (10, N) and (mh >= 0).all()
    avg, free, hx = o.sense_decide(mh, 1e-3)
    assert len(hx) == N // 4 and set(hx) <= set("0123456789ABCDEF")
    busy = np.flatnonzero(free == 0)
    assert len(busy) and abs(int(busy.mean()) - (N // 2 + int(0.2 * N))) <= 3       # tone at +0.2 fs, frequency order
    assert o.hex_conv(free.tolist()) == hx
