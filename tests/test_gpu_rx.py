"""Receive chain against the oracle, stage by stage and end to end, through the C ABI.

Bit-exact: decoded bytes, CRC verdicts, trigger / frame tables, slicer decisions (on the oracle's filtered
stream, which isolates the stage from the 1e-7 differences of the FFT-based channel filter).  Tolerance 1e-4
relative L2 (north star) for float32 samples: filtered stream, equalised symbols."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import ofdm_oracle as o
from helpers import payloads, rel_l2, loopback_capture

pytestmark = pytest.mark.gpu
TOL = 1e-4

CASES = [(512, 200, 128, "bpsk", 8, 40, 0.0), (512, 200, 128, "qpsk", 24, 20, 0.3), (512, 200, 128, "qam16", 24, 25, -0.4),
         (512, 200, 128, "8psk", 12, 30, 0.2), (512, 200, 128, "qam64", 10, 32, 0.1), (1024, 400, 256, "qam64", 12, 30, 1.3),
         (1024, 800, 256, "qam64", 10, 30, -2.2), (4096, 3200, 512, "qam256", 6, 38, 0.3), (256, 104, 64, "qpsk", 12, 30, 0.1),
         (2048, 800, 512, "qam16", 6, 30, 0.2), (128, 56, 32, "qpsk", 20, 30, 0.05),
         (512, 200, 100, "qpsk", 10, 30, 0.15),          # cp not a multiple of the 8 samples a lane owns
         # narrow bands make long channel filters: 241 taps (the 1024-point warp filter with 8 dropped rows) and
         # 321 taps (beyond it: the three-pass 2048-point filter)
         (512, 128, 64, "qpsk", 10, 30, 0.1), (512, 96, 64, "qpsk", 10, 30, -0.1)]


def stage_run(eng, xc, r, lay):
    """Runs every stage on the ORACLE's intermediate of the previous stage and returns the GPU outputs."""
    import torch
    from ofdm_uhd_b200 import _lib
    from gpu_helpers import processed_vectors
    n = len(xc)
    L_, st = eng.L_, eng._stream()
    nvec = len(r.vec_start) + 64
    bufs = eng.rx_alloc(n, taps=True, max_vectors=nvec)
    io = bufs["io"]
    out = {}
    d_x = torch.from_numpy(xc).cuda()
    y_g = torch.empty(n, dtype=torch.complex64, device="cuda")
    _lib.check(L_.ofdm_rx_chan_filter(eng.h, eng._p(d_x), n, eng._p(y_g), st))
    out["y"] = y_g.cpu().numpy()
    d_y = torch.from_numpy(r.y).cuda()
    mf_g = torch.empty(n, dtype=torch.float32, device="cuda")
    fnan = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(L_.ofdm_rx_sync_metric(eng.h, eng._p(d_y), n, eng._p(mf_g), eng._p(fnan), st))
    out["mf"] = mf_g.cpu().numpy()
    out["first_nan"] = int(fnan.item())
    d_mf = torch.from_numpy(r.mf).cuda()
    _lib.check(L_.ofdm_rx_peak_detect(eng.h, eng._p(d_y), eng._p(d_mf), n, eng._p(fnan), C.byref(io), st))
    _lib.check(L_.ofdm_rx_plan(eng.h, n, C.byref(io), st))
    _lib.check(L_.ofdm_rx_demod(eng.h, eng._p(d_y), n, C.byref(io), st))
    _lib.check(L_.ofdm_rx_finish(eng.h, C.byref(io), st))
    res = eng.collect(bufs)
    out["res"] = res
    nf = res.n_frames
    vb_ptr = L_.ofdm_rx_workspace_ptr(eng.h, C.byref(io), n, 5)
    nv_ptr = L_.ofdm_rx_workspace_ptr(eng.h, C.byref(io), n, 6)
    base = bufs["workspace"].data_ptr()
    ws = bufs["workspace"]
    out["vbase"] = ws[vb_ptr - base: vb_ptr - base + 8 * (nf + 1)].view(torch.int64).cpu().numpy()
    out["nvec"] = ws[nv_ptr - base: nv_ptr - base + 4 * nf].view(torch.int32).cpu().numpy()
    nv = len(r.vec_start)
    out["eq"] = bufs["eq_syms"][:nv * lay.occupied_tones].cpu().numpy().reshape(nv, lay.occupied_tones)
    out["sym"] = bufs["sym_idx"][:nv * lay.ncar].cpu().numpy().reshape(nv, lay.ncar)
    out["rows"] = processed_vectors(out["vbase"], res.frame_ndata, out["nvec"])
    return out


@pytest.mark.parametrize("N,occ,cp,mod,nfr,snr,cfo", CASES)
def test_stage_parity(N, occ, cp, mod, nfr, snr, cfo):
    from ofdm_uhd_b200.engine import OfdmEngine
    from gpu_helpers import oracle_demapped
    rng = np.random.default_rng(N + nfr)
    lay = o.Layout(N, occ, cp, mod)
    pay = payloads(rng, nfr)
    _, xc = loopback_capture(lay, pay, snr, cfo, seed=N + 1)
    r = o.rx_demodulate(xc, lay, keep=True)
    assert len(r.packets) >= nfr - 6
    eng = OfdmEngine(N, occ, cp, mod)
    assert np.array_equal(eng.chan_taps(), o.chan_filter_taps(lay))
    g = stage_run(eng, xc, r, lay)
    # channel filter: float32 samples within tolerance
    assert rel_l2(g["y"], r.y) < TOL
    # timing metric on identical input: same op order -> equal up to the last bit of the float64 prefix sums
    finite = np.isfinite(r.mf)
    assert np.array_equal(np.isfinite(g["mf"]), finite)
    assert float(np.max(np.abs(g["mf"][finite] - r.mf[finite]) / (1 + np.abs(r.mf[finite])))) < 1e-5
    assert float(np.mean(g["mf"][finite] != r.mf[finite])) < 0.01
    # peak detector / angle latch / sampler plan on identical input: exact
    res = g["res"]
    assert np.array_equal(res.trig_idx, r.trig)
    assert float(np.max(np.abs(res.trig_ang - r.ang))) < 1e-6
    assert np.array_equal(res.frame_start, r.frame_start) and np.array_equal(res.frame_ndata, r.n_data)
    # demod on identical input: packets and CRC verdicts exact, equalised symbols within tolerance
    assert res.packets == r.packets
    rows = g["rows"]
    assert rel_l2(g["eq"][rows], r.eq[rows]) < TOL
    # slicer decisions of the live sessions' own vectors
    dem = oracle_demapped(r, lay)
    own = set(rows.tolist())
    live_rows = set()
    for f in np.flatnonzero(res.frame_live):
        k = int(min(int(g["nvec"][f]), 1 + int(res.frame_ndata[f])))
        live_rows.update(range(int(g["vbase"][f]), int(g["vbase"][f]) + k))
    checked = 0
    for k, v in enumerate(dem):
        if v in live_rows and v in own:
            assert np.array_equal(g["sym"][v], r.sym_idx[k]), (v,)
            checked += 1
    assert checked >= len(r.packets)
    eng.close()


@pytest.mark.parametrize("N,occ,cp,mod,nfr,snr,cfo", CASES)
def test_full_chain_equals_oracle(N, occ, cp, mod, nfr, snr, cfo):
    """The whole GPU chain (its own FFT-based filter output feeding the detector) against the oracle."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    rng = np.random.default_rng(7 * N + nfr)
    lay = o.Layout(N, occ, cp, mod)
    pay = payloads(rng, nfr)
    _, xc = loopback_capture(lay, pay, snr, cfo, seed=N + 2)
    r = o.rx_demodulate(xc, lay, keep=True)
    eng = OfdmEngine(N, occ, cp, mod)
    res = eng.demodulate(torch.from_numpy(xc).cuda())
    # a trigger may move by one sample where two neighbouring metric values tie to ~1e-7 (flat plateau);
    # decoded bytes and CRC verdicts must still be identical
    assert len(res.trig_idx) == len(r.trig) and int(np.max(np.abs(res.trig_idx - r.trig))) <= 1
    assert res.packets == r.packets
    c = res.counters
    assert c[1] == len(r.packets) and c[2] == sum(1 for ok, _ in r.packets if ok)
    eng.close()


def test_edge_streams():
    """Empty, shorter-than-one-symbol and noise-only inputs; a zero gap (NaN poisoning, C.1)."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(512, 200, 128, "bpsk")
    eng = OfdmEngine(512, 200, 128, "bpsk")
    rng = np.random.default_rng(8)
    for n in (0, 1, 100, 700):
        x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.01).astype(np.complex64)
        res = eng.demodulate(torch.from_numpy(x).cuda() if n else torch.zeros(0, dtype=torch.complex64, device="cuda"))
        assert res.packets == [] and res.n_frames == 0
    noise = ((rng.standard_normal(200000) + 1j * rng.standard_normal(200000)) * 0.05).astype(np.complex64)
    res = eng.demodulate(torch.from_numpy(noise).cuda())
    ref = o.rx_demodulate(noise, lay)
    assert res.packets == ref.packets == [] and np.array_equal(res.trig_idx, ref.trig)
    eng.close()


def test_zero_gap_nan_poisoning_stage_parity():
    """C.1: an all-zero span makes the metric 0/0 = NaN and the detector's average never recovers.  Whether a
    zero INPUT span yields exactly-zero filter output depends on the FFT block alignment of the channel filter
    (true of gr.fft_filter_ccc as well), so the quirk is pinned per stage on the oracle's filtered stream: same
    first NaN index, same triggers (none after the gap); end to end, everything before the gap is identical."""
    import torch
    from ofdm_uhd_b200 import _lib
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(512, 200, 128, "bpsk")
    eng = OfdmEngine(512, 200, 128, "bpsk")
    rng = np.random.default_rng(8)
    x = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in payloads(rng, 2)], lay, 0.25, seed=1)
    burst = o.channel(x, 40, 0.0, 512, seed=1)
    cap = np.concatenate([burst, np.zeros(3000, np.complex64), burst, np.zeros(1500, np.complex64)])
    ref = o.rx_demodulate(cap, lay, keep=True)
    assert np.isnan(ref.mf).any()
    want_nan = int(np.flatnonzero(np.isnan(ref.mf))[0])
    n = len(cap)
    L_, st = eng.L_, eng._stream()
    bufs = eng.rx_alloc(n)
    d_y = torch.from_numpy(ref.y).cuda()
    mf_g = torch.empty(n, dtype=torch.float32, device="cuda")
    fnan = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(L_.ofdm_rx_sync_metric(eng.h, eng._p(d_y), n, eng._p(mf_g), eng._p(fnan), st))
    assert int(fnan.item()) == want_nan
    got = mf_g.cpu().numpy()
    fin = np.isfinite(ref.mf[:want_nan])
    assert np.array_equal(np.isfinite(got[:want_nan]), fin) and np.array_equal(got[:want_nan][~fin], ref.mf[:want_nan][~fin])
    assert float(np.max(np.abs(got[:want_nan][fin] - ref.mf[:want_nan][fin]) / (1 + np.abs(ref.mf[:want_nan][fin])))) < 1e-5
    d_mf = torch.from_numpy(ref.mf).cuda()
    _lib.check(L_.ofdm_rx_peak_detect(eng.h, eng._p(d_y), eng._p(d_mf), n, eng._p(fnan), C.byref(bufs["io"]), st))
    torch.cuda.synchronize()
    nt = int(bufs["n_trig"].item())
    assert np.array_equal(bufs["trig_idx"][:nt].cpu().numpy(), ref.trig) and (ref.trig < want_nan).all()
    res = eng.demodulate(torch.from_numpy(cap).cuda())
    k = len(ref.packets)
    assert k >= 1 and res.packets[:k] == ref.packets
    eng.close()


def test_large_stream_round_trip():
    """Size-independent properties at scale (5 000 QAM16 frames, noise + CFO): every delivered payload with a
    good CRC equals what was sent, almost all frames are delivered, a second run is bit-identical."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    F, psize = 5000, 402
    eng = OfdmEngine(512, 200, 128, "qam16", 0.25, pad_seed=3, max_pkt_bytes=416)
    rng = np.random.default_rng(9)
    body = rng.integers(0, 256, size=(F, psize), dtype=np.uint8)
    body[:, 0] = np.arange(F) >> 8
    body[:, 1] = np.arange(F) & 0xFF
    off = np.arange(F + 1, dtype=np.int64) * psize
    plan = eng.tx_plan(off)
    x = eng.tx_run(plan, torch.from_numpy(body.reshape(-1)).cuda())
    lead = torch.zeros(1300, dtype=torch.complex64, device="cuda")
    cap = torch.cat([lead, x, lead])
    p_sig = float((x.abs() ** 2).mean())
    sigma = (p_sig / 10 ** 2.5 / 2) ** 0.5                                             # 25 dB
    xc = eng.channel(cap, cfo=0.31, sigma=sigma, seed=17)
    res = eng.demodulate(xc)
    ok = [(g, p) for g, p in res.packets if g]
    assert len(res.packets) >= F - 10 and len(ok) >= 0.93 * F          # LS channel estimate from one preamble
    # the oracle on a prefix of the same capture delivers the same packets (ok or not), in the same order
    npre = 1300 + 200 * 6 * 640
    ref = o.rx_demodulate(xc[:npre].cpu().numpy(), o.Layout(512, 200, 128, "qam16"))
    k = len(ref.packets) - 3
    # The first frame after the CFO step decodes to garbage (C.2) and its bogus header length sits on slicer
    # boundaries, where the 1e-7 differences of the FFT-based filter show; from then on the two receivers must
    # deliver the same good packets in the same order.
    assert k > 150
    good_g = [p for g, p in res.packets if g and 10 <= ((p[0] << 8) | p[1]) < k]
    good_o = [p for g, p in ref.packets if g and 10 <= ((p[0] << 8) | p[1]) < k]
    assert good_g == good_o and len(good_o) > 0.9 * (k - 10)
    for _, p in ok:
        k = (p[0] << 8) | p[1]
        assert p == body[k].tobytes()
    res2 = eng.demodulate(xc)
    assert res2.packets == res.packets and np.array_equal(res2.trig_idx, res.trig_idx)
    assert res.counters[2] == len(ok) and res.counters[3] == len(ok) * psize
    eng.close()


@pytest.mark.parametrize("N,occ,cp,mod,nfr,snr,cfo", CASES)
def test_fused_sync_equals_oracle(N, occ, cp, mod, nfr, snr, cfo):
    """ofdm_rx_sync (the fused streaming ofdm_sync_pn kernel for N = 128..512, the two-stage path otherwise) on the
    oracle's filtered stream: trigger indices exact, latched angles within float32 rounding."""
    import torch
    from ofdm_uhd_b200 import _lib
    from ofdm_uhd_b200.engine import OfdmEngine
    rng = np.random.default_rng(3 * N + nfr)
    lay = o.Layout(N, occ, cp, mod)
    _, xc = loopback_capture(lay, payloads(rng, nfr), snr, cfo, seed=N + 5)
    r = o.rx_demodulate(xc, lay, keep=True)
    eng = OfdmEngine(N, occ, cp, mod)
    n = len(xc)
    bufs = eng.rx_alloc(n)
    d_y = torch.from_numpy(r.y).cuda()
    _lib.check(eng.L_.ofdm_rx_sync(eng.h, eng._p(d_y), n, C.byref(bufs["io"]), eng._stream()))
    torch.cuda.synchronize()
    nt = int(bufs["n_trig"].item())
    assert int(bufs["status"].item()) == 0
    assert np.array_equal(bufs["trig_idx"][:nt].cpu().numpy(), r.trig)
    assert float(np.max(np.abs(bufs["trig_ang"][:nt].cpu().numpy() - r.ang))) < 1e-6
    eng.close()


def test_fused_sync_long_stream_segments():
    """Several detector segments (seg_len 65 536) with warm-up and priming steps: 600 QAM16 frames, triggers must
    equal the oracle's sequential detector on the same filtered stream."""
    import torch
    from ofdm_uhd_b200 import _lib
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(512, 200, 128, "qam16")
    rng = np.random.default_rng(77)
    _, xc = loopback_capture(lay, payloads(rng, 600), 22, 0.17, seed=78)
    y = o.chan_filter(xc, o.chan_filter_taps(lay))
    mf, Pr, Pi = o.sync_pn_metric(y, 512, 128)
    trig = o.peak_detect(mf)
    eng = OfdmEngine(512, 200, 128, "qam16")
    n = len(xc)
    assert n > 30 * 65536
    bufs = eng.rx_alloc(n)
    _lib.check(eng.L_.ofdm_rx_sync(eng.h, eng._p(torch.from_numpy(y).cuda()), n, C.byref(bufs["io"]), eng._stream()))
    torch.cuda.synchronize()
    nt = int(bufs["n_trig"].item())
    assert nt == len(trig) and np.array_equal(bufs["trig_idx"][:nt].cpu().numpy(), trig)
    eng.close()


def test_independent_streams_are_order_invariant():
    """BASELINE configs[2] in miniature: independent streams (fft 1024 / occ 400 / cp 256, QAM64, per-stream CFO)
    are the unit of multi-GPU sharding; each stream's result depends only on that stream, whatever rank, order
    or buffer reuse it was processed with, and equals the oracle's."""
    import torch
    from ofdm_uhd_b200 import sharding
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(1024, 400, 256, "qam64")
    eng = OfdmEngine(1024, 400, 256, "qam64")
    rng = np.random.default_rng(21)
    caps, refs = [], []
    for s in range(6):
        _, xc = loopback_capture(lay, payloads(rng, 10), 30, float(rng.uniform(-1.5, 1.5)), seed=100 + s)
        caps.append(xc)
    refs = [o.rx_demodulate(c, lay).packets for c in caps[:2]]
    first = {}
    for world in (1, 2, 4):
        for rank in range(world):
            for s in reversed(sharding.streams_of_rank(6, world, rank)):
                got = eng.demodulate(torch.from_numpy(caps[s]).cuda()).packets
                assert first.setdefault(s, got) == got
    assert first[0] == refs[0] and first[1] == refs[1]
    assert all(sum(1 for g, _ in first[s] if g) >= 7 for s in range(6))
    eng.close()


@pytest.mark.parametrize("F,p_exc", [(1, 0.0), (50, 0.3), (20000, 0.0), (20000, 0.002), (20000, 0.5), (100000, 0.01)])
def test_liveness_walk_fast_and_general(F, p_exc):
    """ofdm_rx_liveness (the sink-liveness stage of ofdm_rx_finish) on synthetic frame tables: the short-exception
    fast path, its overflow hand-over (p_exc = 0.5 flags ~10 000 frames > LIVE_CAP) and the forced general walk all
    equal the sequential orbit of frame 0 under next()."""
    import torch
    from ofdm_uhd_b200 import _lib
    rng = np.random.default_rng(F + int(1000 * p_exc))
    ndata = rng.integers(1, 6, size=F)
    vbase = np.concatenate([[0], np.cumsum(1 + ndata)])[:F].astype(np.int64)
    nvec = (1 + ndata).astype(np.int32)
    exc = rng.random(F) < p_exc
    nvec[exc] = rng.integers(1, 60, size=int(exc.sum())).astype(np.int32)
    if F > 10:
        nvec[F // 2] = np.iinfo(np.int32).max // 2          # a session that never ends swallows the rest
        nvec[exc & (np.arange(F) >= F // 2)] = 3
    live_ref = np.zeros(F, dtype=np.uint8)
    f = 0
    while f < F:
        live_ref[f] = 1
        f = max(f + 1, int(np.searchsorted(vbase, vbase[f] + int(nvec[f]), side="left")))
    L = _lib.lib()
    maxf = F + 7
    d_n = torch.tensor([F], dtype=torch.int32, device="cuda")
    d_vb = torch.from_numpy(vbase).cuda()
    d_nv = torch.from_numpy(nvec).cuda()
    for force in (0, 1):
        scratch = torch.zeros(2 * maxf + 2, dtype=torch.int32, device="cuda")
        live = torch.full((maxf,), 7, dtype=torch.uint8, device="cuda")
        _lib.check(L.ofdm_rx_liveness(d_n.data_ptr(), d_vb.data_ptr(), d_nv.data_ptr(), maxf, scratch.data_ptr(),
                                      live.data_ptr(), force, _lib.stream_ptr()))
        torch.cuda.synchronize()
        got = live.cpu().numpy()
        assert np.array_equal(got[:F], live_ref), (force, int(np.flatnonzero(got[:F] != live_ref)[0]))
        assert not got[F:].any()


@pytest.mark.parametrize("mod,nsym,fo", [("bpsk", 18, 0.0), ("qpsk", 10, 0.3), ("qam64", 4, -0.45)])
def test_fixed_sync_mode_equals_oracle(mod, nsym, fo):
    """The reference's SYNC == "fixed" test mode (ofdm_receiver.py~:108-119): no channel filter, a trigger at the end
    of the first symbol of every nsym-symbol packet, constant frequency offset.  Frames sit back to back from sample
    0; the GPU chain (ofdm_rx_demodulate_fixed) must deliver what the oracle's fixed-mode receiver delivers."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(512, 200, 128, mod)
    rng = np.random.default_rng(nsym)
    pay = payloads(rng, 14)
    pkts = [o.make_packet(p, 1, 1, False) for p in pay]
    x = o.tx_modulate(pkts, lay, 0.25, seed=2)
    assert len(x) == 14 * nsym * 640                       # 402-byte payloads: 18 / 10 / 4 symbols per frame
    tail = np.zeros(3 * 640, dtype=np.complex64)
    xc = o.channel(np.concatenate([x, tail]), 30.0, fo, 512, seed=6, sig_power=float(np.mean(np.abs(x) ** 2)))
    ref = o.rx_demodulate(xc, lay, sync="fixed", nsymbols=nsym, freq_offset=fo)
    assert sum(1 for ok, _ in ref.packets if ok) >= 13
    eng = OfdmEngine(512, 200, 128, mod)
    got = eng.demodulate_fixed(torch.from_numpy(xc).cuda(), nsym, fo)
    assert np.array_equal(got.trig_idx, ref.trig) and np.array_equal(got.trig_ang, ref.ang)
    assert got.packets == ref.packets
    # the same call through the stage entry points
    bufs = eng.rx_alloc(len(xc))
    L_, st = eng.L_, eng._stream()
    d = torch.from_numpy(xc).cuda()
    from ofdm_uhd_b200 import _lib
    _lib.check(L_.ofdm_rx_sync_fixed(eng.h, len(xc), nsym, fo, C.byref(bufs["io"]), st))
    _lib.check(L_.ofdm_rx_plan(eng.h, len(xc), C.byref(bufs["io"]), st))
    _lib.check(L_.ofdm_rx_demod(eng.h, eng._p(d), len(xc), C.byref(bufs["io"]), st))
    _lib.check(L_.ofdm_rx_finish(eng.h, C.byref(bufs["io"]), st))
    assert eng.collect(bufs).packets == ref.packets
    # and a pn run afterwards on the same buffers starts its NCO at rest again
    _, xc2 = loopback_capture(lay, pay[:6], 30, 0.2, seed=3)
    assert eng.demodulate(torch.from_numpy(xc2).cuda()).packets == o.rx_demodulate(xc2, lay).packets
    eng.close()


@pytest.mark.parametrize("N,occ,cp,mod,snr", [(512, 200, 128, "qam64", 14), (512, 200, 128, "qam256", 18),
                                                (1024, 400, 256, "qam256", 40)])
def test_grid_slicer_equals_brute_force(N, occ, cp, mod, snr):
    """The 3 x 3-cell slicer used for square-grid constellations must return, for every carrier, exactly the first
    minimum of the float32 distances over ALL points (ofdm_frame_sink's scan).  Checked on the GPU's own derotated
    symbols at an SNR low enough to scatter them over every cell, beyond the hull, and onto decision boundaries."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(N, occ, cp, mod)
    rng = np.random.default_rng(int(snr))
    _, xc = loopback_capture(lay, payloads(rng, 12), snr, 0.2, seed=77)
    eng = OfdmEngine(N, occ, cp, mod)
    nv = 12 * 8
    bufs = eng.rx_alloc(len(xc), taps=True, max_vectors=nv)
    eng.demodulate_async(torch.from_numpy(xc).cuda(), bufs)
    torch.cuda.synchronize()
    r = bufs["derot_syms"][:nv * lay.ncar].cpu().numpy().reshape(nv, lay.ncar)
    sym = bufs["sym_idx"][:nv * lay.ncar].cpu().numpy().reshape(nv, lay.ncar)
    used = np.abs(r).sum(axis=1) > 0                       # vectors some session demapped
    assert used.sum() >= 10
    c = o.constellation_for(mod).astype(np.complex64)
    rr, ri = r[used].real.astype(np.float32), r[used].imag.astype(np.float32)
    dx = rr[..., None] - c.real.astype(np.float32)
    dy = ri[..., None] - c.imag.astype(np.float32)
    dd = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
    ref = np.argmin(dd, axis=-1)                           # first minimum in index order
    assert np.array_equal(sym[used], ref.astype(np.uint8))
    spread = np.unique(ref).size
    assert spread >= (40 if mod == "qam64" else 100)       # the noise really exercised the table
    eng.close()


@pytest.mark.parametrize("N,occ,cp", [(512, 200, 128), (1024, 400, 256), (4096, 3200, 512), (128, 56, 32)])
def test_streaming_sync_edge_lengths(N, occ, cp):
    """ofdm_rx_sync forces the streaming kernels (one warp per chunk / segment) whatever the length: streams shorter
    than a step, a block, a chunk, with ragged tails and with a zero gap (NaN poisoning) must give the oracle's
    triggers exactly."""
    import torch
    from ofdm_uhd_b200 import _lib
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(N, occ, cp, "qpsk")
    rng = np.random.default_rng(N)
    _, xc = loopback_capture(lay, payloads(rng, 6), 30, 0.1, seed=N + 9)
    y = o.chan_filter(xc, o.chan_filter_taps(lay))
    eng = OfdmEngine(N, occ, cp, "qpsk")
    W = N // 2
    lengths = [1, 7, 255, 256, 257, W - 1, W, W + 1, 2 * W + 3, 5 * W - 1, 16385, len(y) - 3, len(y)]
    for n in sorted(set(v for v in lengths if 0 < v <= len(y))):
        yy = y[:n].copy()
        if n > 6 * W:
            yy[3 * W + 5: 4 * W + 90] = 0                     # an exactly-zero window: 0/0 = NaN from there on (C.1)
        mf, Pr, Pi = o.sync_pn_metric(yy, N, cp)
        trig = o.peak_detect(mf)
        nan = np.flatnonzero(np.isnan(mf))
        if len(nan):
            trig = trig[trig < nan[0]]
        bufs = eng.rx_alloc(n)
        _lib.check(eng.L_.ofdm_rx_sync(eng.h, eng._p(torch.from_numpy(yy).cuda()), n, C.byref(bufs["io"]), eng._stream()))
        _lib.check(eng.L_.ofdm_rx_plan(eng.h, n, C.byref(bufs["io"]), eng._stream()))
        torch.cuda.synchronize()
        nt = int(bufs["n_trig"].item())
        assert np.array_equal(bufs["trig_idx"][:nt].cpu().numpy(), trig), (n,)
    eng.close()


def test_ragged_and_long_packets_dewhiten_crc():
    """unmake_packet on the device for every slice shape of the warp-per-packet CRC (payloads of 0 .. 4091 bytes, one
    with a corrupted byte) and, with small slots, of the staged thread-per-packet CRC: verdicts and bytes must equal
    the oracle's."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(512, 200, 128, "qam16")
    rng = np.random.default_rng(99)
    sizes = [0, 1, 3, 4, 123, 124, 125, 252, 1000, 2047, 4090, 4091, 402, 402]
    pay = [bytes(rng.integers(0, 256, s, dtype=np.uint8)) for s in sizes]
    pkts = [o.make_packet(p, 1, 1, False) for p in pay]
    bad = bytearray(pkts[8]); bad[300] ^= 0x10; pkts[8] = bytes(bad)              # CRC must fail for the 1000-byte one
    x = o.tx_modulate(pkts, lay, 0.25, seed=4)
    lead = np.zeros(2 * lay.sym_len, dtype=np.complex64)
    xc = o.channel(np.concatenate([lead, x, lead]), 45.0, 0.05, 512, seed=8, sig_power=float(np.mean(np.abs(x) ** 2)))
    ref = o.rx_demodulate(xc, lay)
    assert len(ref.packets) >= 13 and sum(1 for ok, _ in ref.packets if not ok) >= 1
    for mpb in (4096, 1016, 420):
        eng = OfdmEngine(512, 200, 128, "qam16", max_pkt_bytes=mpb)
        got = eng.demodulate(torch.from_numpy(xc).cuda())
        assert len(got.packets) == len(ref.packets)
        for (gok, gp), (rok, rp) in zip(got.packets, ref.packets):
            if len(rp) + 4 <= eng.pkt_stride:
                assert (gok, gp) == (rok, rp)
            else:
                assert not gok                                                   # does not fit the slot: reported bad
        eng.close()


def test_full_size_loopback_properties():
    """BASELINE configs[1] at its full size: 100 000 QPSK frames = 1 M OFDM symbols = 640 M samples in one stream, AWGN
    20 dB, a new CFO every 10 000 frames.  The oracle cannot run at this size, so the checks are size-independent:
    every payload delivered with a good CRC is byte-for-byte the one sent under that packet number, nearly all frames
    arrive, the counters agree with the per-frame tables, and a second run is bit-identical (the streaming kernels cover
    the stream with ~3 500 detector segments and ~78 000 metric chunks, so this is also their seam test)."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~30 GB of device memory")
    F, psize = 100000, 402
    eng = OfdmEngine(512, 200, 128, "qpsk", 0.25, pad_seed=11, max_pkt_bytes=416)
    rng = np.random.default_rng(12)
    body = rng.integers(0, 256, size=(F, psize), dtype=np.uint8)
    idx = np.arange(F, dtype=np.uint32)
    body[:, 0], body[:, 1], body[:, 2] = idx >> 16, (idx >> 8) & 0xFF, idx & 0xFF
    plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
    lead = 2 * eng.L
    n = plan.n_samples + 2 * lead
    assert plan.n_samples == 640_000_000
    x = torch.zeros(n, dtype=torch.complex64, device="cuda")
    xs = x[lead:lead + plan.n_samples]
    eng.tx_run(plan, torch.from_numpy(body.reshape(-1)).cuda(), out=xs)
    p_sig = float((xs[:1 << 22].abs() ** 2).mean())
    sigma = (p_sig / 100.0 / 2.0) ** 0.5
    xc = torch.empty_like(x)
    seg = 10000 * 10 * eng.L
    phase = 0.0
    cfos = np.random.default_rng(13).uniform(-0.5, 0.5, size=10)
    for i, cfo in enumerate(cfos):
        lo = 0 if i == 0 else lead + i * seg
        hi = n if i == 9 else lead + (i + 1) * seg
        eng.channel(x[lo:hi], cfo=float(cfo), sigma=sigma, seed=500 + i, phase0=phase, out=xc[lo:hi])
        phase = (phase + 2 * np.pi * cfo / 512 * (hi - lo)) % (2 * np.pi)
    del x
    bufs = eng.rx_alloc(n, max_frames=F + 1024)
    r1 = eng.collect(eng.demodulate_async(xc, bufs), want_packets=False, want_payload=True)
    rows = r1.payload_rows.copy()
    sel, ok, plen = r1.msg_frames, r1.pkt_ok.copy(), r1.pkt_len.copy()
    assert r1.n_frames >= F - 20 and len(sel) >= F - 200
    good = sel[ok[sel] == 1]
    assert len(good) >= 0.995 * F                                   # the first frame after each CFO step may be lost (C.2)
    assert np.all(plen[good] == psize + 4)
    got = rows[good, :psize]
    num = (got[:, 0].astype(np.int64) << 16) | (got[:, 1].astype(np.int64) << 8) | got[:, 2]
    assert np.all(np.diff(num) > 0)                                 # in order, no duplicates
    assert np.array_equal(got, body[num])                           # CRC-good payloads are the ones sent
    assert int(r1.counters[1]) == len(sel) and int(r1.counters[2]) == len(good)
    assert int(r1.counters[3]) == len(good) * psize and int(r1.counters[4]) == n
    trig1 = bufs["trig_idx"][:r1.n_trig].clone()
    r2 = eng.collect(eng.demodulate_async(xc, bufs), want_packets=False, want_payload=True)
    assert r2.n_trig == r1.n_trig and torch.equal(bufs["trig_idx"][:r2.n_trig], trig1)
    assert np.array_equal(r2.pkt_ok, ok) and np.array_equal(r2.payload_rows[good, :psize], got)
    d = np.diff(trig1.cpu().numpy())
    assert abs(np.median(d) - 6400) <= 2 and (np.abs(d - 6400) <= 64).mean() > 0.99   # one trigger per frame (the
    # Schmidl-Cox plateau lets the arg-max wander inside the cyclic prefix at 20 dB)
    eng.close()


def test_cfg5_stream_properties():
    """One stream of BASELINE configs[4] (fft 4096 / occ 3200 / cp 512, QAM256, maximum 4091-byte payloads, 3 symbols
    per frame) long enough for the streaming sync kernels: the wide (N/2 = 8 x 256) metric kernel, the multi-step
    detector ring, the single-buffer N = 4096 demodulator, the 3 x 3-cell slicer and the warp-per-packet CRC all sit
    on this path.  CRC-good payloads must be the ones sent; a second run must be identical."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    F, psize = 6000, 4091
    eng = OfdmEngine(4096, 3200, 512, "qam256", 0.25, pad_seed=2)
    rng = np.random.default_rng(31)
    body = rng.integers(0, 256, size=(F, psize), dtype=np.uint8)
    body[:, 0], body[:, 1] = np.arange(F) >> 8, np.arange(F) & 0xFF
    plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
    assert plan.uniform_syms == 3 and plan.n_samples == F * 3 * 4608
    lead = 2 * eng.L
    n = plan.n_samples + 2 * lead
    x = torch.zeros(n, dtype=torch.complex64, device="cuda")
    xs = x[lead:lead + plan.n_samples]
    eng.tx_run(plan, torch.from_numpy(body.reshape(-1)).cuda(), out=xs)
    p_sig = float((xs[:1 << 22].abs() ** 2).mean())
    xc = eng.channel(x, cfo=0.37, sigma=(p_sig / 10 ** 4.2 / 2) ** 0.5, seed=77)          # 42 dB
    bufs = eng.rx_alloc(n, max_frames=F + 256)
    r1 = eng.collect(eng.demodulate_async(xc, bufs), want_packets=False, want_payload=True)
    sel, ok = r1.msg_frames, r1.pkt_ok.copy()
    good = sel[ok[sel] == 1]
    assert len(sel) >= F - 10 and len(good) >= 0.9 * F
    got = r1.payload_rows[good, :psize].copy()
    num = (got[:, 0].astype(np.int64) << 8) | got[:, 1]
    assert np.all(np.diff(num) > 0) and np.array_equal(got, body[num])
    trig1 = bufs["trig_idx"][:r1.n_trig].clone()
    r2 = eng.collect(eng.demodulate_async(xc, bufs), want_packets=False, want_payload=True)
    assert torch.equal(bufs["trig_idx"][:r2.n_trig], trig1) and np.array_equal(r2.pkt_ok, ok)
    assert np.array_equal(r2.payload_rows[good, :psize], got)
    eng.close()


@pytest.mark.parametrize("N,occ,cp", [(512, 200, 128), (4096, 3200, 512)])
def test_degenerate_inputs_terminate_and_deliver_nothing(N, occ, cp):
    """All-zero, astronomically large, denormal-small, NaN- and Inf-contaminated captures: both sync paths (tile and
    forced streaming) must come back with no trigger, no frame and no message -- the reference's detector is poisoned
    by a NaN metric for the rest of the stream (C.1), and noise alone never crosses its threshold."""
    import torch
    from ofdm_uhd_b200 import _lib
    from ofdm_uhd_b200.engine import OfdmEngine
    eng = OfdmEngine(N, occ, cp, "qpsk")
    n = 300000
    rng = np.random.default_rng(1)
    base = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.01).astype(np.complex64)
    nan_in, inf_in = base.copy(), base.copy()
    nan_in[100000] = np.nan
    inf_in[150000] = np.inf
    cases = {"zeros": np.zeros(n, np.complex64), "huge": (base * 1e20).astype(np.complex64),
             "tiny": (base * 1e-28).astype(np.complex64), "noise": base, "nan": nan_in, "inf": inf_in}
    for name, x in cases.items():
        d = torch.from_numpy(x).cuda()
        r = eng.demodulate(d)
        assert (r.n_trig, r.n_frames, len(r.packets)) == (0, 0, 0), name
        bufs = eng.rx_alloc(n)
        io, st = C.byref(bufs["io"]), eng._stream()
        _lib.check(eng.L_.ofdm_rx_sync(eng.h, eng._p(d), n, io, st))            # streaming kernels, forced
        _lib.check(eng.L_.ofdm_rx_plan(eng.h, n, io, st))
        _lib.check(eng.L_.ofdm_rx_demod(eng.h, eng._p(d), n, io, st))
        _lib.check(eng.L_.ofdm_rx_finish(eng.h, io, st))
        r = eng.collect(bufs)
        assert (r.n_trig, r.n_frames, len(r.packets)) == (0, 0, 0), name
    eng.close()


def test_packed_math_selftest():
    """Every packed-fp32 helper of the decision paths returns, bit for bit, what its scalar definition returns
    (ptxas contracts packed multiply -> packed add pairs; common.cuh keeps them apart)."""
    from ofdm_uhd_b200 import _lib
    L_ = _lib.lib()
    out = (C.c_int64 * 2)()
    _lib.check(L_.ofdm_selftest_packed_math(0, 1 << 24, 12345, out), "selftest")
    assert out[0] == 0, "packed helpers differ from their scalar definitions (mask 0x%x) in %d of %d cases" % (out[1], out[0], 1 << 24)


def test_sampler_emits_1001_data_vectors():
    """`if (d_timeout-- == 0)`: a lone trigger followed by more than 1001 symbols yields a frame of exactly
    1001 data vectors -- plan kernel == oracle sampler, through the fixed synchroniser (one trigger per 1200 symbols)."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    N, occ, cp = 128, 56, 32
    L = N + cp
    n = 1300 * L
    rng = np.random.default_rng(3)
    xc = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.01).astype(np.complex64)
    lay = o.Layout(N, occ, cp, "qpsk")
    trig, ang = o.sync_fixed(n, N, cp, 1200, 0.0)
    plan = o.plan_frames(trig, ang, n, N, L)
    assert list(plan.n_data[:1]) == [1001]
    eng = OfdmEngine(N, occ, cp, "qpsk")
    got = eng.demodulate_fixed(torch.from_numpy(xc).cuda(), 1200, 0.0)
    assert np.array_equal(got.trig_idx, trig)
    assert np.array_equal(got.frame_ndata, plan.n_data) and got.frame_ndata[0] == 1001
    assert np.array_equal(got.frame_start, trig[plan.frame_trig] - N + 1)
    eng.close()


@pytest.mark.parametrize("sync,gain,N,occ,cp,mod", [("pnac", 8.0, 512, 200, 128, "qpsk"), ("pnac", 10.0, 512, 200, 128, "bpsk"),
                                                     ("ml", 1.0, 512, 200, 128, "qpsk"), ("ml", 1.0, 1024, 400, 256, "qam16"),
                                                     ("pnac", 8.0, 256, 104, 64, "qpsk")])
def test_alternative_synchronisers_equal_oracle(sync, gain, N, occ, cp, mod):
    """SYNC == "pnac" / "ml" of ofdm_receiver.py~:89-107 (never taken in the reference: SYNC is hard-coded to "pn"):
    triggers, angles, the NCO's own event list (ml) and the delivered packets are the oracle's.  pnac compares a fourth
    power with a second power of the input, so (as upstream's own docstring says) it only works at a suitable signal
    level: the capture is scaled by `gain`."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(N, occ, cp, mod)
    rng = np.random.default_rng(17)
    pay = payloads(rng, 8)
    _, xc = loopback_capture(lay, pay, 30, 0.05, seed=31)
    xc = (xc * np.float32(gain)).astype(np.complex64)
    ref = o.rx_demodulate(xc, lay, keep=True, sync=sync, snr_db=30.0)
    eng = OfdmEngine(N, occ, cp, mod)
    bufs = eng.rx_alloc(len(xc), max_frames=max(64, 4 * len(ref.trig) + 64))
    got = eng.collect(eng.demodulate_async(torch.from_numpy(xc).cuda(), bufs, sync=sync, snr_db=30.0))
    if sync == "ml":
        ev_idx, ev_ang = eng.nco_events(bufs, len(xc))
        assert np.array_equal(ev_idx, ref.trig) and np.allclose(ev_ang, ref.ang, atol=2e-4)      # one event per OFDM symbol
        assert len(ev_idx) >= 3 * len(got.trig_idx) > 0       # one NCO event per symbol, one timing trigger per frame
        assert np.array_equal(got.frame_start, ref.frame_start) and np.array_equal(got.frame_ndata, ref.n_data)
    else:
        # indices exact; the angles come from the float32 FFT correlator here and a float64 FIR in the oracle
        assert np.array_equal(got.trig_idx, ref.trig) and np.allclose(got.trig_ang, ref.ang, atol=2e-4)
        assert np.array_equal(got.frame_start, ref.frame_start) and np.array_equal(got.frame_ndata, ref.n_data)
    assert got.packets == ref.packets
    assert sum(1 for ok, _ in got.packets if ok) >= (6 if sync == "ml" or gain == 8.0 else 1)
    # the same through the reference-shaped class (options.sync)
    from types import SimpleNamespace
    from ofdm_uhd_b200 import ofdm
    seen = []
    d = ofdm.ofdm_demod(SimpleNamespace(modulation=mod, fft_length=N, occupied_tones=occ, cp_length=cp, snr=30.0, verbose=False,
                                        log=False, sync=sync), callback=lambda ok, p: seen.append((ok, p)))
    d.feed(xc)
    d.wait(30)
    assert seen == ref.packets
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N,occ,cp,mod,nsym", [(512, 200, 128, "qpsk", 10), (1024, 400, 256, "qam16", 4), (512, 400, 64, "qpsk", 6)])
def test_acquisition_on_silent_frames(N, occ, cp, mod, nsym):
    """ofdm_frame_acquisition on a flagged vector with an all-zero spectrum: no shift of the coarse search has a positive
    correlation, the upstream search index stays 0 (delta = -zero_left), the estimate is 0/0 -- and the block goes on
    equalising the following data vectors with it (digital_swig.py:4316-4330 as restated in oracle.FrameAcquisition).
    Fixed synchroniser, two frames of a back-to-back capture zeroed: the packet list must equal the oracle's (the
    warp-plan acquisition kernel re-parks the spectrum around bin 0 for such a frame)."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(N, occ, cp, mod)
    rng = np.random.default_rng(N + nsym)
    pay = payloads(rng, 10)
    x = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, 0.25, seed=2)
    L = N + cp
    assert len(x) == 10 * nsym * L
    xc = o.channel(np.concatenate([x, np.zeros(3 * L, np.complex64)]), 30.0, 0.1, N, seed=8,
                   sig_power=float(np.mean(np.abs(x) ** 2)))
    for f in (3, 7):
        xc[f * nsym * L:(f + 1) * nsym * L] = 0
    ref = o.rx_demodulate(xc, lay, sync="fixed", nsymbols=nsym, freq_offset=0.1)
    assert sum(1 for ok, _ in ref.packets if ok) == 8
    eng = OfdmEngine(N, occ, cp, mod)
    got = eng.demodulate_fixed(torch.from_numpy(xc).cuda(), nsym, 0.1)
    assert got.packets == ref.packets
