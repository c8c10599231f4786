"""Many independent streams in one call (ofdm_tx_modulate_streams / ofdm_rx_demodulate_batch): every stream of the
batch gets what a call of its own gets, and what the oracle delivers (the reference runs one flowgraph per stream,
benchmark_ofdm_rx.py:42-87; BASELINE configs[2] is 64 of them)."""
import numpy as np
import pytest

from oracle import ofdm_oracle as o
from helpers import payloads, rel_l2, loopback_capture

pytestmark = pytest.mark.gpu


def _pad4(n):
    return (n + 3) // 4 * 4


@pytest.mark.parametrize("N,occ,cp,mod,snr", [(512, 200, 128, "qpsk", 25), (1024, 400, 256, "qam64", 30),
                                               (128, 56, 32, "qpsk", 30), (4096, 3200, 512, "qam256", 38)])
def test_rx_batch_equals_single_calls_and_oracle(N, occ, cp, mod, snr):
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(N, occ, cp, mod)
    eng = OfdmEngine(N, occ, cp, mod)
    rng = np.random.default_rng(N)
    nfr = [5, 1, 9, 0, 3] if N <= 1024 else [3, 0, 2]
    caps = []
    for s, k in enumerate(nfr):
        if k == 0:                                   # a stream of noise only
            caps.append((rng.standard_normal(3 * lay.sym_len + 5) * 0.01 + 1j * rng.standard_normal(3 * lay.sym_len + 5) * 0.01).astype(np.complex64))
            continue
        _, xc = loopback_capture(lay, payloads(rng, k), snr, float(rng.uniform(-0.4, 0.4)), seed=50 + s,
                                 lead=N + 37 + 11 * s, tail=4 * lay.sym_len + 3 * s)
        caps.append(xc)
    # ragged lengths, stream starts padded to multiples of 4 samples (gap samples belong to no stream ... they are
    # the tail of the previous one, so pad with that stream's own noise floor instead: keep offsets exact)
    off = np.zeros(len(caps) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in caps], out=off[1:])
    x = torch.from_numpy(np.concatenate(caps)).cuda()
    bufs = eng.rx_alloc_batch(off, max_frames=64)
    got = eng.collect_batch(eng.demodulate_batch_async(x, bufs))
    assert len(got) == len(caps)
    for s, c in enumerate(caps):
        one = eng.demodulate(torch.from_numpy(c).cuda())
        assert got[s].packets == one.packets, "stream %d" % s
        assert np.array_equal(got[s].trig_idx, one.trig_idx) and np.array_equal(got[s].trig_ang, one.trig_ang)
        assert np.array_equal(got[s].frame_start, one.frame_start) and np.array_equal(got[s].frame_ndata, one.frame_ndata)
        assert np.array_equal(got[s].counters, one.counters)
        if N <= 1024 or s == 0:
            assert got[s].packets == o.rx_demodulate(c, lay).packets, "stream %d vs oracle" % s
    assert sum(len(g.packets) for g in got) >= sum(nfr) // 2        # a bogus first header may swallow followers (C.2); the oracle agrees
    # second run on the same buffers with the streams in another order: nothing leaks from one stream into the next
    order = list(reversed(range(len(caps))))
    off2 = np.zeros(len(caps) + 1, dtype=np.int64)
    np.cumsum([len(caps[s]) for s in order], out=off2[1:])
    x2 = torch.from_numpy(np.concatenate([caps[s] for s in order])).cuda()
    got2 = eng.collect_batch(eng.demodulate_batch_async(x2, eng.rx_alloc_batch(off2, max_frames=64)))
    for k, s in enumerate(order):
        assert got2[k].packets == got[s].packets
    eng.close()


def test_rx_batch_rejects_layout_without_streaming_sync():
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    eng = OfdmEngine(64, 48, 16, "bpsk")
    off = np.array([0, 4000, 8000], dtype=np.int64)
    bufs = eng.rx_alloc_batch(off, max_frames=16)
    with pytest.raises(RuntimeError, match="streaming synchroniser"):
        eng.demodulate_batch_async(torch.zeros(8000, dtype=torch.complex64, device="cuda"), bufs)
    eng.close()


@pytest.mark.parametrize("N,occ,cp,mod", [(512, 200, 128, "8psk"), (1024, 400, 256, "qam64")])
def test_tx_streams_equal_oracle(N, occ, cp, mod):
    """ofdm_tx_modulate_streams: ragged packets of three streams in one launch, each stream written at its own
    offset, frames numbered per stream, pad symbols from pad_seed + s."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(N, occ, cp, mod)
    eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=9)
    rng = np.random.default_rng(4)
    per = [[bytes(rng.integers(0, 256, int(k), dtype=np.uint8)) for k in rng.integers(0, 900, size=cnt)] for cnt in (4, 1, 6)]
    flat = [p for st in per for p in st]
    poff = np.zeros(len(flat) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in flat], out=poff[1:])
    ref = [o.tx_modulate([o.make_packet(p, 1, 1, False) for p in st], lay, 0.25, seed=9 + s) for s, st in enumerate(per)]
    gap = 1000
    out_off, pos = [], 64
    for r in ref:
        out_off.append(pos)
        pos += len(r) + gap
    sf = np.concatenate([[0], np.cumsum([len(st) for st in per])])
    plan = eng.tx_plan(poff, stream_frame0=sf, stream_out_off=np.array(out_off))
    out = torch.zeros(pos, dtype=torch.complex64, device="cuda")
    raw = torch.from_numpy(np.frombuffer(b"".join(flat), dtype=np.uint8).copy()).cuda()
    eng.tx_run(plan, raw, out=out)
    torch.cuda.synchronize()
    g = out.cpu().numpy()
    mask = np.ones(pos, dtype=bool)
    for s, r in enumerate(ref):
        seg = g[out_off[s]:out_off[s] + len(r)]
        assert rel_l2(seg, r) < 1e-4, "stream %d" % s
        mask[out_off[s]:out_off[s] + len(r)] = False
    assert not g[mask].any()                          # nothing written between the streams
    eng.close()


@pytest.mark.parametrize("batch", [False, True])
def test_dense_delivery_equals_collect(batch):
    """ofdm_rx_compact: the dense hand-over (messages back to back + bit-packed verdicts) delivers exactly what
    collect() assembles from the per-frame slots, for a single stream and for a batch, ragged payloads and a bad CRC
    included; with an expected size too small the rest is reported as overflow."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    lay = o.Layout(512, 200, 128, "qam16")
    eng = OfdmEngine(512, 200, 128, "qam16", max_pkt_bytes=1024)
    rng = np.random.default_rng(8)
    caps = []
    for s in range(3 if batch else 1):
        pay = [bytes(rng.integers(0, 256, int(k), dtype=np.uint8)) for k in rng.integers(0, 1000, size=7 + s)]
        pk = [o.make_packet(p, 1, 1, False) for p in pay]
        bad = bytearray(pk[3]); bad[9] ^= 0x40; pk[3] = bytes(bad)                       # CRC must fail for this one
        x = o.tx_modulate(pk, lay, 0.25, seed=1)
        xin = np.concatenate([np.zeros(700, np.complex64), x, np.zeros(3000, np.complex64)])
        caps.append(o.channel(xin, 35.0, 0.1 * s, 512, seed=60 + s, sig_power=float(np.mean(np.abs(x) ** 2))))
    if batch:
        off = np.zeros(len(caps) + 1, dtype=np.int64)
        np.cumsum([len(c) for c in caps], out=off[1:])
        bufs = eng.rx_alloc_batch(off, max_frames=32)
        eng.demodulate_batch_async(torch.from_numpy(np.concatenate(caps)).cuda(), bufs)
        ref = [p for r in eng.collect_batch(bufs) for p in r.packets]
    else:
        bufs = eng.rx_alloc(len(caps[0]))
        eng.demodulate_async(torch.from_numpy(caps[0]).cuda(), bufs)
        ref = eng.collect(bufs).packets
    got, info = eng.deliver(bufs)
    assert got == ref and len(ref) >= 6 and not all(ok for ok, _ in ref) and info["overflow"] == 0
    assert info["bytes_device"] == sum(len(p) + 4 for _, p in ref)
    few, info2 = eng.deliver(bufs, expect_msgs=4, expect_bytes=1 << 20)
    assert few == ref[:4] and info2["overflow"] == len(ref) - 4
    eng.close()
