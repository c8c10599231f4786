"""The reference-facing Python surface on the GPU: transmit_path.send_pkt -> channel_model -> receive_path
callback, driven the way benchmark_ofdm_tx.py / benchmark_ofdm_rx.py drive the reference."""
import struct
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import ofdm_oracle as o

pytestmark = pytest.mark.gpu


def options(**kw):
    d = dict(modulation="bpsk", fft_length=512, occupied_tones=200, cp_length=128, snr=30, verbose=False, log=False,
             tx_amplitude=0.25, samples_per_symbol=2)
    d.update(kw)
    return SimpleNamespace(**d)


@pytest.mark.parametrize("mod", ["bpsk", "qpsk", "qam16"])
def test_send_pkt_to_rx_callback(mod):
    from ofdm_uhd_b200 import transmit_path, receive_path, channel_model
    opts = options(modulation=mod)
    got = []

    def rx_callback(ok, payload):                                  # benchmark_ofdm_rx.py:50-61
        (pktno,) = struct.unpack("!H", payload[0:2])
        got.append((ok, pktno, payload))

    tx = transmit_path.transmit_path(opts, pad_seed=21)
    rx = receive_path.receive_path(rx_callback, opts)
    lay = o.Layout(512, 200, 128, mod)
    chan = channel_model.channel_model(tx.ofdm_tx._engine, noise_voltage=0.003, frequency_offset=0.2, seed=4,
                                       lead_in=1200, tail=2600)
    caps = []
    tx.connect(chan)
    chan.connect(lambda smp: caps.append(smp.cpu().numpy()))
    chan.connect(rx)
    rng = np.random.default_rng(5)
    sent = []
    for pktno in range(30):
        data = bytes(rng.integers(0, 256, 398, dtype=np.uint8))
        payload = struct.pack("!H", pktno & 0xFFFF) + struct.pack("!H", 0) + data       # benchmark_ofdm_tx.py:117
        sent.append(payload)
        tx.send_pkt(payload, False, "FE7F")                                             # 3 positional args (:60)
    tx.send_pkt(eof=True)
    rx.wait(timeout=60)
    good = [g for g in got if g[0]]
    # the first frame after the CFO step may be lost and a bogus header can swallow followers (C.2): the
    # callbacks must be exactly what the oracle's receiver delivers for the same capture, in the same order
    ref = o.rx_demodulate(np.concatenate(caps), lay)
    assert [(g[0], g[2]) for g in got] == ref.packets
    assert len(good) >= 22 and [g[1] for g in good] == sorted(g[1] for g in good)
    for ok, pktno, payload in good:
        assert payload == sent[pktno]
    # the samples the transmit path produced are the oracle's
    x = tx.ofdm_tx._engine
    assert tx.carrier_map_old == "FE7F" and tx._tx_amplitude == 0.25
    tx.set_tx_amplitude(7.0)
    assert tx._tx_amplitude == 1


def test_constructor_errors():
    from ofdm_uhd_b200 import ofdm
    with pytest.raises(KeyError):
        ofdm.ofdm_mod(options(modulation="qam4"))                  # ofdm.py:92
    with pytest.raises(ValueError):
        ofdm.ofdm_mod(options(occupied_tones=600))                 # upstream std::invalid_argument
    m = ofdm.ofdm_mod(options(modulation="qam8"))                  # "qam8" is accepted (ofdm.py:88)
    assert m._arity == 8
    with pytest.raises(ValueError):
        m.send_pkt(bytes(4093))                                    # ofdm_packet_utils.py:124-126


def test_ofdm_mod_samples_equal_oracle():
    import torch
    from ofdm_uhd_b200 import ofdm
    opts = options(modulation="qpsk")
    m = ofdm.ofdm_mod(opts, msgq_limit=2, pad_for_usrp=True, pad_seed=33)
    pay = [bytes([i]) * (50 + 7 * i) for i in range(6)]
    chunks = []
    m.connect(lambda s: chunks.append(s.cpu().numpy()))
    for p in pay:
        m.send_pkt(p)
    m.send_pkt(eof=True)
    got = np.concatenate(chunks)
    lay = o.Layout(512, 200, 128, "qpsk")
    want = o.tx_modulate([o.make_packet(p, 1, 1, True) for p in pay], lay, 1.0, seed=33)   # ofdm_mod: 1/sqrt(N) only
    assert got.shape == want.shape and float(np.max(np.abs(got - want))) < 1e-5


def test_custom_carrier_map_round_trip():
    """a non-default hex data-carrier mask (the ctor argument of the commented call at
    ofdm.py:103-104, honoured on request by transmit_path.send_pkt) -- samples and decoded packets equal the
    oracle's for the same mask; by default send_pkt ignores the map like the reference."""
    import torch
    from ofdm_uhd_b200 import transmit_path, ofdm
    from ofdm_uhd_b200.engine import OfdmEngine
    cmap = "F00FF00F"
    opts = options(modulation="qpsk")
    lay = o.Layout(512, 200, 128, "qpsk", carrier_map=cmap)
    assert lay.ncar == 184
    rng = np.random.default_rng(8)
    pay = [struct.pack("!HH", i, 0) + bytes(rng.integers(0, 256, 200, dtype=np.uint8)) for i in range(10)]
    # default: map accepted and ignored
    tx = transmit_path.transmit_path(opts, pad_seed=3)
    out = []
    tx.connect(lambda smp: out.append(smp.cpu().numpy()))
    for p in pay:
        tx.send_pkt(p, False, cmap)
    tx.send_pkt(eof=True)
    ref_default = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], o.Layout(512, 200, 128, "qpsk"), 0.25, seed=3)
    assert float(np.max(np.abs(np.concatenate(out) - ref_default))) < 1e-5
    # honoured
    tx = transmit_path.transmit_path(opts, pad_seed=3, honor_carrier_map=True)
    out = []
    tx.connect(lambda smp: out.append(smp.cpu().numpy()))
    for p in pay:
        tx.send_pkt(p, False, cmap)
    tx.send_pkt(eof=True)
    x = np.concatenate(out)
    want = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, 0.25, seed=3)
    assert x.shape == want.shape and float(np.max(np.abs(x - want))) < 1e-5
    lead = np.zeros(700, np.complex64)
    cap = o.channel(np.concatenate([lead, want, np.zeros(2600, np.complex64)]), 30, 0.15, 512, seed=4,
                    sig_power=float(np.mean(np.abs(want) ** 2)))
    ref = o.rx_demodulate(cap, lay)
    eng = OfdmEngine(512, 200, 128, "qpsk", carrier_map=cmap)
    got = eng.demodulate(torch.from_numpy(cap).cuda())
    assert got.packets == ref.packets and sum(1 for g, _ in got.packets if g) >= 8
    with pytest.raises(ValueError):
        OfdmEngine(512, 200, 128, "qpsk", carrier_map="XYZ")


def test_log_option_dumps_stage_taps(tmp_path, monkeypatch):
    """options.log (ofdm.py:123-131,253-254; ofdm_receiver.py~:144-152): raw float32 I/Q dumps under the
    reference's file names; the channel-filter dump equals the oracle's filtered stream."""
    import torch
    from ofdm_uhd_b200 import ofdm
    monkeypatch.chdir(tmp_path)
    lay = o.Layout(512, 200, 128, "bpsk")
    rng = np.random.default_rng(2)
    pay = [struct.pack("!HH", i, 0) + bytes(rng.integers(0, 256, 100, dtype=np.uint8)) for i in range(4)]
    x = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, 0.25, seed=0)
    cap = o.channel(np.concatenate([np.zeros(700, np.complex64), x, np.zeros(2600, np.complex64)]), 35, 0.0, 512, seed=3,
                    sig_power=float(np.mean(np.abs(x) ** 2)))
    got = []
    d = ofdm.ofdm_demod(options(log=True), callback=lambda ok, p: got.append((ok, p)))
    d.feed(cap)
    d.wait(30)
    ref = o.rx_demodulate(cap, lay, keep=True)
    assert got == ref.packets
    y = np.fromfile("ofdm_receiver-chan_filt_c.dat", dtype=np.complex64)
    assert y.shape == ref.y.shape and np.linalg.norm(y - ref.y) / np.linalg.norm(ref.y) < 1e-4
    flags = np.fromfile("ofdm_receiver-found_corr_b.dat", dtype=np.uint8)
    assert np.array_equal(flags, ref.flags)
    eq = np.fromfile("ofdm_receiver-frame_acq_c.dat", dtype=np.complex64).reshape(-1, 200)
    assert eq.shape[0] == len(ref.flags)
    sink = np.fromfile("ofdm_frame_sink_c.dat", dtype=np.complex64).reshape(-1, 200)
    assert sink.shape[0] == len(ref.derot) and np.linalg.norm(sink[:, :198] - np.array(ref.derot)) / np.linalg.norm(np.array(ref.derot)) < 1e-4
    # the taps between the channel filter and the frame acquisition (ofdm_receiver.py~:145,148-151)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    plan = o.plan_frames(ref.trig, ref.ang, len(cap), 512, 640)
    ph = o.nco_phase_at(plan, np.arange(len(cap)), 512)
    nco_ref = (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)
    nco = np.fromfile("ofdm_receiver-nco_c.dat", dtype=np.complex64)
    sigmix = np.fromfile("ofdm_receiver-sigmix_c.dat", dtype=np.complex64)
    # (the NCO integrates the latched angle, itself a function of y (2e-7 from the oracle's float64 FIR), over the
    # whole capture: a relative 1e-5 at the end of 18 660 samples is the size of that, not a kernel tolerance)
    assert nco.shape == nco_ref.shape and rel(nco, nco_ref) < 3e-5 and rel(sigmix, ref.y * nco_ref) < 1e-4
    samp = np.fromfile("ofdm_receiver-sampler_c.dat", dtype=np.complex64).reshape(-1, 512)
    samp_ref = np.array([o.derotate(ref.y, st + np.arange(512), plan, 512) for st in ref.vec_start])
    assert samp.shape == samp_ref.shape and rel(samp, samp_ref) < 1e-4
    fft = np.fromfile("ofdm_receiver-fft_out_c.dat", dtype=np.complex64).reshape(-1, 512)
    fft_ref = np.array([o._fft_shift(v) for v in samp_ref])
    assert fft.shape == fft_ref.shape and rel(fft, fft_ref) < 1e-4
    assert rel(eq, ref.eq) < 1e-4
    # transmit side (ofdm.py:123-131): mapper output, the stream behind insert_preamble, IFFT output, samples
    mod = ofdm.ofdm_mod(options(log=True), pad_for_usrp=False, pad_seed=0)
    for p in pay:
        mod.send_pkt(p)
    out = mod.flush().cpu().numpy()
    pk = [o.make_packet(p, 1, 1, False) for p in pay]
    X = [o.tx_symbols_freq(q, lay, f, 0) for f, q in enumerate(pk)]
    pre = np.fromfile("ofdm_preambles.dat", dtype=np.complex64).reshape(-1, 512)
    assert np.array_equal(pre, np.concatenate(X))
    mapper = np.fromfile("ofdm_mapper_c.dat", dtype=np.complex64).reshape(-1, 512)
    assert np.array_equal(mapper, np.concatenate([x_[1:] for x_ in X]))
    ifft = np.fromfile("ofdm_ifft_c.dat", dtype=np.complex64).reshape(-1, 512)
    assert rel(ifft, o._ifft_unnorm(np.concatenate(X))) < 1e-5
    cpa = np.fromfile("ofdm_cp_adder_c.dat", dtype=np.complex64)
    assert np.array_equal(cpa, out) and rel(out, o.tx_modulate(pk, lay, 1.0, seed=0)) < 1e-5


def test_rendezvous_over_the_modem():
    """The 920 MHz rendezvous of secondary_tx.py:54-73 / secondary_rx.py:51-85 carried by the GPU modem in loopback:
    synchronisation packets announce a frequency, the receiver state machine follows it."""
    from ofdm_uhd_b200 import transmit_path, receive_path, channel_model, rendezvous as rv
    opts = options(modulation="qpsk")
    tuned = []
    state = rv.secondary_receiver(set_center_freq=tuned.append)
    tx = transmit_path.transmit_path(opts, pad_seed=3)
    rx = receive_path.receive_path(state.rx_callback, opts)
    chan = channel_model.channel_model(tx.ofdm_tx._engine, noise_voltage=0.003, frequency_offset=0.1, seed=9,
                                       lead_in=1200, tail=2600)
    tx.connect(chan)
    chan.connect(rx)
    assert rv.synchronization(tx.send_pkt, 903700000) == 100
    tx.send_pkt(eof=True)
    rx.wait(timeout=60)
    assert tuned == [903700000] and state.sync == 0 and state.n_right >= 95 and state.n_rcvd >= state.n_right


def test_collect_begin_end_equals_collect():
    """The split (asynchronous) collect returns what collect() returns, also with two receive calls in flight on
    two streams and two buffer sets."""
    import torch
    from ofdm_uhd_b200.engine import OfdmEngine
    from helpers import payloads, loopback_capture
    lay = o.Layout(512, 200, 128, "qpsk")
    rng = np.random.default_rng(8)
    caps = [loopback_capture(lay, payloads(rng, 12), 25, cfo, seed=40 + i)[1] for i, cfo in enumerate((0.1, -0.3))]
    eng = OfdmEngine(512, 200, 128, "qpsk")
    refs = []
    for c in caps:
        r = eng.demodulate(torch.from_numpy(c).cuda())
        r.payload_rows = r.payload_rows.copy()             # views of pinned staging: the next collect reuses it
        refs.append(r)
    lanes = []
    for c in caps:
        lanes.append((torch.cuda.Stream(), torch.from_numpy(c).cuda(), eng.rx_alloc(len(c), fresh=True)))
    torch.cuda.synchronize()
    tickets = []
    for st, x, bufs in lanes:
        with torch.cuda.stream(st):
            eng.demodulate_async(x, bufs)
            tickets.append(eng.collect_begin(bufs))
    for t, ref in zip(tickets, refs):
        r = eng.collect_end(t)
        assert r.n_frames == ref.n_frames and np.array_equal(r.pkt_ok, ref.pkt_ok) and np.array_equal(r.pkt_len, ref.pkt_len)
        assert np.array_equal(r.frame_live, ref.frame_live) and np.array_equal(r.counters, ref.counters)
        assert np.array_equal(r.msg_frames, ref.msg_frames)
        for f in r.msg_frames:
            ln = int(r.pkt_len[f])
            assert r.payload_rows[f, :ln].tobytes() == ref.payload_rows[f, :ln].tobytes()
    eng.close()


@pytest.mark.parametrize("chunk", [50_000, 131_072, 400_001])
def test_feed_stream_equals_whole_stream(chunk):
    """ofdm_demod.feed_stream on consecutive buffers of one capture delivers what feed() delivers on the whole capture:
    every packet once, in order, wherever the buffer boundaries cut the frames."""
    from helpers import payloads, loopback_capture
    from ofdm_uhd_b200 import receive_path
    lay = o.Layout(512, 200, 128, "qpsk")
    rng = np.random.default_rng(3)
    pay = payloads(rng, 150)
    _, xc = loopback_capture(lay, pay, 28, 0.12, seed=15)
    opts = options(modulation="qpsk")
    whole, parts = [], []
    rx1 = receive_path.receive_path(lambda ok, p: whole.append((ok, p)), opts)
    rx1.feed(xc)
    rx1.wait(timeout=60)
    rx2 = receive_path.receive_path(lambda ok, p: parts.append((ok, p)), opts)
    for a in range(0, len(xc), chunk):
        rx2.feed_stream(xc[a:a + chunk])
    rx2.wait(timeout=60)
    assert len(whole) >= 145
    good_w = [p for ok, p in whole if ok]
    good_p = [p for ok, p in parts if ok]
    assert good_p == good_w                                  # same good packets, same order, none twice
    assert len(parts) == len(whole)
    # radio-sized buffers queued into batches: one receiver pass per 300 000 new samples, then a flush for the rest
    batched = []
    rx3 = receive_path.receive_path(lambda ok, p: batched.append((ok, p)), opts)
    rx3.ofdm_rx.stream_batch_samples = 300_000
    passes = 0
    for a in range(0, len(xc), 4096):
        passes += rx3.feed_stream(xc[a:a + 4096]) is not None
    passes += rx3.flush_stream() is not None
    rx3.wait(timeout=60)
    assert 2 <= passes <= len(xc) // 300_000 + 1
    assert [p for ok, p in batched if ok] == good_w and len(batched) == len(whole)
    # the same through the dense hand-over (rx_callback_batch)
    dense = []
    rx4 = receive_path.receive_path(None, opts)
    rx4.set_batch_callback(lambda ok, data, off: dense.extend((bool(ok[k]), data[off[k]:off[k + 1] - 4].tobytes()) for k in range(len(ok))))
    rx4.ofdm_rx.stream_batch_samples = 250_000
    for a in range(0, len(xc), chunk):
        rx4.feed_stream(xc[a:a + chunk])
    rx4.flush_stream()
    assert dense == whole


def test_benchmark_loopback_example_script(tmp_path):
    """examples/benchmark_ofdm_loopback.py drives the flat-imported modules (import ofdm, transmit_path, receive_path)
    the way benchmark_ofdm_tx.py / benchmark_ofdm_rx.py drive the reference's: a file goes through the modem and
    comes back byte for byte."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "tx.bin"
    data = bytes(np.random.default_rng(1).integers(0, 256, 30000, dtype=np.uint8))
    src.write_bytes(data)
    spec = importlib.util.spec_from_file_location("bench_loop", os.path.join(root, "examples", "benchmark_ofdm_loopback.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = tmp_path / "rx.bin"
    pktno, n_rcvd, n_right, sent, got = mod.main(["-m", "qpsk", "-s", "402", "-M", "0.01", "--snr", "30", "--cfo", "0.1",
                                                   "--from-file", str(src), "--to-file", str(out)])
    assert pktno == 20 + -(-len(data) // 400) and sent == data
    assert n_right >= pktno - 2 and out.read_bytes() == got
    # the first frame after the CFO step may be lost (C.2): it is one of the 20 filler packets, so the file is whole
    assert got == data


def test_secondary_user_example_script():
    """examples/secondary_loopback.py: sense -> hop decision -> 920 MHz rendezvous -> file on the new frequency, the loop
    of secondary_tx.py / secondary_rx.py, with every numeric stage on the GPU."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("sec_loop", os.path.join(root, "examples", "secondary_loopback.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    new_freq, tuned, data, got, state = mod.main(seed=3, verbose=False)
    assert tuned == [new_freq] and new_freq % 100000 == 0 and abs(new_freq - 905e6) > 1e6     # away from the primary
    # the receiver stores packets 21..70 (the reference's pktno window, secondary_rx.py:58-75): the 40 file packets, then
    # the first ten of the transmitter's trailing filler packets -- as the reference pair would
    assert got[:len(data)] == data and got[len(data):] == b"This is also Garbage data" * 10
    assert state.n_right >= state.n_rcvd - 3


def test_usrp_path_shims_loopback_and_retune():
    """The scripts' own structure -- tb.txpath = usrp_transmit_path(options), tb.rxpath = usrp_receive_path(cb, options),
    tb.txpath.send_pkt(payload, eof, carrier_map), tb.rxpath.u.u.set_center_freq(freq, 0) -- over the loop-back medium:
    packets arrive while both ends are on the same frequency and stop arriving after the receiver retunes."""
    from ofdm_uhd_b200 import usrp_transmit_path, usrp_receive_path, loopback_air
    loopback_air.AIR.reset(noise_voltage=0.003, frequency_offset=0.1, seed=5)
    opts = options(modulation="qpsk", tx_freq=905e6, rx_freq=905e6)
    got = []
    rxpath = usrp_receive_path.usrp_receive_path(lambda ok, p: got.append((ok, p)), opts)
    txpath = usrp_transmit_path.usrp_transmit_path(opts, pad_seed=1)
    pay = [struct.pack("!HH", i, 0) + bytes([i]) * 100 for i in range(12)]
    for p in pay:
        txpath.send_pkt(p, False, "FE7F")
    txpath.send_pkt(eof=True)
    rxpath.wait(timeout=60)
    good = [p for ok, p in got if ok]
    assert len(good) >= 11 and all(p in pay for p in good)
    rxpath.u.u.set_center_freq(920e6, 0)                     # secondary_rx.py:85
    n0 = len(got)
    for p in pay:
        txpath.send_pkt(p, False, "FE7F")
    txpath.flush()
    rxpath.wait(timeout=10)
    assert len(got) == n0 and rxpath.u.history == [905e6, 920e6]


def test_bulk_send_pkts_and_batch_callback_equal_per_packet_surface():
    """transmit_path.send_pkts (make_packets_kernel on the device) produces the samples a loop of send_pkt calls
    produces, and receive_path.set_batch_callback hands over, in one call, exactly the (ok, payload) sequence the
    per-packet rx_callback sees."""
    import torch
    from ofdm_uhd_b200 import transmit_path, receive_path, channel_model
    opts = options(modulation="qpsk")
    rng = np.random.default_rng(12)
    sent = [struct.pack("!HH", k, 0) + bytes(rng.integers(0, 256, int(rng.integers(0, 700)), dtype=np.uint8)) for k in range(40)]
    outs = {}
    for mode in ("loop", "bulk"):
        tx = transmit_path.transmit_path(opts, pad_seed=3)
        tx.connect(lambda smp, mode=mode: outs.setdefault(mode, smp.cpu().numpy()))
        if mode == "loop":
            for p in sent:
                tx.send_pkt(p, False, "FE7F")
            tx.send_pkt(eof=True)
        else:
            tx.send_pkts(sent)
    assert np.array_equal(outs["loop"], outs["bulk"])
    with pytest.raises(ValueError):
        tx.send_pkts([b"x" * 4092])
    # receive side
    eng = tx.ofdm_tx._engine
    chan = channel_model.channel_model(eng, noise_voltage=0.004, frequency_offset=0.15, seed=2, lead_in=1300, tail=2600)
    cap = chan.process(torch.from_numpy(outs["bulk"]).cuda())
    per_packet, bulk = [], []
    rx = receive_path.receive_path(lambda ok, payload: per_packet.append((ok, payload)), opts)
    rx.feed(cap)
    rx.wait(timeout=60)
    rx.set_batch_callback(lambda ok, data, off: bulk.extend((bool(ok[k]), data[off[k]:off[k + 1] - 4].tobytes()) for k in range(len(ok))))
    rx.feed(cap)
    assert bulk == per_packet and sum(1 for ok, _ in bulk if ok) >= 36
    rx.set_batch_callback(None)
    per_packet.clear()
    rx.feed(cap)
    rx.wait(timeout=60)
    assert per_packet == bulk
