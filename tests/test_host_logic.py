"""Host-side mirror of the reference interface: packet utils against the oracle, option parsing, queue shim,
flat imports, sharding helpers (single process) -- nothing here needs a GPU."""
import optparse
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from oracle import ofdm_oracle as o
from ofdm_uhd_b200 import ofdm_packet_utils as pu, sensing, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packet_utils_equal_oracle():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 402, 1020, 4091):
        p = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        for pad in (False, True):
            if pad and n == 4091:
                continue
            assert pu.make_packet(p, 1, 1, pad) == o.make_packet(p, 1, 1, pad)
        pkt = pu.make_packet(p, 1, 1, False)
        assert pu.unmake_packet(pkt[4:-1]) == (True, p)
        bad = bytearray(pkt[4:-1])
        if bad:
            bad[0] ^= 1
            assert pu.unmake_packet(bytes(bad))[0] is False
    assert pu.unmake_packet(b"abc") == (False, b"")
    assert pu.whiten(pu.whiten(b"hello", 3), 3) == b"hello" and pu.dewhiten(pu.whiten(b"x", 0), 0) == b"x"
    with pytest.raises(ValueError):
        pu.make_packet(bytes(4093), 1, 1, False)
    assert pu.make_packet("abc", 1, 1, False) == pu.make_packet(b"abc", 1, 1, False)   # py2-style str payloads
    assert pu._npadding_bytes(411, 1, 1) == 5 and pu._npadding_bytes(416, 1, 1) == 0


def test_add_options_like_the_reference_scripts():
    pytest.importorskip("torch")
    from ofdm_uhd_b200 import ofdm, transmit_path, receive_path
    parser = optparse.OptionParser(conflict_handler="resolve")
    expert = parser.add_option_group("Expert")
    parser.add_option("", "--snr", type="float", default=30)
    transmit_path.transmit_path.add_options(parser, expert)
    receive_path.receive_path.add_options(parser, expert)
    ofdm.ofdm_mod.add_options(parser, expert)
    ofdm.ofdm_demod.add_options(parser, expert)
    opts, args = parser.parse_args([])
    assert (opts.modulation, opts.fft_length, opts.occupied_tones, opts.cp_length) == ("bpsk", 512, 200, 128)
    assert opts.tx_amplitude == 0.25 and opts.samples_per_symbol == 2 and opts.log is False and opts.verbose is False
    opts, _ = parser.parse_args(["-m", "qam16", "--fft-length", "1024", "--tx-amplitude", "0.5", "-v"])
    assert opts.modulation == "qam16" and opts.fft_length == 1024 and opts.tx_amplitude == 0.5 and opts.verbose


def test_msg_queue_shim():
    pytest.importorskip("torch")
    from ofdm_uhd_b200 import ofdm
    hits = []
    q = ofdm.msg_queue(2, on_full=lambda: hits.append(q.count()) or q.flush())
    for i in range(5):
        q.insert_tail(ofdm.message_from_string(bytes([i])))
    assert hits == [2, 2] and q.count() == 1 and q.delete_head().to_string() == b"\x04" and q.empty_p()
    m = ofdm.message(1)
    assert m.type() == 1 and m.length() == 0


def test_flat_import_like_the_reference_directory():
    code = ("import sys; sys.path.insert(0, %r); import psk, qam, ofdm_packet_utils, ofdm, transmit_path, receive_path;"
            "print(ofdm.ofdm_mod.__name__, transmit_path.transmit_path.__name__, len(ofdm.known_symbols_4512_3))"
            % os.path.join(ROOT, "ofdm_uhd_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["ofdm_mod", "transmit_path", "4512"]


def test_sensing_decision_helpers():
    assert sensing.hex_conv([0, 0, 0, 0] + [1] * 12) == "0FFF"
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, 256).tolist()
    assert sensing.hex_conv(bits) == o.hex_conv(bits)
    avg = rng.random(1024) * 1e-5
    avg[400:430] = 1e-7
    assert sensing.best_band(avg) == o.best_band(avg)
    assert sensing.busy_count([1] * 1024, 905e6, 25e6, 1024) == 0


def test_sharding_helpers():
    assert sharding.streams_of_rank(64, 8, 3) == list(range(3, 64, 8))
    cover = []
    for r in range(3):
        lo, hi = sharding.split_frames(10, 3, r)
        cover += list(range(lo, hi))
    assert cover == list(range(10))
    assert sharding.as_dict(range(8))["crc_ok"] == 2


def test_rendezvous_protocol_state_machines():
    """secondary_tx.py:54-73,345-381 / secondary_rx.py:51-85 without a radio: the receiver follows the announced
    frequency, stores packets 21..70 of the source, and falls back to 920 MHz after 10 unmarked packets."""
    import io
    import struct
    from ofdm_uhd_b200 import rendezvous as rv
    sent = []
    send = lambda payload, eof=False, cmap="FE7F": sent.append(payload)
    assert rv.next_tx_frequency(1, 905000000) == rv.SYNC_FREQ and rv.next_tx_frequency(0, 905000000) == 905000000
    assert rv.synchronization(send, 903700000) == 100
    assert len(sent) == 100 and struct.unpack('!HHL', sent[0]) == (150, 11111, 903700000)
    assert struct.unpack('!H', sent[-1][:2])[0] == 249
    tuned = []
    sink = io.BytesIO()
    rx = rv.secondary_receiver(set_center_freq=tuned.append, sink=sink)
    rx.rx_callback(False, sent[0])                       # bad CRC: counted, not followed
    assert rx.sync == 1 and rx.n_rcvd == 1 and rx.n_right == 0 and not tuned
    for p in sent[1:]:
        rx.rx_callback(True, p)
    assert tuned == [903700000] and rx.sync == 0 and rx.n_right == 99
    source = bytes(range(256)) * 40
    sent.clear()
    n = rv.run_transmitter(send, source, 204)
    npk = len(source) // 204
    assert struct.unpack('!H', sent[20][4:6])[0] == npk
    assert len(sent) == 21 + -(-len(source) // 200) + 20 and n == sum(len(p) for p in sent[:-20])
    for p in sent:
        rx.rx_callback(True, p)
    assert rx.no_packets == npk + 20
    assert sink.getvalue() == source[:50 * 200]          # packets 21..70 only (the reference's pktno window)
    for _ in range(9):
        rx.rx_callback(True, b"\x00\x01\x00\x00junk")
    assert rx.freq == 903700000
    rx.rx_callback(True, b"xy")                          # short payload counts as unmarked
    assert rx.sync == 1 and rx.freq == rv.SYNC_FREQ and tuned[-1] == rv.SYNC_FREQ


def test_usrp_path_shims_options_and_missing_frequency():
    """usrp_transmit_path / usrp_receive_path keep the reference's option surface (-f sets both frequencies,
    usrp_transmit_path.py:28-38) and its SystemExit when no frequency is given (:54-56, usrp_receive_path.py:53-55)."""
    sys.path.insert(0, os.path.join(ROOT, "ofdm_uhd_b200"))
    try:
        import importlib
        utx = importlib.import_module("usrp_transmit_path")
        urx = importlib.import_module("usrp_receive_path")
        ofdm = importlib.import_module("ofdm")
    finally:
        sys.path.pop(0)
    parser = optparse.OptionParser(conflict_handler="resolve")
    expert = parser.add_option_group("Expert")
    utx.add_options(parser, expert)
    urx.add_options(parser, expert)
    ofdm.ofdm_mod.add_options(parser, expert)
    ofdm.ofdm_demod.add_options(parser, expert)
    opts, args = parser.parse_args(["-f", "905e6", "-m", "qpsk", "--tx-amplitude", "0.3"])
    assert opts.tx_freq == 905e6 and opts.rx_freq == 905e6 and opts.modulation == "qpsk" and opts.tx_amplitude == 0.3
    opts2, _ = parser.parse_args([])
    assert opts2.tx_freq is None
    with pytest.raises(SystemExit):
        utx.usrp_transmit_path(opts2)
    with pytest.raises(SystemExit):
        urx.usrp_receive_path(lambda ok, p: None, opts2)


def test_watcher_thread_survives_a_failing_callback():
    """An exception in the user's rx callback must not end the watcher: later packets are still delivered and wait()
    returns."""
    from ofdm_uhd_b200 import ofdm
    seen = []

    def cb(ok, payload):
        if payload == b"boom":
            raise RuntimeError("user callback failed")
        seen.append(payload)

    q = ofdm.msg_queue()
    w = ofdm._queue_watcher_thread(q, cb)
    for p in (b"a", b"boom", b"b"):
        q.insert_tail(ofdm._rx_message(True, p))
    w.drain(timeout=10)
    assert seen == [b"a", b"b"] and w.errors == 1 and w.is_alive()


def test_fft_core_host_emulation(tmp_path):
    """csrc/fft.cuh compiled as plain C++ (OFDM_HOST_EMUL): every plan -- the three-pass Stockham plans and the
    radix-32 warp plans -- against a float64 DFT, forward and inverse (replaces gr.fft_vcc, ofdm.py:112,
    ofdm_receiver.py~:126)."""
    import os, shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = os.path.join(os.path.dirname(__file__), "host_emul", "fft_check.cpp")
    exe = str(tmp_path / "fft_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", src, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "warp plan N=1024" in r.stdout


def test_crc_slicing_host_emulation(tmp_path):
    """crc32_step4 (csrc/common.cuh, four bytes per step over the slicing tables of ofdm_create) equals the
    byte-at-a-time CRC-32 of digital.crc32 (digital_swig.py:3151-3168) on random buffers."""
    import os, shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = os.path.join(os.path.dirname(__file__), "host_emul", "crc_check.cpp")
    exe = str(tmp_path / "crc_check")
    subprocess.run(["g++", "-O2", "-std=c++17", src, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_metric_math_host_emulation(tmp_path):
    """The two arithmetic arguments of the metric kernels (ofdm_sync_pn's window sums and |P|^2 / R^2, ofdm_receiver.py~:97-101),
    restated lane by lane on the CPU: the division sequence of fdiv_inrange (common.cuh) is the correctly rounded quotient for
    in-window operands whatever reciprocal within 2 ulp it starts from, and bfly_scan (rx_sync_stream.cu) yields the exact
    exclusive sums of the lanes in front / behind with NaN confinement."""
    import os, shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = os.path.join(os.path.dirname(__file__), "host_emul", "metric_math_check.cpp")
    exe = str(tmp_path / "metric_math_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", src, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "division sequence ok" in r.stdout and "bfly_scan ok" in r.stdout, r.stdout + r.stderr
