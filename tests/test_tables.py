"""The only data in the reference tree that pin tables exactly, checked against the
oracle AND the product modules: whitening mask, known symbols, constellations, header layout, docstring examples."""
import hashlib
import math

import numpy as np
import pytest

from oracle import ofdm_oracle as o
from ofdm_uhd_b200 import ofdm_packet_utils as pu, psk, qam


def test_whitening_mask_matches_reference(golden):
    assert np.array_equal(o.whitening_mask(), golden["mask"])
    assert np.array_equal(pu.random_mask_vec8, golden["mask"])
    assert pu.random_mask_tuple[:8] == (255, 63, 0, 16, 0, 12, 0, 5)
    assert pu.random_mask_tuple[-8:] == (88, 90, 186, 187, 51, 51, 255, 63)
    assert hashlib.sha256(bytes(golden["mask"])).hexdigest() == \
        "330ef53e31fbdc2de2c91c8fee3e89a6308bd68c31a63e4dbe77f3b4a0297e80"


def test_known_symbols_match_reference(golden):
    assert np.array_equal(o.known_symbols_4512(), golden["known"].astype(np.int32))
    pytest.importorskip("torch")
    from ofdm_uhd_b200 import ofdm
    assert ofdm.known_symbols_4512_3 == golden["known"].astype(int).tolist()


def test_constellations_match_reference(golden):
    for m in (2, 4, 8):
        assert np.allclose(o.psk_gray_constellation(m), golden["psk_gray_%d" % m], atol=1e-15)
        assert np.allclose(psk.gray_constellation[m], golden["psk_gray_%d" % m], atol=1e-15)
        assert np.allclose(psk.constellation[m], golden["psk_%d" % m], atol=1e-15)
    for m in (4, 8, 16, 64, 256):
        assert np.allclose(o.qam_constellation(m), golden["qam_%d" % m], atol=1e-15)
        assert np.allclose(qam.constellation[m], golden["qam_%d" % m], atol=1e-15)
    assert psk.binary_to_gray[8] == [0, 1, 3, 2, 7, 6, 4, 5] and psk.gray_to_binary[8] == [0, 1, 3, 2, 6, 7, 5, 4]
    assert qam.binary_to_gray[16] == list(range(16)) and 256 not in qam.binary_to_ungray


def test_qpsk_rotation_literal():
    c = o.constellation_for("qpsk")
    assert abs(abs(c[0]) - math.hypot(0.707, 0.707)) < 1e-7          # 0.707+0.707j, not exp(j*pi/4) (C.6)
    with pytest.raises(KeyError):
        o.constellation_for("qam4")                                   # ofdm.py:92
    assert len(o.constellation_for("qam8")) == 8
    for mod, p in (("qam16", 10 / 9), ("qam64", 6 / 7)):
        assert abs(float(np.mean(np.abs(o.constellation_for(mod).astype(np.complex128)) ** 2)) - p) < 1e-6


def test_make_header_matches_reference(golden):
    for (ln, off), want in zip(golden["hdr_in"].tolist(), golden["hdr_out"]):
        assert o.make_header(ln, off) == bytes(want)
        assert pu.make_header(ln, off) == bytes(want)


def test_crc32_known_answer():
    assert o.crc32_gr(b"123456789") == 0xFC891918                    # CRC-32/BZIP2 check value (A.1)
    assert pu.crc32(b"123456789") == 0xFC891918
    assert o.crc32_gr(b"") == 0 and pu.crc32(b"") == 0


def test_crc32_mirrored_zlib_equals_table_form():
    """pu.crc32 computes digital.crc32 (digital_swig.py:3151-3168) as zlib's CRC-32 on bit-mirrored bytes; it must equal the
    byte-at-a-time table form and the oracle on ragged random buffers (0 .. 4095 bytes)."""
    rng = np.random.default_rng(5)
    for n in list(range(0, 40)) + [255, 256, 257, 402, 1500, 4091, 4095]:
        for _ in range(8):
            b = bytes(rng.integers(0, 256, n, dtype=np.uint8))
            assert pu.crc32(b) == pu._crc32_table(b) == o.crc32_gr(b), n


def test_docstring_examples():
    assert pu.conv_packed_binary_string_to_1_0_string(b"\xAF") == "10101111"       # ofdm_packet_utils.py:29
    assert pu.conv_1_0_string_to_packed_binary_string("10101111") == (b"\xAF", False)  # :42
    assert pu.conv_1_0_string_to_packed_binary_string("101") == (b"\x05", True)
    with pytest.raises(ValueError):
        pu.conv_1_0_string_to_packed_binary_string("12")
    assert pu.is_1_0_string("0101") and not pu.is_1_0_string(b"01") and not pu.is_1_0_string("2")
    assert pu.string_to_hex_list(b"\x01\xff") == ["0x1", "0xff"]
    assert o.hex_conv([0, 0, 0, 0] + [1] * 12) == "0FFF"                            # final_hex_conv.py:37-39


def test_carrier_map_and_layout_constants():
    lay = o.Layout(512, 200, 128, "bpsk")
    assert lay.zl == 156 and lay.ncar == 198
    assert sorted(set(range(156, 356)) - set(lay.tx_map.tolist())) == [255, 256]
    assert sorted(set(range(200)) - set(lay.sink_map.tolist())) == [99, 100]
    taps = o.chan_filter_taps(lay)
    assert len(taps) == 155 and abs(float(taps.sum()) - 1) < 1e-6 and np.allclose(taps, taps[::-1])
    assert len(o.chan_filter_taps(o.Layout(4096, 3200, 512, "qam256"))) == 77
    assert len(o.chan_filter_taps(o.Layout(1024, 800, 256, "qam64"))) == 77
    # Appendix B: symbols per frame for the 402-byte payload (411-byte packet)
    for mod, nd in (("bpsk", 17), ("qpsk", 9), ("qam16", 5)):
        assert o.Layout(512, 200, 128, mod).n_data_syms(411) == nd
    assert o.Layout(1024, 400, 256, "qam64").n_data_syms(411) == 2
    assert o.Layout(4096, 3200, 512, "qam256").n_data_syms(4100) == 2
    # known symbol: odd absolute bins are zero
    assert all(lay.ks[i] == 0 for i in range(200) if (156 + i) & 1) and all(abs(lay.ks[i]) == 1 for i in range(0, 200, 2))


def test_blackmanharris_window():
    w = o.blackmanharris(1024)
    assert w.dtype == np.float32 and abs(float(w.max()) - 1.0) < 1e-3 and float(w.min()) < 1e-4
    # the (i+0.5)/(N-1) argument of GNU Radio 3.x makes it asymmetric by one sample (A.13)
    assert np.allclose(w[:-1], w[:-1][::-1], atol=1e-6) and int(np.argmax(w)) in (510, 511, 512)


def test_sense_decision_equals_reference_console_logs():
    """tests/golden/reference_sense_logs.npz holds what the reference's own sense_loop printed on the authors' radio
    (output.txt, output_with_detection.txt: 32 sweeps x 256 bins of frequency / 10-dwell average / free flag, and the
    carrier map hex_conv() made of them).  The oracle's decision half and the host mirror must reproduce every flag,
    every hex string and every printed bin frequency."""
    import os
    from ofdm_uhd_b200 import sensing
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sense_logs.npz"))
    avg, flag, hexes = d["avg"], d["flag"], d["hex"]
    assert avg.shape == (32, 256) and float(np.min(np.abs(avg - 1e-4))) > 1e-7        # nothing sits on the threshold
    for s in range(len(hexes)):
        dwell = np.zeros((1, 256), dtype=np.float32)
        dwell[0, (np.arange(256) + 128) % 256] = avg[s].astype(np.float32)           # frequency order -> FFT order
        a, free, hx = o.sense_decide(dwell, 1e-4)                                     # threshold of that script variant
        assert np.array_equal(free, flag[s]) and hx == str(hexes[s])
        assert sensing.hex_conv(list(flag[s])) == str(hexes[s]) and o.hex_conv(list(flag[s])) == str(hexes[s])
    # Python 2 printed the bin frequencies with str(float) = 12 significant digits
    for i in range(256):
        t = "%.12g" % sensing.sensed_frequency(900e6, 100e6 / 16, 256, i)
        if "." not in t and "e" not in t:
            t += ".0"
        assert t == str(d["freq_txt"][0][i])
