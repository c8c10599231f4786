"""Packet framing with the function names and semantics of the reference's ofdm_packet_utils.py
(/root/reference/ofdm_packet_utils.py:27-191), for Python 3 ``bytes``.

This is the single-packet host API (what ``send_pkt`` and the rx watcher use for one packet at a
time); batches go through the CUDA kernels behind ``ofdm_make_packets`` / ``ofdm_rx_finish``, which
produce byte-identical results (tests/test_gpu_tx.py)."""
import math
import struct
import zlib

import numpy

__all__ = ["conv_packed_binary_string_to_1_0_string", "conv_1_0_string_to_packed_binary_string", "is_1_0_string",
           "string_to_hex_list", "whiten", "dewhiten", "make_header", "make_packet", "unmake_packet",
           "random_mask_tuple", "random_mask_vec8", "crc32", "gen_and_append_crc32", "check_crc32"]


def _as_bytes(s):
    if isinstance(s, str):
        return s.encode("latin-1")
    return bytes(s)


def conv_packed_binary_string_to_1_0_string(s):
    """b'\\xAF' --> '10101111' (reference :27-38)."""
    return "".join(format(b, "08b") for b in _as_bytes(s))


def is_1_0_string(s):
    """reference :72-78"""
    return isinstance(s, str) and all(ch in "01" for ch in s)


def conv_1_0_string_to_packed_binary_string(s):
    """'10101111' -> (b'\\xAF', False); the flag says whether leading zeros were added (reference :40-69)."""
    if not is_1_0_string(s):
        raise ValueError("Input must be a string containing only 0's and 1's")
    padded = len(s) % 8 != 0
    if padded:
        s = "0" * (8 - len(s) % 8) + s
    return bytes(int(s[i:i + 8], 2) for i in range(0, len(s), 8)), padded


def string_to_hex_list(s):
    """reference :80-81"""
    return [hex(b) for b in _as_bytes(s)]


def _pn15_mask():
    """The 4096-byte whitening table (reference :194-451).  Its comment says "output of a 15-bit LFSR":
    it is PN15 (x^15 + x^14 + 1) started from fourteen ones, packed LSB first for 4094 bytes; the
    table's last two bytes repeat its first two.  tests/test_tables.py pins the SHA-256."""
    nbits = 4094 * 8
    bits = numpy.zeros(nbits, dtype=numpy.uint8)
    bits[:14] = 1
    for n in range(15, nbits):
        bits[n] = bits[n - 14] ^ bits[n - 15]
    table = numpy.packbits(bits, bitorder="little")
    return numpy.concatenate([table, table[:2]])


random_mask_vec8 = _pn15_mask()
random_mask_tuple = tuple(int(v) for v in random_mask_vec8)

_CRC_POLY = 0x04C11DB7
_CRC_TABLE = []
for _i in range(256):
    _c = _i << 24
    for _ in range(8):
        _c = ((_c << 1) ^ _CRC_POLY) & 0xFFFFFFFF if _c & 0x80000000 else (_c << 1) & 0xFFFFFFFF
    _CRC_TABLE.append(_c)


def _crc32_table(data):
    """The byte-at-a-time table form (what digital.crc32 does); crc32() below must equal it (tests/test_tables.py)."""
    reg = 0xFFFFFFFF
    for b in _as_bytes(data):
        reg = (_CRC_TABLE[(b ^ (reg >> 24)) & 0xFF] ^ (reg << 8)) & 0xFFFFFFFF
    return reg ^ 0xFFFFFFFF


_REV8 = bytes(int("{:08b}".format(_i)[::-1], 2) for _i in range(256))


def crc32(data):
    """gnuradio digital.crc32 (reference digital_swig.py:3151-3168): register initialised to all ones,
    MSB-first polynomial 0x04C11DB7, transmitted value is the one's complement.  That is zlib's CRC-32 with the bit
    order mirrored (same polynomial, same all-ones init / final xor): mirror every input byte, take zlib.crc32 (C speed
    instead of a Python loop per byte: the per-packet send_pkt / rx_callback surface spends its time here), mirror the
    32-bit result."""
    z = zlib.crc32(_as_bytes(data).translate(_REV8)) & 0xFFFFFFFF
    return int.from_bytes(z.to_bytes(4, "little").translate(_REV8), "big")


def gen_and_append_crc32(s):
    s = _as_bytes(s)
    return s + struct.pack("!I", crc32(s))


def check_crc32(s):
    s = _as_bytes(s)
    if len(s) < 4:
        return False, b""
    body = s[:-4]
    return crc32(body) == struct.unpack("!I", s[-4:])[0], body


def whiten(s, o):
    """XOR with the PN table starting at offset ``o`` (reference :84-87)."""
    sa = numpy.frombuffer(_as_bytes(s), dtype=numpy.uint8)
    return (sa ^ random_mask_vec8[o:len(sa) + o]).tobytes()


def dewhiten(s, o):
    """reference :89-90 (self inverse)"""
    return whiten(s, o)


def make_header(payload_len, whitener_offset=0):
    """Offset in the upper nibble, length in the lower 12 bits, sent twice (reference :93-97)."""
    val = ((whitener_offset & 0xF) << 12) | (payload_len & 0x0FFF)
    return struct.pack("!HH", val, val)


def _npadding_bytes(pkt_byte_len, samples_per_symbol, bits_per_symbol):
    """Padding so that the modulated packet is a multiple of 128 samples (reference :145-166)."""
    modulus = 128
    byte_modulus = math.lcm(modulus // 8, samples_per_symbol) * bits_per_symbol // samples_per_symbol
    r = pkt_byte_len % byte_modulus
    return 0 if r == 0 else byte_modulus - r


def make_packet(payload, samples_per_symbol, bits_per_symbol, pad_for_usrp=True, whitener_offset=0, whitening=True):
    """header || whiten(payload || crc32 || 0x55 [|| 0x55 * npad]) (reference :99-143)."""
    if not whitener_offset >= 0 and whitener_offset < 16:          # the reference's test, operator precedence included
        raise ValueError("whitener_offset must be between 0 and 15, inclusive (%i)" % (whitener_offset,))
    payload_with_crc = gen_and_append_crc32(payload)
    L = len(payload_with_crc)
    MAXLEN = len(random_mask_tuple)
    if L > MAXLEN:
        raise ValueError("len(payload) must be in [0, %d]" % (MAXLEN,))
    pkt_hd = make_header(L, whitener_offset)
    pkt_dt = payload_with_crc + b"\x55"
    if pad_for_usrp:
        pkt_dt += b"\x55" * _npadding_bytes(len(pkt_hd) + len(pkt_dt), samples_per_symbol, bits_per_symbol)
    return pkt_hd + (whiten(pkt_dt, whitener_offset) if whitening else pkt_dt)


def unmake_packet(whitened_payload_with_crc, whitener_offset=0, dewhitening=1):
    """Return (ok, payload) (reference :169-191)."""
    data = _as_bytes(whitened_payload_with_crc)
    if dewhitening:
        data = dewhiten(data, whitener_offset)
    return check_crc32(data)
