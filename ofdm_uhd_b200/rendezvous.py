"""The secondary-user rendezvous protocol of the reference scripts, radio-free (py3).

Transmit side (/root/reference/secondary_tx.py): after a hop decision the transmitter retunes to the 920 MHz
rendezvous channel and sends 100 synchronisation packets that carry the new operating frequency
(``synchronization``, :54-73); ``transmitter_control`` (:30-52) picks the channel; ``run_transmiter`` (:345-381)
frames a byte source into numbered packets.  Receive side (/root/reference/secondary_rx.py:51-85): the
``rx_callback`` state machine follows the announced frequency and falls back to 920 MHz after 10 packets
without the 11111 marker.  ``send_pkt`` is any callable with ``transmit_path.send_pkt``'s signature, so the
same code drives ``ofdm_uhd_b200.transmit_path`` in loopback or a real radio front end.
"""
import math
import struct

SYNC_FREQ = 920 * 10 ** 6            # secondary_tx.py:37, secondary_rx.py:80
PREAMBLE = 11111                     # secondary_tx.py:60, secondary_rx.py:54
SYNC_PKTNO = 150                     # secondary_tx.py:58, secondary_rx.py:66
N_SYNC_PACKETS = 100                 # secondary_tx.py:61


def data_payload(pktno: int, data: bytes) -> bytes:
    """secondary_tx.py:63,376: '!H' packet number, '!H' marker, then the data."""
    return struct.pack('!H', pktno & 0xffff) + struct.pack('!H', PREAMBLE & 0xffff) + data


def sync_payload(pktno: int, frequency: int) -> bytes:
    """secondary_tx.py:62-63: the new operating frequency as '!L'."""
    return data_payload(pktno, struct.pack('!L', int(frequency) & 0xffffffff))


def next_tx_frequency(sync: int, frequency: int) -> int:
    """transmitter_control (secondary_tx.py:36-39): rendezvous channel while a hop is being announced."""
    return SYNC_FREQ if sync == 1 else int(frequency)


def synchronization(send_pkt, frequency: int, carrier_map: str = "FE7F") -> int:
    """secondary_tx.py:54-73: 100 synchronisation packets numbered from 150.  Returns the packet count."""
    n, pktno = 0, SYNC_PKTNO
    while n < N_SYNC_PACKETS:
        send_pkt(sync_payload(pktno, frequency), False, carrier_map)
        n += 1
        pktno += 1
    return n


def run_transmitter(send_pkt, source: bytes, pkt_size: int, carrier_map: str = "FE7F") -> int:
    """secondary_tx.py:345-381 with the file replaced by a byte string: 20 filler packets, packet 20 carries the
    packet count, then ``pkt_size - 4`` source bytes per packet, and 20 trailing filler packets at the end of the
    source.  Returns the payload bytes sent before the trailer (the reference's ``n``)."""
    file_size = len(source)
    no_packets = int(math.ceil(file_size // pkt_size))           # py2 integer division inside ceil, as written
    n, pktno, pos = 0, 0, 0
    while True:
        if pktno < 20:
            data = b"This is Garbage data"
        elif pktno == 20:
            data = struct.pack('!H', no_packets & 0xffff)
        else:
            data = source[pos:pos + pkt_size - 4]
            pos += len(data)
            if data == b'':
                for _ in range(20):
                    send_pkt(data_payload(pktno, b"This is also Garbage data"), False, carrier_map)
                    pktno += 1
                break
        payload = data_payload(pktno, data)
        send_pkt(payload, False, carrier_map)
        n += len(payload)
        pktno += 1
    return n


class secondary_receiver:
    """The rx_callback closure of secondary_rx.py:51-85 as an object: ``set_center_freq(freq)`` stands for
    ``tb.rxpath.u.u.set_center_freq(freq, 0)``, ``sink`` for the rx.txt file."""

    def __init__(self, set_center_freq=None, sink=None, verbose=False):
        self.n_rcvd = 0
        self.n_right = 0
        self.shift = 0
        self.sync = 1
        self.no_packets = 50
        self.freq = SYNC_FREQ
        self.set_center_freq = set_center_freq or (lambda f: None)
        self.sink = sink
        self.verbose = verbose

    def _tune(self, freq):
        self.freq = int(freq)
        self.set_center_freq(self.freq)

    def rx_callback(self, ok, payload):
        # a payload shorter than 4 bytes makes the reference's struct.unpack raise inside the watcher thread
        # (SURVEY C.9); here it simply counts as a packet without the marker
        preamble = struct.unpack('!H', payload[2:4])[0] if len(payload) >= 4 else None
        if preamble == PREAMBLE:
            self.n_rcvd += 1
            self.shift = 0
            (pktno,) = struct.unpack('!H', payload[0:2])
            if pktno == 20:
                if len(payload) >= 6:
                    self.no_packets = struct.unpack('!H', payload[4:6])[0] + 20
            elif pktno < 20:
                pass
            elif 70 < pktno < SYNC_PKTNO:
                pass
            elif pktno >= SYNC_PKTNO:
                if self.sync == 1 and ok and len(payload) >= 8:
                    (freq,) = struct.unpack('!L', payload[4:8])
                    self._tune(freq)
                    self.sync = 0
            elif self.sink is not None:
                self.sink.write(payload[4:])
            if ok:
                self.n_right += 1
            if self.verbose:
                print("ok: %r \t pktno: %d \t n_rcvd: %d \t n_right: %d" % (ok, pktno, self.n_rcvd, self.n_right))
        else:
            self.shift += 1
            if self.shift >= 10:
                self.sync = 1
                self._tune(SYNC_FREQ)
