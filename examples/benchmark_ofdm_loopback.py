#!/usr/bin/env python
"""benchmark_ofdm_tx.py + benchmark_ofdm_rx.py of the reference in one process, with the radio replaced by the
synthetic channel: the same option parsing (transmit_path / receive_path / ofdm_mod / ofdm_demod add_options), the same
packet format (struct '!H' pktno, '!H' preamble 0, data; benchmark_ofdm_tx.py:106-121) and the same rx_callback
(benchmark_ofdm_rx.py:50-61) -- only `import ofdm, transmit_path, receive_path` now resolves to ofdm_uhd_b200/.

    python examples/benchmark_ofdm_loopback.py -m qpsk -s 402 -M 0.004 --snr 25 --cfo 0.2 [--from-file F] [--to-file G]
"""
import math
import os
import struct
import sys
from optparse import OptionParser

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "ofdm_uhd_b200"))       # the flat script-directory layout

import ofdm                                                                   # noqa: E402  (the B200 modules)
import transmit_path                                                          # noqa: E402
import receive_path                                                           # noqa: E402
import channel_model                                                          # noqa: E402


def main(argv=None):
    n_rcvd = n_right = 0
    sink = []

    def rx_callback(ok, payload):                                             # benchmark_ofdm_rx.py:50-61
        nonlocal n_rcvd, n_right
        if len(payload) < 4:
            return
        (preamble,) = struct.unpack('!H', payload[2:4])
        if preamble == 0:
            n_rcvd += 1
            (pktno,) = struct.unpack('!H', payload[0:2])
            if pktno > 19 and ok:
                sink.append(payload[4:])
            if ok:
                n_right += 1
            if options.verbose:
                print("ok: %r \t pktno: %d \t n_rcvd: %d \t n_right: %d" % (ok, pktno, n_rcvd, n_right))

    parser = OptionParser(conflict_handler="resolve")
    expert_grp = parser.add_option_group("Expert")
    parser.add_option("-s", "--size", type="float", default=1024, help="set packet size [default=%default]")
    parser.add_option("-M", "--megabytes", type="float", default=0.01,
                      help="set megabytes to transmit (x 10e6 bytes, benchmark_ofdm_tx.py:96) [default=%default]")
    parser.add_option("", "--from-file", default=None, help="use file for packet contents")
    parser.add_option("", "--to-file", default=None, help="write the received file contents here")
    parser.add_option("", "--snr", type="float", default=30, help="set the SNR of the channel in dB [default=%default]")
    parser.add_option("", "--cfo", type="float", default=0.0, help="carrier frequency offset in subcarrier spacings")
    transmit_path.transmit_path.add_options(parser, expert_grp)
    receive_path.receive_path.add_options(parser, expert_grp)
    ofdm.ofdm_mod.add_options(parser, expert_grp)
    ofdm.ofdm_demod.add_options(parser, expert_grp)
    (options, args) = parser.parse_args(argv)
    if len(args) != 0:
        parser.print_help()
        sys.exit(1)

    tx = transmit_path.transmit_path(options)                                 # tb.txpath of benchmark_ofdm_tx.py:42
    rx = receive_path.receive_path(rx_callback, options)                      # tb.rxpath of benchmark_ofdm_rx.py:39
    sig_rms = options.tx_amplitude * math.sqrt(options.occupied_tones / float(options.fft_length))
    chan = channel_model.channel_model(tx.ofdm_tx._engine, noise_voltage=sig_rms / 10 ** (options.snr / 20.0) / math.sqrt(2),
                                       frequency_offset=options.cfo, seed=1, lead_in=2 * 640, tail=4 * 640)
    tx.connect(chan)
    chan.connect(rx)

    def send_pkt(payload=b'', carrier_map="FE7F", eof=False):                 # benchmark_ofdm_tx.py:59-60
        return tx.send_pkt(payload, eof, carrier_map)

    source = open(options.from_file, 'rb').read() if options.from_file else os.urandom(int(1e4))
    nbytes = int(10e6 * options.megabytes)
    pkt_size = int(options.size)
    n = pktno = pos = 0
    sent_data = []
    while n < nbytes:
        if pktno < 20:
            data = b"This is Garbage data"
        else:
            data = source[pos:pos + pkt_size - 2]
            pos += len(data)
            if data == b'':
                break
            sent_data.append(data)
        payload = struct.pack('!H', pktno & 0xffff) + struct.pack('!H', 0) + data
        send_pkt(payload)
        n += len(payload)
        pktno += 1
    send_pkt(eof=True)
    rx.wait(timeout=120)
    got = b"".join(sink)
    if options.to_file:
        open(options.to_file, 'wb').write(got)
    print("sent %d packets (%d bytes of file data), received %d, CRC ok %d, file bytes recovered %d"
          % (pktno, sum(len(d) for d in sent_data), n_rcvd, n_right, len(got)))
    return pktno, n_rcvd, n_right, b"".join(sent_data), got


if __name__ == '__main__':
    main()
