"""usrp_transmit_path with the reference's interface (/root/reference/usrp_transmit_path.py:28-62) and no radio: the
UHD sink is replaced by the loop-back medium of loopback_air.py.  Every public attribute of the inner transmit_path is
forwarded, which is how the scripts find ``tb.txpath.send_pkt``."""
import sys

try:
    from . import transmit_path, loopback_air
except ImportError:
    import transmit_path
    import loopback_air


def add_freq_option(parser):
    """-f / --freq sets both tx_freq and rx_freq (usrp_transmit_path.py:28-38)."""
    def freq_callback(option, opt_str, value, parser):
        parser.values.rx_freq = value
        parser.values.tx_freq = value

    if not parser.has_option('--freq'):
        parser.add_option('-f', '--freq', type="float", action="callback", callback=freq_callback,
                          help="set Tx and/or Rx frequency to FREQ [default=%default]", metavar="FREQ")


def add_options(parser, expert):
    add_freq_option(parser)
    transmit_path.transmit_path.add_options(parser, expert)
    expert.add_option("", "--tx-freq", type="float", default=None,
                      help="set transmit frequency to FREQ [default=%default]", metavar="FREQ")
    parser.add_option("-v", "--verbose", action="store_true", default=False)
    for flag, name in (("-a", "--args"), ("", "--spec"), ("-A", "--antenna")):          # uhd_interface.py:144-157, inert here
        if not parser.has_option(name):
            parser.add_option(flag, name, type="string", default=None, help="accepted and ignored (no UHD device)")
    if not parser.has_option("--tx-gain"):
        parser.add_option("", "--tx-gain", type="float", default=None, help="accepted and ignored (no UHD device)")
    if not parser.has_option("--log"):
        parser.add_option("", "--log", action="store_true", default=False,
                          help="Log all parts of flow graph to file (CAUTION: lots of data)")


class usrp_transmit_path:
    def __init__(self, options, **kw):
        if options.tx_freq is None:
            sys.stderr.write("-f FREQ or --freq FREQ or --tx-freq FREQ must be specified\n")
            raise SystemExit
        tx_path = transmit_path.transmit_path(options, **kw)
        for attr in dir(tx_path):                                       # forward the methods
            if not attr.startswith('_') and not hasattr(self, attr):
                setattr(self, attr, getattr(tx_path, attr))
        self.sink = self.u = loopback_air.tuner(options.tx_freq)
        tx_path.connect(loopback_air.AIR.make_sink(self.u, tx_path.ofdm_tx._engine))
