"""usrp_receive_path with the reference's interface (/root/reference/usrp_receive_path.py:28-75) and no radio: the UHD
source is the loop-back medium of loopback_air.py; ``self.u.u.set_center_freq(freq, 0)`` (secondary_rx.py:71,85) retunes
which transmitters this receiver hears."""
import sys

try:
    from . import receive_path, loopback_air
    from .usrp_transmit_path import add_freq_option
except ImportError:
    import receive_path
    import loopback_air
    from usrp_transmit_path import add_freq_option


def add_options(parser, expert):
    add_freq_option(parser)
    receive_path.receive_path.add_options(parser, expert)
    expert.add_option("", "--rx-freq", type="float", default=None,
                      help="set Rx frequency to FREQ [default=%default]", metavar="FREQ")
    parser.add_option("-v", "--verbose", action="store_true", default=False)
    for flag, name in (("-a", "--args"), ("", "--spec"), ("-A", "--antenna")):          # uhd_interface.py:196-210, inert here
        if not parser.has_option(name):
            parser.add_option(flag, name, type="string", default=None, help="accepted and ignored (no UHD device)")
    if not parser.has_option("--rx-gain"):
        parser.add_option("", "--rx-gain", type="float", default=None, help="accepted and ignored (no UHD device)")


class usrp_receive_path:
    def __init__(self, rx_callback, options, **kw):
        if options.rx_freq is None:
            sys.stderr.write("-f FREQ or --freq FREQ or --rx-freq FREQ must be specified\n")
            raise SystemExit
        rx_path = receive_path.receive_path(rx_callback, options, **kw)
        for attr in dir(rx_path):                                       # forward the methods
            if not attr.startswith('_') and not hasattr(self, attr):
                setattr(self, attr, getattr(rx_path, attr))
        self.u = loopback_air.tuner(options.rx_freq)
        loopback_air.AIR.attach_receiver(self.u, rx_path)
