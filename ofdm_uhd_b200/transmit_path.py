"""transmit_path with the interface of the reference's transmit_path.py (/root/reference/transmit_path.py:35-84):
``ofdm_mod`` followed by the digital amplitude stage, ``send_pkt(payload, eof, carrier_map_new)`` and
``set_tx_amplitude`` (clamped to [0, 1]).  The amplitude multiply is fused into the transmit kernel."""
import copy

try:
    from . import ofdm
except ImportError:
    import ofdm


class _amp_stage:
    """Stands where ``gr.multiply_const_cc`` stood (transmit_path.py:48): ``set_k`` forwards to the kernel."""

    def __init__(self, owner):
        self._owner = owner            # the ofdm_mod whose kernel applies the amplitude
        self._k = 1.0

    def set_k(self, k):
        self._k = k
        self._owner._amp = k
        self._owner._engine.set_tx_amplitude(k)

    def k(self):
        return self._k


# (flag, long name, group, keyword arguments) of the options the reference's transmit_path registers
# (/root/reference/transmit_path.py:72-76); "eng_float" becomes plain float: no gnuradio.eng_option here
_OPTIONS = (
    ("", "--tx-amplitude", "normal", dict(type="float", default=0.250, metavar="AMPL",
                                          help="set transmitter digital amplitude: 0 <= AMPL < 1 [default=%default]")),
    ("-v", "--verbose", "normal", dict(action="store_true", default=False)),
    ("-S", "--samples-per-symbol", "expert", dict(type="int", default=2, help="set samples/symbol [default=%default]")),
    ("", "--log", "expert", dict(action="store_true", default=False,
                                 help="Log all parts of flow graph to file (CAUTION: lots of data)")),
)


class transmit_path:
    """Packets in, complex64 cuda sample tensors out (to whatever was ``connect()``-ed).

    ``honor_carrier_map``: the reference's ``send_pkt`` receives the sensed carrier map and drops it (the
    ``reset_carrier_map`` call is commented out, transmit_path.py:66-70); True applies it to the modulator
    (SURVEY.md section 8f-1)."""

    def __init__(self, options, device=None, pad_seed=0, batch_limit=4096, honor_carrier_map=False):
        opts = copy.copy(options)                    # the caller's object stays untouched, as in the reference
        self._verbose = opts.verbose
        self._samples_per_symbol = opts.samples_per_symbol        # read and reported only (transmit_path.py:45)
        self._honor_carrier_map = bool(honor_carrier_map)
        self.carrier_map_old = ""
        self.ofdm_tx = ofdm.ofdm_mod(opts, msgq_limit=4, pad_for_usrp=False, device=device, pad_seed=pad_seed,
                                     batch_limit=batch_limit)
        self.amp = _amp_stage(self.ofdm_tx)
        self.set_tx_amplitude(opts.tx_amplitude)
        if self._verbose:
            self._print_verbage()

    # -- plumbing that replaces hier_block2.connect(self.ofdm_tx, self.amp, self) --
    def connect(self, sink):
        return self.ofdm_tx.connect(sink)

    def flush(self):
        return self.ofdm_tx.flush()

    def set_tx_amplitude(self, ampl):
        """Digital amplitude of the samples, clamped to [0, 1] (transmit_path.py:56-62)."""
        self._tx_amplitude = max(0.0, min(ampl, 1))
        self.amp.set_k(self._tx_amplitude)

    def send_pkt(self, payload='', eof=False, carrier_map_new="FE7F"):
        changed = carrier_map_new != self.carrier_map_old
        if changed and self._honor_carrier_map and not eof:
            self.ofdm_tx.reset_carrier_map(carrier_map_new)
        if changed:
            self.carrier_map_old = carrier_map_new
        return self.ofdm_tx.send_pkt(payload, eof)

    def send_pkts(self, payloads, carrier_map_new="FE7F"):
        """Bulk twin of send_pkt: the list is framed by make_packets_kernel and modulated in one pass."""
        if carrier_map_new != self.carrier_map_old:
            if self._honor_carrier_map:
                self.ofdm_tx.reset_carrier_map(carrier_map_new)
            self.carrier_map_old = carrier_map_new
        return self.ofdm_tx.send_pkts(payloads)

    @staticmethod
    def add_options(normal, expert):
        groups = {"normal": normal, "expert": expert}
        for short, long_, grp, kw in _OPTIONS:
            groups[grp].add_option(short, long_, **kw)

    def _print_verbage(self):
        for label, value in (("Tx amplitude     %s", self._tx_amplitude), ("samples/symbol:  %3d", self._samples_per_symbol)):
            print(label % value)
