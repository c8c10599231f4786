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


class transmit_path:
    def __init__(self, options, device=None, pad_seed=0, batch_limit=4096, honor_carrier_map=False):
        options = copy.copy(options)    # make a copy so we can destructively modify

        self._verbose = options.verbose
        self._tx_amplitude = options.tx_amplitude                 # digital amplitude sent to the radio
        self._samples_per_symbol = options.samples_per_symbol     # parsed, unused (as in the reference)

        self.ofdm_tx = ofdm.ofdm_mod(options, msgq_limit=4, pad_for_usrp=False, device=device, pad_seed=pad_seed,
                                     batch_limit=batch_limit)
        self.amp = _amp_stage(self.ofdm_tx)
        self.set_tx_amplitude(self._tx_amplitude)
        self.carrier_map_old = ""
        self._honor_carrier_map = bool(honor_carrier_map)
        if self._verbose:
            self._print_verbage()

    def connect(self, sink):
        return self.ofdm_tx.connect(sink)

    def flush(self):
        return self.ofdm_tx.flush()

    def set_tx_amplitude(self, ampl):
        """
        Sets the transmit amplitude sent to the radio
        @param: ampl 0 <= ampl < 1.
        """
        self._tx_amplitude = max(0.0, min(ampl, 1))
        self.amp.set_k(self._tx_amplitude)

    def send_pkt(self, payload='', eof=False, carrier_map_new="FE7F"):
        # the reference accepts the map and ignores it (reset_carrier_map is commented out, :66-70);
        # honor_carrier_map=True restores the intended behaviour
        if carrier_map_new != self.carrier_map_old:
            if self._honor_carrier_map and not eof:
                self.ofdm_tx.reset_carrier_map(carrier_map_new)
            self.carrier_map_old = carrier_map_new
        return self.ofdm_tx.send_pkt(payload, eof)

    @staticmethod
    def add_options(normal, expert):
        normal.add_option("", "--tx-amplitude", type="float", default=0.250, metavar="AMPL",
                          help="set transmitter digital amplitude: 0 <= AMPL < 1 [default=%default]")
        normal.add_option("-v", "--verbose", action="store_true", default=False)
        expert.add_option("-S", "--samples-per-symbol", type="int", default=2,
                          help="set samples/symbol [default=%default]")
        expert.add_option("", "--log", action="store_true", default=False,
                          help="Log all parts of flow graph to file (CAUTION: lots of data)")

    def _print_verbage(self):
        print("Tx amplitude     %s" % (self._tx_amplitude))
        print("samples/symbol:  %3d" % (self._samples_per_symbol))
