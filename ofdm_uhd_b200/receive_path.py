"""receive_path with the interface of the reference's receive_path.py (/root/reference/receive_path.py:29-58):
wraps ``ofdm_demod`` and forwards ``rx_callback(ok, payload)``."""
import copy

try:
    from . import ofdm
except ImportError:
    import ofdm


class receive_path:
    def __init__(self, rx_callback, options, device=None, max_pkt_bytes=4096):
        options = copy.copy(options)    # make a copy so we can destructively modify

        self._verbose = options.verbose
        self._log = options.log
        self._rx_callback = rx_callback      # this callback is fired when there's a packet available

        self.ofdm_rx = ofdm.ofdm_demod(options, callback=self._rx_callback, device=device,
                                       max_pkt_bytes=max_pkt_bytes)
        if self._verbose:
            self._print_verbage()
        # a radio-sized buffer is a microsecond of receiver time: the per-buffer call goes straight to the demodulator
        self.feed_stream = self.ofdm_rx.feed_stream

    def feed(self, samples, max_frames=None):
        """Run the receiver on one buffer of complex64 baseband samples."""
        return self.ofdm_rx.feed(samples, max_frames=max_frames)

    def feed_stream(self, samples, max_frames=None, flush=False):
        """feed() for consecutive buffers of one continuous stream (see ofdm_demod.feed_stream)."""
        return self.ofdm_rx.feed_stream(samples, max_frames, flush)

    def flush_stream(self, max_frames=None):
        return self.ofdm_rx.flush_stream(max_frames=max_frames)

    def set_batch_callback(self, fn):
        """rx_callback_batch(ok[], bytes, offsets): one call per feed() instead of one per packet (ofdm_demod.set_batch_callback)."""
        self.ofdm_rx.set_batch_callback(fn)

    def wait(self, timeout=None):
        self.ofdm_rx.wait(timeout)

    @staticmethod
    def add_options(normal, expert):
        normal.add_option("-v", "--verbose", action="store_true", default=False)
        expert.add_option("-S", "--samples-per-symbol", type="int", default=2,
                          help="set samples/symbol [default=%default]")
        expert.add_option("", "--log", action="store_true", default=False,
                          help="Log all parts of flow graph to files (CAUTION: lots of data)")

    def _print_verbage(self):
        print("\nReceive Path:")
