"""The medium between the radio-free ``usrp_transmit_path`` and ``usrp_receive_path`` shims: every transmitter's samples
go through the synthetic channel and reach the receivers that are tuned to the transmitter's frequency at that moment
(the USRP / UHD layer itself -- uhd_interface.py, generic_usrp.py -- is out of scope: SURVEY.md section 8)."""
import threading


class Air:
    def __init__(self):
        self.receivers = []                      # (tuner, rx_path)
        self.noise_voltage = 0.003
        self.frequency_offset = 0.0
        self.lead_in = 1280
        self.tail = 2560
        self.seed = 0
        self._lock = threading.Lock()

    def reset(self, **kw):
        self.receivers = []
        for k, v in kw.items():
            setattr(self, k, v)

    def attach_receiver(self, tuner, rx_path):
        self.receivers.append((tuner, rx_path))

    def make_sink(self, tuner, engine):
        """Callable for transmit_path.connect(): channel, then every receiver on the same frequency."""
        try:
            from . import channel_model
        except ImportError:
            import channel_model
        chan = channel_model.channel_model(engine, noise_voltage=self.noise_voltage, frequency_offset=self.frequency_offset,
                                           seed=self.seed, lead_in=self.lead_in, tail=self.tail)

        def sink(samples):
            out = chan.process(samples)
            with self._lock:
                for rt, rx in list(self.receivers):
                    if rt.freq == tuner.freq:
                        rx.feed(out)
        return sink


AIR = Air()


class tuner:
    """What ``self.u`` / ``self.u.u`` of the reference's paths offer to the scripts: set_center_freq(freq, chan)."""

    def __init__(self, freq):
        self.freq = freq
        self.u = self                               # secondary_rx.py:71 reaches tb.rxpath.u.u.set_center_freq(freq, 0)
        self.history = [freq]

    def set_center_freq(self, freq, chan=0):
        self.freq = freq
        self.history.append(freq)
        return True

    set_freq = set_center_freq
