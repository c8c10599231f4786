"""Synthetic channel for the loopback drivers (BASELINE configs[0]: benchmark_ofdm_tx -> channel_model ->
benchmark_ofdm_rx).  It stands where the USRP sink/source pair stood (usrp_transmit_path.py /
usrp_receive_path.py, out of scope): AWGN + carrier frequency offset, computed by a CUDA kernel."""
import math


class channel_model:
    def __init__(self, engine, noise_voltage=0.0, frequency_offset=0.0, seed=0, lead_in=0, tail=0):
        """
        @param engine: an OfdmEngine (any layout with the same fft_length)
        @param noise_voltage: standard deviation of each of the I and Q noise components
        @param frequency_offset: carrier offset in subcarrier spacings
        @param lead_in / tail: noise-only samples before / after each buffer, so that the first preamble's
               correlation window starts on noise and the detector can close the last frame's peak
        """
        self.engine = engine
        self.noise_voltage = float(noise_voltage)
        self.frequency_offset = float(frequency_offset)
        self.seed = int(seed)
        self.lead_in = int(lead_in)
        self.tail = int(tail)
        self._calls = 0
        self._phase = 0.0
        self._sinks = []

    @staticmethod
    def noise_voltage_for_snr(snr_db, signal_power):
        return math.sqrt(signal_power / (10.0 ** (snr_db / 10.0)) / 2.0)

    def connect(self, sink):
        self._sinks.append(sink)
        return sink

    def process(self, samples):
        import torch
        if self.lead_in or self.tail:
            z = samples.new_zeros
            samples = torch.cat([z(self.lead_in), samples, z(self.tail)])
        out = self.engine.channel(samples, cfo=self.frequency_offset, sigma=self.noise_voltage,
                                  seed=self.seed + 0x9E3779B9 * self._calls, phase0=self._phase)
        self._phase = (self._phase + 2.0 * math.pi * self.frequency_offset / self.engine.N * samples.numel()) % (2.0 * math.pi)
        self._calls += 1
        return out

    def feed(self, samples):
        out = self.process(samples)
        for s in self._sinks:
            (s.feed if hasattr(s, "feed") else s)(out)
        return out
