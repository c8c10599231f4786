"""Spectrum sensing with the structure of the reference's secondary-user scripts
(/root/reference/secondary_tx.py:146-331, sensing_and_tramsmitting.py:154-374, predictive_sense.py:36-268,
usrp_fft_save.py:44-62): Blackman-Harris windowed FFT -> |X|^2 -> per-dwell max-hold (bin_statistics_f) ->
10-dwell mean -> threshold -> frequency-order carrier map -> hex string.  The radio / file source is a
complex64 capture tensor; every numeric stage runs in CUDA kernels (ofdm_sense / ofdm_sense_decide)."""
import math

import numpy as np

try:
    from .engine import SenseEngine
except ImportError:
    from engine import SenseEngine


class parse_msg:
    """The message bin_statistics_f posts (secondary_tx.py:137-144): centre frequency, length, float32 data."""

    def __init__(self, center_freq, data):
        self.center_freq = center_freq
        self.vlen = len(data)
        self.data = data
        self.raw_data = np.asarray(data, dtype=np.float32).tobytes()


class sensor:
    """secondary_tx.sensor (secondary_tx.py:146-226) without the radio: ``options`` needs fft_size, decim,
    tune_delay, dwell_delay (seconds) and optionally sense_bins."""

    def __init__(self, options, device=None, shift=False):
        self.fft_size = int(options.fft_size)
        self.ofdm_bins = int(getattr(options, "sense_bins", 128))
        self.samp_rate = 100e6 / options.decim                       # secondary_tx.py:175 (C.10: not 100**6)
        self.min_freq = 905 * 10 ** 6 - (10 * 10 ** 6)
        self.max_freq = 905 * 10 ** 6 + (10 * 10 ** 6)
        self.min_center_freq = (self.min_freq + self.max_freq) / 2
        self.freq_step = 0
        self.next_freq = self.min_center_freq
        self.tune_delay = max(0, int(round(options.tune_delay * self.samp_rate / self.fft_size)))    # in fft_frames
        self.dwell_delay = max(1, int(round(options.dwell_delay * self.samp_rate / self.fft_size)))  # in fft_frames
        self.shift = bool(shift)
        self.engine = SenseEngine(self.fft_size, device=device)
        self.msgq_limit = 16                                          # gr.msg_queue(16), secondary_tx.py:198

    def set_next_freq(self):
        target = self.next_freq
        self.next_freq = self.next_freq + self.freq_step
        return target

    def dwell_vectors(self, capture):
        """All dwell max-hold vectors of a capture: cuda float32 [n_dwell, fft_size]."""
        return self.engine.maxhold(capture, self.tune_delay, self.dwell_delay, shift=self.shift)

    def messages(self, capture):
        mh = self.dwell_vectors(capture).cpu().numpy()
        return [parse_msg(self.set_next_freq(), row) for row in mh]


def hex_conv(thrshold_inorder):
    """secondary_tx.py:306-331: 4 bins per hex digit, first bin = least significant bit, upper case."""
    digits = "0123456789ABCDEF"
    out = []
    for i in range(0, len(thrshold_inorder) - 3, 4):
        v = sum((1 << j) for j in range(4) if thrshold_inorder[i + j] == 1)
        out.append(digits[v])
    return "".join(out)


def sense_decision(tb, capture, threshold=0.001, avg_iterations=10):
    """The averaging / threshold half of ``sense_loop`` (secondary_tx.py:228-266) for every complete group of
    ``avg_iterations`` dwells in ``capture``.  Returns a list of (avg_inorder float64[N], free_inorder uint8[N],
    hexa_thr str)."""
    mh = tb.dwell_vectors(capture)
    out = []
    for g in range(mh.shape[0] // avg_iterations):
        out.append(tb.engine.decide(mh[g * avg_iterations:(g + 1) * avg_iterations], threshold))
    return out


def busy_count(thrshold_inorder, frequency, samp_rate, size, ref_freq=8925 * 10 ** 5):
    """secondary_tx.py:268-277: occupied bins in the 32-bin window around the operating frequency."""
    required_index = int(math.ceil((frequency - ref_freq) * size / samp_rate))
    window = thrshold_inorder[required_index - 16:required_index + 16]
    return int(sum(1 for v in window if v == 0))


def best_band(avg_inorder, lo=200, span=17):
    """secondary_tx.py:284-295: centre (+8) of the quietest 17-bin window over bins lo .. size-217."""
    size = len(avg_inorder)
    if size - 217 <= lo:
        return -1
    power_temp, index = 50.0, -1
    for i in range(lo, size - 217):
        power = 0.0
        for j in range(span):
            power = power + float(avg_inorder[i + j])
        if power < power_temp:
            power_temp, index = power, i + 8
    return index


def sensed_frequency(center_freq, samp_rate, size, index):
    """sensed_freq[index] of secondary_tx.py:250-262: the bin frequencies are built by repeated addition of
    ``usr/size`` from ``center - res*((size/2)-1)`` (integer size/2), so the same sequence of float adds is kept."""
    res = samp_rate / size
    p = center_freq - res * ((size // 2) - 1)
    for _ in range(int(index)):
        p = p + res
    return p


def hop_decision(tb, avg_inorder_dev, free_dev, frequency, center_freq, ref_freq=8925 * 10 ** 5, busy_limit=9):
    """The second half of ``sense_loop`` (secondary_tx.py:268-300) on the device tensors of
    ``SenseEngine.decide_device``: the busy count of the 32-bin window around ``frequency`` and, when at least
    ``busy_limit`` of them are occupied ('Primary Transmission detected'), the new operating frequency = centre of
    the quietest 17-bin band rounded up to 100 kHz.  Returns (busy, new_frequency or None)."""
    size = tb.fft_size
    required_index = int(math.ceil((frequency - ref_freq) * size / tb.samp_rate))
    busy, index, _ = tb.engine.hop(avg_inorder_dev, free_dev, required_index)
    if busy < busy_limit or index < 0:
        return busy, None
    return busy, int(1e5 * math.ceil(sensed_frequency(center_freq, tb.samp_rate, size, index) / 1e5))


def fft_save(capture, path="fft_data", fft_size=512, device=None, append=False):
    """usrp_fft_save.fft (usrp_fft_save.py:44-62) without the radio: stream_to_vector(512) ->
    fft_vcc(512, True, blackmanharris(512), shift=True) -> file_sink: the windowed, shifted spectra of consecutive
    ``fft_size``-sample frames written as raw interleaved float32 (the layout utils/read_complex_binary.m:39-46 reads).
    ``capture``: complex64 cuda tensor or NumPy array.  Returns the number of frames written."""
    import torch
    eng = SenseEngine(int(fft_size), device=device)
    try:
        x = capture if isinstance(capture, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(capture, dtype=np.complex64))
        if x.device.type != "cuda":
            x = x.to(eng.dev)
        sp = eng.spectra(x.contiguous(), shift=True)
        with open(path, "ab" if append else "wb") as f:
            f.write(sp.cpu().numpy().astype(np.complex64).tobytes())
        return int(sp.shape[0])
    finally:
        eng.close()
