"""K_SENSE (window + FFT + |X|^2 + max-hold) and the decision kernel against the oracle."""
import numpy as np
import pytest

from oracle import ofdm_oracle as o
from helpers import rel_l2

pytestmark = pytest.mark.gpu


def capture(N, nfr, seed, tone=0.13, amp=0.01):
    rng = np.random.default_rng(seed)
    x = ((rng.standard_normal(nfr * N) + 1j * rng.standard_normal(nfr * N)) * 1e-3).astype(np.complex64)
    x += (amp * np.exp(2j * np.pi * tone * np.arange(nfr * N))).astype(np.complex64)
    return x


@pytest.mark.parametrize("N", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("shift", [False, True])
def test_maxhold_and_decision(N, shift):
    import torch
    from ofdm_uhd_b200.engine import SenseEngine
    x = capture(N, 130, N)
    se = SenseEngine(N)
    dx = torch.from_numpy(x).cuda()
    for tune, dwell in ((0, 12), (3, 10), (0, 1)):
        mh = se.maxhold(dx, tune, dwell, shift=shift)
        ref = o.sense_maxhold(x, N, tune, dwell, shift=shift)
        assert tuple(mh.shape) == ref.shape
        assert rel_l2(mh.cpu().numpy(), ref) < 1e-4
    mh = se.maxhold(dx, 0, 12, shift=False)
    avg, free, hx = se.decide(mh[:10], 1e-3)
    a2, f2, h2 = o.sense_decide(mh[:10].cpu().numpy(), 1e-3)           # same dwell vectors -> exact
    assert np.array_equal(avg, a2) and np.array_equal(free, f2) and hx == h2
    sp = se.spectra(dx[:6 * N], shift=shift).cpu().numpy()
    assert rel_l2(sp, o.sense_fft(x[:6 * N], N, shift)) < 1e-4
    se.close()


def test_wideband_capture_bands_detected():
    """BASELINE configs[3] in miniature: 1024-pt sensing of white noise plus three occupied bands; the hex carrier
    map must flag exactly those bands, and must equal the oracle's map computed from the same capture."""
    import torch
    from ofdm_uhd_b200.engine import SenseEngine
    N, nfr = 1024, 12 * 10 * 4
    rng = np.random.default_rng(5)
    sig = np.sqrt(5e-6 / 1024 / 2)
    X = (rng.standard_normal((nfr, N)) + 1j * rng.standard_normal((nfr, N))) * sig * np.sqrt(N)
    bands = [(100, 140), (400, 416), (800, 900)]                      # in shifted (frequency) order
    for lo, hi in bands:
        X[:, lo:hi] *= 10.0                                           # +20 dB
    x = np.fft.ifft(np.fft.ifftshift(X, axes=1), axis=1).reshape(-1).astype(np.complex64)
    se = SenseEngine(N)
    mh = se.maxhold(torch.from_numpy(x).cuda(), 0, 12, shift=False)
    assert mh.shape[0] == 40
    ref = o.sense_maxhold(x, N, 0, 12, shift=False)
    assert rel_l2(mh.cpu().numpy(), ref) < 1e-4
    thr = 2e-5
    for g in range(4):
        avg, free, hx = se.decide(mh[10 * g:10 * g + 10], thr)
        _, f2, h2 = o.sense_decide(ref[10 * g:10 * g + 10], thr)
        assert hx == h2 and np.array_equal(free, f2)
        busy = np.flatnonzero(free == 0)
        inside = np.zeros(N, bool)
        for lo, hi in bands:
            inside[lo:hi] = True
        assert inside[busy].mean() > 0.9 and (free[inside] == 0).mean() > 0.7
    se.close()


@pytest.mark.parametrize("N", [512, 1024, 2048])
def test_hop_decision_on_device(N):
    """ofdm_sense_hop (secondary_tx.py:268-295) against the ORACLE's best_band (oracle/ofdm_oracle.py) and the reference's
    own slice arithmetic: busy count of the 32-bin window with Python slice semantics (also for windows that run off
    either end) and the quietest 17-bin band, exact (same left-to-right double sums, first strict minimum).  The host
    transcription in sensing.py is held to the same oracle."""
    import math
    from types import SimpleNamespace
    import torch
    from ofdm_uhd_b200 import sensing
    rng = np.random.default_rng(N)
    tb = sensing.sensor(SimpleNamespace(fft_size=N, decim=4, tune_delay=0.0, dwell_delay=1e-3))
    for trial in range(6):
        avg = rng.random(N) * (10.0 if trial == 3 else 1e-3)          # trial 3: no window sums below 50
        if trial == 4:
            avg[300:330] = avg[700:730] = 0.0                         # tie between two all-zero windows: first wins
        free = (rng.random(N) > 0.4).astype(np.uint8)
        d_avg, d_free = torch.from_numpy(avg).cuda(), torch.from_numpy(free).cuda()
        for ri in (-40, -3, 5, 16, N // 2, N - 10, N + 40):
            busy, index, wlen = tb.engine.hop(d_avg, d_free, ri)
            win = list(free)[ri - 16:ri + 16]
            assert wlen == len(win) and busy == sum(1 for v in win if v == 0)
            assert index == o.best_band(avg) == sensing.best_band(avg)
        freq = 905 * 10 ** 6
        ri = int(math.ceil((freq - 8925 * 10 ** 5) * N / tb.samp_rate))
        busy, newf = sensing.hop_decision(tb, d_avg, d_free, freq, 905e6)
        assert busy == sensing.busy_count(list(free), freq, tb.samp_rate, N)
        idx = o.best_band(avg)
        if busy >= 9 and idx >= 0:
            assert newf == int(1e5 * math.ceil(sensing.sensed_frequency(905e6, tb.samp_rate, N, idx) / 1e5))
            assert abs(newf - (905e6 + (idx - N / 2 + 1) * tb.samp_rate / N)) <= 1e5
        else:
            assert newf is None
    tb.engine.close()


def test_decide_kernel_equals_reference_console_logs():
    """ofdm_sense_decide on the averages the reference itself printed (tests/golden/reference_sense_logs.npz): flags and
    carrier-map hex strings must be the ones in the logs; ofdm_sense_hop's busy count must equal the count of 0 flags
    in the printed window."""
    import os
    import torch
    from ofdm_uhd_b200.engine import SenseEngine
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sense_logs.npz"))
    se = SenseEngine(256)
    for s in range(len(d["hex"])):
        dwell = np.zeros((1, 256), dtype=np.float32)
        dwell[0, (np.arange(256) + 128) % 256] = d["avg"][s].astype(np.float32)
        avg, free, hx = se.decide(torch.from_numpy(dwell).cuda(), 1e-4)
        assert hx == str(d["hex"][s]) and np.array_equal(free, d["flag"][s])
        a_d, f_d, _ = se.decide_device(torch.from_numpy(dwell).cuda(), 1e-4)
        busy, _, wlen = se.hop(a_d, f_d, 128)
        assert wlen == 32 and busy == int((d["flag"][s][112:144] == 0).sum())
    se.close()


def test_full_size_capture_properties():
    """BASELINE configs[3] at its full size: 100 M samples = 97 656 frames of 1024, Blackman-Harris, dwell 12.
    Size-independent properties: scaling the capture by 2 scales every max-hold power by exactly 4 (powers of two are
    exact in float32), a capture made of two identical halves gives two identical halves of dwell vectors, and the
    decision over any 10 dwells is the oracle's on those same dwell vectors."""
    import torch
    from ofdm_uhd_b200.engine import SenseEngine
    N, dwell = 1024, 12
    nfr = 100_000_000 // N
    per_half = (nfr // 2) // dwell * dwell
    g = torch.Generator(device="cuda").manual_seed(5)
    half = torch.randn(per_half * N, 2, device="cuda", generator=g) * 1.5e-3
    t = torch.arange(per_half * N, device="cuda", dtype=torch.float32)
    half[:, 0] += 0.01 * torch.cos(2 * np.pi * 0.21 * t)
    half[:, 1] += 0.01 * torch.sin(2 * np.pi * 0.21 * t)
    x = torch.view_as_complex(torch.cat([half, half]).contiguous())
    se = SenseEngine(N)
    mh = se.maxhold(x, 0, dwell)
    nd = mh.shape[0]
    assert nd == 2 * per_half // dwell and nd >= 8100
    assert torch.equal(mh[:nd // 2], mh[nd // 2:])
    mh2 = se.maxhold(x * 2, 0, dwell)
    assert torch.equal(mh2, mh * 4)
    for g0 in (0, nd // 3, nd - 10):
        avg, free, hx = se.decide(mh[g0:g0 + 10], 1e-3)
        a2, f2, h2 = o.sense_decide(mh[g0:g0 + 10].cpu().numpy(), 1e-3)
        assert hx == h2 and np.array_equal(free, f2) and np.array_equal(avg, a2)
        assert (free == 0).sum() >= 1 and free[(N // 2 + int(0.21 * N)) % N] == 0       # the tone is seen as occupied
    se.close()


def test_fft_save_file_format(tmp_path):
    """sensing.fft_save = usrp_fft_save.py's flowgraph: the file holds fftshift(fft(blackmanharris(512) * frame)) of every
    512-sample frame as raw interleaved float32 (what utils/read_complex_binary.m reads back)."""
    from ofdm_uhd_b200 import sensing
    x = capture(512, 9, 3)
    path = tmp_path / "fft_data"
    n = sensing.fft_save(x[:9 * 512 - 100], str(path))                 # the incomplete last frame is dropped
    assert n == 8
    got = np.fromfile(str(path), dtype=np.complex64).reshape(n, 512)
    assert rel_l2(got, o.sense_fft(x[:8 * 512], 512, True)) < 1e-4
    assert sensing.fft_save(x[:512], str(path), append=True) == 1 and path.stat().st_size == 9 * 512 * 8
