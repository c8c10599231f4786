"""K_TX and the packet kernels against the oracle (bit-exact bytes, samples within 1e-4 relative L2 -- the
north-star tolerance; observed ~1e-7), plus size-independent properties at BASELINE sizes."""
import struct

import numpy as np
import pytest

from oracle import ofdm_oracle as o
from helpers import payloads, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4

LAYOUTS = [(512, 200, 128), (1024, 400, 256), (1024, 800, 256), (4096, 3200, 512), (256, 104, 64), (2048, 800, 512),
           (64, 24, 16), (128, 56, 32)]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("pad", [False, True])
def test_make_packets_bit_exact(torch_cuda, pad):
    from ofdm_uhd_b200.engine import OfdmEngine
    from gpu_helpers import tx_gpu
    rng = np.random.default_rng(1)
    lens = [0, 1, 2, 15, 16, 17, 398, 402, 1020, 4075 if pad else 4091, 7, 0]
    pay = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in lens]
    eng = OfdmEngine(512, 200, 128, "bpsk")
    _, plan = tx_gpu(eng, pay, pad_for_usrp=pad)
    got = plan.pkts.cpu().numpy().tobytes()
    want = b"".join(o.make_packet(p, 1, 1, pad) for p in pay)
    assert got == want
    eng.close()


@pytest.mark.parametrize("mod", ["bpsk", "qpsk", "8psk", "qam16", "qam64", "qam256"])
@pytest.mark.parametrize("N,occ,cp", LAYOUTS)
def test_tx_samples_match_oracle(torch_cuda, N, occ, cp, mod):
    from ofdm_uhd_b200.engine import OfdmEngine
    from gpu_helpers import tx_gpu
    lay = o.Layout(N, occ, cp, mod)
    if lay.ncar * lay.nbits < 32:
        pytest.skip("fewer than 32 bits per OFDM symbol: the header would span symbols (rejected by ofdm_create)")
    rng = np.random.default_rng(N + len(mod))
    lens = [398, 0, 1, 33, 402, 219, 700]                     # ragged: exercises partial groups and pad symbols
    pay = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in lens]
    eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=99)
    x, _ = tx_gpu(eng, pay, first_frame=5)
    want = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, 0.25, seed=99, first_frame=5)
    got = x.cpu().numpy()
    assert got.shape == want.shape
    assert rel_l2(got, want) < TOL
    assert float(np.max(np.abs(got - want))) < 1e-5
    eng.close()


def test_tx_amplitude_and_max_payload(torch_cuda):
    from ofdm_uhd_b200.engine import OfdmEngine
    from gpu_helpers import tx_gpu
    lay = o.Layout(4096, 3200, 512, "qam256")
    rng = np.random.default_rng(3)
    pay = [bytes(rng.integers(0, 256, 4091, dtype=np.uint8)), b""]
    eng = OfdmEngine(4096, 3200, 512, "qam256", 0.25, pad_seed=1)
    for amp, eff in ((0.7, 0.7), (1.5, 1.0), (-1.0, 0.0)):         # set_tx_amplitude clamps to [0, 1]
        eng.set_tx_amplitude(amp)
        x, _ = tx_gpu(eng, pay)
        want = o.tx_modulate([o.make_packet(p, 1, 1, False) for p in pay], lay, eff, seed=1)
        got = x.cpu().numpy()
        assert got.shape == want.shape == ((3 + 2) * 4608,)
        assert float(np.max(np.abs(got - want))) < 1e-5
    with pytest.raises(ValueError):
        tx_gpu(eng, [bytes(4093)])
    eng.close()


def test_tx_full_size_properties(torch_cuda):
    """BASELINE configs[1] scale (QAM16, 1/8 of the 1 M symbols to keep the test short): the cyclic prefix equals
    the symbol tail, every frame starts with the same preamble, re-running is idempotent."""
    torch = torch_cuda
    from ofdm_uhd_b200.engine import OfdmEngine
    F, psize = 20000, 402
    eng = OfdmEngine(512, 200, 128, "qam16", 0.25, pad_seed=4)
    rng = np.random.default_rng(4)
    body = rng.integers(0, 256, size=F * psize, dtype=np.uint8)
    off = np.arange(F + 1, dtype=np.int64) * psize
    plan = eng.tx_plan(off)
    d = torch.from_numpy(body).cuda()
    x = eng.tx_run(plan, d)
    nsym = plan.uniform_syms
    assert nsym == 6 and x.numel() == F * nsym * 640
    v = x.view(F, nsym, 640)
    assert torch.equal(v[:, :, :128], v[:, :, 512:])             # ofdm_cyclic_prefixer
    assert torch.equal(v[:, 0, :], v[0:1, 0, :].expand(F, 640))   # ofdm_insert_preamble
    x2 = eng.tx_run(plan, d).clone()
    assert torch.equal(x, x2)
    p = float((v[:, 1:, :].abs() ** 2).mean())
    want = 0.25 ** 2 * 198 * (10 / 9) / 512                       # amp^2 * ncar * E|c|^2 / N
    assert abs(p / want - 1) < 0.02
    eng.close()


@pytest.mark.parametrize("pad", [False, True])
def test_payload_limit_matches_host_make_packet(torch_cuda, pad):
    """The whitened body payload || crc32 || 0x55 [|| padding] must fit the 4096-byte PN table: the longest payload the
    reference's make_packet frames is 4091 bytes (4087 with pad_for_usrp); one byte more raises on the host
    (ofdm_packet_utils.py:117-135: the numpy XOR against the table) and must raise here too, not wrap the table.
    At the limit the device framing is byte-identical to the host's."""
    from ofdm_uhd_b200 import ofdm_packet_utils
    from ofdm_uhd_b200.engine import OfdmEngine
    from gpu_helpers import tx_gpu
    limit = 4087 if pad else 4091
    rng = np.random.default_rng(9)
    eng = OfdmEngine(512, 200, 128, "qpsk")
    ok = bytes(rng.integers(0, 256, limit, dtype=np.uint8))
    _, plan = tx_gpu(eng, [ok], pad_for_usrp=pad)
    assert plan.pkts.cpu().numpy().tobytes() == ofdm_packet_utils.make_packet(ok, 1, 1, pad)
    too_long = ok + b"\x00" * (1 if not pad else 5)              # the padded length moves in steps of 16
    with pytest.raises(ValueError):
        ofdm_packet_utils.make_packet(too_long, 1, 1, pad)
    with pytest.raises(ValueError):
        eng.tx_plan(np.array([0, len(too_long)], dtype=np.int64), pad_for_usrp=pad)
    with pytest.raises(ValueError):
        eng.make_packets(None, np.array([0, len(too_long)], dtype=np.int64), pad_for_usrp=pad)
    eng.close()
