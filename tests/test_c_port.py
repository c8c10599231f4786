"""The C port of the oracle (CPU baseline of bench.py) against the NumPy oracle: same triggers, same angles,
same packets and CRC verdicts, transmit samples within float32 rounding."""
import numpy as np
import pytest

from oracle import ofdm_oracle as o
from helpers import payloads, loopback_capture


@pytest.fixture(scope="module")
def cport():
    from oracle import c_port
    c_port.build()
    return c_port


CASES = [(512, 200, 128, "bpsk", 40, 0.0), (512, 200, 128, "qpsk", 20, 0.3), (512, 200, 128, "8psk", 30, 0.2),
         (512, 200, 128, "qam16", 25, -0.4), (1024, 400, 256, "qam64", 30, 1.3), (4096, 3200, 512, "qam256", 38, 0.3)]


@pytest.mark.parametrize("N,occ,cp,mod,snr,cfo", CASES)
def test_c_port_equals_numpy_oracle(cport, N, occ, cp, mod, snr, cfo):
    rng = np.random.default_rng(5)
    lay = o.Layout(N, occ, cp, mod)
    pay = payloads(rng, 10)
    cfg = cport.make_cfg(N, occ, cp, mod, 0.25, 77)
    pk = [o.make_packet(p, 1, 1, False) for p in pay]
    assert float(np.max(np.abs(o.tx_modulate(pk, lay, 0.25, seed=77) - cport.tx(cfg, pk)))) < 1e-6
    _, xc = loopback_capture(lay, pay, snr, cfo, seed=9)
    r = o.rx_demodulate(xc, lay)
    pkts, trig, ang, counts = cport.rx(cfg, xc)
    assert np.array_equal(trig, r.trig)
    assert float(np.max(np.abs(ang - r.ang))) < 1e-6
    assert pkts == r.packets and len(pkts) >= 8


def test_c_port_timed_loopback(cport):
    r = cport.time_loopback("qpsk", frames=40, snr=20.0, threads=2)
    assert r["kind"] == "port" and r["cores"] == 2 and r["value"] > 0
    assert "packets ok" in r["sample"]
