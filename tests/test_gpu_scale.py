"""Parity at BASELINE scale, with the C port of the oracle (oracle/ofdm_oracle_c.c, proven equal to ofdm_oracle.py by
tests/test_c_port.py) as the checker: the receiver's (ok, payload) list must be identical to the oracle's on thousands
of frames of every BASELINE modem configuration.  Trigger indices agree except for
rare one-sample moves: the GPU filters with an FFT, the oracle with a float64 FIR sum, so y differs by ~1e-7 and the
arg-max of a flat metric plateau can move by one sample (bytes are unaffected; per-stage tests feed both sides the
same y and demand exact triggers)."""
import threading

import numpy as np
import pytest

from oracle import c_port

pytestmark = pytest.mark.gpu


def _capture(eng, torch, payloads, snr, cfo, seed):
    """payload rows -> GPU make_packets + TX -> [lead | frames | tail] -> GPU channel kernel (the bench's recipe)."""
    F, psize = payloads.shape
    plan = eng.tx_plan(np.arange(F + 1, dtype=np.int64) * psize)
    lead = 2 * eng.L
    x = torch.zeros(plan.n_samples + 2 * lead, dtype=torch.complex64, device="cuda")
    eng.tx_run(plan, torch.from_numpy(payloads.reshape(-1)).cuda(), out=x[lead:lead + plan.n_samples])
    p_sig = float((x[lead:lead + plan.n_samples].abs() ** 2).mean().item())
    sigma = (p_sig / (10 ** (snr / 10.0)) / 2.0) ** 0.5
    return eng.channel(x, cfo=cfo, sigma=sigma, seed=seed), plan.n_samples // F


CASES = {
    # name: (N, occ, cp, mod, payload bytes, frames, snr dB, cfo)
    "cfg1-bpsk-40dB-1000-back-to-back": (512, 200, 128, "bpsk", 402, 1000, 40.0, 0.0),
    "cfg2-qpsk-20dB-bench-capture": (512, 200, 128, "qpsk", 402, 4000, 20.0, None),      # None: the bench's first CFO
    "cfg2-qam16-20dB-bench-capture": (512, 200, 128, "qam16", 402, 4000, 20.0, None),
    "cfg3-qam64-30dB-full-stream": (1024, 400, 256, "qam64", 402, 4096, 30.0, 1.37),
    "cfg5-qam256-35dB": (4096, 3200, 512, "qam256", 4091, 1200, 35.0, -0.31),
}


def test_receiver_equals_oracle_at_scale():
    import torch
    import bench
    from ofdm_uhd_b200.engine import OfdmEngine
    caps, got = {}, {}
    for name, (N, occ, cp, mod, psize, F, snr, cfo) in CASES.items():
        eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=20260102)
        pay = bench.make_payloads(F, psize, 20260102)
        if cfo is None:
            cfo = float(bench.bench_cfos(100000, 0)[0])
        xc, _ = _capture(eng, torch, pay, snr, cfo, seed=991)
        r = eng.demodulate(xc, max_frames=F + 64)
        caps[name] = xc.cpu().numpy()
        got[name] = (r.packets, r.trig_idx.copy(), r.trig_ang.copy(), pay, r.msg_frames.copy())
        eng.close()
        del xc
    torch.cuda.empty_cache()
    ref = {}

    def work(name):
        N, occ, cp, mod, psize, F, snr, cfo = CASES[name]
        ref[name] = c_port.rx(c_port.make_cfg(N, occ, cp, mod), caps[name], max_pkts=F + 64)

    th = [threading.Thread(target=work, args=(n,)) for n in CASES]
    for t in th:
        t.start()
    for t in th:
        t.join()
    report, differing = {}, {}
    for name, (N, occ, cp, mod, psize, F, snr, cfo) in CASES.items():
        packets, trig, ang, pay, msg_frames = got[name]
        rpk, rtrig, rang, _ = ref[name]
        n_ok = sum(1 for ok, _ in packets if ok)
        report[name] = (len(packets), n_ok, F)
        # CRC verdicts identical message by message, CRC-good payloads identical byte for byte -- for every frame whose
        # trigger index is the oracle's.  Where the arg-max of the (flat-topped) timing metric broke a near-tie the other
        # way (the two filters' 1e-7 difference in y), the FFT window sits one sample off: one sample of the next symbol's
        # cyclic prefix leaks in, which a QAM64 frame at 30 dB may or may not survive.  Those frames are counted, not
        # compared.  A message that FAILS its CRC on both sides may differ in a few bits as well (its symbols sit anywhere
        # relative to the slicer's boundaries); bytes of good packets on unmoved frames never differ.
        assert len(packets) == len(rpk), "%s: %d messages vs the oracle's %d" % (name, len(packets), len(rpk))
        assert len(trig) == len(rtrig)
        moved = set(np.flatnonzero(trig != rtrig).tolist())
        first_ok = int(np.searchsorted(trig, N))              # frames are the triggers the sampler can see
        exempt = {i for i, f in enumerate(msg_frames) if (first_ok + int(f)) in moved}
        keep = [i for i in range(len(packets)) if i not in exempt]
        assert [packets[i][0] for i in keep] == [rpk[i][0] for i in keep], "%s: CRC verdicts differ" % name
        assert all(packets[i][1] == rpk[i][1] for i in keep if packets[i][0]), "%s: a CRC-good payload differs" % name
        bad_diff = [(i, sum(bin(a ^ b).count("1") for a, b in zip(packets[i][1], rpk[i][1])) if len(packets[i][1]) == len(rpk[i][1]) else -1)
                    for i in keep if not packets[i][0] and packets[i][1] != rpk[i][1]]
        n_bad = sum(1 for ok, _ in packets if not ok)
        differing[name] = (len(bad_diff), n_bad, len(exempt), sum(1 for i in exempt if packets[i] != rpk[i]))
        assert all(nb >= 0 for _, nb in bad_diff), "%s: a failed message has another length" % name
        assert len(bad_diff) <= max(1, n_bad // 5), "%s: %d of %d CRC-failed messages differ: %s" % (name, len(bad_diff), n_bad, bad_diff[:8])
        # the oracle decodes what was sent: every CRC-good payload is one of the transmitted ones, in order
        nos = [int.from_bytes(p[:2], "big") for ok, p in packets if ok]
        assert nos == sorted(nos) and all(bytes(pay[k]) == p for k, (ok, p) in zip(nos, [q for q in packets if q[0]]))
        moved = np.flatnonzero(trig != rtrig)
        assert len(moved) <= max(2, len(trig) // 50) and (np.abs(trig - rtrig) <= max(1, cp // 64)).all(), name    # ~1 % at N = 4096 (cp 512: a flatter, wider top)
        assert np.abs(ang - rang)[np.setdiff1d(np.arange(len(trig)), moved)].max(initial=0) < 2e-3
    # delivery rates of the surveyed configurations (identical for oracle and receiver by the equality above)
    assert report["cfg1-bpsk-40dB-1000-back-to-back"][1] >= 990
    assert report["cfg2-qpsk-20dB-bench-capture"][1] >= 3950 and report["cfg2-qam16-20dB-bench-capture"][1] >= 3000
    assert report["cfg3-qam64-30dB-full-stream"][1] >= 3500
    print("delivery (messages, crc ok, sent):", report)
    print("(CRC-failed messages whose bytes differ, CRC-failed in all, messages on moved triggers, of those differing):", differing)
    # at the bench's own operating point the lists are identical outright
    assert differing["cfg2-qpsk-20dB-bench-capture"][0] == 0 and differing["cfg1-bpsk-40dB-1000-back-to-back"][0] == 0
