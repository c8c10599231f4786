"""Stage-by-stage GPU-vs-oracle diagnostics (prints, never asserts).  Run on a B200:
    python tests/gpu_bringup.py [quick]
"""
import os
import struct
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                        # noqa: E402
from oracle import ofdm_oracle as o                                 # noqa: E402
from ofdm_uhd_b200.engine import OfdmEngine, SenseEngine            # noqa: E402
from ofdm_uhd_b200 import _lib                                      # noqa: E402
import ctypes as C                                                  # noqa: E402


def rel(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def payloads(rng, k, size=398):
    return [struct.pack("!HH", i & 0xFFFF, 0) + bytes(rng.integers(0, 256, size, dtype=np.uint8)) for i in range(k)]


def run_case(N, occ, cp, mod, nfr, snr, cfo, seed=5, psize=398):
    print("=" * 100)
    print("case N=%d occ=%d cp=%d mod=%s frames=%d snr=%g cfo=%g" % (N, occ, cp, mod, nfr, snr, cfo), flush=True)
    rng = np.random.default_rng(seed)
    lay = o.Layout(N, occ, cp, mod)
    eng = OfdmEngine(N, occ, cp, mod, 0.25, pad_seed=77)
    pay = payloads(rng, nfr, psize)
    # ---- packets
    pkts_o = [o.make_packet(p, 1, 1, False) for p in pay]
    off = np.zeros(nfr + 1, dtype=np.int64)
    np.cumsum([len(p) for p in pay], out=off[1:])
    d_pay = torch.from_numpy(np.frombuffer(b"".join(pay), dtype=np.uint8).copy()).cuda()
    d_pk, pk_off, d_koff = eng.make_packets(d_pay, off, pad_for_usrp=False)
    torch.cuda.synchronize()
    same = d_pk.cpu().numpy().tobytes() == b"".join(pkts_o)
    print("make_packets identical:", same)
    # ---- tx
    x_o = o.tx_modulate(pkts_o, lay, 0.25, seed=77)
    x_g = eng.modulate(d_pk, pk_off, d_koff)
    torch.cuda.synchronize()
    xg = x_g.cpu().numpy()
    print("tx samples:", len(xg), len(x_o), "rel-L2", rel(xg, x_o) if len(xg) == len(x_o) else "LEN MISMATCH")
    if len(xg) == len(x_o):
        d = np.abs(xg - x_o)
        print("   max abs diff %.3g at %d (rms signal %.3g)" % (d.max(), int(d.argmax()), np.sqrt(np.mean(np.abs(x_o) ** 2))))
    # ---- channel (oracle's, so both receivers see identical samples); noise tail so the last frame closes
    tail = np.zeros(4 * (N + cp), dtype=np.complex64)
    xin = np.concatenate([np.zeros(N + 37, dtype=np.complex64), x_o, tail])
    xc = o.channel(xin, snr, cfo, N, seed=seed + 1, sig_power=float(np.mean(np.abs(x_o) ** 2)))
    t0 = time.time()
    r = o.rx_demodulate(xc, lay, keep=True)
    print("oracle rx: %.2fs trig=%d frames=%d pkts=%d ok=%d" % (time.time() - t0, len(r.trig), len(r.frame_start),
                                                                len(r.packets), sum(1 for g, _ in r.packets if g)))
    n = len(xc)
    d_x = torch.from_numpy(xc).cuda()
    nvec = int(len(r.vec_start)) + 64
    bufs = eng.rx_alloc(n, taps=True, max_vectors=nvec)
    io = bufs["io"]
    st = eng._stream()
    L_ = eng.L_
    # stage 1: filter
    y_g = torch.empty(n, dtype=torch.complex64, device="cuda")
    _lib.check(L_.ofdm_rx_chan_filter(eng.h, eng._p(d_x), n, eng._p(y_g), st))
    torch.cuda.synchronize()
    print("chan taps equal:", np.array_equal(eng.chan_taps(), o.chan_filter_taps(lay)), "ntaps", eng.ntaps)
    print("chan_filter rel-L2 vs oracle:", rel(y_g.cpu().numpy(), r.y))
    # stage 2: metric on the ORACLE's y (isolates the stage)
    d_y = torch.from_numpy(r.y).cuda()
    mf_g = torch.empty(n, dtype=torch.float32, device="cuda")
    fnan = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(L_.ofdm_rx_sync_metric(eng.h, eng._p(d_y), n, eng._p(mf_g), eng._p(fnan), st))
    torch.cuda.synchronize()
    mfg = mf_g.cpu().numpy()
    neq = int(np.sum(mfg != r.mf))
    print("sync_metric (oracle y): mismatching samples %d / %d, max abs diff %.3g, first_nan %d" %
          (neq, n, float(np.nanmax(np.abs(mfg - r.mf))), int(fnan.item())))
    if neq:
        bad = np.flatnonzero(mfg != r.mf)[:8]
        print("   first bad idx", bad, mfg[bad], r.mf[bad])
    # stage 3: peak detect on the ORACLE's mf and y
    d_mf = torch.from_numpy(r.mf).cuda()
    _lib.check(L_.ofdm_rx_peak_detect(eng.h, eng._p(d_y), eng._p(d_mf), n, eng._p(fnan), C.byref(io), st))
    torch.cuda.synchronize()
    nt = int(bufs["n_trig"].item())
    tg = bufs["trig_idx"][:nt].cpu().numpy()
    ag = bufs["trig_ang"][:nt].cpu().numpy()
    print("peak_detect (oracle mf): n_trig %d vs %d, identical idx: %s, status %d" %
          (nt, len(r.trig), np.array_equal(tg, r.trig), int(bufs["status"].item())))
    if nt == len(r.trig) and nt:
        print("   angle max abs diff %.3g" % float(np.max(np.abs(ag - r.ang))))
    else:
        print("   gpu", tg[:10], "oracle", r.trig[:10])
    # stage 4: plan
    _lib.check(L_.ofdm_rx_plan(eng.h, n, C.byref(io), st))
    torch.cuda.synchronize()
    nf = int(bufs["n_frames"].item())
    fs = bufs["frame_start"][:nf].cpu().numpy()
    fd = bufs["frame_ndata"][:nf].cpu().numpy()
    print("plan: frames %d vs %d, starts equal %s, ndata equal %s" %
          (nf, len(r.frame_start), np.array_equal(fs, r.frame_start), np.array_equal(fd, r.n_data)))
    # stage 5+6: demod + finish on the oracle's y
    _lib.check(L_.ofdm_rx_demod(eng.h, eng._p(d_y), n, C.byref(io), st))
    _lib.check(L_.ofdm_rx_finish(eng.h, C.byref(io), st))
    torch.cuda.synchronize()
    res = eng.collect(bufs)
    print("demod (oracle y): packets %d vs %d ; ok %d vs %d ; identical list: %s" %
          (len(res.packets), len(r.packets), sum(1 for g, _ in res.packets if g), sum(1 for g, _ in r.packets if g),
           res.packets == r.packets))
    print("   live %s status %s" % (res.frame_live[:12], res.frame_status[:12]))
    nv = len(r.vec_start)
    if nv and nf == len(r.frame_start):
        eqg = bufs["eq_syms"][:nv * occ].cpu().numpy().reshape(nv, occ)
        print("   eq symbols rel-L2: %.3g" % rel(eqg, r.eq))
        # slicer decisions of demapped vectors, in order
        sidx = bufs["sym_idx"][:nv * lay.ncar].cpu().numpy().reshape(nv, lay.ncar)
        # oracle sym_log holds only demapped vectors; find which vectors those are by replaying the sink states
        dem = demapped_vectors(r, lay)
        if len(dem) == len(r.sym_idx):
            mism = sum(int(np.sum(sidx[v] != r.sym_idx[k])) for k, v in enumerate(dem))
            tot = len(dem) * lay.ncar
            print("   slicer decisions: %d mismatches / %d" % (mism, tot))
        else:
            print("   (could not align demapped vectors: %d vs %d)" % (len(dem), len(r.sym_idx)))
    # full chain on the GPU's own filter output
    res2 = eng.demodulate(d_x)
    print("full GPU chain: trig equal %s ; packets %d ok %d ; identical to oracle: %s ; counters %s" %
          (np.array_equal(res2.trig_idx, r.trig), len(res2.packets), sum(1 for g, _ in res2.packets if g),
           res2.packets == r.packets, res2.counters.tolist()))
    eng.close()


def demapped_vectors(r, lay):
    """Indices (into the sampler's vector stream) of the vectors the oracle's sink demapped."""
    sink = o.FrameSink(lay)
    out = []
    for v in range(len(r.flags)):
        before = len(sink.sym_log)
        sink.work(r.eq[v], int(r.flags[v]))
        if len(sink.sym_log) > before:
            out.append(v)
    return out


def run_sense(N, shift):
    print("=" * 100)
    print("sense N=%d shift=%s" % (N, shift))
    rng = np.random.default_rng(3)
    nfr = 240
    x = ((rng.standard_normal(nfr * N) + 1j * rng.standard_normal(nfr * N)) * 1e-3).astype(np.complex64)
    k = np.arange(nfr * N)
    x += (0.01 * np.exp(2j * np.pi * 0.13 * k)).astype(np.complex64)
    se = SenseEngine(N)
    dx = torch.from_numpy(x).cuda()
    mh = se.maxhold(dx, 2, 10, shift=shift)
    torch.cuda.synchronize()
    ref = o.sense_maxhold(x, N, 2, 10, shift=shift)
    print("maxhold shape", tuple(mh.shape), ref.shape, "rel-L2", rel(mh.cpu().numpy(), ref))
    avg, free, hx = se.decide(mh[:10], 1e-3)
    a2, f2, h2 = o.sense_decide(mh[:10].cpu().numpy(), 1e-3)
    print("decide: avg equal %s free equal %s hex equal %s" % (np.array_equal(avg, a2), np.array_equal(free, f2), hx == h2))
    sp = se.spectra(dx[:8 * N], shift=True).cpu().numpy()
    print("spectra rel-L2", rel(sp, o.sense_fft(x[:8 * N], N, True)))
    se.close()


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print(torch.cuda.get_device_name(0))
    cases = [(512, 200, 128, "bpsk", 6, 40, 0.0), (512, 200, 128, "qpsk", 24, 20, 0.3),
             (512, 200, 128, "qam16", 24, 25, -0.4), (512, 200, 128, "8psk", 12, 30, 0.2),
             (1024, 400, 256, "qam64", 12, 30, 1.3), (4096, 3200, 512, "qam256", 6, 35, 0.3),
             (256, 104, 64, "qpsk", 12, 30, 0.1), (2048, 800, 512, "qam16", 6, 30, 0.2)]
    if quick:
        cases = cases[:2]
    for c in cases:
        try:
            run_case(*c)
        except Exception:
            traceback.print_exc()
    for N, sh in ((1024, False), (512, True), (256, False), (2048, False), (4096, False), (64, False), (128, True)):
        try:
            run_sense(N, sh)
        except Exception:
            traceback.print_exc()
