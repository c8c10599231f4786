#!/usr/bin/env python
"""bench.py -- OFDM Msamples/s, fused mod+demod at fft_length=512 (BASELINE.json metric).

A step = one pass of the hot path over one batch: make_packets + K_TX (packets -> samples) and the whole
receive chain (samples -> packets) on a resident noisy capture of the same batch.  Workload = BASELINE
configs[1]: 512/200/128 QPSK, 1 M OFDM symbols (100 000 frames x 10 symbols, 402-byte payloads), AWGN at
20 dB, CFO ~ U(-0.5, 0.5) subcarriers redrawn every 10 000 frames.  A "sample" is one complex64 sample of the
(fft+cp)-per-symbol stream that went through both mod and demod; algorithmic traffic is 16 B per sample
(SURVEY.md section 8d).  The synthetic channel runs outside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F] [--mod qpsk]
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KERNELS_PER_STEP = 16   # make_packets, tx, chan_filter, stream_init, metric_chunk, detect_seg, seg_scan, trig_gather, plan_init,
                        # plan_local, plan_offset, demod, next, liveness_fast, liveness (general walk, idle), crc


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_label, n_samples):
    """dram__bytes_read+write of one launch of the named kernel(s) from the committed ncu capture (bytes)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            k = json.load(f)["kernels"]
        tot = 0.0
        for name, v in k.items():
            if name in kernel_label:
                tot += v["traffic_bytes"] * (n_samples / 640e6)
        return tot or None
    except Exception:
        return None


def issue_floor(n_samples, sm_mhz):
    """Instruction-issue floor of one step: warp instructions of the five stream kernels (ncu smsp__inst_executed.sum
    of the committed capture, scaled to this launch's samples) over 148 SMs x 4 schedulers x 1 instruction per clock."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            k = json.load(f)["kernels"]
        inst = sum(v["inst_executed"] for v in k.values()) * (n_samples / 640e6)
        peak = 148 * 4 * float(sm_mhz) * 1e6
        return {"warp_inst_per_step": inst, "peak_warp_inst_per_s": peak, "floor_ms": inst / peak * 1e3,
                "source": "profiles/r01_traffic.json (ncu smsp__inst_executed.sum per kernel)"}
    except Exception:
        return None


def make_payloads(n_frames, size, seed):
    """payload = struct.pack('!HH', pktno, 0) + random bytes, like benchmark_ofdm_tx.py:117."""
    rng = np.random.Generator(np.random.Philox(seed))
    body = rng.integers(0, 256, size=(n_frames, size), dtype=np.uint8)
    pktno = np.arange(n_frames, dtype=np.uint32) & 0xFFFF
    body[:, 0] = pktno >> 8
    body[:, 1] = pktno & 0xFF
    body[:, 2] = 0
    body[:, 3] = 0
    return body


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ofdm_uhd_b200.engine import OfdmEngine
    from ofdm_uhd_b200 import _lib
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    N, occ, cp = 512, 200, 128
    F = args.frames
    psize = 402
    eng = OfdmEngine(N, occ, cp, args.mod, 0.25, device=local, pad_seed=20260102 + rank, max_pkt_bytes=psize + 4 + 12)
    body = make_payloads(F, psize, 20260102 + rank)
    h_pay = torch.from_numpy(body.reshape(-1)).pin_memory()
    pay_off = np.arange(F + 1, dtype=np.int64) * psize
    d_pay = h_pay.to(dev)
    nsym = eng.frame_symbols(psize + 9)
    n_sig = F * nsym * eng.L
    lead = 2 * eng.L
    n = n_sig + 2 * lead
    seg_frames = 10000
    rng = np.random.Generator(np.random.Philox(777 + rank))
    cfos = rng.uniform(-0.5, 0.5, size=(F + seg_frames - 1) // seg_frames)

    x = torch.zeros(n, dtype=torch.complex64, device=dev)           # [lead-in | frames | tail]
    xc = torch.empty(n, dtype=torch.complex64, device=dev)
    xs = x[lead:lead + n_sig]

    plan = eng.tx_plan(pay_off, pad_for_usrp=False)

    def tx_step():
        eng.tx_run(plan, d_pay, out=xs)

    def channel_pass():
        p_sig = float((xs[:min(n_sig, 4 << 20)].abs() ** 2).mean().item())
        sigma = (p_sig / (10 ** (args.snr / 10.0)) / 2.0) ** 0.5
        phase = 0.0
        pos = 0
        seg_samples = seg_frames * nsym * eng.L
        for i, cfo in enumerate(cfos):
            lo = 0 if i == 0 else lead + i * seg_samples
            hi = n if i == len(cfos) - 1 else lead + (i + 1) * seg_samples
            eng.channel(x[lo:hi], cfo=float(cfo), sigma=sigma, seed=991 + 131 * i + rank, phase0=phase, out=xc[lo:hi])
            phase = (phase + 2 * np.pi * cfo / N * (hi - lo)) % (2 * np.pi)
            pos = hi
        return sigma

    bufs = eng.rx_alloc(n, max_frames=F + 1024)

    def rx_step():
        eng.demodulate_async(xc, bufs)

    # ---- setup: one TX + channel so that the capture exists -------------------------------------
    tx_step()
    sigma = channel_pass()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tx_step()
        rx_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        tx_step()
        ev[2 * k + 1].record()
        rx_step()
        ev[2 * k + 2].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    t_mod = float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]))
    t_dem = float(np.mean([ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)]))
    res = eng.collect(bufs, want_packets=False)
    n_ok = int(res.counters[2])

    # ---- per-kernel timing pass (for the roofline of the dominant kernel) ------------------------
    L_, st, io = eng.L_, eng._stream(), bufs["io"]
    y = eng.ws_view(bufs, 0, n)
    mf = eng.ws_view(bufs, 1, n)
    fnan = torch.zeros(1, dtype=torch.int64, device=dev)
    stages = {
        "make_packets_kernel": lambda: L_.ofdm_make_packets(eng.h, eng._p(d_pay), eng._p(plan.d_payload_off), F, 1,
                                                            eng._p(plan.pkts), eng._p(plan.d_pkt_off), st),
        "tx_kernel": lambda: L_.ofdm_tx_modulate_batch(eng.h, eng._p(plan.pkts), eng._p(plan.d_pkt_off), F, 0, None,
                                                       plan.total_syms, plan.uniform_syms, eng._p(xs), st),
        "chan_filter_kernel": lambda: L_.ofdm_rx_chan_filter(eng.h, eng._p(xc), n, eng._p(y), st),
        "metric_chunk_kernel+detect_seg_kernel(+scan,gather)": lambda: L_.ofdm_rx_sync(eng.h, eng._p(y), n, C.byref(io), st),
        "plan_kernel": lambda: L_.ofdm_rx_plan(eng.h, n, C.byref(io), st),
        "demod_kernel": lambda: L_.ofdm_rx_demod(eng.h, eng._p(y), n, C.byref(io), st),
        "liveness+crc": lambda: L_.ofdm_rx_finish(eng.h, C.byref(io), st),
    }
    kern_ms = {}
    for name, fn in stages.items():
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        kern_ms[name] = a.elapsed_time(b) / reps

    # ---- end to end through the host API: pinned host payloads in, host results out -------------
    # Two steps in flight (one stream + buffer set each), as a streaming user of the engine would run it: the
    # host->device copy of step k+1 and the device->host copy of step k-1 overlap step k's kernels.  Every step
    # still copies its own 40 MB of payloads in and its own verdicts + payload bytes out inside the timed region.
    class Lane:
        pass

    def channel_pass_fixed(x_, xc_):
        phase = 0.0
        seg_samples = seg_frames * nsym * eng.L
        for i, cfo in enumerate(cfos):
            lo = 0 if i == 0 else lead + i * seg_samples
            hi = n if i == len(cfos) - 1 else lead + (i + 1) * seg_samples
            eng.channel(x_[lo:hi], cfo=float(cfo), sigma=sigma, seed=991 + 131 * i + rank, phase0=phase, out=xc_[lo:hi])
            phase = (phase + 2 * np.pi * cfo / N * (hi - lo)) % (2 * np.pi)

    lanes = []
    for li in range(2):
        ln = Lane()
        ln.stream = torch.cuda.Stream(device=dev)
        if li == 0:
            ln.x, ln.xc, ln.bufs, ln.plan = x, xc, bufs, plan
        else:
            ln.x = torch.zeros(n, dtype=torch.complex64, device=dev)
            ln.xc = torch.empty(n, dtype=torch.complex64, device=dev)
            ln.bufs = eng.rx_alloc(n, max_frames=F + 1024, fresh=True)     # a second, independent buffer set
            ln.plan = eng.tx_plan(pay_off, pad_for_usrp=False)
        ln.xs = ln.x[lead:lead + n_sig]
        ln.d_pay = torch.empty_like(d_pay)
        ln.ticket = None
        lanes.append(ln)
    torch.cuda.synchronize()

    def e2e_launch(ln):
        with torch.cuda.stream(ln.stream):
            ln.d_pay.copy_(h_pay, non_blocking=True)
            eng.tx_run(ln.plan, ln.d_pay, out=ln.xs)
            channel_pass_fixed(ln.x, ln.xc)
            eng.demodulate_async(ln.xc, ln.bufs)
            ln.ticket = eng.collect_begin(ln.bufs, want_payload=True)

    trace = [] if os.environ.get("OFDM_E2E_TRACE") else None

    def e2e_run(k_steps):
        last = None
        for k in range(k_steps):
            t_a = time.perf_counter()
            e2e_launch(lanes[k % 2])
            t_b = time.perf_counter()
            if k > 0:
                last = eng.collect_end(lanes[(k - 1) % 2].ticket)
            if trace is not None:
                trace.append((t_b - t_a, time.perf_counter() - t_b))
        return eng.collect_end(lanes[(k_steps - 1) % 2].ticket)

    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps)
    r2 = e2e_run(e2e_steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if trace is not None:
        sys.stderr.write("rank %d e2e %.2f ms/step; per step (enqueue ms, wait ms): %s\n" % (
            rank, e2e_s * 1e3, " ".join("(%.2f,%.2f)" % (a * 1e3, b * 1e3) for a, b in trace[-e2e_steps:])))
    d2h = int(r2.payload_bytes_copied + r2.meta_bytes_copied)
    e2e_ok = int(r2.counters[2])

    # ---- reduce over ranks --------------------------------------------------------------------
    t = torch.tensor([total_ms, t_mod, t_dem, e2e_s], dtype=torch.float64, device=dev)
    cnt = torch.tensor([int(res.counters[0]), int(res.counters[1]), n_ok, n_sig], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)          # NCCL: per-rank BER/CRC statistics
    total_ms, t_mod, t_dem, e2e_s = [float(v) for v in t.tolist()]
    frames_all, msgs_all, ok_all, samples_all = [int(v) for v in cnt.tolist()]

    if rank == 0:
        peak, peak_src = peaks()
        ms_per_step = total_ms / args.steps
        value = samples_all / (ms_per_step * 1e-3) / 1e6
        dom = max(kern_ms, key=kern_ms.get)
        alg_bytes = 8.0 * n_sig          # every kernel of the chain streams the capture once: 8 B per sample
        dom_gbs = alg_bytes / (kern_ms[dom] * 1e-3) / 1e9
        traffic = ncu_traffic(dom, n_sig)
        step_gbs = 16.0 * n_sig / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "OFDM Msamples/s mod+demod (fft=512)", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: fft 512 / occ 200 / cp 128, %s, %d frames x %d symbols = %d OFDM "
                                   "symbols per GPU, 402-byte payloads, AWGN %g dB, CFO U(-0.5,0.5) per 10k frames"
                                   % (args.mod, F, nsym, F * nsym, args.snr),
                       "samples_per_gpu": n_sig, "sharding": "one independent stream per rank, no collective on the path",
                       "l2": "inputs (%.1f GB per pass) larger than L2" % (8.0 * n_sig / 1e9)},
            "ms_mod": t_mod, "ms_demod": t_dem,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                         "frac": dom_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "traffic_source": "profiles/r01_traffic.json (ncu --set full, dram read+write of one launch at "
                                           "640 M samples, scaled to this launch's samples)" if traffic else None,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "algorithmic_bytes_per_sample": 16},
            "kernels_ms": kern_ms,
            "e2e": {"value": samples_all / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h_pay.numel()),
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "crc_ok_last_step": e2e_ok,
                    "what": "pinned host payloads -> make_packets -> K_TX -> channel kernel -> receive chain -> ok flags + "
                    "payload bytes on the host (OfdmEngine API), two steps in flight on two streams; every step's "
                    "copies are inside the timed region"},
            "gpu_launches": KERNELS_PER_STEP * args.steps,
            "clocks": clocks,
            "parity": {"frames": frames_all, "messages": msgs_all, "crc_ok": ok_all, "sent": F * world},
        }
        fl = issue_floor(n_sig, (clocks or {}).get("sm_max_mhz") or 1965.0)
        if fl:
            fl["frac"] = fl["floor_ms"] / ms_per_step      # share of the step explained by pure instruction issue
            line["issue_roofline"] = fl
        line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, frames=None, threads=None):
    """The oracle port timed on this box's host cores on a bounded sample of the same workload."""
    from oracle import cpu_ref
    return cpu_ref.time_loopback(args.mod, frames or args.cpu_frames, args.snr, threads=threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import cpu_ref
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_ref.time_loopback(args.mod, args.cpu_frames, args.snr)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["ms"] for r in vals]))
    base = dict(vals[-1])
    base["value"] = v
    line = {"impl": "reference", "metric": "OFDM Msamples/s mod+demod (fft=512)", "value": v, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] layout (fft 512 / occ 200 / cp 128, %s, AWGN %g dB), bounded sample: %s"
                                   % (args.mod, args.snr, base["sample"])},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=100000)
    ap.add_argument("--mod", default="qpsk")
    ap.add_argument("--snr", type=float, default=20.0)
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the bounded CPU sample (0 = auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
