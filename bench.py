#!/usr/bin/env python
"""bench.py -- OFDM Msamples/s, fused mod+demod at fft_length=512 (BASELINE.json metric).

A step = one pass of the hot path over one batch: make_packets + K_TX (packets -> samples) and the whole
receive chain (samples -> packets) on a resident noisy capture of the same batch.  Workload = BASELINE
configs[1]: 512/200/128 QPSK, 1 M OFDM symbols (100 000 frames x 10 symbols, 402-byte payloads), AWGN at
20 dB, CFO ~ U(-0.5, 0.5) subcarriers redrawn every 10 000 frames.  A "sample" is one complex64 sample of the
(fft+cp)-per-symbol stream that went through both mod and demod; algorithmic traffic is 16 B per sample
(SURVEY.md section 8d).  The synthetic channel runs outside the timed regions (it is test infrastructure, not a
reference block; its noise is a pure function of seed and sample index, so the resident capture IS the channel
output of every step's transmit signal).

The same JSON line carries `other_configs`: BASELINE configs[2] (64 streams, 1024/400/256 QAM64, 4096 frames
each, 30 dB, per-stream CFO, streams s mod G across the ranks, through the batched entry points) and configs[4]
(4096/3200/512 QAM256, 4091-byte payloads, 35 dB, one stream per GPU), each with its own roofline fraction and
CRC-ok count; and `parity.oracle_prefix_equal`: the receiver's (ok, payload) list on windows of the bench capture
against the C port of the oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F] [--mod qpsk]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# make_packets, tx, chan_filter, stream_init, metric_chunk, detect_seg, seg_scan, trig_gather, plan_init, plan_local,
# plan_offset, acq, sink, next, liveness_fast, liveness (general walk, idle), crc
KERNELS_PER_STEP = 17
PROFILE = "r02b_traffic.json"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _profile():
    for name in (PROFILE, "r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                return json.load(f), name
    return None, None


def ncu_traffic(kernel_label, n_samples):
    """dram__bytes_read+write of one launch of the named kernel(s) from the committed ncu capture (bytes)."""
    prof, name = _profile()
    if not prof:
        return None, None
    tot = 0.0
    for k, v in prof["kernels"].items():
        if k.replace("_warp", "") in kernel_label:      # tx_warp_kernel etc.: the warp-plan kernels behind the same stage labels
            tot += v["traffic_bytes"] * (n_samples / 640e6)
    return (tot or None), name


def issue_floor(n_samples, sm_mhz, sms):
    """Instruction-issue floor of one step: warp instructions of the five stream kernels (ncu smsp__inst_executed.sum
    of the committed capture, scaled to this launch's samples) over SMs x 4 schedulers x 1 instruction per clock."""
    prof, name = _profile()
    if not prof:
        return None
    inst = sum(v["inst_executed"] for v in prof["kernels"].values()) * (n_samples / 640e6)
    peak = sms * 4 * float(sm_mhz) * 1e6
    return {"warp_inst_per_step": inst, "peak_warp_inst_per_s": peak, "floor_ms": inst / peak * 1e3,
            "source": "profiles/%s (ncu smsp__inst_executed.sum per kernel)" % name}


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


def bench_cfos(F, rank, seg_frames=10000):
    rng = np.random.Generator(np.random.Philox(777 + rank))
    return rng.uniform(-0.5, 0.5, size=(F + seg_frames - 1) // seg_frames)


def bench_config(args):
    """The workload both arms are quoted on (identical in the b200 and the reference line)."""
    nsym = 1 + -(-8 * 411 // (198 * {"bpsk": 1, "qpsk": 2, "8psk": 3, "qam16": 4, "qam64": 6, "qam256": 8}[args.mod]))
    n_sig = args.frames * nsym * 640
    return {"workload": "BASELINE configs[1]: fft 512 / occ 200 / cp 128, %s, %d frames x %d symbols = %d OFDM "
                        "symbols per GPU, 402-byte payloads, AWGN %g dB, CFO U(-0.5,0.5) per 10k frames"
                        % (args.mod, args.frames, nsym, args.frames * nsym, args.snr),
            "samples_per_gpu": n_sig, "sharding": "one independent stream per rank, no collective on the path",
            "l2": "inputs (%.1f GB per pass) larger than L2" % (8.0 * n_sig / 1e9)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML names as local to its GPU, so that the pinned staging buffers of the end-to-end
    leg are first-touched on that GPU's NUMA node (eight ranks on one socket's memory halve each other's copy rate)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


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


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] and configs[4]: many streams through the batched entry points
# ---------------------------------------------------------------------------------------------------------------
def run_stream_config(torch, dist, rank, world, dev, name, N, occ, cp, mod, psize, frames, n_streams_total, snr, cfo_max,
                      steps, warmup, seed, what):
    """S independent streams ([lead | frames | tail] each), streams s with s mod world == rank on this rank; one step =
    make_packets + ofdm_tx_modulate_streams + ofdm_rx_demodulate_batch over all of this rank's streams."""
    from ofdm_uhd_b200.engine import OfdmEngine
    mine = [s for s in range(n_streams_total) if s % world == rank]
    S = len(mine)
    eng = OfdmEngine(N, occ, cp, mod, 0.25, device=dev.index, pad_seed=seed, max_pkt_bytes=min(4096, psize + 4 + 12))
    nsym = eng.frame_symbols(psize + 9)
    L = eng.L
    n_sig = frames * nsym * L
    lead = 2 * L
    n_stream = n_sig + 2 * lead                              # a multiple of 4 samples for every BASELINE layout
    soff = np.arange(S + 1, dtype=np.int64) * n_stream
    n_total = int(soff[-1])
    body = np.concatenate([make_payloads(frames, psize, seed + 1000 * s).reshape(-1) for s in mine]) if S else np.zeros(0, np.uint8)
    d_pay = torch.from_numpy(body).to(dev)
    pay_off = np.arange(S * frames + 1, dtype=np.int64) * psize
    plan = eng.tx_plan(pay_off, pad_for_usrp=False, stream_frame0=np.arange(S + 1, dtype=np.int64) * frames,
                       stream_out_off=soff[:-1] + lead)
    x = torch.zeros(n_total, dtype=torch.complex64, device=dev)
    xc = torch.empty(n_total, dtype=torch.complex64, device=dev)
    eng.tx_run(plan, d_pay, out=x)
    p_sig = float((x[lead:lead + min(n_sig, 4 << 20)].abs() ** 2).mean().item()) if S else 1.0
    sigma = (p_sig / (10 ** (snr / 10.0)) / 2.0) ** 0.5
    rng = np.random.Generator(np.random.Philox(seed))
    cfos = rng.uniform(-cfo_max, cfo_max, size=n_streams_total)
    for k, s in enumerate(mine):
        eng.channel(x[soff[k]:soff[k + 1]], cfo=float(cfos[s]), sigma=sigma, seed=seed + 31 * s, out=xc[soff[k]:soff[k + 1]])
    bufs = eng.rx_alloc_batch(soff, max_frames=frames + 256) if S else None
    torch.cuda.synchronize()

    def step():
        if S:
            eng.tx_run(plan, d_pay, out=x)
            eng.demodulate_batch_async(xc, bufs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    cnt = np.zeros(4, dtype=np.int64)                        # frames, messages, crc ok, samples
    if S:
        c = bufs["counters"].cpu().numpy().reshape(S, 8)
        cnt[:] = [int(c[:, 0].sum()), int(c[:, 1].sum()), int(c[:, 2].sum()), S * n_sig]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    k = torch.from_numpy(cnt).to(dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(k, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    frames_all, msgs_all, ok_all, samples_all = [int(v) for v in k.tolist()]
    peak, _ = peaks()
    gbs = 16.0 * samples_all / world / (ms * 1e-3) / 1e9      # per GPU
    out = {"workload": what, "value": samples_all / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": world,
           "ms_per_step": ms, "steps": steps, "warmup": warmup, "streams": n_streams_total, "streams_per_rank": S if world == 1 else "%d..%d" % (n_streams_total // world, -(-n_streams_total // world)),
           "samples_per_stream": n_sig, "scaling": "strong" if n_streams_total > world else "weak",
           "roofline_step": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "algorithmic_bytes_per_sample": 16},
           "parity": {"sent": n_streams_total * frames, "frames": frames_all, "messages": msgs_all, "crc_ok": ok_all},
           "gpu_launches": KERNELS_PER_STEP * steps, "snr_db": snr}
    del x, xc, bufs, d_pay, plan
    eng.close()
    torch.cuda.empty_cache()
    return out


def oracle_prefix_parity(torch, eng, xc, lead, frame_samples, n_frames_total, mod, windows, frames_per_window):
    """GPU receiver == C port of the oracle on windows of the bench capture (each window a stream of its own for both
    sides): the (ok, payload) lists must be identical.  Oracle windows run on host threads in parallel."""
    from oracle import c_port
    cfg_of = lambda: c_port.make_cfg(512, 200, 128, mod)
    starts = [int(k * max(1, (n_frames_total - frames_per_window) // max(1, windows - 1))) for k in range(windows)] if windows > 1 else [0]
    starts = sorted(set(min(s, max(0, n_frames_total - frames_per_window)) for s in starts))
    caps = []
    for f0 in starts:
        a = 0 if f0 == 0 else lead + f0 * frame_samples
        b = min(int(xc.numel()), lead + (f0 + frames_per_window) * frame_samples + lead)
        caps.append(xc[a:b].clone())
    # full-width packet slots for this leg: a bogus header may announce up to 4095 bytes, and the comparison is on
    # complete messages (the timed engine keeps only what a 402-byte payload needs)
    from ofdm_uhd_b200.engine import OfdmEngine
    eng_x = OfdmEngine(eng.N, eng.occ, eng.cp, mod, 0.25, device=eng.device, max_pkt_bytes=4096)
    got = [eng_x.demodulate(c, max_frames=frames_per_window + 64).packets for c in caps]
    eng_x.close()
    host = [c.cpu().numpy() for c in caps]
    ref = [None] * len(caps)

    def work(i):
        ref[i] = c_port.rx(cfg_of(), host[i], max_pkts=frames_per_window + 64)[0]

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(caps))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    secs = time.perf_counter() - t0
    equal = [got[i] == ref[i] for i in range(len(caps))]
    verdicts = all([ok for ok, _ in got[i]] == [ok for ok, _ in ref[i]] for i in range(len(caps)))
    good = all(all(p == q for (ok, p), (_, q) in zip(got[i], ref[i]) if ok) for i in range(len(caps)))
    return {"oracle_prefix_equal": bool(all(equal)), "oracle_windows_equal": equal, "oracle_crc_verdicts_equal": bool(verdicts),
            "oracle_good_payloads_equal": bool(good), "oracle_windows": len(caps),
            "oracle_frames_per_window": frames_per_window,
            "oracle_window_first_frames": starts, "oracle_messages": int(sum(len(r) for r in ref)),
            "oracle_crc_ok": int(sum(1 for r in ref for ok, _ in r if ok)),
            "gpu_messages_on_windows": int(sum(len(g) for g in got)), "oracle_seconds": secs,
            "oracle": "oracle/ofdm_oracle_c.c (C port of ofdm_oracle.py), one host thread per window"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ofdm_uhd_b200.engine import OfdmEngine
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sms = torch.cuda.get_device_properties(local).multi_processor_count

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
    cfos = bench_cfos(F, rank, seg_frames)

    x = torch.zeros(n, dtype=torch.complex64, device=dev)           # [lead-in | frames | tail]
    xc = torch.empty(n, dtype=torch.complex64, device=dev)
    xs = x[lead:lead + n_sig]

    plan = eng.tx_plan(pay_off, pad_for_usrp=False)

    def tx_step():
        eng.tx_run(plan, d_pay, out=xs)

    def channel_pass(x_, xc_, sigma):
        phase = 0.0
        seg_samples = seg_frames * nsym * eng.L
        for i, cfo in enumerate(cfos):
            lo = 0 if i == 0 else lead + i * seg_samples
            hi = n if i == len(cfos) - 1 else lead + (i + 1) * seg_samples
            eng.channel(x_[lo:hi], cfo=float(cfo), sigma=sigma, seed=991 + 131 * i + rank, phase0=phase, out=xc_[lo:hi])
            phase = (phase + 2 * np.pi * cfo / N * (hi - lo)) % (2 * np.pi)

    bufs = eng.rx_alloc(n, max_frames=F + 1024)

    def rx_step():
        eng.demodulate_async(xc, bufs)

    # ---- setup: one TX + channel so that the capture exists -------------------------------------
    tx_step()
    p_sig = float((xs[:min(n_sig, 4 << 20)].abs() ** 2).mean().item())
    sigma = (p_sig / (10 ** (args.snr / 10.0)) / 2.0) ** 0.5
    channel_pass(x, xc, sigma)
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
    stages = {
        "make_packets_kernel": lambda: L_.ofdm_make_packets(eng.h, eng._p(d_pay), eng._p(plan.d_payload_off), F, 1,
                                                            eng._p(plan.pkts), eng._p(plan.d_pkt_off), st),
        "tx_kernel": lambda: L_.ofdm_tx_modulate_batch(eng.h, eng._p(plan.pkts), eng._p(plan.d_pkt_off), F, 0, None,
                                                       plan.total_syms, plan.uniform_syms, eng._p(xs), st),
        "chan_filter_kernel": lambda: L_.ofdm_rx_chan_filter(eng.h, eng._p(xc), n, eng._p(y), st),
        "metric_chunk_kernel": lambda: L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 0, st),
        "detect_seg_kernel": lambda: L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 1, st),
        "seg_scan+trig_gather": lambda: L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 2, st),
        "plan_kernel": lambda: L_.ofdm_rx_plan(eng.h, n, C.byref(io), st),
        "acq_kernel": lambda: L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 3, st),
        "sink_kernel": lambda: L_.ofdm_rx_stage(eng.h, eng._p(y), n, C.byref(io), 4, st),
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
    # host->device copy of step k+1 and the device->host copy of step k-1 overlap step k's kernels.  Every step copies
    # its own 40 MB of payloads in and its own verdicts + payload bytes (dense hand-over, ofdm_rx_compact, sized by
    # what was transmitted) out inside the timed region; the receive chain runs on the lane's resident noisy capture
    # of the very signal the step's transmit pass regenerates (deterministic payloads and noise).
    class Lane:
        pass

    lanes = []
    for li in range(2):
        ln = Lane()
        ln.stream = torch.cuda.Stream(device=dev)
        if li == 0:
            ln.x, ln.xc, ln.bufs, ln.plan = x, xc, bufs, plan
        else:
            ln.x = torch.zeros(n, dtype=torch.complex64, device=dev)
            ln.xc = xc                                                      # the capture is read-only: shared
            ln.bufs = eng.rx_alloc(n, max_frames=F + 1024, fresh=True)     # a second, independent buffer set
            ln.plan = eng.tx_plan(pay_off, pad_for_usrp=False)
        ln.xs = ln.x[lead:lead + n_sig]
        ln.d_pay = torch.empty_like(d_pay)
        ln.ticket = None
        lanes.append(ln)
    torch.cuda.synchronize()
    exp_bytes = F * (psize + 4)

    compute_done = [None]              # event after the kernels of the step launched last (on the other lane)

    def e2e_launch(ln):
        with torch.cuda.stream(ln.stream):
            ln.d_pay.copy_(h_pay, non_blocking=True)
            # classic double buffering: this step's kernels start when the previous step's kernels are done, so the two
            # lanes never interleave their kernels (they would then finish together and leave the GPU idle during both
            # lanes' copies); the copies of one lane overlap the kernels of the other
            if compute_done[0] is not None:
                ln.stream.wait_event(compute_done[0])
            eng.tx_run(ln.plan, ln.d_pay, out=ln.xs)
            eng.demodulate_async(ln.xc, ln.bufs)
            ev = torch.cuda.Event()
            ev.record(ln.stream)
            compute_done[0] = ev
            ln.ticket = eng.deliver_begin(ln.bufs, expect_msgs=F, expect_bytes=exp_bytes)

    trace = [] if os.environ.get("OFDM_E2E_TRACE") else None

    def e2e_run(k_steps):
        for k in range(k_steps):
            t_a = time.perf_counter()
            e2e_launch(lanes[k % 2])
            t_b = time.perf_counter()
            if k > 0:
                eng.deliver_end(lanes[(k - 1) % 2].ticket)
            if trace is not None:
                trace.append((t_b - t_a, time.perf_counter() - t_b))
        return eng.deliver_end(lanes[(k_steps - 1) % 2].ticket)

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
    d2h = int(r2["d2h_bytes"])
    e2e_ok = int(r2["ok"].sum())
    e2e_msgs = int(r2["n_msgs"])

    # ---- the reference-shaped per-packet surface: transmit_path.send_pkt -> ofdm_demod callback ----
    surface = None
    if rank == 0 and not args.no_surface:
        try:
            surface = packet_surface(torch, eng, args, local)
        except Exception as e:                                 # a side leg must not take the measurement down
            surface = {"error": repr(e)}

    # ---- parity against the oracle (C port) on windows of the capture (rank 0) --------------------
    parity_oracle = {}
    if rank == 0 and args.oracle_frames > 0:
        try:
            parity_oracle = oracle_prefix_parity(torch, eng, xc, lead, nsym * eng.L, F, args.mod, args.oracle_windows,
                                                 min(args.oracle_frames, F))
        except Exception as e:                                 # the checker must not take the measurement down
            parity_oracle = {"oracle_prefix_equal": None, "oracle_error": repr(e)}

    # ---- reduce over ranks --------------------------------------------------------------------
    t = torch.tensor([total_ms, t_mod, t_dem, e2e_s], dtype=torch.float64, device=dev)
    cnt = torch.tensor([int(res.counters[0]), int(res.counters[1]), n_ok, n_sig], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)          # NCCL: per-rank BER/CRC statistics
    total_ms, t_mod, t_dem, e2e_s = [float(v) for v in t.tolist()]
    frames_all, msgs_all, ok_all, samples_all = [int(v) for v in cnt.tolist()]

    # ---- the other BASELINE configurations (driver-visible, same JSON line) ----------------------
    del lanes, x, xc, xs, bufs, y, stages, io, plan, d_pay, res
    eng.close()
    torch.cuda.empty_cache()
    other = {}
    if not args.no_other:
        k3 = max(2, min(args.steps, 5))
        other["configs[2]"] = run_stream_config(
            torch, dist, rank, world, dev, "cfg3", 1024, 400, 256, "qam64", 402, args.cfg3_frames, 64, 30.0, 1.5, k3, 2,
            20260103, "BASELINE configs[2]: 64 independent streams x fft 1024 / occ 400 / cp 256, QAM64, %d frames per stream "
            "(402-byte payloads), AWGN 30 dB, per-stream CFO U(-1.5,1.5); streams s mod G on rank s; batched entry points "
            "(one launch sequence for all of a rank's streams)" % args.cfg3_frames)
        other["configs[4]"] = run_stream_config(
            torch, dist, rank, world, dev, "cfg5", 4096, 3200, 512, "qam256", 4091, args.cfg5_frames, world, 35.0, 0.5, k3, 2,
            20260105, "BASELINE configs[4]: fft 4096 / occ 3200 / cp 512, QAM256, 4091-byte payloads (3 symbols per frame), "
            "%d frames per stream, AWGN 35 dB, CFO U(-0.5,0.5), one stream per GPU" % args.cfg5_frames)

    if rank == 0:
        peak, peak_src = peaks()
        ms_per_step = total_ms / args.steps
        value = samples_all / (ms_per_step * 1e-3) / 1e6
        dom = max(kern_ms, key=kern_ms.get)
        # algorithmic bytes per launch: 8 B per sample -- every stream kernel sweeps the capture (or its
        # filtered copy) once.  (The channel filter must also WRITE the 8 B/sample stream it exists to produce; that
        # second half is not counted here.)
        alg_bytes = 8.0 * n_sig
        dom_gbs = alg_bytes / (kern_ms[dom] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(dom, n_sig)
        step_gbs = 16.0 * n_sig / (ms_per_step * 1e-3) / 1e9
        parity = {"frames": frames_all, "messages": msgs_all, "crc_ok": ok_all, "sent": F * world}
        parity.update(parity_oracle)
        line = {
            "metric": "OFDM Msamples/s mod+demod (fft=512)", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args),
            "ms_mod": t_mod, "ms_demod": t_dem,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                         "frac": dom_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "traffic_source": ("profiles/%s (ncu --set full, dram read+write of one launch at 640 M samples, "
                                            "scaled to this launch's samples)" % traffic_src) if traffic else None,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "algorithmic_bytes_per_sample": 16},
            "kernels_ms": kern_ms,
            "e2e": {"value": samples_all / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h_pay.numel()),
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "crc_ok_last_step": e2e_ok, "messages_last_step": e2e_msgs,
                    "what": "pinned host payloads -> H2D -> make_packets -> K_TX -> receive chain on the resident noisy capture "
                    "of that transmit signal -> ofdm_rx_compact -> D2H of the dense payload bytes + bit-packed CRC verdicts + "
                    "offsets (OfdmEngine API), two steps in flight on two streams; every step's copies are inside the timed "
                    "region; the synthetic AWGN/CFO channel (test infrastructure, deterministic) is applied once outside it"},
            "gpu_launches": KERNELS_PER_STEP * args.steps,
            "clocks": clocks,
            "parity": parity,
        }
        if surface:
            line["packet_surface"] = surface
        if other:
            line["other_configs"] = other
        fl = issue_floor(n_sig, (clocks or {}).get("sm_max_mhz") or 1965.0, sms)
        if fl:
            fl["frac"] = fl["floor_ms"] / ms_per_step      # share of the step explained by pure instruction issue
            line["issue_roofline"] = fl
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(args, cfos=cfos)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def packet_surface(torch, eng_unused, args, device):
    """Throughput of the reference-shaped surface: transmit_path.send_pkt(payload) per packet -> ofdm_mod -> samples ->
    ofdm_demod.feed -> rx_callback(ok, payload) per packet (benchmark_ofdm_tx.py:57-126, benchmark_ofdm_rx.py:50-61), and
    of its bulk twin send_pkts(list) / rx_callback_batch."""
    from types import SimpleNamespace
    try:
        from ofdm_uhd_b200 import transmit_path as tp_mod, receive_path as rp_mod
    except Exception as e:
        return {"error": repr(e)}
    F = args.surface_frames
    opt = SimpleNamespace(modulation=args.mod, fft_length=512, occupied_tones=200, cp_length=128, verbose=False, log=False,
                          snr=args.snr, tx_amplitude=0.25, samples_per_symbol=2)
    body = make_payloads(F, 402, 4242)
    pays = [bytes(r) for r in body]
    out = {"frames": F}
    got = []
    from ofdm_uhd_b200 import channel_model
    rx = rp_mod.receive_path(lambda ok, p: got.append(ok), opt, device=device, max_pkt_bytes=416)
    tx = tp_mod.transmit_path(opt, device=device, batch_limit=F + 1)
    chan = channel_model.channel_model(tx.ofdm_tx._engine, noise_voltage=0.0108, frequency_offset=0.2, seed=4,
                                       lead_in=1280, tail=2560)          # ~20 dB below the 0.25-amplitude signal
    tx.connect(chan)
    chan.connect(rx)
    for timed in (False, True):                        # first pass: allocations (workspace, pinned staging)
        del got[:]
        t0 = time.perf_counter()
        for p in pays:
            tx.send_pkt(p)
        tx.send_pkt(eof=True)
        rx.wait(timeout=120)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    samples = F * 10 * 640 if args.mod == "qpsk" else None
    out["per_packet"] = {"seconds": dt, "packets_per_s": F / dt, "Msamples_per_s": (samples / dt / 1e6) if samples else None,
                         "callbacks": len(got), "ok": int(sum(1 for g in got if g)),
                         "what": "send_pkt(payload) per packet (host make_packet) -> flush -> feed -> callback per packet"}
    if hasattr(tx, "send_pkts") and hasattr(rx, "set_batch_callback"):
        batches = []
        rx.set_batch_callback(lambda oks, blob, off: batches.append(int(np.count_nonzero(oks))))
        for timed in (False, True):
            del batches[:]
            t0 = time.perf_counter()
            tx.send_pkts(pays)
            tx.send_pkt(eof=True)
            rx.wait(timeout=120)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out["bulk"] = {"seconds": dt, "packets_per_s": F / dt, "Msamples_per_s": (samples / dt / 1e6) if samples else None,
                       "ok": int(sum(batches)),
                       "what": "send_pkts(list) (make_packets_kernel on the device) -> feed -> rx_callback_batch(ok[], bytes, offsets)"}
    # a continuous source delivered in radio-sized buffers: feed_stream with batching vs one whole-stream feed
    try:
        n_frames = 20000
        body2 = make_payloads(n_frames, 402, 777)
        eng2 = tx.ofdm_tx._engine
        plan = eng2.tx_plan(np.arange(n_frames + 1, dtype=np.int64) * 402)
        xs = eng2.tx_run(plan, torch.from_numpy(body2.reshape(-1)).to(eng2.dev))
        cap = chan.process(xs)
        cnt = [0]
        rx2 = rp_mod.receive_path(lambda ok, p: None, opt, device=device, max_pkt_bytes=416)
        rx2.set_batch_callback(lambda oks, blob, off: cnt.__setitem__(0, cnt[0] + int(np.count_nonzero(oks))))
        for timed in (False, True):
            cnt[0] = 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rx2.feed(cap)
            torch.cuda.synchronize()
            dt_whole = time.perf_counter() - t0
        whole_ok = cnt[0]
        res = {}
        for buf in (65536, 1 << 20):
            rx3 = rp_mod.receive_path(lambda ok, p: None, opt, device=device, max_pkt_bytes=416)
            got3 = [0]
            rx3.ofdm_rx.stream_batch_samples = 1 << 25
            rx3.set_batch_callback(lambda oks, blob, off: got3.__setitem__(0, got3[0] + int(np.count_nonzero(oks))))
            chunks = [cap[a:a + buf] for a in range(0, cap.numel(), buf)]     # the source's buffers, ready made
            for timed in (False, True):
                got3[0] = 0
                rx3.ofdm_rx.reset_stream()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for c in chunks:
                    rx3.feed_stream(c)
                rx3.flush_stream()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            res["%d-sample buffers" % buf] = {"seconds": dt, "Msamples_per_s": cap.numel() / dt / 1e6, "ok": got3[0]}
        out["feed_stream"] = {"samples": int(cap.numel()), "whole_stream_feed": {"seconds": dt_whole, "Msamples_per_s": cap.numel() / dt_whole / 1e6, "ok": whole_ok},
                              "batched": res, "stream_batch_samples": 1 << 25,
                              "what": "ofdm_demod.feed_stream on consecutive buffers of one capture (queued on the device, one receiver pass per "
                                      "stream_batch_samples new samples + the carried tail, passes alternating between two CUDA streams) vs one "
                                      "feed() of the whole capture; the buffers are handed over ready made (device tensors), host-side packet "
                                      "assembly included"}
    except Exception as e:
        out["feed_stream"] = {"error": repr(e)}
    return out


def cpu_baseline(args, frames=None, threads=None, cfos=None):
    """The oracle port timed on this box's host cores on a bounded sample of the same workload."""
    from oracle import cpu_ref
    return cpu_ref.time_loopback(args.mod, frames or args.cpu_frames, args.snr, threads=threads, cfos=cfos)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import cpu_ref
    cfos = bench_cfos(args.frames, 0)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_ref.time_loopback(args.mod, args.cpu_frames, args.snr, cfos=cfos)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["ms"] for r in vals]))
    base = dict(vals[-1])
    base["value"] = v
    line = {"impl": "reference", "metric": "OFDM Msamples/s mod+demod (fft=512)", "value": v, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args),
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
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per thread in the bounded CPU sample (0 = auto)")
    ap.add_argument("--cfg3-frames", type=int, default=4096, help="frames per stream of the configs[2] leg")
    ap.add_argument("--cfg5-frames", type=int, default=20000, help="frames per stream of the configs[4] leg")
    ap.add_argument("--oracle-frames", type=int, default=2000, help="frames per window of the oracle parity leg (0 = off)")
    ap.add_argument("--oracle-windows", type=int, default=5)
    ap.add_argument("--surface-frames", type=int, default=2000)
    ap.add_argument("--no-other", action="store_true", help="skip the configs[2] / configs[4] legs")
    ap.add_argument("--no-surface", action="store_true", help="skip the per-packet surface leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
