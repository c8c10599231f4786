"""ctypes wrapper of oracle/libofdm_oracle.so (the C port of ofdm_oracle.py).  TEST / BENCH INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import ofdm_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libofdm_oracle.so")
_lib = None


class OcCfg(C.Structure):
    _fields_ = [("N", C.c_int32), ("occ", C.c_int32), ("cp", C.c_int32), ("M", C.c_int32),
                ("constellation", C.POINTER(C.c_float)), ("amp", C.c_float), ("pad_seed", C.c_uint64)]


STAMP = LIB + ".stamp"


def _cpu_stamp() -> str:
    """The library is built -march=native: a copy built on another CPU model (this container vs the GPU box) must be
    rebuilt, not loaded.  Stamp = hash of the host's CPU flags + the Makefile."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    with open(os.path.join(HERE, "Makefile")) as f:
        mk = f.read()
    return hashlib.sha256((flags + mk).encode()).hexdigest()[:16]


def build():
    subprocess.run(["make", "-s", "-C", HERE, "clean", "all"], check=True)
    with open(STAMP, "w") as f:
        f.write(_cpu_stamp())


def _fresh() -> bool:
    if not os.path.exists(LIB):
        return False
    if os.path.getmtime(os.path.join(HERE, "ofdm_oracle_c.c")) > os.path.getmtime(LIB):
        return False
    try:
        with open(STAMP) as f:
            return f.read().strip() == _cpu_stamp()
    except OSError:
        return False


def available() -> bool:
    """True when the C port can be used here (built for this CPU, or gcc + make are present to build it)."""
    import shutil
    return _fresh() or (shutil.which("gcc") is not None and shutil.which("make") is not None)


def lib():
    global _lib
    if _lib is None:
        if not _fresh():
            build()
        L = C.CDLL(LIB)
        L.oc_tx.restype = C.c_int64
        L.oc_tx.argtypes = [C.POINTER(OcCfg), C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
        L.oc_rx.restype = C.c_int64
        L.oc_rx.argtypes = [C.POINTER(OcCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                            C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.oc_frame_symbols.argtypes = [C.POINTER(OcCfg), C.c_int]
        L.oc_make_packet.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oc_loopback_mt.argtypes = [C.POINTER(OcCfg), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.oc_loopback_mt2.argtypes = [C.POINTER(OcCfg), C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_cfg(N, occ, cp, mod, amp=0.25, pad_seed=0):
    const = np.ascontiguousarray(o.constellation_for(mod).view(np.float32))
    cfg = OcCfg(N, occ, cp, len(const) // 2, const.ctypes.data_as(C.POINTER(C.c_float)), amp, pad_seed)
    cfg._keep = const
    return cfg


def tx(cfg, pkts, first_frame=0):
    L = lib()
    off = np.zeros(len(pkts) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in pkts], out=off[1:])
    raw = np.frombuffer(b"".join(pkts), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    nsym = sum(L.oc_frame_symbols(C.byref(cfg), len(p)) for p in pkts)
    out = np.zeros(nsym * (cfg.N + cfg.cp), dtype=np.complex64)
    n = L.oc_tx(C.byref(cfg), raw.ctypes.data, off.ctypes.data, len(pkts), first_frame, out.ctypes.data)
    return out[:n]


def rx(cfg, x, max_pkts=None):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.complex64)
    n = len(x)
    max_pkts = max_pkts or n // (cfg.N + cfg.cp) + 16
    max_trig = n // 2 + 2
    trig = np.zeros(max_trig, dtype=np.int64)
    ang = np.zeros(max_trig, dtype=np.float32)
    stride = 4096
    pb = np.zeros(max_pkts * stride, dtype=np.uint8)
    plen = np.zeros(max_pkts, dtype=np.int32)
    pok = np.zeros(max_pkts, dtype=np.uint8)
    counts = np.zeros(3, dtype=np.int64)
    npk = L.oc_rx(C.byref(cfg), x.ctypes.data, n, trig.ctypes.data, ang.ctypes.data, max_trig, pb.ctypes.data, stride,
                  plen.ctypes.data, pok.ctypes.data, max_pkts, counts.ctypes.data)
    pkts = []
    for k in range(min(int(npk), max_pkts)):
        ln = int(plen[k])
        body = pb[k * stride:k * stride + ln].tobytes()
        pkts.append((bool(pok[k]), body[:-4] if ln >= 4 else b""))
    nt = int(counts[0])
    return pkts, trig[:nt], ang[:nt], counts


def time_loopback(mod="qpsk", frames=0, snr=20.0, threads=None, cfos=None):
    """Timed loopback of the bench workload on ``threads`` host threads, one independent stream each.  ``cfos``: the
    bench capture's per-10 000-frame carrier offsets; thread t works on a piece of segment t (mod their number)."""
    L = lib()
    threads = threads or (os.cpu_count() or 1)
    cfg = make_cfg(512, 200, 128, mod)
    frames = frames or 4000                                 # per thread: ~5 s of work on every core
    out = np.zeros(6, dtype=np.float64)
    if cfos is None or len(cfos) == 0:
        cfos = [0.2]
    cf = np.ascontiguousarray([float(cfos[t % len(cfos)]) for t in range(threads)], dtype=np.float64)
    L.oc_loopback_mt2(C.byref(cfg), frames, 402, threads, snr, cf.ctypes.data, out.ctypes.data)
    samples, secs, npk, nok = out[0], out[1], out[2], out[3]
    return {"value": float(samples / secs / 1e6), "unit": "Msamples/s", "cores": int(threads), "kind": "port",
            "ms": float(secs * 1e3), "flags": "gcc -O3 -march=native -ffp-contract=off, pthreads",
            "sample": "%d threads x %d frames (%d samples in all) of the bench workload (each thread a piece of one "
                      "10 000-frame CFO segment, offsets %s) through the C port of the oracle "
                      "(oracle/ofdm_oracle_c.c, one independent stream per thread); %d/%d packets ok; "
                      "t_mod %.0f ms + t_demod %.0f ms" % (threads, frames, int(samples),
                                                            ",".join("%.3f" % v for v in cf[:4]) + ("..." if threads > 4 else ""),
                                                            int(nok), threads * frames, out[4] * 1e3, out[5] * 1e3)}
