"""Device engine: owns an ``ofdm_handle`` and the torch buffers handed to libofdm_b200.so.

This is the host-side glue the reference-facing classes in ofdm.py share.  It never computes
samples or bits itself -- every stage is a CUDA kernel behind the C ABI (include/ofdm_b200.h).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

try:
    from . import _lib, psk, qam
except ImportError:                      # flat import (directory on sys.path, like the reference)
    import _lib
    import psk
    import qam

MODS = {"bpsk": 2, "qpsk": 4, "8psk": 8, "qam8": 8, "qam16": 16, "qam64": 64, "qam256": 256}   # ofdm.py:88


def rotated_constellation(modulation: str) -> List[complex]:
    """The constellation exactly as the reference builds it (ofdm.py:88-101)."""
    arity = MODS[modulation]                          # KeyError for unknown names, like the reference
    rot = 1
    if modulation == "qpsk":
        rot = (0.707 + 0.707j)
    if modulation.find("psk") >= 0:
        return [pt * rot for pt in psk.gray_constellation[arity]]
    return [pt * rot for pt in qam.constellation[arity]]


@dataclass
class TxPlan:
    """Offsets of one batch of frames, resident on the device (built once, reused every step)."""
    n_frames: int
    payload_off: np.ndarray
    pkt_off: np.ndarray
    d_payload_off: object
    d_pkt_off: object
    d_sym_off: object
    pkts: object                 # uint8 cuda tensor holding the framed packets
    total_syms: int
    uniform_syms: int
    n_samples: int
    pad_for_usrp: bool
    # several streams in one batch (ofdm_tx_modulate_streams): frames [stream_frame0[s], stream_frame0[s+1]) go to
    # out[stream_out_off[s]:]; None for the classic back-to-back batch
    n_streams: int = 0
    d_stream_frame0: object = None
    d_stream_out_off: object = None


@dataclass
class RxBatch:
    """Host view of one ofdm_rx_demodulate call."""
    n_trig: int
    n_frames: int
    status_bits: int
    trig_idx: np.ndarray
    trig_ang: np.ndarray
    frame_start: np.ndarray
    frame_ndata: np.ndarray
    frame_live: np.ndarray
    frame_status: np.ndarray
    pkt_len: np.ndarray
    pkt_ok: np.ndarray
    counters: np.ndarray
    packets: List[Tuple[bool, bytes]]          # (ok, payload) in arrival order, what the callback sees
    msg_frames: Optional[np.ndarray] = None    # frame index of every delivered message
    payload_rows: Optional[np.ndarray] = None  # uint8 [n_frames, pkt_stride] dewhitened payload || crc, by frame slot
    payload_bytes_copied: int = 0
    meta_bytes_copied: int = 0


class OfdmEngine:
    def __init__(self, fft_length=512, occupied_tones=200, cp_length=128, modulation="bpsk", tx_amplitude=0.25,
                 device: Optional[int] = None, pad_seed: int = 0, max_pkt_bytes: int = 4096,
                 carrier_map: Optional[str] = None):
        import torch
        self.torch = torch
        self.L_ = _lib.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        self.modulation = modulation
        const = np.array(rotated_constellation(modulation), dtype=np.complex128).astype(np.complex64)
        self.constellation = const
        flat = np.ascontiguousarray(const.view(np.float32))
        cfg = _lib.OfdmCfg(fft_length, occupied_tones, cp_length, len(const),
                           flat.ctypes.data_as(C.POINTER(C.c_float)), float(tx_amplitude), self.device,
                           int(pad_seed) & 0xFFFFFFFFFFFFFFFF, int(max_pkt_bytes),
                           carrier_map.encode("ascii") if carrier_map else None)
        self.carrier_map = carrier_map or "FE7F"
        h = self.L_.ofdm_create(C.byref(cfg))
        if not h:
            raise ValueError("ofdm_create: " + self.L_.ofdm_last_error().decode())
        self.h = C.c_void_p(h)
        lay = (C.c_int32 * 8)()
        _lib.check(self.L_.ofdm_get_layout(self.h, lay), "get_layout")
        self.N, self.occ, self.cp = fft_length, occupied_tones, cp_length
        self.zl, self.ncar, self.nbits, self.L, self.ntaps, self.nos, self.pkt_stride = [int(v) for v in lay[:7]]
        self._ws = None
        self._ws_key = None

    def close(self):
        if getattr(self, "h", None):
            self.L_.ofdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def set_tx_amplitude(self, ampl: float):
        _lib.check(self.L_.ofdm_set_tx_amplitude(self.h, float(ampl)))

    def chan_taps(self) -> np.ndarray:
        buf = (C.c_float * 512)()
        n = _lib.check(self.L_.ofdm_get_chan_taps(self.h, buf, 512))
        return np.array(buf[:n], dtype=np.float32)

    def frame_symbols(self, pkt_len: int) -> int:
        return int(self.L_.ofdm_frame_symbols(self.h, int(pkt_len)))

    # ------------------------------------------------------------------ transmit
    @staticmethod
    def _packet_lengths(plen: np.ndarray, pad_for_usrp: bool) -> np.ndarray:
        """Framed length of every packet; ValueError where the reference's make_packet raises one
        (ofdm_packet_utils.py:117-135): the whitened body payload || crc32 || 0x55 [|| padding] must fit the
        4096-byte PN table, i.e. len(payload) <= 4091 (4087 with pad_for_usrp)."""
        klen = plen + 9
        if pad_for_usrp:
            klen = (klen + 15) // 16 * 16
        if len(klen) and int(klen.max()) - 4 > 4096:
            raise ValueError("len(payload) must be in [0, %d]: payload || crc32 || 0x55%s exceeds the 4096-byte "
                             "whitening table" % (4087 if pad_for_usrp else 4091, " || padding" if pad_for_usrp else ""))
        return klen

    def tx_plan(self, payload_off: np.ndarray, pad_for_usrp: bool = False, stream_frame0=None,
                stream_out_off=None) -> TxPlan:
        """Everything about a batch that does not depend on the payload bytes.  With ``stream_frame0`` (int64 [S+1],
        frame index where each stream starts) and ``stream_out_off`` (int64 [S], sample offset of each stream's first
        symbol in the output buffer) the batch is the transmit side of S independent streams."""
        torch = self.torch
        payload_off = np.ascontiguousarray(payload_off, dtype=np.int64)
        plen = np.diff(payload_off)
        F = len(plen)
        klen = self._packet_lengths(plen, pad_for_usrp)
        pkt_off = np.zeros(F + 1, dtype=np.int64)
        np.cumsum(klen, out=pkt_off[1:])
        per = self.ncar * self.nbits
        nsym = 1 + np.maximum(1, -(-(8 * klen) // per))
        uniform = int(nsym[0]) if F and bool((nsym == nsym[0]).all()) else 0
        total = int(nsym.sum())
        d_sym_off = None
        if not uniform and F:
            sym_off = np.zeros(F + 1, dtype=np.int64)
            np.cumsum(nsym, out=sym_off[1:])
            d_sym_off = torch.from_numpy(sym_off).to(self.dev)
        plan = TxPlan(F, payload_off, pkt_off, torch.from_numpy(payload_off).to(self.dev),
                      torch.from_numpy(pkt_off).to(self.dev), d_sym_off,
                      torch.empty(int(pkt_off[-1]), dtype=torch.uint8, device=self.dev), total, uniform,
                      total * self.L, bool(pad_for_usrp))
        if stream_frame0 is not None:
            sf = np.ascontiguousarray(stream_frame0, dtype=np.int64)
            so = np.ascontiguousarray(stream_out_off, dtype=np.int64)
            if len(sf) != len(so) + 1 or sf[0] != 0 or sf[-1] != F or (np.diff(sf) < 0).any():
                raise ValueError("tx_plan: stream_frame0 must ascend from 0 to the number of frames, one entry per stream + 1")
            sym0 = np.concatenate([[0], np.cumsum(nsym)])[sf]
            plan.n_streams = len(so)
            plan.d_stream_frame0 = torch.from_numpy(sf).to(self.dev)
            plan.d_stream_out_off = torch.from_numpy(so).to(self.dev)
            plan.n_samples = int((so + (sym0[1:] - sym0[:-1]) * self.L).max()) if len(so) else 0
        return plan

    def tx_run(self, plan: TxPlan, payload, out=None, first_frame: int = 0, whitening: bool = True):
        """make_packets + K_TX for a prepared batch: two kernel launches, nothing else."""
        torch = self.torch
        if plan.n_frames == 0:
            return torch.zeros(0, dtype=torch.complex64, device=self.dev)
        if out is None:
            out = torch.empty(plan.n_samples, dtype=torch.complex64, device=self.dev)
        elif out.numel() < plan.n_samples:
            raise ValueError("tx_run: output buffer too small")
        st = self._stream()
        _lib.check(self.L_.ofdm_make_packets(self.h, self._p(payload), self._p(plan.d_payload_off), plan.n_frames,
                                             int(whitening), self._p(plan.pkts), self._p(plan.d_pkt_off), st),
                   "make_packets")
        if plan.n_streams:
            _lib.check(self.L_.ofdm_tx_modulate_streams(self.h, self._p(plan.pkts), self._p(plan.d_pkt_off), plan.n_frames,
                                                        int(first_frame), self._p(plan.d_sym_off), plan.total_syms,
                                                        plan.uniform_syms, self._p(plan.d_stream_frame0),
                                                        self._p(plan.d_stream_out_off), plan.n_streams, self._p(out), st),
                       "tx_modulate_streams")
        else:
            _lib.check(self.L_.ofdm_tx_modulate_batch(self.h, self._p(plan.pkts), self._p(plan.d_pkt_off), plan.n_frames,
                                                      int(first_frame), self._p(plan.d_sym_off), plan.total_syms,
                                                      plan.uniform_syms, self._p(out), st), "tx_modulate_batch")
        return out[:plan.n_samples]

    def make_packets(self, payload, payload_off: np.ndarray, pad_for_usrp: bool = False, whitening: bool = True):
        """payload: uint8 cuda tensor of the concatenated payloads; payload_off: host int64 [F+1].
        Returns (pkts uint8 cuda tensor, pkt_off host int64 [F+1])."""
        torch = self.torch
        payload_off = np.ascontiguousarray(payload_off, dtype=np.int64)
        plen = np.diff(payload_off)
        klen = self._packet_lengths(plen, pad_for_usrp)
        pkt_off = np.zeros(len(plen) + 1, dtype=np.int64)
        np.cumsum(klen, out=pkt_off[1:])
        pkts = torch.empty(int(pkt_off[-1]), dtype=torch.uint8, device=self.dev)
        d_poff = torch.from_numpy(payload_off).to(self.dev, non_blocking=True)
        d_koff = torch.from_numpy(pkt_off).to(self.dev, non_blocking=True)
        _lib.check(self.L_.ofdm_make_packets(self.h, self._p(payload), self._p(d_poff), len(plen), int(whitening),
                                             self._p(pkts), self._p(d_koff), self._stream()), "make_packets")
        return pkts, pkt_off, d_koff

    def modulate(self, pkts, pkt_off: np.ndarray, d_pkt_off=None, first_frame: int = 0, out=None, taps: Optional[dict] = None):
        """pkts: uint8 cuda tensor of concatenated packets; pkt_off: host int64 [F+1].
        Returns complex64 cuda tensor with all frames back to back.  ``taps``: a dict that receives the reference's
        options.log stage outputs (ofdm.py:123-129) as complex64 cuda tensors: "mapper" [data symbols, N],
        "preambles" [symbols, N], "ifft" [symbols, N]."""
        torch = self.torch
        pkt_off = np.ascontiguousarray(pkt_off, dtype=np.int64)
        F = len(pkt_off) - 1
        if F <= 0:
            return torch.zeros(0, dtype=torch.complex64, device=self.dev)
        plen = np.diff(pkt_off)
        per = self.ncar * self.nbits
        nsym = 1 + np.maximum(1, -(-(8 * plen) // per))
        uniform = int(nsym[0]) if bool((nsym == nsym[0]).all()) else 0
        total = int(nsym.sum())
        if d_pkt_off is None:
            d_pkt_off = torch.from_numpy(pkt_off).to(self.dev, non_blocking=True)
        d_sym_off = None
        if not uniform:
            sym_off = np.zeros(F + 1, dtype=np.int64)
            np.cumsum(nsym, out=sym_off[1:])
            d_sym_off = torch.from_numpy(sym_off).to(self.dev, non_blocking=True)
        if out is None:
            out = torch.empty(total * self.L, dtype=torch.complex64, device=self.dev)
        elif out.numel() < total * self.L:
            raise ValueError("modulate: output buffer too small")
        if taps is not None:
            taps["mapper"] = torch.zeros((total - F, self.N), dtype=torch.complex64, device=self.dev)
            taps["preambles"] = torch.zeros((total, self.N), dtype=torch.complex64, device=self.dev)
            taps["ifft"] = torch.zeros((total, self.N), dtype=torch.complex64, device=self.dev)
            _lib.check(self.L_.ofdm_tx_modulate_taps(self.h, self._p(pkts), self._p(d_pkt_off), F, int(first_frame),
                                                     self._p(d_sym_off), total, uniform, self._p(out), self._p(taps["mapper"]),
                                                     self._p(taps["preambles"]), self._p(taps["ifft"]), self._stream()),
                       "tx_modulate_taps")
            return out[:total * self.L]
        _lib.check(self.L_.ofdm_tx_modulate_batch(self.h, self._p(pkts), self._p(d_pkt_off), F, int(first_frame),
                                                  self._p(d_sym_off), total, uniform, self._p(out), self._stream()),
                   "tx_modulate_batch")
        return out[:total * self.L]

    def channel(self, x, cfo: float = 0.0, sigma: float = 0.0, seed: int = 0, phase0: float = 0.0, out=None):
        torch = self.torch
        if out is None:
            out = torch.empty_like(x)
        _lib.check(self.L_.ofdm_channel(self.h, self._p(x), x.numel(), float(cfo), float(phase0), float(sigma),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, self._p(out), self._stream()), "channel")
        return out

    # ------------------------------------------------------------------ receive
    def rx_alloc(self, n: int, max_frames: Optional[int] = None, taps: bool = False, max_vectors: int = 0,
                 fresh: bool = False):
        """Allocate (and cache) the output arrays + workspace of one receive call.  ``fresh=True`` always allocates an
        independent buffer set (for callers that keep several receive calls in flight, one stream each)."""
        torch = self.torch
        if max_frames is None:
            max_frames = max(64, int(n // self.L) + 64)
        key = (int(n), int(max_frames), taps, int(max_vectors))
        if self._ws_key == key and not fresh:
            return self._ws
        # a cached set that is large enough serves a shorter stream as well (the workspace is carved by the n of the
        # call): consecutive feed_stream passes of slightly different lengths must not re-allocate gigabytes
        if (not fresh and self._ws_key is not None and not taps and self._ws_key[2] == taps and self._ws_key[0] >= n
                and self._ws_key[1] == int(max_frames) and self._ws_key[0] <= 2 * int(n) + (1 << 20)):
            return self._ws
        dev = self.dev
        need = int(self.L_.ofdm_rx_workspace_bytes(self.h, int(n), int(max_frames)))
        b = {}
        b["workspace"] = torch.empty(need, dtype=torch.uint8, device=dev)
        b["status"] = torch.zeros(1, dtype=torch.int32, device=dev)
        b["n_trig"] = torch.zeros(1, dtype=torch.int32, device=dev)
        b["trig_idx"] = torch.zeros(max_frames, dtype=torch.int64, device=dev)
        b["trig_ang"] = torch.zeros(max_frames, dtype=torch.float32, device=dev)
        b["n_frames"] = torch.zeros(1, dtype=torch.int32, device=dev)
        b["frame_start"] = torch.zeros(max_frames, dtype=torch.int64, device=dev)
        b["frame_ndata"] = torch.zeros(max_frames, dtype=torch.int32, device=dev)
        b["frame_live"] = torch.zeros(max_frames, dtype=torch.uint8, device=dev)
        b["frame_status"] = torch.zeros(max_frames, dtype=torch.uint8, device=dev)
        b["pkt_len"] = torch.zeros(max_frames, dtype=torch.int32, device=dev)
        b["pkt_ok"] = torch.zeros(max_frames, dtype=torch.uint8, device=dev)
        b["pkt_bytes"] = torch.zeros(max_frames * self.pkt_stride, dtype=torch.uint8, device=dev)
        b["counters"] = torch.zeros(8, dtype=torch.int64, device=dev)
        if taps:
            b["eq_syms"] = torch.zeros(max_vectors * self.occ, dtype=torch.complex64, device=dev)
            b["sym_idx"] = torch.zeros(max_vectors * self.ncar, dtype=torch.uint8, device=dev)
            b["derot_syms"] = torch.zeros(max_vectors * self.ncar, dtype=torch.complex64, device=dev)
            if taps == "all":                      # options.log: the vector taps in front of the frame acquisition too
                b["fft_out"] = torch.zeros(max_vectors * self.N, dtype=torch.complex64, device=dev)
                b["sampler_out"] = torch.zeros(max_vectors * self.N, dtype=torch.complex64, device=dev)
        io = _lib.RxIo()
        io.max_frames = int(max_frames)
        io.pkt_stride = self.pkt_stride
        io.workspace = b["workspace"].data_ptr()
        io.workspace_bytes = need
        for k in ("status", "n_trig", "trig_idx", "trig_ang", "n_frames", "frame_start", "frame_ndata", "frame_live",
                  "frame_status", "pkt_len", "pkt_ok", "pkt_bytes", "counters"):
            setattr(io, k, b[k].data_ptr())
        io.eq_syms = b["eq_syms"].data_ptr() if taps else None
        io.sym_idx = b["sym_idx"].data_ptr() if taps else None
        io.derot_syms = b["derot_syms"].data_ptr() if taps else None
        io.max_vectors = int(max_vectors) if taps else 0
        io.fft_out = b["fft_out"].data_ptr() if "fft_out" in b else None
        io.sampler_out = b["sampler_out"].data_ptr() if "sampler_out" in b else None
        b["io"] = io
        b["n"] = int(n)
        if not fresh:
            self._ws, self._ws_key = b, key
        return b

    def demodulate_async(self, x, bufs=None, sync: str = "pn", nsymbols: int = 18, freq_offset: float = 0.0,
                         snr_db: float = 30.0, **kw):
        """Run the whole receive chain on the current stream; returns the buffer dict (no sync).
        ``sync`` selects the synchroniser like the SYNC constant of ofdm_receiver.py~:89-119: "pn" (the live one),
        "pnac", "ml" (``snr_db`` = options.snr feeds its rho), or "fixed" -- the reference's test mode: no channel filter,
        a trigger every ``nsymbols`` symbols, constant frequency offset ``freq_offset`` (subcarrier spacings)."""
        n = int(x.numel())
        if bufs is None:
            bufs = self.rx_alloc(n, **kw)
        if sync == "pn":
            _lib.check(self.L_.ofdm_rx_demodulate(self.h, self._p(x), n, C.byref(bufs["io"]), self._stream()),
                       "rx_demodulate")
        elif sync == "fixed":
            _lib.check(self.L_.ofdm_rx_demodulate_fixed(self.h, self._p(x), n, int(nsymbols), float(freq_offset),
                                                        C.byref(bufs["io"]), self._stream()), "rx_demodulate_fixed")
        elif sync in ("pnac", "ml"):
            need = int(self.L_.ofdm_rx_sync_alt_scratch_bytes(self.h, n))
            sc = bufs.get("_alt_scratch")
            if sc is None or sc.numel() < need:
                sc = bufs["_alt_scratch"] = self.torch.empty(need, dtype=self.torch.uint8, device=self.dev)
            _lib.check(self.L_.ofdm_rx_demodulate_alt(self.h, self._p(x), n, sync.encode("ascii"), float(snr_db),
                                                      C.byref(bufs["io"]), self._p(sc), sc.numel(), self._stream()),
                       "rx_demodulate_alt")
        else:
            raise ValueError("sync %r: the reference names 'pn', 'ml', 'pnac' and 'fixed'" % (sync,))
        return bufs

    def nco_events(self, bufs, n: int):
        """(indices, angles) of the NCO's own event list after a sync="ml" call (every detector peak)."""
        torch = self.torch
        base = bufs["workspace"].data_ptr()
        ptr = lambda which: int(self.L_.ofdm_rx_workspace_ptr(self.h, C.byref(bufs["io"]), int(n), which)) - base
        torch.cuda.current_stream(self.dev).synchronize()
        k = int(bufs["workspace"][ptr(7):ptr(7) + 4].view(torch.int32).item())
        if k < 0:
            return None
        idx = bufs["workspace"][ptr(8):ptr(8) + 8 * k].view(torch.int64).cpu().numpy()
        ang = bufs["workspace"][ptr(9):ptr(9) + 4 * k].view(torch.float32).cpu().numpy()
        return idx, ang

    # ---- many independent streams in one call (ofdm_rx_demodulate_batch) ----
    def rx_alloc_batch(self, stream_off, max_frames: int):
        """Output arrays + workspace for S streams lying back to back in one sample buffer (stream s = samples
        [stream_off[s], stream_off[s+1])).  Every per-frame array is [S, max_frames]; scalars are [S]."""
        torch = self.torch
        dev = self.dev
        so = np.ascontiguousarray(stream_off, dtype=np.int64)
        S = len(so) - 1
        if S < 1 or (np.diff(so) < 0).any():
            raise ValueError("rx_alloc_batch: stream_off must ascend and hold at least two entries")
        n_total, n_max = int(so[-1]), int(np.diff(so).max())
        mf = int(max_frames)
        need = int(self.L_.ofdm_rx_workspace_bytes_batch(self.h, S, n_total, n_max, mf))
        b = {"workspace": torch.empty(need, dtype=torch.uint8, device=dev)}
        for k, dt, per in (("status", torch.int32, 1), ("n_trig", torch.int32, 1), ("n_frames", torch.int32, 1),
                           ("trig_idx", torch.int64, mf), ("trig_ang", torch.float32, mf), ("frame_start", torch.int64, mf),
                           ("frame_ndata", torch.int32, mf), ("frame_live", torch.uint8, mf), ("frame_status", torch.uint8, mf),
                           ("pkt_len", torch.int32, mf), ("pkt_ok", torch.uint8, mf), ("pkt_bytes", torch.uint8, mf * self.pkt_stride),
                           ("counters", torch.int64, 8)):
            b[k] = torch.zeros(S * per, dtype=dt, device=dev)
        io = _lib.RxIo()
        io.max_frames = mf
        io.pkt_stride = self.pkt_stride
        io.workspace = b["workspace"].data_ptr()
        io.workspace_bytes = need
        for k in ("status", "n_trig", "trig_idx", "trig_ang", "n_frames", "frame_start", "frame_ndata", "frame_live",
                  "frame_status", "pkt_len", "pkt_ok", "pkt_bytes", "counters"):
            setattr(io, k, b[k].data_ptr())
        io.eq_syms = io.sym_idx = io.derot_syms = None
        io.max_vectors = 0
        b.update(io=io, n=n_total, S=S, n_max=n_max, max_frames=mf, stream_off=so,
                 d_stream_off=torch.from_numpy(so).to(dev))
        return b

    def demodulate_batch_async(self, x, bufs):
        """The whole receive chain on S streams at once, on the current stream (no sync)."""
        if int(x.numel()) < bufs["n"]:
            raise ValueError("demodulate_batch: sample buffer shorter than stream_off[-1]")
        _lib.check(self.L_.ofdm_rx_demodulate_batch(self.h, self._p(x), self._p(bufs["d_stream_off"]), bufs["S"], bufs["n"],
                                                    bufs["n_max"], C.byref(bufs["io"]), self._stream()), "rx_demodulate_batch")
        return bufs

    def collect_batch(self, bufs, want_packets: bool = True) -> List[RxBatch]:
        """Synchronise and return one RxBatch per stream (what S separate demodulate() calls would return)."""
        torch = self.torch
        S, mf = bufs["S"], bufs["max_frames"]
        torch.cuda.current_stream(self.dev).synchronize()
        host = {k: bufs[k].cpu().numpy() for k in ("status", "n_trig", "n_frames", "counters", "frame_live", "frame_status",
                                                   "pkt_len", "pkt_ok")}
        if int(host["status"].max(initial=0)):
            raise RuntimeError("receive: capacity overflow in stream %d (status bits 0x%x): raise max_frames"
                               % (int(np.argmax(host["status"] != 0)), int(host["status"].max())))
        if want_packets:
            for k in ("trig_idx", "trig_ang", "frame_start", "frame_ndata"):
                host[k] = bufs[k].cpu().numpy()
        out = []
        z = np.zeros(0)
        for s in range(S):
            nt, nf = int(host["n_trig"][s]), int(host["n_frames"][s])
            sl = slice(s * mf, s * mf + nf)
            live, fstat, plen, pok = (host[k][sl] for k in ("frame_live", "frame_status", "pkt_len", "pkt_ok"))
            sel = np.flatnonzero((live == 1) & (fstat == 2))
            packets, rows = [], None
            if want_packets:
                rows = bufs["pkt_bytes"][s * mf * self.pkt_stride:(s * mf + nf) * self.pkt_stride].cpu().numpy()
                rows = rows.reshape(nf, self.pkt_stride)
                for f in sel:
                    ln = int(plen[f])
                    body = rows[f, :min(ln, self.pkt_stride)].tobytes()
                    packets.append((bool(pok[f]), body[:-4] if ln >= 4 else b""))
                tl = slice(s * mf, s * mf + nt)
                extra = (host["trig_idx"][tl].copy(), host["trig_ang"][tl].copy(), host["frame_start"][sl].copy(),
                         host["frame_ndata"][sl].copy())
            else:
                extra = (z, z, z, z)
            out.append(RxBatch(nt, nf, 0, *extra, live.copy(), fstat.copy(), plen.copy(), pok.copy(),
                               host["counters"][8 * s:8 * s + 8].copy(), packets, sel, rows))
        return out

    def nco_taps(self, bufs, n: int):
        """(nco, sigmix) per-sample streams of the last receive call on ``bufs`` (ofdm_receiver.py~:150-151)."""
        torch = self.torch
        nco = torch.empty(n, dtype=torch.complex64, device=self.dev)
        mix = torch.empty(n, dtype=torch.complex64, device=self.dev)
        _lib.check(self.L_.ofdm_rx_nco_taps(self.h, self._p(self.ws_view(bufs, 0, n)), int(n), C.byref(bufs["io"]), self._p(nco),
                                            self._p(mix), self._stream()), "rx_nco_taps")
        return nco, mix

    def ws_view(self, bufs, which: int, n: int):
        """Tensor views of the workspace taps (0: filtered stream y, 1: timing metric mf)."""
        torch = self.torch
        ptr = self.L_.ofdm_rx_workspace_ptr(self.h, C.byref(bufs["io"]), int(n), int(which))
        base = bufs["workspace"].data_ptr()
        off = int(ptr) - base
        if which == 0:
            return bufs["workspace"][off:off + 8 * n].view(torch.complex64)
        if which == 1:
            return bufs["workspace"][off:off + 4 * n].view(torch.float32)
        raise ValueError(which)

    def _pinned(self, bufs, key, like, count):
        """Pinned host mirror of a result array (allocated once per buffer set): device->host copies into
        pinned memory run at PCIe speed and can be queued asynchronously."""
        pool = bufs.setdefault("_pinned", {})
        t = pool.get(key)
        if t is None or t.numel() < count or t.dtype != like.dtype:
            t = self.torch.empty(max(count, 1), dtype=like.dtype, pin_memory=True)
            pool[key] = t
        return t

    def collect(self, bufs, want_packets: bool = True, want_payload: bool = True) -> RxBatch:
        """Synchronise and bring one receive call's results to the host.  The returned arrays are views of
        pinned staging buffers that the next collect() on the same buffer set overwrites."""
        torch = self.torch
        stream = torch.cuda.current_stream(self.dev)
        hd = self._pinned(bufs, "head", bufs["n_trig"], 3)
        hd[0:1].copy_(bufs["n_trig"], non_blocking=True)
        hd[1:2].copy_(bufs["n_frames"], non_blocking=True)
        hd[2:3].copy_(bufs["status"], non_blocking=True)
        hc = self._pinned(bufs, "counters", bufs["counters"], 8)
        hc.copy_(bufs["counters"], non_blocking=True)
        stream.synchronize()
        nt, nf, st = int(hd[0]), int(hd[1]), int(hd[2])
        if st:
            raise RuntimeError("receive: capacity overflow (status bits 0x%x): raise max_frames" % st)
        host = {}

        def fetch(key, m):
            t = self._pinned(bufs, key, bufs[key], m)
            if m:
                t[:m].copy_(bufs[key][:m], non_blocking=True)
            host[key] = t[:m].numpy()

        for k in ("frame_live", "frame_status", "pkt_len", "pkt_ok"):
            fetch(k, nf)
        if want_packets:
            fetch("trig_idx", nt)
            fetch("trig_ang", nt)
            fetch("frame_start", nf)
            fetch("frame_ndata", nf)
        rows_all = None
        if (want_payload or want_packets) and nf:
            # every slot up to n_frames in one pinned copy; the (few) undelivered rows are dropped on the host
            pb = self._pinned(bufs, "pkt_bytes", bufs["pkt_bytes"], nf * self.pkt_stride)
            pb[:nf * self.pkt_stride].copy_(bufs["pkt_bytes"][:nf * self.pkt_stride], non_blocking=True)
            rows_all = pb[:nf * self.pkt_stride].numpy().reshape(nf, self.pkt_stride)
        stream.synchronize()
        live, fstat = host["frame_live"], host["frame_status"]
        plen, pok = host["pkt_len"], host["pkt_ok"]
        meta = 4 * nf + 3 * nf + 64 + 12
        sel = np.flatnonzero((live == 1) & (fstat == 2))
        packets: List[Tuple[bool, bytes]] = []
        rows = None
        copied = 0
        if rows_all is not None:
            copied = int(rows_all.size)
            rows = rows_all                      # indexed by frame slot; msg_frames lists the delivered ones
            if want_packets:
                for f in sel:
                    ln = int(plen[f])
                    body = rows[f, :min(ln, self.pkt_stride)].tobytes()
                    packets.append((bool(pok[f]), body[:-4] if ln >= 4 else b""))
        if want_packets:
            trig_idx, trig_ang = host["trig_idx"].copy(), host["trig_ang"].copy()
            fstart, fnd = host["frame_start"].copy(), host["frame_ndata"].copy()
        else:
            trig_idx = trig_ang = fstart = fnd = np.zeros(0)
        return RxBatch(nt, nf, st, trig_idx, trig_ang, fstart, fnd, live.copy(), fstat.copy(), plen.copy(), pok.copy(),
                       hc.numpy().copy(), packets, sel, rows, copied, meta)

    def demodulate(self, x, **kw) -> RxBatch:
        return self.collect(self.demodulate_async(x, **kw))

    def demodulate_fixed(self, x, nsymbols: int = 18, freq_offset: float = 0.0, **kw) -> RxBatch:
        return self.collect(self.demodulate_async(x, sync="fixed", nsymbols=nsymbols, freq_offset=freq_offset, **kw))

    # ---- split collect: lets a caller keep several receive calls in flight (one stream + buffer set each) ----
    def collect_begin(self, bufs, want_payload: bool = True):
        """Queue, on the current stream, the device->host copies of one receive call's results into the buffer
        set's pinned staging arrays and return a ticket for :meth:`collect_end`.  n_frames is not known on the
        host yet, so every per-frame array travels at its allocated (max_frames) size."""
        torch = self.torch
        stream = torch.cuda.current_stream(self.dev)
        mf = int(bufs["io"].max_frames)
        hd = self._pinned(bufs, "head", bufs["n_trig"], 3)
        hd[0:1].copy_(bufs["n_trig"], non_blocking=True)
        hd[1:2].copy_(bufs["n_frames"], non_blocking=True)
        hd[2:3].copy_(bufs["status"], non_blocking=True)
        hc = self._pinned(bufs, "counters", bufs["counters"], 8)
        hc.copy_(bufs["counters"], non_blocking=True)
        host = {}
        nbytes = 12 + 64
        for k in ("frame_live", "frame_status", "pkt_len", "pkt_ok"):
            t = self._pinned(bufs, k, bufs[k], mf)
            t[:mf].copy_(bufs[k][:mf], non_blocking=True)
            host[k] = t
            nbytes += mf * t.element_size()
        pb = None
        if want_payload:
            pb = self._pinned(bufs, "pkt_bytes", bufs["pkt_bytes"], mf * self.pkt_stride)
            pb[:mf * self.pkt_stride].copy_(bufs["pkt_bytes"][:mf * self.pkt_stride], non_blocking=True)
            nbytes += mf * self.pkt_stride
        ev = torch.cuda.Event()
        ev.record(stream)
        return {"event": ev, "head": hd, "counters": hc, "host": host, "rows": pb, "max_frames": mf, "bytes": nbytes}

    def collect_end(self, ticket) -> RxBatch:
        """Wait for the copies of :meth:`collect_begin` and assemble the batch (arrays are views of the pinned
        staging buffers: the next collect on the same buffer set overwrites them)."""
        ticket["event"].synchronize()
        hd = ticket["head"]
        nt, nf, st = int(hd[0]), int(hd[1]), int(hd[2])
        if st:
            raise RuntimeError("receive: capacity overflow (status bits 0x%x): raise max_frames" % st)
        h = ticket["host"]
        live, fstat = h["frame_live"][:nf].numpy(), h["frame_status"][:nf].numpy()
        plen, pok = h["pkt_len"][:nf].numpy(), h["pkt_ok"][:nf].numpy()
        sel = np.flatnonzero((live == 1) & (fstat == 2))
        rows = None
        copied = 0
        if ticket["rows"] is not None:
            rows = ticket["rows"][:nf * self.pkt_stride].numpy().reshape(nf, self.pkt_stride)
            copied = ticket["max_frames"] * self.pkt_stride
        z = np.zeros(0)
        return RxBatch(nt, nf, st, z, z, z, z, live, fstat, plen, pok, ticket["counters"].numpy().copy(), [], sel, rows,
                       copied, ticket["bytes"] - copied)


    # ---- dense hand-over of the delivered messages (ofdm_rx_compact) ----
    def deliver_begin(self, bufs, expect_msgs: Optional[int] = None, expect_bytes: Optional[int] = None,
                      frame_starts: bool = False):
        """Queue, on the current stream, the device-side packing of every delivered message of one receive call
        (single stream or batch) into a dense byte array and the device->host copies of the result; returns a ticket
        for :meth:`deliver_end`.  The copies are sized by what the caller expects to come back (``expect_msgs``
        messages, ``expect_bytes`` payload+crc bytes -- e.g. what it transmitted), not by max_frames * pkt_stride;
        messages beyond that are reported by ``overflow`` in the result (``deliver_end(complete=True)`` fetches them
        with a second, exactly sized copy).  ``frame_starts``: also bring back the first sample of every message's
        frame (single stream only), for callers that place messages on a continuous stream."""
        torch = self.torch
        S = int(bufs.get("S", 1))
        mf = int(bufs["io"].max_frames)
        slots = S * mf
        d = bufs.get("_dense")
        if d is None:
            dev = self.dev
            d = {"bytes": torch.empty(slots * self.pkt_stride, dtype=torch.uint8, device=dev),
                 "off": torch.empty(slots + 1, dtype=torch.int64, device=dev),
                 "frame": torch.empty(slots, dtype=torch.int32, device=dev),
                 "ok": torch.empty((slots + 31) // 32, dtype=torch.int32, device=dev),
                 "totals": torch.empty(3, dtype=torch.int64, device=dev),
                 "scratch": torch.empty(2 * ((slots + 1023) // 1024), dtype=torch.int64, device=dev)}
            bufs["_dense"] = d
        _lib.check(self.L_.ofdm_rx_compact(self.h, C.byref(bufs["io"]), S, self._p(d["bytes"]), d["bytes"].numel(),
                                           self._p(d["off"]), self._p(d["frame"]), self._p(d["ok"]), self._p(d["totals"]),
                                           self._p(d["scratch"]), self._stream()), "rx_compact")
        nm = slots if expect_msgs is None else min(int(expect_msgs), slots)
        nb = d["bytes"].numel() if expect_bytes is None else min(int(expect_bytes), d["bytes"].numel())
        host = {}
        nbytes = 0
        items = [("totals", d["totals"], 3), ("counters", bufs["counters"], 8 * S), ("status", bufs["status"], S),
                 ("ok", d["ok"], (nm + 31) // 32), ("off", d["off"], nm + 1), ("frame", d["frame"], nm),
                 ("bytes", d["bytes"], nb)]
        if frame_starts and S == 1 and nm:
            # frame_start[frame[k]] gathered on the device (entries past the message count are clamped, never read)
            idx = d["frame"][:nm].to(torch.int64).clamp_(0, mf - 1)
            items.append(("frame_start", bufs["frame_start"][idx], nm))
        for key, src, cnt in items:
            t = self._pinned(bufs, "dense_" + key, src, cnt)
            if cnt:
                t[:cnt].copy_(src[:cnt], non_blocking=True)
            host[key] = t[:cnt]
            nbytes += cnt * t.element_size()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        return {"event": ev, "host": host, "nm": nm, "nb": nb, "d2h_bytes": nbytes, "S": S, "bufs": bufs,
                "frame_starts": frame_starts}

    def deliver_end(self, ticket, complete: bool = False):
        """Wait for :meth:`deliver_begin`; returns a dict: n_msgs, ok (bool [n]), off (int64 [n+1]), frame (int32 [n]),
        data (uint8, dense payload||crc bytes), counters ([S, 8]), overflow (messages or bytes beyond the expected
        sizes), d2h_bytes.  Arrays are views of pinned staging buffers reused by the next call on the buffer set."""
        ticket["event"].synchronize()
        h = ticket["host"]
        if int(h["status"].numpy().max(initial=0)):
            raise RuntimeError("receive: capacity overflow (status bits 0x%x): raise max_frames" % int(h["status"].numpy().max()))
        n_all, b_all = int(h["totals"][0]), int(h["totals"][1])
        n = min(n_all, ticket["nm"])
        off = h["off"].numpy()[:n + 1]
        while n > 0 and off[n] > ticket["nb"]:           # messages whose bytes were not copied
            n -= 1
        if complete and n < n_all:
            # more came back than the caller expected: one more copy, sized exactly (the device arrays are intact)
            t2 = self.deliver_begin(ticket["bufs"], expect_msgs=n_all, expect_bytes=b_all, frame_starts=ticket["frame_starts"])
            r = self.deliver_end(t2)
            r["d2h_bytes"] += ticket["d2h_bytes"]
            r["refetched"] = True
            return r
        okw = h["ok"].numpy().view(np.uint32)
        ok = ((okw[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool) if n else np.zeros(0, bool)
        out = {"n_msgs": n, "n_msgs_device": n_all, "ok": ok, "off": off[:n + 1], "frame": h["frame"].numpy()[:n],
               "data": h["bytes"].numpy(), "counters": h["counters"].numpy().reshape(ticket["S"], 8).copy(),
               "overflow": n_all - n, "d2h_bytes": ticket["d2h_bytes"], "bytes_device": b_all}
        if "frame_start" in h:
            out["frame_start"] = h["frame_start"].numpy()[:n]
        elif ticket.get("frame_starts"):
            out["frame_start"] = np.zeros(0, np.int64)
        return out

    def deliver(self, bufs, **kw):
        """(ok, payload) of every delivered message, in stream / arrival order, through the dense hand-over."""
        r = self.deliver_end(self.deliver_begin(bufs, **kw))
        data, off = r["data"], r["off"]
        out = []
        for m in range(r["n_msgs"]):
            body = data[off[m]:off[m + 1]].tobytes()
            out.append((bool(r["ok"][m]), body[:-4] if len(body) >= 4 else b""))
        return out, r


class SenseEngine:
    """stream_to_vector -> fft_vcc(N, True, blackmanharris) -> complex_to_mag_squared -> bin_statistics_f."""

    def __init__(self, fft_size: int, device: Optional[int] = None):
        import torch
        self.torch = torch
        self.L_ = _lib.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        self.N = int(fft_size)
        s = self.L_.ofdm_sense_create(self.N, self.device)
        if not s:
            raise ValueError("ofdm_sense_create: " + self.L_.ofdm_last_error().decode())
        self.s = C.c_void_p(s)

    def close(self):
        if getattr(self, "s", None):
            self.L_.ofdm_sense_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def maxhold(self, x, tune_delay: int, dwell_delay: int, shift: bool = False, out=None):
        torch = self.torch
        nfr = int(x.numel()) // self.N
        nd = nfr // (tune_delay + dwell_delay)
        if out is None:
            out = torch.empty((nd, self.N), dtype=torch.float32, device=self.dev)
        _lib.check(self.L_.ofdm_sense(self.s, C.c_void_p(x.data_ptr()), nfr, int(shift), int(tune_delay),
                                      int(dwell_delay), C.c_void_p(out.data_ptr()), self._stream()), "sense")
        return out

    def spectra(self, x, shift: bool = True):
        torch = self.torch
        nfr = int(x.numel()) // self.N
        out = torch.empty((nfr, self.N), dtype=torch.complex64, device=self.dev)
        _lib.check(self.L_.ofdm_sense_fft(self.s, C.c_void_p(x.data_ptr()), nfr, int(shift),
                                          C.c_void_p(out.data_ptr()), self._stream()), "sense_fft")
        return out

    def decide_device(self, maxhold, threshold: float = 0.001):
        """:meth:`decide` without the host copies: (avg_inorder float64[N], free uint8[N], hex uint8[N/4]) cuda tensors."""
        torch = self.torch
        n_avg = int(maxhold.shape[0])
        avg = torch.empty(self.N, dtype=torch.float64, device=self.dev)
        free = torch.empty(self.N, dtype=torch.uint8, device=self.dev)
        hx = torch.empty(self.N // 4, dtype=torch.uint8, device=self.dev)
        _lib.check(self.L_.ofdm_sense_decide(self.s, C.c_void_p(maxhold.data_ptr()), n_avg, float(threshold),
                                             C.c_void_p(avg.data_ptr()), C.c_void_p(free.data_ptr()),
                                             C.c_void_p(hx.data_ptr()), self._stream()), "sense_decide")
        return avg, free, hx

    def decide(self, maxhold, threshold: float = 0.001):
        """Mean of the dwell vectors, threshold, frequency-order swap, hex map (secondary_tx.py:237-266)."""
        avg, free, hx = self.decide_device(maxhold, threshold)
        return avg.cpu().numpy(), free.cpu().numpy(), bytes(hx.cpu().numpy()).decode("ascii")

    def hop(self, avg_inorder, free_bits, required_index: int):
        """Busy-bin count around the operating frequency and the quietest 17-bin band (secondary_tx.py:268-295),
        on the device tensors of :meth:`decide_device`.  Returns (busy, index or -1, window length)."""
        torch = self.torch
        out = torch.empty(3, dtype=torch.int32, device=self.dev)
        _lib.check(self.L_.ofdm_sense_hop(self.s, C.c_void_p(avg_inorder.data_ptr()), C.c_void_p(free_bits.data_ptr()),
                                          int(required_index), C.c_void_p(out.data_ptr()), self._stream()), "sense_hop")
        busy, index, wlen = [int(v) for v in out.cpu().tolist()]
        return busy, index, wlen
