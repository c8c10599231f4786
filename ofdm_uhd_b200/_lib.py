"""ctypes binding of libofdm_b200.so (include/ofdm_b200.h).

PyTorch tensors are only the buffer carrier: every call receives ``tensor.data_ptr()`` and the
current CUDA stream.  There is no CPU fallback: if the library or a CUDA device is missing,
:func:`lib` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libofdm_b200.so")

EXPORTS = [
    "ofdm_last_error", "ofdm_version", "ofdm_selftest_packed_math", "ofdm_create", "ofdm_destroy", "ofdm_set_tx_amplitude", "ofdm_get_layout",
    "ofdm_get_chan_taps", "ofdm_packet_len", "ofdm_make_packets", "ofdm_frame_symbols", "ofdm_tx_modulate_batch", "ofdm_tx_modulate_taps", "ofdm_tx_modulate_streams",
    "ofdm_rx_workspace_bytes", "ofdm_rx_chan_filter", "ofdm_rx_sync_metric", "ofdm_rx_peak_detect", "ofdm_rx_sync",
    "ofdm_rx_plan",
    "ofdm_rx_demod", "ofdm_rx_stage", "ofdm_rx_finish", "ofdm_rx_liveness", "ofdm_rx_sync_alt_scratch_bytes", "ofdm_rx_sync_alt", "ofdm_rx_demodulate_alt", "ofdm_rx_sync_fixed", "ofdm_rx_demodulate_fixed", "ofdm_rx_demodulate", "ofdm_rx_workspace_bytes_batch", "ofdm_rx_demodulate_batch", "ofdm_rx_nco_taps", "ofdm_rx_compact", "ofdm_rx_workspace_ptr", "ofdm_channel",
    "ofdm_sense_create", "ofdm_sense_destroy", "ofdm_sense", "ofdm_sense_fft", "ofdm_sense_decide", "ofdm_sense_hop",
]


class OfdmCfg(C.Structure):
    _fields_ = [("fft_length", C.c_int32), ("occupied_tones", C.c_int32), ("cp_length", C.c_int32),
                ("constellation_size", C.c_int32), ("host_constellation", C.POINTER(C.c_float)),
                ("tx_amplitude", C.c_float), ("device", C.c_int32), ("pad_seed", C.c_uint64),
                ("max_pkt_bytes", C.c_int32), ("host_carrier_map", C.c_char_p)]


class RxIo(C.Structure):
    _fields_ = [("max_frames", C.c_int32), ("pkt_stride", C.c_int32), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("status", C.c_void_p), ("n_trig", C.c_void_p),
                ("trig_idx", C.c_void_p), ("trig_ang", C.c_void_p), ("n_frames", C.c_void_p),
                ("frame_start", C.c_void_p), ("frame_ndata", C.c_void_p), ("frame_live", C.c_void_p),
                ("frame_status", C.c_void_p), ("pkt_len", C.c_void_p), ("pkt_ok", C.c_void_p),
                ("pkt_bytes", C.c_void_p), ("counters", C.c_void_p), ("eq_syms", C.c_void_p),
                ("sym_idx", C.c_void_p), ("derot_syms", C.c_void_p), ("max_vectors", C.c_int64),
                ("fft_out", C.c_void_p), ("sampler_out", C.c_void_p)]


_lib: Optional[C.CDLL] = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the shared library and declare the prototypes (no CUDA call is made)."""
    if not os.path.exists(path):
        raise RuntimeError("libofdm_b200.so is not built (%s); run `python -m ofdm_uhd_b200._build` or "
                           "__graft_entry__.build() -- there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i32, i64, u64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
    L.ofdm_last_error.restype = C.c_char_p
    L.ofdm_version.restype = C.c_int
    L.ofdm_selftest_packed_math.argtypes = [i32, i64, u64, C.POINTER(i64)]
    L.ofdm_create.restype = vp
    L.ofdm_create.argtypes = [C.POINTER(OfdmCfg)]
    L.ofdm_destroy.argtypes = [vp]
    L.ofdm_destroy.restype = None
    L.ofdm_set_tx_amplitude.argtypes = [vp, f32]
    L.ofdm_get_layout.argtypes = [vp, C.POINTER(i32)]
    L.ofdm_get_chan_taps.argtypes = [vp, C.POINTER(f32), i32]
    L.ofdm_packet_len.argtypes = [i32, C.c_int]
    L.ofdm_packet_len.restype = i32
    L.ofdm_make_packets.argtypes = [vp, vp, vp, i32, C.c_int, vp, vp, vp]
    L.ofdm_frame_symbols.argtypes = [vp, i32]
    L.ofdm_frame_symbols.restype = i32
    L.ofdm_tx_modulate_batch.argtypes = [vp, vp, vp, i32, i64, vp, i64, i32, vp, vp]
    L.ofdm_tx_modulate_taps.argtypes = [vp, vp, vp, i32, i64, vp, i64, i32, vp, vp, vp, vp, vp]
    L.ofdm_rx_nco_taps.argtypes = [vp, vp, i64, C.POINTER(RxIo), vp, vp, vp]
    L.ofdm_tx_modulate_streams.argtypes = [vp, vp, vp, i32, i64, vp, i64, i32, vp, vp, i32, vp, vp]
    L.ofdm_rx_workspace_bytes.argtypes = [vp, i64, i32]
    L.ofdm_rx_workspace_bytes.restype = C.c_size_t
    L.ofdm_rx_chan_filter.argtypes = [vp, vp, i64, vp, vp]
    L.ofdm_rx_sync_metric.argtypes = [vp, vp, i64, vp, vp, vp]
    L.ofdm_rx_peak_detect.argtypes = [vp, vp, vp, i64, vp, C.POINTER(RxIo), vp]
    L.ofdm_rx_sync.argtypes = [vp, vp, i64, C.POINTER(RxIo), vp]
    L.ofdm_rx_plan.argtypes = [vp, i64, C.POINTER(RxIo), vp]
    L.ofdm_rx_demod.argtypes = [vp, vp, i64, C.POINTER(RxIo), vp]
    L.ofdm_rx_stage.argtypes = [vp, vp, i64, C.POINTER(RxIo), i32, vp]
    L.ofdm_rx_finish.argtypes = [vp, C.POINTER(RxIo), vp]
    L.ofdm_rx_liveness.argtypes = [vp, vp, vp, i32, vp, vp, C.c_int, vp]
    L.ofdm_rx_sync_alt_scratch_bytes.argtypes = [vp, i64]
    L.ofdm_rx_sync_alt_scratch_bytes.restype = C.c_size_t
    L.ofdm_rx_sync_alt.argtypes = [vp, vp, i64, C.c_char_p, f32, C.POINTER(RxIo), vp, C.c_size_t, vp]
    L.ofdm_rx_demodulate_alt.argtypes = [vp, vp, i64, C.c_char_p, f32, C.POINTER(RxIo), vp, C.c_size_t, vp]
    L.ofdm_rx_sync_fixed.argtypes = [vp, i64, i32, f32, C.POINTER(RxIo), vp]
    L.ofdm_rx_demodulate_fixed.argtypes = [vp, vp, i64, i32, f32, C.POINTER(RxIo), vp]
    L.ofdm_rx_demodulate.argtypes = [vp, vp, i64, C.POINTER(RxIo), vp]
    L.ofdm_rx_workspace_bytes_batch.argtypes = [vp, i32, i64, i64, i32]
    L.ofdm_rx_workspace_bytes_batch.restype = C.c_size_t
    L.ofdm_rx_demodulate_batch.argtypes = [vp, vp, vp, i32, i64, i64, C.POINTER(RxIo), vp]
    L.ofdm_rx_compact.argtypes = [vp, C.POINTER(RxIo), i32, vp, i64, vp, vp, vp, vp, vp, vp]
    L.ofdm_rx_workspace_ptr.argtypes = [vp, C.POINTER(RxIo), i64, C.c_int]
    L.ofdm_rx_workspace_ptr.restype = vp
    L.ofdm_channel.argtypes = [vp, vp, i64, f32, f64, f32, u64, vp, vp]
    L.ofdm_sense_create.argtypes = [i32, i32]
    L.ofdm_sense_create.restype = vp
    L.ofdm_sense_destroy.argtypes = [vp]
    L.ofdm_sense_destroy.restype = None
    L.ofdm_sense.argtypes = [vp, vp, i64, C.c_int, i32, i32, vp, vp]
    L.ofdm_sense_fft.argtypes = [vp, vp, i64, C.c_int, vp, vp]
    L.ofdm_sense_decide.argtypes = [vp, vp, i32, f64, vp, vp, vp, vp]
    L.ofdm_sense_hop.argtypes = [vp, vp, vp, i32, vp, vp]
    return L


def lib() -> C.CDLL:
    """The library, for compute calls: requires a CUDA device (fails loudly otherwise)."""
    global _lib
    if _lib is None:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ofdm_uhd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib = load_library()
    return _lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise RuntimeError("libofdm_b200 %s failed (%d): %s" % (what, rc, lib().ofdm_last_error().decode()))
    return rc


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
