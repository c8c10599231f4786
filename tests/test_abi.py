"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/ofdm_b200.h
declares (no compute call is made here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ofdm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from ofdm_uhd_b200 import _lib
    L = _lib.load_library(built_lib)
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libofdm_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names
    assert L.ofdm_version() >= 100


def test_host_only_entry_points(built_lib):
    from ofdm_uhd_b200 import _lib
    L = _lib.load_library(built_lib)
    # header(4) + payload + crc(4) + 0x55, optionally padded to 16 (ofdm_packet_utils.py:128-135,145-166)
    assert L.ofdm_packet_len(402, 0) == 411 and L.ofdm_packet_len(402, 1) == 416 and L.ofdm_packet_len(0, 0) == 9


def test_sass_has_no_library_fft(built_lib):
    import subprocess
    out = subprocess.run(["nm", "-D", "--undefined-only", built_lib], capture_output=True, text=True).stdout
    assert "cufft" not in out.lower()


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from types import SimpleNamespace
    from ofdm_uhd_b200 import ofdm
    opt = SimpleNamespace(modulation="bpsk", fft_length=512, occupied_tones=200, cp_length=128, verbose=False,
                          log=False, snr=30)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ofdm.ofdm_mod(opt)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ofdm.ofdm_demod(opt)


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of ofdm_cfg and ofdm_rx_io as a C compiler lays them out == the ctypes mirrors in _lib.py
    (a binding with a field missing hands the library a short struct)."""
    import ctypes as C
    import subprocess
    from ofdm_uhd_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ofdm_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ofdm_cfg), offsetof(ofdm_cfg, host_carrier_map), '
                   'sizeof(ofdm_rx_io), offsetof(ofdm_rx_io, counters), offsetof(ofdm_rx_io, max_vectors), '
                   'offsetof(ofdm_rx_io, sampler_out)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.OfdmCfg), _lib.OfdmCfg.host_carrier_map.offset, C.sizeof(_lib.RxIo), _lib.RxIo.counters.offset,
            _lib.RxIo.max_vectors.offset, _lib.RxIo.sampler_out.offset]
    assert got == want
