"""Regenerates tests/golden/reference_tables.npz from the reference tree (run in the build container only:
/root/reference does not exist on the GPU box).

What can be taken from the reference itself:
  * psk.py / qam.py are pure-math modules that run unchanged under Python 3 -> constellation tables;
  * ofdm_packet_utils.py and ofdm.py are Python 2 syntax, so only their DATA is taken, by parsing the literals:
    random_mask_tuple (ofdm_packet_utils.py:195-451) and known_symbols_4512_3 (ofdm.py:310-325);
  * make_header (ofdm_packet_utils.py:93-97) is executed from its own source text (it is py3-compatible).
"""
import ast
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tables.npz")


def literal(path, name, close):
    src = open(os.path.join(REF, path)).read()
    i = src.index(name + " = ")
    j = src.index(close, i)
    return ast.literal_eval(src[i + len(name) + 3:j + 1])


def run_module(path, patch=None):
    src = open(os.path.join(REF, path)).read()
    if patch:
        src = patch(src)
    ns = {}
    exec(compile(src, path, "exec"), ns)
    return ns


def main():
    out = {}
    out["mask"] = np.array(literal("ofdm_packet_utils.py", "random_mask_tuple", ")"), dtype=np.uint8)
    out["known"] = np.array(literal("ofdm.py", "known_symbols_4512_3", "]"), dtype=np.int8)
    psk = run_module("psk.py")
    qam = run_module("qam.py")
    for m in (2, 4, 8):
        out["psk_gray_%d" % m] = np.array(psk["gray_constellation"][m], dtype=np.complex128)
        out["psk_%d" % m] = np.array(psk["constellation"][m], dtype=np.complex128)
    for m in (4, 8, 16, 64, 256):
        out["qam_%d" % m] = np.array(qam["constellation"][m], dtype=np.complex128)
    src = open(os.path.join(REF, "ofdm_packet_utils.py")).read()
    fn = re.search(r"def make_header\(.*?\n    return struct.pack\('!HH', val, val\)\n", src, re.S).group(0)
    ns = {}
    exec("import struct\n" + fn, ns)
    cases = [(0, 0), (406, 0), (4095, 0), (4096, 3), (123, 15), (17, 255)]
    out["hdr_in"] = np.array(cases, dtype=np.int64)
    out["hdr_out"] = np.array([list(ns["make_header"](a, b)) for a, b in cases], dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
