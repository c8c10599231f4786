"""Regenerates tests/golden/reference_sense_logs.npz from the console logs of the reference's own over-the-air
sensing runs (run in the build container only: /root/reference does not exist on the GPU box).

output.txt, output_with_detection.txt and crap.txt are what sense_loop (secondary_tx.py:228-300 and its
sensing_and_tramsmitting variants) printed on the authors' USRP2: for every sweep, 256 lines
``<bin frequency> <10-dwell average power> <free flag>`` in frequency order (the loop at :252-262 prints
``p, moving_avg_data[...], thrshold[...]``), then the carrier map ``hex_conv(thrshold_inorder)``.  They are the only
recorded outputs of the reference's hot path in its tree, so they pin the decision half of the sensing path
(threshold sense, frequency ordering, hex packing, the bin-frequency arithmetic)."""
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_sense_logs.npz")
LINE = re.compile(r"^(\d+\.\d+) (\S+) ([01])$")
MAP = re.compile(r"^(?:Carrier map =\s+|carrier_map_new=\s*)?([0-9A-F]{64})$")


def sweeps(path):
    out = []
    rows = []
    for raw in open(os.path.join(REF, path), errors="replace"):
        s = raw.strip()
        m = LINE.match(s)
        if m:
            rows.append((m.group(1), float(m.group(2)), int(m.group(3))))
            continue
        h = MAP.match(s)
        if h and len(rows) == 256:
            out.append((rows, h.group(1)))
        if s:
            rows = [] if not m else rows
    return out


def main():
    freq_txt, avg, flag, hexes, src = [], [], [], [], []
    for f in ("output.txt", "output_with_detection.txt", "crap.txt"):
        for rows, hx in sweeps(f):
            freq_txt.append([r[0] for r in rows])
            avg.append([r[1] for r in rows])
            flag.append([r[2] for r in rows])
            hexes.append(hx)
            src.append(f)
    np.savez_compressed(OUT, freq_txt=np.array(freq_txt), avg=np.array(avg, dtype=np.float64),
                        flag=np.array(flag, dtype=np.uint8), hex=np.array(hexes), source=np.array(src))
    print("%d sweeps -> %s" % (len(hexes), OUT))


if __name__ == "__main__":
    main()
