"""QAM constellation tables with the names and contents of the reference's qam.py
(/root/reference/qam.py:28-113): symbol i is already Gray labelled, so the code maps are identities."""
import math


def _gray_level(bits):
    """Magnitude level (odd integer) selected by the Gray-coded magnitude bits (reference qam.py:40-56)."""
    n = len(bits)
    total = 0.0
    for shift in range(n):
        parity = 0
        for b in bits[:n - shift]:
            parity = abs(b - parity)
        total += parity * 2.0 ** (shift + 1)
    return total + 1


def make_constellation(m):
    """Bit k-1 is the sign of I, bit k-2 the sign of Q, the remaining bits alternate between the I and Q
    magnitudes; the result is scaled so the largest coordinate is 1 (reference qam.py:28-64)."""
    k = int(round(math.log2(m)))
    raw = []
    biggest = 1
    for sym in range(m):
        msb_first = [(sym >> (k - 1 - pos)) & 1 for pos in range(k)]
        re = (2 * msb_first[0] - 1) * _gray_level(msb_first[2::2])
        im = (2 * msb_first[1] - 1) * _gray_level(msb_first[3::2])
        biggest = max(biggest, re, im)
        raw.append((re, im))
    return [complex(re / biggest, im / biggest) for re, im in raw]


constellation = {m: make_constellation(m) for m in (4, 8, 16, 64, 256)}

binary_to_gray = {m: list(range(m)) for m in (4, 8, 16, 64, 256)}
gray_to_binary = {m: list(range(m)) for m in (4, 8, 16, 64, 256)}
binary_to_ungray = {m: list(range(m)) for m in (4, 8, 16, 64)}
ungray_to_binary = {m: list(range(m)) for m in (4, 8, 16, 64)}
