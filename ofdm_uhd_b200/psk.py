"""PSK constellation tables with the names and contents of the reference's psk.py
(/root/reference/psk.py:26-94): Gray-coded M-PSK for M in {2,4,8} plus the counting-order tables."""
import cmath
import math


def make_gray_constellation(m):
    """Gray-coded M-PSK (reference psk.py:26-43).  With (b0,b1,b2) the bits of the symbol index,
    MSB first and right-aligned in three positions, the point sits at
    -(2*b0-1) * 2*pi/m * (b0 + |b1-b2| + 2*b1)."""
    k = int(round(math.log2(m)))
    points = []
    for sym in range(m):
        b = [0, 0, 0]
        for pos in range(k):
            b[3 - k + pos] = (sym >> (k - 1 - pos)) & 1
        theta = -(2 * b[0] - 1) * (2 * math.pi / m) * (b[0] + abs(b[1] - b[2]) + 2 * b[1])
        points.append(complex(math.cos(theta), math.sin(theta)))
    return points


def make_constellation(m):
    """Points in counting order around the unit circle (reference psk.py:46-47)."""
    return [cmath.exp(2j * math.pi * i / m) for i in range(m)]


constellation = {m: make_constellation(m) for m in (2, 4, 8)}
gray_constellation = {m: make_gray_constellation(m) for m in (2, 4, 8)}

binary_to_gray = {2: list(range(2)), 4: [0, 1, 3, 2], 8: [0, 1, 3, 2, 7, 6, 4, 5]}
gray_to_binary = {2: list(range(2)), 4: [0, 1, 3, 2], 8: [0, 1, 3, 2, 6, 7, 5, 4]}
binary_to_ungray = {m: list(range(m)) for m in (2, 4, 8)}
ungray_to_binary = {m: list(range(m)) for m in (2, 4, 8)}
