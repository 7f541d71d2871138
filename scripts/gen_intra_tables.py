#!/usr/bin/env python
"""Generates the Intra8x8 directional-predictor table of kernels.cu (c_i8_pred): for mode m and sample (x, y) of the
8x8 block, pred = (F[a] + 2 F[b] + F[c] + 2) >> 2 with F the 25 filtered reference samples p'
(F[7 - i] = p'(-1, i), F[8] = p'(-1, -1), F[12 + i] = p'(i, -1): both runs word aligned); entry = a | b << 8 | c << 16.  Two-tap averages are
(a, b, a), copies (a, a, a), the "(s + 3 t + 2) >> 2" end cases (a, b, b).  Restates the nine branches of spec 8.3.2.2
(reference: decoder/intra_prediction.cc:449-621); mode 2 (DC) is computed in the kernel, not tabulated.
Usage: gen_intra_tables.py > table.inc"""
N = 8
T = lambda i: 8 if i == -1 else 12 + i          # i = -1 .. 15
L = lambda i: 7 - i          # i = -1 .. 7


def taps(mode, x, y):
    n = N
    two = lambda a, b: (a, b, a)
    one = lambda a: (a, a, a)
    if mode == 0: return one(T(x))
    if mode == 1: return one(L(y))
    if mode == 2: return (0, 0, 0)
    if mode == 3:
        if x == n - 1 and y == n - 1: return (T(x + y), T(x + y + 1), T(x + y + 1))
        return (T(x + y), T(x + y + 1), T(x + y + 2))
    if mode == 4:
        if x > y: return (T(x - y - 2), T(x - y - 1), T(x - y))
        if x < y: return (L(y - x - 2), L(y - x - 1), L(y - x))
        return (T(0), T(-1), L(0))
    if mode == 5:
        z = 2 * x - y
        if z >= 0 and z % 2 == 0: return two(T(x - (y >> 1) - 1), T(x - (y >> 1)))
        if z >= 0: return (T(x - (y >> 1) - 2), T(x - (y >> 1) - 1), T(x - (y >> 1)))
        if z == -1: return (L(0), T(-1), T(0))
        return (L(y - 2 * x - 1), L(y - 2 * x - 2), L(y - 2 * x - 3))
    if mode == 6:
        z = 2 * y - x
        if z >= 0 and z % 2 == 0: return two(L(y - (x >> 1) - 1), L(y - (x >> 1)))
        if z >= 0: return (L(y - (x >> 1) - 2), L(y - (x >> 1) - 1), L(y - (x >> 1)))
        if z == -1: return (L(0), T(-1), T(0))
        return (T(x - 2 * y - 1), T(x - 2 * y - 2), T(x - 2 * y - 3))
    if mode == 7:
        if y % 2 == 0: return two(T(x + (y >> 1)), T(x + (y >> 1) + 1))
        return (T(x + (y >> 1)), T(x + (y >> 1) + 1), T(x + (y >> 1) + 2))
    z, m = x + 2 * y, 2 * n - 3
    if z < m and z % 2 == 0: return two(L(y + (x >> 1)), L(y + (x >> 1) + 1))
    if z < m: return (L(y + (x >> 1)), L(y + (x >> 1) + 1), L(y + (x >> 1) + 2))
    if z == m: return (L(n - 2), L(n - 1), L(n - 1))
    return one(L(n - 1))


for mode in range(9):
    row = []
    for y in range(N):
        for x in range(N):
            a, b, c = taps(mode, x, y)
            assert 0 <= min(a, b, c) and max(a, b, c) <= 27, (mode, x, y, a, b, c)
            row.append("0x%06X" % (a | b << 8 | c << 16))
    for k in range(0, 64, 16):
        print("    " + ", ".join(row[k:k + 16]) + ",")
