"""Exact check (rational arithmetic) that K1's table-free division RN(x c1 + RN(x c2)), x = 4 v, equals numpy's
float32(v) / float32(255) for every v in 0..255.  c1 = RN(1/1020) = 0x3a808081, c2 = RN(1/1020 - c1) = 0xae7efeff."""
import math
from fractions import Fraction

import numpy as np


def rn32(fr: Fraction) -> float:
    if fr == 0:
        return 0.0
    e = math.floor(math.log2(abs(fr)))
    while Fraction(2) ** e > abs(fr):
        e -= 1
    while Fraction(2) ** (e + 1) <= abs(fr):
        e += 1
    ulp = Fraction(2) ** (e - 23)
    q = fr / ulp
    n = math.floor(q)
    r = q - n
    if r > Fraction(1, 2) or (r == Fraction(1, 2) and n % 2 == 1):
        n += 1
    return float(n * ulp)


def main():
    c1 = float(np.array([0x3a808081], np.uint32).view(np.float32)[0])
    c2 = float(np.array([0xae7efeff], np.uint32).view(np.float32)[0])
    assert c1 == rn32(Fraction(1, 1020)) and c2 == rn32(Fraction(1, 1020) - Fraction(c1))
    bad = []
    for v in range(256):
        x = Fraction(4 * v)
        y = rn32(x * Fraction(c1) + Fraction(rn32(x * Fraction(c2))))
        if np.float32(y) != np.float32(v) / np.float32(255.0):
            bad.append(v)
    print("mismatches:", bad)
    return not bad


if __name__ == "__main__":
    raise SystemExit(0 if main() else 1)
