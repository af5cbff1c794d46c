#!/usr/bin/env python
"""Exhaustive check, in exact rational arithmetic, of the table-free pixel normalisation of the fused crop + conv1 kernel
(vittracker_b200/csrc/vt_stem.cu: normalize_px / normalize_px2).

The reference computes ((v / 255.0) - mean) / std on fp32 tensors, every step rounded to fp32 (lib/test/tracker/data_utils.py:8-14).
The kernel replaces both divisions by constants with   q = a r;  q' = fma(fma(-d, q, a), r, q),  r = fl(1 / d)   and this script
proves that q' equals the correctly rounded quotient for every pixel value 0..255 and every channel - the domain is 768 inputs,
so the proof is a loop.  fp32 rounding (nearest, ties to even) and FMA are emulated with fractions.Fraction."""
from __future__ import annotations

from fractions import Fraction as Fr

import numpy as np

F32 = np.float32
MEAN = [F32(0.485), F32(0.456), F32(0.406)]
STD = [F32(0.229), F32(0.224), F32(0.225)]


def fl(x) -> np.float32:
    """Round an exact rational to the nearest float32, ties to even."""
    x = Fr(x)
    if x == 0:
        return F32(0.0)
    f = F32(float(x))                          # Fraction -> double is correctly rounded; fix a possible double rounding below
    best = None
    for c in (np.nextafter(f, F32(-np.inf)), f, np.nextafter(f, F32(np.inf))):
        key = (abs(Fr(float(c)) - x), int(np.array([c], dtype=F32).view(np.uint32)[0]) & 1)
        if best is None or key < best[0]:
            best = (key, c)
    return best[1]


def ex(a) -> Fr:
    return Fr(float(a))


def fma(a, b, c):
    return fl(ex(a) * ex(b) + ex(c))


def bits(a) -> int:
    return int(np.array([a], dtype=F32).view(np.uint32)[0])


def kernel_formula(v: int, ch: int) -> np.float32:
    a = F32(v)
    r255 = fl(Fr(1, 255))
    q = fl(ex(a) * ex(r255))
    x = fma(fma(F32(-255.0), q, a), r255, q)
    y = fl(ex(x) - ex(MEAN[ch]))
    rs = fl(1 / ex(STD[ch]))
    z = fl(ex(y) * ex(rs))
    return fma(fma(-STD[ch], z, y), rs, z)


def reference_formula(v: int, ch: int) -> np.float32:
    x = fl(Fr(v, 255))
    y = fl(ex(x) - ex(MEAN[ch]))
    return fl(ex(y) / ex(STD[ch]))


def check() -> int:
    """Number of (value, channel) pairs where the kernel's formula differs from the reference's three rounded steps."""
    bad = 0
    for ch in range(3):
        v = np.arange(256, dtype=F32)
        numpy_ref = ((v / F32(255.0)) - MEAN[ch]) / STD[ch]          # the emulation itself against NumPy's fp32 arithmetic
        for i in range(256):
            want = reference_formula(i, ch)
            assert bits(want) == bits(numpy_ref[i])
            bad += bits(kernel_formula(i, ch)) != bits(want)
    return bad


if __name__ == "__main__":
    n = check()
    print(f"{n} mismatches over 768 (value, channel) pairs")
    raise SystemExit(1 if n else 0)
