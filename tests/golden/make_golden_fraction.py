"""Generates tests/golden/fraction_golden.npz from the REFERENCE ITSELF (oracle/_ref: unmodified src/decompose.h):
allelicFraction(c, tr, bc) on synthetic two-allele traces.

    python tests/golden/make_golden_fraction.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(rng, nbc, frac, ndiff, style):
    """Two alleles mixed frac : 1-frac in signal space at `ndiff` differing positions (plus noise channels)."""
    ns = 12 * nbc + 40
    tr = rng.integers(0, 30, size=(4, ns)).astype(np.int32)
    pos = (12 * np.arange(nbc) + 10 + rng.integers(-2, 3, nbc)).astype(np.int32)
    pri = bytearray(rng.choice(list(b"ACGT"), nbc).astype(np.uint8).tobytes())
    sec = bytearray(pri)
    diff = set(int(x) for x in rng.choice(np.arange(nbc), size=min(ndiff, nbc), replace=False))
    for j in range(nbc):
        a = b"ACGT".index(pri[j])
        h = int(rng.integers(800, 1400))
        if j in diff:
            b = (a + int(rng.integers(1, 4))) % 4
            sec[j] = b"ACGT"[b]
            tr[a, pos[j]] += int(h * frac)
            tr[b, pos[j]] += int(h * (1 - frac))
            if style == 1:                                # a third allele leaks in
                tr[(a + 2) % 4 if (a + 2) % 4 != b else (a + 1) % 4, pos[j]] += int(h * 0.15)
        else:
            tr[a, pos[j]] += h
    if style == 2:                                        # non-ACGT codes among the differing positions
        for j in list(diff)[:3]:
            sec[j] = ord("N")
        for j in list(diff)[3:5]:
            pri[j] = ord("R")
    if style == 3:                                        # a dead position: all four channels zero -> NaN profile
        j = sorted(diff)[len(diff) // 2]
        tr[:, pos[j]] = 0
    return tr, pos, bytes(pri), bytes(sec)


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(5150)
    d = {}
    specs = [(300, 0.5, 40, 0, 20, 20), (500, 0.7, 120, 0, 50, 50), (450, 0.33, 200, 1, 50, 50), (200, 0.9, 10, 0, 0, 0), (400, 0.6, 60, 2, 30, 10),
             (350, 0.5, 0, 0, 50, 50), (120, 0.4, 60, 0, 50, 50), (300, 0.25, 80, 3, 20, 20), (800, 0.55, 500, 1, 50, 50), (250, 0.02, 50, 0, 10, 10),
             (250, 0.98, 50, 0, 10, 10), (40, 0.6, 20, 0, 15, 15)]
    d["n"] = np.int64(len(specs))
    for i, (nbc, frac, ndiff, style, tl, trr) in enumerate(specs):
        tr, pos, pri, sec = case(rng, nbc, frac, ndiff, style)
        a1, a2 = ref.allelic_fraction(tr, pos, pri, sec, tl, trr)
        print(i, nbc, frac, ndiff, style, "->", a1, a2)
        d[f"tr{i}"], d[f"pos{i}"] = tr, pos
        d[f"pri{i}"], d[f"sec{i}"] = np.frombuffer(pri, np.uint8), np.frombuffer(sec, np.uint8)
        d[f"cfg{i}"] = np.array([tl, trr], np.int64)
        d[f"out{i}"] = np.array([a1, a2], np.float64)
    np.savez_compressed(os.path.join(OUT, "fraction_golden.npz"), **d)


if __name__ == "__main__":
    main()
