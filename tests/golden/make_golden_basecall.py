"""Generates tests/golden/basecall_golden.npz from the REFERENCE's basecall() (src/abif.h:408-511, via oracle/_ref):
synthetic chromatograms (Gaussian peaks every ~12 samples, noise, heterozygous double peaks, dropped / crowded positions).

    python tests/golden/make_golden_basecall.py       (build container only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def chromatogram(rng, nb, style):
    spacing = 12
    ns = spacing * nb + 60
    x = np.arange(ns)[None, :]
    tr = rng.integers(0, 18, size=(4, ns)).astype(np.float64)
    ploc = []
    for j in range(nb):
        c = 20 + spacing * j + int(rng.integers(-2, 3))
        b = int(rng.integers(0, 4))
        h = float(rng.integers(400, 1500))
        tr[b] += h * np.exp(-0.5 * ((x[0] - c) / 2.2) ** 2)
        u = rng.random()
        if u < 0.15:                                   # heterozygous: second peak, sometimes slightly shifted
            b2 = (b + int(rng.integers(1, 4))) % 4
            tr[b2] += h * float(rng.uniform(0.2, 1.0)) * np.exp(-0.5 * ((x[0] - c - int(rng.integers(-1, 2))) / 2.2) ** 2)
        elif u < 0.2 and style == 1:                   # three or four channels up: 'N'
            for b2 in range(4):
                tr[b2] += h * 0.8 * np.exp(-0.5 * ((x[0] - c) / 2.5) ** 2)
        elif u < 0.25 and style == 2:                  # flat region: no local maximum, midpoint fallback
            tr[:, c - 5: c + 6] = 3
        p = c + int(rng.integers(-1, 2))
        if style == 2 and rng.random() < 0.05 and ploc:
            p = ploc[-1]                               # duplicated position: empty window, dropped by the reference
        ploc.append(max(p, ploc[-1] if ploc else 0))
    return np.round(tr).astype(np.int32), np.array(ploc, np.int32)


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(424242)
    d = {}
    n = 16
    d["n"] = np.int64(n)
    for i in range(n):
        nb = [1, 2, 5, 60, 250, 700][i % 6]
        tr, ploc = chromatogram(rng, nb, i % 3)
        ratio = [0.33, 0.2, 0.5][i % 3]
        r = ref.basecall(tr, ploc, ratio)
        d[f"tr{i}"], d[f"ploc{i}"], d[f"ratio{i}"] = tr, ploc, np.float32(ratio)
        d[f"pos{i}"] = r["bcPos"]
        for k in ("primary", "secondary", "consensus"):
            d[f"{k}{i}"] = np.frombuffer(r[k], np.uint8)
        print(i, nb, len(r["bcPos"]), r["primary"][:30], r["secondary"][:30])
    np.savez_compressed(os.path.join(OUT, "basecall_golden.npz"), **d)
    print("wrote", os.path.getsize(os.path.join(OUT, "basecall_golden.npz")))


if __name__ == "__main__":
    main()
