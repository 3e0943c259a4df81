"""Generates tests/golden/writers_golden.npz from the REFERENCE ITSELF (oracle/_ref): plotAlignment (src/fmindex.h:329-420)
and writeDecomposition (src/decompose.h:621-627) outputs for synthetic alignments.

    python tests/golden/make_golden_writers.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def rows(rng, L):
    r0, r1 = bytearray(), bytearray()
    for _ in range(L):
        u = rng.random()
        a = b"ACGTN"[int(rng.integers(0, 5))]
        if u < 0.07:
            r0.append(ord("-")); r1.append(a)
        elif u < 0.14:
            r0.append(a); r1.append(ord("-"))
        elif u < 0.25:
            r0.append(a); r1.append(b"ACGT"[int(rng.integers(0, 4))])
        else:
            r0.append(a); r1.append(a)
    return bytes(r0), bytes(r1)


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(12)
    d = {}
    cases = [(130, b"chr7", 1000, True, 0, 60), (75, b"chrX", 0, False, 0, 60), (420, b"ref", 52, True, 1, 60), (400, b"ref", 52, False, 2, 80),
             (260, b"", 7, True, 3, 60), (0, b"c", 5, True, 0, 60), (60, b"chr1", 4_000_000_000, False, 0, 20), (361, b"wildtype", 0, True, 1, 100)]
    d["n"] = np.int64(len(cases))
    for i, (L, chr_name, pos, fw, key, ll) in enumerate(cases):
        r0, r1 = rows(rng, L)
        rl = len(r1) - r1.count(b"-")
        score = int(rng.integers(-500, 3000))
        a1a2 = (float(rng.integers(0, 101)) * 0.01, float(rng.random()))
        txt = ref.plot_alignment(r0, r1, chr_name, pos, rl, fw, score, key, a1a2, ll)
        d[f"r0_{i}"], d[f"r1_{i}"], d[f"chr{i}"] = (np.frombuffer(x, np.uint8) for x in (r0, r1, chr_name))
        d[f"cfg{i}"] = np.array([pos, rl, int(fw), score, key, ll], np.int64)
        d[f"a1a2_{i}"] = np.array(a1a2, np.float64)
        d[f"txt{i}"] = np.frombuffer(txt, np.uint8)
    dc = np.array([[0, 31], [1, 27], [-1, 99], [12, 3], [-30, 0]], np.int32)
    d["decomp"] = dc
    d["decomp_txt"] = np.frombuffer(ref.write_decomposition(dc), np.uint8)
    np.savez_compressed(os.path.join(OUT, "writers_golden.npz"), **d)
    print(bytes(d["txt1"]).decode()[:400])


if __name__ == "__main__":
    main()
