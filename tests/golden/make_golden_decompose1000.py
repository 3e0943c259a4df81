"""Goldens for decomposeAlleles at the CLI-default maxindel = 1000 (reference src/indigo.h) on config-3-sized
inputs: ~1 kb heterozygous traces against 4 kb windows, with long deletions / insertions and the ins x del fallback grid
(reference src/decompose.h:288-313). Every stored output comes from the reference's own decomposeAlleles (oracle/_ref).
Writes tests/golden/decompose1000_golden.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader               # noqa: E402
from tracy_b200 import decompose, synth  # noqa: E402

IUP = {frozenset("AG"): "R", frozenset("CT"): "Y", frozenset("CG"): "S", frozenset("AT"): "W", frozenset("GT"): "K", frozenset("AC"): "M"}


def main():
    ref, port = loader.ref(), loader.port()
    assert ref is not None
    rng = np.random.default_rng(1000)
    kept, per_mode = [], {}
    for t in range(60):
        nref = 4000
        refseq = synth.random_seq(rng, nref)
        start = int(rng.integers(200, 1500))
        L = int(rng.integers(850, 1000))
        bp_pos = int(rng.integers(150, 400))

        def allele(ins, dl):
            return bytearray((refseq[start:start + bp_pos] + synth.random_seq(rng, ins) + refseq[start + bp_pos + dl:])[:L])
        style = t % 4
        i1 = d1 = i2 = d2 = 0
        if style == 0: d2 = int(rng.integers(40, 600))
        elif style == 1: i2 = int(rng.integers(40, 250))
        elif style == 2: i1 = int(rng.integers(5, 60)); i2 = int(rng.integers(5, 60)); d2 = int(rng.integers(20, 300))
        a1, a2 = allele(i1, d1), allele(i2, d2)
        if style == 3:
            for _ in range(6):
                a2[int(rng.integers(0, L))] = b"ACGT"[int(rng.integers(0, 4))]
        pri, sec = bytearray(), bytearray()
        for x, y in zip(a1, a2):
            if x == y:
                pri.append(x); sec.append(x if rng.random() > 0.02 else ord("N"))
            else:
                hi, lo = (x, y) if rng.random() < 0.5 else (y, x)
                pri.append(hi)
                sec.append(lo)
        pri, sec = bytes(pri), bytes(sec)
        trimL, trimR = int(rng.integers(0, 50)), int(rng.integers(0, 50))
        score, r0, r1 = ref.gotoh(pri[trimL:L - trimR], refseq, 1, 0, (3, -5, -10, -4))
        breakpoint_ = max(bp_pos - trimL, 1)
        sweep = lambda *a: port.decompose_sweep(*a)
        mode = decompose.decompose_alleles(r0, r1, pri, sec, trimL, trimR, 1000, 5, breakpoint_, nref, sweep)[3]["mode"]
        if per_mode.get(mode, 0) >= 3:
            continue
        per_mode[mode] = per_mode.get(mode, 0) + 1
        p2, s2, dcp = ref.decompose_alleles(r0, r1, pri, sec, trimL, trimR, 1000, 5, breakpoint_, nref)
        kept.append(dict(row0=r0, row1=r1, pri=pri, sec=sec, trimL=trimL, trimR=trimR, maxindel=1000, madc=5, bp=breakpoint_, nref=nref,
                         pri_out=p2, sec_out=s2, dcp=dcp))
    print("branches kept:", per_mode)
    d = {"n": np.int64(len(kept))}
    for i, c in enumerate(kept):
        for k in ("row0", "row1", "pri", "sec", "pri_out", "sec_out"):
            d[f"{k}{i}"] = np.frombuffer(c[k], np.uint8)
        d[f"dcp{i}"] = np.asarray(c["dcp"], np.int64)
        d[f"cfg{i}"] = np.array([c["trimL"], c["trimR"], c["maxindel"], c["madc"], c["bp"], c["nref"]], np.int64)
    out = os.path.join(ROOT, "tests", "golden", "decompose1000_golden.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out), "bytes,", len(kept), "cases")


if __name__ == "__main__":
    main()
