"""Generates tests/golden/drivers_golden.npz: the DP sequence of indigo() (reference src/indigo.h:190-388) composed from
the REFERENCE's own functions (oracle/_ref behind oracle/ref_bridge.cpp), one trace at a time, on synthetic heterozygous
traces. tests/test_gpu_drivers.py replays the same inputs through drivers.decompose_batch on the GPU.

    python tests/golden/make_golden_drivers.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tracy_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SC = (3, -5, -10, -4)
IUP = {frozenset("AG"): "R", frozenset("CT"): "Y", frozenset("CG"): "S", frozenset("AT"): "W", frozenset("GT"): "K", frozenset("AC"): "M"}


def het_trace(rng, refseq, start, L, bp_pos, ins, dl, snvs, rc):
    """Two alleles of refseq[start:start+L], the second with an indel at bp_pos; mixed 60/40 in signal space."""
    a1 = bytearray(refseq[start:start + L])
    a2 = bytearray((refseq[start:start + bp_pos] + synth.random_seq(rng, ins) + refseq[start + bp_pos + dl:])[:L])
    for _ in range(snvs):
        a2[int(rng.integers(0, len(a2)))] = b"ACGT"[int(rng.integers(0, 4))]
    if rc:
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        a1, a2 = bytearray(bytes(a1).translate(comp)[::-1]), bytearray(bytes(a2).translate(comp)[::-1])
    nbc = min(len(a1), len(a2))
    ns = 12 * nbc + 40
    tr = rng.integers(0, 20, size=(4, ns)).astype(np.int32)
    pos = (12 * np.arange(nbc) + 10 + rng.integers(-2, 3, nbc)).astype(np.int32)
    pri, sec = bytearray(), bytearray()
    for j in range(nbc):
        x, y = b"ACGT".index(a1[j]), b"ACGT".index(a2[j])
        hx, hy = int(rng.integers(700, 1200)), int(rng.integers(350, 650))
        tr[x, pos[j]] += hx
        tr[y, pos[j]] += hy
        pri.append(a1[j])
        if x == y:
            sec.append(a1[j] if rng.random() > 0.03 else ord("N"))
        elif rng.random() < 0.15:
            sec.append(ord(IUP[frozenset(chr(a1[j]) + chr(a2[j]))]))
        else:
            sec.append(a2[j])
    return tr, pos, bytes(pri), bytes(sec)


def reference_decompose(ref, tr, pos, pri, sec, refseq, tl, trr, maxindel, madc):
    """indigo() for one trace, every step a call into the reference's own code."""
    prof = ref.create_profile(tr, pos, pri, sec, tl, trr)
    bp = ref.find_breakpoint(prof)
    fwdp = ref.onehot(refseq)
    gs_f = ref.gotoh_score(prof, fwdp, 1, 0, SC)
    gs_r = ref.gotoh_score(prof, ref.revcomp_profile(fwdp), 1, 0, SC)
    fw = gs_f > gs_r
    rsl = refseq if fw else ref.reverse_complement(refseq)
    score, r0, r1 = ref.gotoh(prof, rsl, 1, 0, SC)
    seqsize = float(prof.shape[1])
    if score <= seqsize * 0.35 * SC[0] + seqsize * (1 - 0.35) * SC[1]:
        return None
    if not bp[0]:
        bp = ref.find_homozygous_breakpoint(r0, r1)
        if bp is None:
            return None
    p2, s2, dcp = ref.decompose_alleles(r0, r1, pri, sec, tl, trr, maxindel, madc, bp[2], len(rsl))
    sd = ref.generate_secondary_decomposed(tr, pos, p2, s2)

    def tseq(s):
        return s if tl + trr + 1 >= len(s) else s[tl: len(s) - trr]
    out = dict(forward=fw, refslice=rsl, score=score, row0=r0, row1=r1, bp=np.array(bp, np.float64), primary=p2, secondary=s2, secDecompose=sd, decomp=dcp)
    # allelicFraction (src/indigo.h:350); short reads that trimmedSeq() leaves untrimmed index bcPos out of bounds in the reference
    out["frac"] = np.array(ref.allelic_fraction(tr, pos, p2, sd, tl, trr) if tl + trr + 1 < len(p2) else (0.5, 0.5), np.float64)
    for name, q in (("align1", tseq(p2)), ("align2", tseq(sd))):
        _, a0, a1 = ref.gotoh(q, rsl, 1, 0, SC)
        sl, npos = ref.trim_reference_slice(a0, a1, rsl, fw, 0, tl, trr)
        sc2, f0, f1 = ref.gotoh(q, sl, 1, 0, SC)
        out[name] = (sc2, f0, f1, sl, npos)
    sc3, g0, g1 = ref.gotoh(tseq(p2), tseq(sd), 0, 0, SC)
    out["align3"] = (sc3, g0, g1, tseq(sd), 0)
    return out


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(31337)
    d = {}
    n = 14
    d["n"] = np.int64(n)
    for i in range(n):
        nref = int(rng.integers(700, 1100))
        refseq = synth.random_seq(rng, nref)
        start, L, bp_pos = int(rng.integers(20, 120)), int(rng.integers(380, 520)), int(rng.integers(120, 260))
        style = i % 7
        ins, dl, snvs = [(0, 12), (9, 0), (0, 0), (5, 3), (0, 27), (14, 0), (0, 0)][style] + ((4 if style in (2, 6) else 1),)
        tr, pos, pri, sec = het_trace(rng, refseq, start, L, bp_pos, ins, dl, snvs, rc=bool(i % 2))
        if style == 6:                                   # an unrelated trace: indigo() gives up on the score threshold
            refseq = synth.random_seq(rng, nref)
        tl, trr, maxindel = [(20, 20, 30), (50, 50, 1000), (0, 10, 30)][i % 3]
        want = reference_decompose(ref, tr, pos, pri, sec, refseq, tl, trr, maxindel, 5)
        d[f"tr{i}"], d[f"pos{i}"] = tr, pos
        d[f"pri{i}"], d[f"sec{i}"], d[f"ref{i}"] = (np.frombuffer(x, np.uint8) for x in (pri, sec, refseq))
        d[f"cfg{i}"] = np.array([tl, trr, maxindel, 5, 0 if want is None else 1], np.int64)
        print(i, "none" if want is None else (want["forward"], want["score"], want["bp"], len(want["decomp"])))
        if want is None:
            continue
        d[f"fw{i}"] = np.array([want["forward"], want["score"]], np.int64)
        d[f"bp{i}"] = want["bp"]
        for k in ("refslice", "row0", "row1", "primary", "secondary", "secDecompose"):
            d[f"{k}{i}"] = np.frombuffer(want[k], np.uint8)
        d[f"decomp{i}"] = want["decomp"].astype(np.int32)
        d[f"frac{i}"] = want["frac"]
        for name in ("align1", "align2", "align3"):
            sc_, a0, a1, sl, npos = want[name]
            d[f"{name}_s{i}"] = np.array([sc_, npos], np.int64)
            d[f"{name}_r0{i}"], d[f"{name}_r1{i}"], d[f"{name}_sl{i}"] = (np.frombuffer(x, np.uint8) for x in (a0, a1, sl))
    np.savez_compressed(os.path.join(OUT, "drivers_golden.npz"), **d)
    print("wrote drivers_golden.npz", os.path.getsize(os.path.join(OUT, "drivers_golden.npz")))


if __name__ == "__main__":
    main()
