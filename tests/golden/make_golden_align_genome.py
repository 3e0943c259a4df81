"""Generates tests/golden/align_genome_golden.npz: sage() against an indexed genome (reference src/sage.h:216-222, :258-260,
:311) composed one trace at a time from the REFERENCE's own functions (oracle/_ref): createProfile, getReferenceSlice over
sdsl's FM-index (faidx served by the bridge's in-memory stand-in), gotoh, trimReferenceSlice, gotoh.

    python tests/golden/make_golden_align_genome.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SC = (3, -5, -10, -4)
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def sanger(rng, seq):
    """A clean synthetic trace of `seq`: peaks every ~12 samples, a weak second peak now and then."""
    nbc = len(seq)
    ns = 12 * nbc + 40
    tr = rng.integers(0, 20, size=(4, ns)).astype(np.int32)
    pos = (12 * np.arange(nbc) + 10 + rng.integers(-2, 3, nbc)).astype(np.int32)
    for j, ch in enumerate(seq):
        tr[b"ACGT".index(ch), pos[j]] += int(rng.integers(700, 1200))
        if rng.random() < 0.05:
            tr[int(rng.integers(0, 4)), pos[j]] += int(rng.integers(100, 400))
    return tr, pos


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(2718)
    names = [b"chr1", b"chr2"]
    seqs = [bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8)) for n in (9000, 6000)]
    text = b"\n".join(seqs) + b"\n"
    h = ref.fm_build(text)
    ref.set_genome(names, seqs)
    d = dict(text=np.frombuffer(text, np.uint8), n=np.int64(10))
    tl, trr, kmer, maxindel, ms = 50, 50, 15, 300, 3
    d["cfg"] = np.array([tl, trr, kmer, maxindel, ms], np.int64)
    for i in range(10):
        c = i % 2
        L = int(rng.integers(350, 700))
        p = int(rng.integers(0, len(seqs[c]) - L)) if i not in (4, 5) else (0 if i == 4 else len(seqs[c]) - L)
        s = bytearray(seqs[c][p:p + L])
        for q in rng.integers(60, L - 60, 5):
            s[q] = b"ACGT"[int(rng.integers(0, 4))]
        if i % 3 == 0:
            del s[200:203]
        s = bytes(s)
        if i == 9:
            s = bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8))     # unanchorable
        if i % 2:
            s = s.translate(COMP)[::-1]
        tr, pos = sanger(rng, s)
        bc = ref.basecall(tr, pos, 0.33)
        bcpos, pri, sec, cons = bc["bcPos"], bc["primary"], bc["secondary"], bc["consensus"]
        full = ref.create_profile(tr, bcpos, pri, sec, 0, 0)
        trimmed = ref.create_profile(tr, bcpos, pri, sec, tl, trr)
        d[f"tr{i}"], d[f"ploc{i}"] = tr, pos
        g = ref.get_reference_slice(h, 0, cons, tl, trr, kmer, maxindel, ms)
        d[f"ok{i}"] = np.int64(g["ok"])
        print(i, g["ok"], g["forward"], g["kmersupport"], g["pos"], g["chr"], len(g["refslice"]))
        if not g["ok"]:
            continue
        _, r0, r1 = ref.gotoh(trimmed, g["refslice"], 1, 0, SC)
        sl, npos = ref.trim_reference_slice(r0, r1, g["refslice"], g["forward"], g["pos"], tl, trr)
        score, f0, f1 = ref.gotoh(full, sl, 1, 0, SC)
        d[f"meta{i}"] = np.array([g["forward"], names.index(g["chr"]), g["kmersupport"], npos, score], np.int64)
        for k, v in (("refslice", sl), ("row0", f0), ("row1", f1)):
            d[f"{k}{i}"] = np.frombuffer(v, np.uint8)
    ref.fm_free(h)
    np.savez_compressed(os.path.join(OUT, "align_genome_golden.npz"), **d)


if __name__ == "__main__":
    main()
