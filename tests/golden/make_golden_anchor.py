"""Generates tests/golden/anchor_golden.npz from the REFERENCE ITSELF (oracle/_ref/libtracy_ref.so: unmodified
src/fmindex.h over the vendored sdsl csa_wt<>, htslib's faidx replaced by the in-memory stand-in of oracle/ref_bridge.cpp)
for SURVEY section 8f rank 2: scanSequence, findMaxFreq, getReferenceSlice. Run in the build container:

    python tests/golden/make_golden_anchor.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")
IUPAC = b"RYSWKMBDHV"


def genome(rng):
    """Three sequences with a 2.5 kb segment duplicated between chrA and chrB, an N run and a few IUPAC codes."""
    seqs = [bytearray(rng.choice(list(b"ACGT"), n).astype(np.uint8).tobytes()) for n in (30000, 20000, 9000)]
    seqs[1][4000:6500] = seqs[0][11000:13500]            # repeat: only non-unique k-mers anchor here
    seqs[0][20000:20300] = b"N" * 300
    for s in seqs:
        for p in rng.integers(0, len(s), 6):
            s[p] = IUPAC[int(rng.integers(0, len(IUPAC)))]
    seqs[2][100:2600] = seqs[2][3000:5500]               # tandem-ish duplicate inside one sequence
    return [b"chrA", b"chrB", b"chrC"], [bytes(s) for s in seqs]


def mutate(rng, s, sub=0.01, iupac=0.02, nrate=0.005, indel=0.003):
    out = bytearray()
    for ch in s:
        u = rng.random()
        if u < indel:
            continue
        if u < 2 * indel:
            out.append(b"ACGT"[int(rng.integers(0, 4))])
        v = rng.random()
        if v < sub:
            ch = b"ACGT"[int(rng.integers(0, 4))]
        elif v < sub + iupac:
            ch = IUPAC[int(rng.integers(0, len(IUPAC)))]
        elif v < sub + iupac + nrate:
            ch = ord("N")
        out.append(ch)
    return bytes(out)


def traces(rng, seqs):
    T = []
    for i in range(26):
        c = int(rng.integers(0, 3))
        L = int(rng.integers(200, 1300))
        p = int(rng.integers(0, len(seqs[c]) - L))
        t = mutate(rng, seqs[c][p:p + L])
        if i % 2:
            t = t.translate(COMP)[::-1]
        T.append(t)
    T.append(mutate(rng, seqs[1][4100:6400], iupac=0.0))          # inside the repeat: unique pass fails
    T.append(mutate(rng, seqs[0][11200:13300]).translate(COMP)[::-1])
    T.append(mutate(rng, seqs[2][200:2500], sub=0.0, iupac=0.0, nrate=0.0, indel=0.0))
    T.append(bytes(rng.choice(list(b"ACGT"), 700).astype(np.uint8).tobytes()))   # not in the genome
    T.append(seqs[0][100:140])                                   # shorter than the trims
    T.append(seqs[0][500:510])                                   # shorter than a k-mer
    T.append(seqs[0][0:600])                                     # sequence start: hits at position - k < trim
    T.append(seqs[2][-500:])                                     # sequence end
    T.append(seqs[0][-300:] )                                    # ends at the '\n' joint
    T.append(b"N" * 400)
    T.append(seqs[1][9000:9800].lower())                         # lower case never matches an upper-cased text
    T.append(seqs[0][19800:20600])                               # spans the N run
    return T


def main():
    ref = loader.ref()
    assert ref is not None, "needs oracle/_ref/libtracy_ref.so (build container)"
    rng = np.random.default_rng(4242)
    names, seqs = genome(rng)
    text = b"\n".join(seqs) + b"\n"                              # what `tracy index` dumps (src/index.h:104-116)
    tr = traces(rng, seqs)
    out = dict(text=np.frombuffer(text, np.uint8), names=np.frombuffer(b"\n".join(names), np.uint8), ntraces=len(tr))
    for i, t in enumerate(tr):
        out[f"cons{i}"] = np.frombuffer(t, np.uint8)
    cfgs = [(50, 50, 15, 1000, 3), (0, 0, 12, 300, 3), (20, 70, 16, 50, 10), (50, 50, 9, 1000, 3)]
    out["cfgs"] = np.array(cfgs, np.int32)
    h = ref.fm_build(text)
    ref.set_genome(names, seqs)
    for ci, (tl, trr, k, mi, ms) in enumerate(cfgs):
        rows, slices = [], []
        for i, t in enumerate(tr):
            rv = ref.reverse_complement(t)
            per = []
            for uniq in (True, False):
                hf, gf, ff = ref.scan_sequence(h, t, tl, trr, k, uniq)
                hr, gr, fr = ref.scan_sequence(h, rv, trr, tl, k, uniq)
                per += [len(hf), gf, ff, len(hr), gr, fr]
            r = ref.get_reference_slice(h, 0, t, tl, trr, k, mi, ms)
            rows.append(per + [int(r["ok"]), int(r["forward"]), r["kmersupport"], r["pos"], names.index(r["chr"]) if r["ok"] else -1, len(r["refslice"]) if r["ok"] else 0])
            slices.append(r["refslice"] if r["ok"] else b"")
        out[f"rows{ci}"] = np.array(rows, np.int64)
        out[f"slices{ci}"] = np.frombuffer(b"\n".join(slices), np.uint8)
    ref.fm_free(h)
    # single-sequence (FASTA / wild-type) reference: filetype 1 keeps the whole refslice (src/fmindex.h:300-306)
    h1 = ref.fm_build(seqs[0])
    rows = []
    for i, t in enumerate(tr[:12]):
        r = ref.get_reference_slice(h1, 1, t, 50, 50, 15, 1000, 3, refslice=seqs[0])
        rows.append([int(r["ok"]), int(r["forward"]), r["kmersupport"]])
    out["fasta_rows"] = np.array(rows, np.int64)
    ref.fm_free(h1)
    np.savez_compressed(os.path.join(OUT, "anchor_golden.npz"), **out)
    for ci in range(len(cfgs)):
        r = out[f"rows{ci}"]
        print("cfg", cfgs[ci], "anchored", int(r[:, 12].sum()), "of", len(r), "| forward", int(r[:, 13].sum()), "| unique-pass decided",
              int(sum(1 for x in r if (x[2] >= cfgs[ci][4] and x[2] > 2 * x[5]) or (x[5] >= cfgs[ci][4] and x[5] > 2 * x[2]))))


if __name__ == "__main__":
    main()
