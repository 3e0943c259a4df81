"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libtracy_ref.so = unmodified tracy headers,
see oracle/ref_bridge.cpp). Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Upstream has no tests or fixtures for this path; these vectors are what pins the oracle and the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tracy_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SCORES = [(3, -5, -10, -4), (5, -4, -10, -1), (2, -3, -5, -2), (1, -1, -1, -1), (7, -6, -12, 0)]


def gotoh_cases(ref):
    rng = np.random.default_rng(20261017)
    cases = []
    shapes = [(1, 1), (1, 9), (9, 1), (2, 3), (15, 17), (16, 16), (17, 33), (31, 64), (33, 100), (64, 31), (100, 200), (130, 97),
              (257, 300), (300, 120), (513, 70), (40, 700)]
    for i, (m, n) in enumerate(shapes * 3):
        kind = ["ps", "pp", "ss"][(i // len(shapes)) % 3]
        hf, vf = (i >> 0) & 1, (i >> 1) & 1
        sc = SCORES[i % len(SCORES)]
        if kind == "ss":
            a = synth.random_seq(rng, m, b"ACGTN")
            b = synth.random_seq(rng, n, b"ACGTN") if i % 2 else (synth.mutate_seq(rng, (a * (n // m + 1))[:n], 0.1, 0.05) or b"A")
        else:
            a = synth.random_profile(rng, m, ["trace", "ties", "msa"][i % 3])
            if kind == "ps":
                if i % 2:
                    b = synth.random_seq(rng, n, b"ACGTNacgtn-RYX" if i % 4 == 1 else b"ACGT")
                else:  # related sequence: consensus of the profile embedded in random flanks
                    cons = bytes(b"ACGTNN"[int(k)] for k in np.argmax(a, axis=0))
                    core = synth.mutate_seq(rng, cons.replace(b"N", b"A"), 0.03, 0.03)
                    fl = max(n - len(core), 0)
                    b = (synth.random_seq(rng, fl // 2) + core + synth.random_seq(rng, fl - fl // 2))[:n] or b"A"
            else:
                b = synth.random_profile(rng, n, ["msa", "trace", "ties"][i % 3])
        score, r0, r1 = ref.gotoh(a, b, hf, vf, sc)
        assert ref.gotoh_score(a, b, hf, vf, sc) == score
        cases.append(dict(kind=kind, a=a, b=b, hf=hf, vf=vf, sc=sc, score=score, row0=r0, row1=r1))
    # two config-2 shaped pairs (1 kb x 4 kb, 3/-5/-10/-4, <true,false>)
    prof, win = synth.align_batch(2, 1000, 4000, seed=44, rc_frac=0.0)
    for k in range(2):
        b = bytes(win[k])
        score, r0, r1 = ref.gotoh(prof[k], b, 1, 0, SCORES[0])
        cases.append(dict(kind="ps", a=prof[k], b=b, hf=1, vf=0, sc=SCORES[0], score=score, row0=r0, row1=r1))
    return cases


def save_gotoh(cases):
    d = {"n": np.int64(len(cases))}
    for i, c in enumerate(cases):
        d[f"kind{i}"] = np.frombuffer(c["kind"].encode(), np.uint8)
        d[f"a{i}"] = c["a"] if isinstance(c["a"], np.ndarray) else np.frombuffer(c["a"], np.uint8)
        d[f"b{i}"] = c["b"] if isinstance(c["b"], np.ndarray) else np.frombuffer(c["b"], np.uint8)
        d[f"cfg{i}"] = np.array([c["hf"], c["vf"], *c["sc"], c["score"]], np.int64)
        d[f"row0_{i}"] = np.frombuffer(c["row0"], np.uint8)
        d[f"row1_{i}"] = np.frombuffer(c["row1"], np.uint8)
    np.savez_compressed(os.path.join(OUT, "gotoh_golden.npz"), **d)


def decompose_cases(ref):
    """Heterozygous-indel style inputs pushed through the reference's decomposeAlleles (src/decompose.h:179-376).
    Both alleles are built from the reference window with their own (insertion, deletion) at the breakpoint; candidates are
    drawn at random and up to 8 per decision branch (deletion / insertion / complex grid / no indel) are kept. The branch
    label only steers the selection -- every stored output comes from the reference."""
    from tracy_b200 import decompose
    port = loader.port()
    rng = np.random.default_rng(7)
    IUP = {frozenset("AG"): "R", frozenset("CT"): "Y", frozenset("CG"): "S", frozenset("AT"): "W", frozenset("GT"): "K", frozenset("AC"): "M"}
    kept, per_mode = [], {}
    for t in range(400):
        nref = int(rng.integers(500, 900))
        refseq = synth.random_seq(rng, nref)
        start = int(rng.integers(20, 60))
        L = int(rng.integers(300, 420))
        bp_pos = int(rng.integers(80, 200))

        def allele(ins, dl):
            return bytearray((refseq[start:start + bp_pos] + synth.random_seq(rng, ins) + refseq[start + bp_pos + dl:])[:L])

        style = t % 5
        i1 = d1 = i2 = d2 = 0
        if style == 0: d2 = int(rng.integers(1, 25))
        elif style == 1: i2 = int(rng.integers(1, 25))
        elif style == 2: d1 = int(rng.integers(1, 12)); d2 = d1 + int(rng.integers(1, 12))
        elif style == 3: i1 = int(rng.integers(1, 12)); i2 = int(rng.integers(1, 12)); d2 = int(rng.integers(0, 8))
        a1, a2 = allele(i1, d1), allele(i2, d2)
        if style == 4:
            for _ in range(4):
                a2[int(rng.integers(0, L))] = b"ACGT"[int(rng.integers(0, 4))]
        pri, sec = bytearray(), bytearray()
        for x, y in zip(a1, a2):
            if x == y:
                pri.append(x); sec.append(x if rng.random() > 0.02 else ord("N"))
            else:
                # basecall(): primary = the higher peak, secondary = iupac(leftover) (reference src/abif.h:470-483)
                hi, lo = (x, y) if rng.random() < 0.5 else (y, x)
                pri.append(hi)
                if rng.random() < 0.1:
                    third = [b for b in b"ACGT" if b not in (x, y)][int(rng.integers(0, 2))]
                    sec.append(ord(IUP[frozenset(chr(lo) + chr(third))]))
                else:
                    sec.append(lo)
            if rng.random() < 0.01:
                pri[-1] = ord("N")
        pri, sec = bytes(pri), bytes(sec)
        trimL, trimR = int(rng.integers(0, 30)), int(rng.integers(0, 30))
        score, r0, r1 = ref.gotoh(pri[trimL:L - trimR], refseq, 1, 0, (3, -5, -10, -4))   # the mixed primary, like the real pipeline
        breakpoint_ = max(bp_pos - trimL, 1)
        maxindel = [30, 1000, 10, 30][t % 4]
        sweep = lambda *a: port.decompose_sweep(*a)
        mode = decompose.decompose_alleles(r0, r1, pri, sec, trimL, trimR, maxindel, 5, breakpoint_, nref, sweep)[3]["mode"]
        if per_mode.get(mode, 0) >= 8:
            continue
        per_mode[mode] = per_mode.get(mode, 0) + 1
        p2, s2, dcp = ref.decompose_alleles(r0, r1, pri, sec, trimL, trimR, maxindel, 5, breakpoint_, nref)
        kept.append(dict(row0=r0, row1=r1, pri=pri, sec=sec, trimL=trimL, trimR=trimR, maxindel=maxindel, madc=5, bp=breakpoint_,
                         nref=nref, pri_out=p2, sec_out=s2, dcp=dcp))
    print("decompose branches kept:", per_mode)
    return kept


def save_decompose(cases):
    d = {"n": np.int64(len(cases))}
    for i, c in enumerate(cases):
        for k in ("row0", "row1", "pri", "sec", "pri_out", "sec_out"):
            d[f"{k}{i}"] = np.frombuffer(c[k], np.uint8)
        d[f"cfg{i}"] = np.array([c["trimL"], c["trimR"], c["maxindel"], c["madc"], c["bp"], c["nref"]], np.int64)
        d[f"dcp{i}"] = c["dcp"].astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "decompose_golden.npz"), **d)


def main():
    ref = loader.ref()
    assert ref is not None, "oracle/_ref/libtracy_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    g = gotoh_cases(ref)
    save_gotoh(g)
    d = decompose_cases(ref)
    save_decompose(d)
    print(f"wrote {len(g)} gotoh cases and {len(d)} decompose cases to {OUT}")


if __name__ == "__main__":
    main()
