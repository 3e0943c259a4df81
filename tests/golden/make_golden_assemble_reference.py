"""Goldens of the reference-guided branch of assemble() (src/assemble.h:163-292): the DP sequence composed from the reference's own
functions in oracle/ref_bridge.cpp (ref_assemble_reference), inputs included, for tests/test_glue.py and tests/test_gpu_glue.py.
Run in the build container: python tests/golden/make_golden_assemble_reference.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tracy_b200 import synth  # noqa: E402

SC = (3, -5, -10, -4)

if __name__ == "__main__":
    ref = loader.ref()
    assert ref is not None, "needs the reference build (oracle/_ref/libtracy_ref.so)"
    rng = np.random.default_rng(61)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    clean = bytes.maketrans(b"nN-x", b"ACGT")
    out = {}
    ncase = 8
    for it in range(ncase):
        L = int(rng.integers(150, 320))
        reference = synth.random_seq(rng, L)
        if it % 3 == 0:
            reference = reference[:50] + b"nN-x" + reference[54:]
        n = int(rng.integers(1, 7))
        profs = []
        for i in range(n):
            a = int(rng.integers(0, L - 80))
            s = synth.mutate_seq(rng, reference[a: a + int(rng.integers(60, 130))].translate(clean), 0.03, 0.02)
            if rng.random() < 0.25:
                s = synth.random_seq(rng, len(s))                      # unrelated: below the match threshold
            if rng.random() < 0.5:
                s = s.translate(comp)[::-1]
            profs.append(synth.profile_from_seq(rng, s, 0.15))
        inc = bool(it % 2)
        w = ref.assemble_reference(profs, reference, SC, 0.5, 0.5, inc)
        out[f"ref{it}"] = np.frombuffer(reference, np.uint8)
        out[f"n{it}"] = np.int64(n)
        out[f"inc{it}"] = np.int64(inc)
        for i, p in enumerate(profs):
            out[f"p{it}_{i}"] = p
        out[f"rows{it}"] = w["rows"]
        out[f"idx{it}"] = np.array(w["idx"], np.int64)
        out[f"fwd{it}"] = np.array(w["forward"], np.int64)
        for k in ("gapped", "cons", "qual"):
            out[f"{k}{it}"] = np.frombuffer(w[k], np.uint8)
    out["ncase"] = np.int64(ncase)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "assemble_reference_golden.npz"), **out)
    print("wrote assemble_reference_golden.npz:", ncase, "cases,", sum(len(out[f"idx{i}"]) for i in range(ncase)), "traces kept")
