"""Golden P.json documents of `tracy decompose` from the reference's traceAlleleAlignJsonOut (src/json.h:260-381, over callVariants
and std::sort as in src/indigo.h:405-446) through oracle/ref_bridge.cpp, for tests/test_writers.py.
Run in the build container: python tests/golden/make_golden_decompose_json.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_golden_variants", os.path.join(ROOT, "tests", "golden", "make_golden_variants.py"))
_var = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_var)


def cases(seed, n):
    """Seeded inputs: two allele alignments (sometimes identical: homozygous calls), an allele1-vs-allele2 alignment, a trace long
    enough for every variant's basecall index in either orientation, a decomposition table, breakpoint and fractions."""
    rng = np.random.default_rng(seed)
    made = 0
    while made < n:
        a1 = _var.rand_align(rng, int(rng.integers(5, 60)))
        same = rng.random() < 0.3
        a2 = a1 if same else _var.rand_align(rng, int(rng.integers(5, 60)))
        pos1, pos2 = int(rng.integers(1, 10 ** 5)), int(rng.integers(1, 10 ** 5))
        als = [(a1[0], a1[1], b"chr3", pos1), (a2[0], a2[1], b"chr3", pos1 if same else pos2)]
        nb_al = max(sum(c != 45 for c in a1[0]), sum(c != 45 for c in a2[0]))
        tl, tr = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        nb = nb_al + tl + tr + 3
        ns = nb * 12 + 40
        acgt = rng.integers(0, 3000, (4, ns)).astype(np.int32)
        bcpos = (np.arange(nb) * 12 + int(rng.integers(3, 20))).astype(np.int32)
        qual = rng.integers(0, 61, nb).astype(np.uint8)
        pri = bytes(rng.choice(list(b"ACGT"), nb).astype(np.uint8))
        sec = bytes(np.where(rng.random(nb) < 0.7, np.frombuffer(pri, np.uint8), rng.choice(list(b"ACGTRYSWKM"), nb)).astype(np.uint8))
        a3 = _var.rand_align(rng, int(rng.integers(0, 50)))
        c = dict(als=als, acgt=acgt, bcpos=bcpos, qual=qual, pri=pri, sec=sec,
                 cfg=dict(trim_left=tl, trim_right=tr, qual_cut=int(rng.integers(0, 60)), pratio=0.33, input="/data/run/trace%d.ab1" % made, genome="ref.fa.gz"),
                 allele1=(a1[0], a1[1], b"chr3", pos1, bool(made % 2), int(rng.integers(-500, 3000))),
                 allele2=(a2[0], a2[1], b"chr3", als[1][3], bool(made % 3), int(rng.integers(-500, 3000))),
                 align3=(a3[0], a3[1], int(rng.integers(-100, 900))), decomp=[(int(i), int(rng.integers(0, 200))) for i in range(-5, 6)],
                 indelshift=bool(made % 2), breakpoint=int(rng.integers(0, nb - tl)), a1a2=(float(rng.random()), float(rng.random())))
        yield c
        made += 1


def reference_json(ref, c):
    return ref.decompose_json(c["cfg"], c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["als"], c["allele1"], c["allele2"], c["align3"], c["decomp"],
                              c["indelshift"], c["breakpoint"], c["a1a2"], sort=True)


if __name__ == "__main__":
    ref = loader.ref()
    assert ref is not None, "needs the reference build (oracle/_ref/libtracy_ref.so)"
    out = {}
    n = 10
    for i, c in enumerate(cases(21, n)):
        out[f"json{i}"] = np.frombuffer(reference_json(ref, c), np.uint8)
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "decompose_json_golden.npz"), **out)
    print("wrote decompose_json_golden.npz,", n, "cases")
