"""Golden variant calls of the reference's callVariants / insertVariant / variantType (src/variants.h:34-138) through
oracle/ref_bridge.cpp for tests/test_variants.py. Run in the build container: python tests/golden/make_golden_variants.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402


def rand_align(rng, n):
    r0, r1 = bytearray(), bytearray()
    for _ in range(n):
        u = rng.random()
        b = b"ACGTN"[int(rng.integers(0, 5 if rng.random() < 0.05 else 4))]
        if u < 0.8:
            r0.append(b); r1.append(b if rng.random() < 0.9 else b"ACGT"[int(rng.integers(0, 4))])
        elif u < 0.9:
            r0 += b"-"; r1.append(b)
        else:
            r0.append(b); r1 += b"-"
    lead, trail = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    r0 = bytearray(b"-" * lead) + r0 + b"-" * trail
    r1 = bytearray(bytes(rng.choice(list(b"ACGT"), lead).astype(np.uint8))) + r1 + bytes(rng.choice(list(b"ACGT"), trail).astype(np.uint8))
    return bytes(r0), bytes(r1)


def cases(seed, n):
    """Seeded groups of 1-3 alignments that share one variant vector (the two alleles of indigo(), src/indigo.h:405-422); a
    later alignment often repeats the first one, so that homozygous calls (gt + 1) occur."""
    rng = np.random.default_rng(seed)
    for _ in range(n):
        base = rand_align(rng, int(rng.integers(0, 80)))
        als = []
        for k in range(int(rng.integers(1, 4))):
            a = base if (k and rng.random() < 0.5) else rand_align(rng, int(rng.integers(0, 80)))
            als.append((a[0], a[1], b"chr3" if rng.random() < 0.8 else b"chrX", int(rng.integers(0, 3)) if rng.random() < 0.2 else int(rng.integers(0, 10 ** 6))))
        yield als


if __name__ == "__main__":
    ref = loader.ref()
    assert ref is not None, "needs the reference build (oracle/_ref/libtracy_ref.so)"
    out = [ref.call_variants(als) for als in cases(11, 60)]
    with open(os.path.join(ROOT, "tests", "golden", "variants_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote variants_golden.json:", sum(len(x) for x in out), "variants in", len(out), "cases")
