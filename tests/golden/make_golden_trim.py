"""Golden outputs of the reference's estimateQualities / findBestTraceSection (src/abif.h:164-253) and trimTrace (src/trim.h:35-99)
through oracle/ref_bridge.cpp for tests/test_trim.py. Run in the build container: python tests/golden/make_golden_trim.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402


def cases(seed, n_cases):
    """Seeded basecall tracks: 2..700 basecalls, regular or irregular peak distances, few or many ambiguous secondary calls, noisy
    ends; a trimming stringency and trim counts for the basecall trimming."""
    rng = np.random.default_rng(seed)
    for it in range(n_cases):
        n = int(rng.integers(12, 700)) if it % 5 else int(rng.integers(2, 12))
        gaps = rng.integers(6, 18, n)
        if it % 3 == 0:
            gaps[rng.integers(0, n, max(1, n // 20))] = rng.integers(1, 60, max(1, n // 20))
        bcpos = (np.cumsum(gaps) + int(rng.integers(0, 30))).astype(np.int32)
        pamb = [0.02, 0.1, 0.4][it % 3]
        sec = bytes(np.where(rng.random(n) < pamb, rng.choice(list(b"RYSWKMN"), n), rng.choice(list(b"ACGT"), n)).astype(np.uint8))
        if n > 60 and it % 2:
            k = int(rng.integers(5, 30))
            sec = bytes(rng.choice(list(b"RYSWKMN"), k).astype(np.uint8)) + sec[k:-k] + bytes(rng.choice(list(b"RYSWKMN"), k).astype(np.uint8))
        yield dict(bcpos=bcpos, sec=sec, stringency=float([0, 0.5, 1, 2, 5][it % 5]), nsamples=int(bcpos[-1]) + int(rng.integers(-5, 20)),
                   tl=int(rng.integers(0, n // 2 + 1)), tr=int(rng.integers(0, n // 2 + 1)))


def reference_outputs(ref, c):
    q, best = ref.estimate_qualities(c["bcpos"], c["sec"], c["sec"])
    left, right = ref.trim_trace(c["bcpos"], c["sec"], c["stringency"])
    kept = ref.trim_basecalls(c["nsamples"], c["bcpos"], q, c["sec"], c["sec"], c["sec"], c["tl"], c["tr"])
    return dict(qual=[int(x) for x in q], best=best, left=left, right=right, kept_bcpos=[int(x) for x in kept[0]], kept_primary=kept[2].decode("latin-1"))


def snp_cases(seed, n_cases):
    """Seeded inputs of nearestSNP (src/trim.h:11-33): calls with no, few or many heterozygous positions, any trims, any start."""
    rng = np.random.default_rng(seed)
    for it in range(n_cases):
        n = int(rng.integers(1, 120))
        pri = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
        p = [0, 0.02, 0.2][it % 3]
        sec = bytes(np.where(rng.random(n) < p, rng.choice(list(b"ACGTRY"), n), np.frombuffer(pri, np.uint8)).astype(np.uint8))
        yield pri, sec, int(rng.integers(0, n // 2 + 2)), int(rng.integers(0, n // 2 + 2)), int(rng.integers(0, n))


if __name__ == "__main__":
    ref = loader.ref()
    assert ref is not None, "needs the reference build (oracle/_ref/libtracy_ref.so)"
    out = [reference_outputs(ref, c) for c in cases(31, 40)]
    with open(os.path.join(ROOT, "tests", "golden", "trim_golden.json"), "w") as f:
        json.dump(out, f)
    snp = [ref.nearest_snp(*c) for c in snp_cases(32, 300)]
    with open(os.path.join(ROOT, "tests", "golden", "nearest_snp_golden.json"), "w") as f:
        json.dump(snp, f)
    print("wrote trim_golden.json:", len(out), "cases; nearest_snp_golden.json:", len(snp), "cases")
