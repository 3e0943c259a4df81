"""Golden outputs of the two reference functions inside assemble()'s output section -- alignedTraceByRow (src/json.h:220-246) and
reverseComplementTrace (src/trim.h:102-151) -- through oracle/ref_bridge.cpp, for tests/test_writers.py.
Run in the build container: python tests/golden/make_golden_assemble_writers.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402


def cases(seed, n_cases):
    rng = np.random.default_rng(seed)
    for it in range(n_cases):
        n = int(rng.integers(1, 60))
        ns = n * 12 + int(rng.integers(5, 40))
        acgt = rng.integers(0, 3000, (4, ns)).astype(np.int32)
        bcpos = np.sort(rng.choice(ns, n, replace=False)).astype(np.int32)
        if it % 9 == 0 and n > 2:
            bcpos[n // 2] = bcpos[n // 2 - 1]
        qual = rng.integers(0, 61, n).astype(np.uint8)
        pri = bytes(rng.choice(list(b"ACGTNRYKMSWBDHVUXn-"), n).astype(np.uint8))
        sec = bytes(rng.choice(list(b"ACGTNRYKMSW"), n).astype(np.uint8))
        nrow, ncol = int(rng.integers(1, 6)), int(rng.integers(1, 50))
        rows = rng.choice(list(b"ACGT-"), (nrow, ncol)).astype(np.uint8)
        r = int(rng.integers(0, nrow))
        if it % 10 == 0:
            rows[r, :] = 45
        if it % 3 == 0:
            rows[r, :int(rng.integers(0, ncol))] = 45
        if it % 4 == 0:
            rows[r, ncol - int(rng.integers(0, ncol)):] = 45
        yield dict(acgt=acgt, bcpos=bcpos, qual=qual, pri=pri, sec=sec, rows=rows, row=r, name="trace_%d" % it, fwd=bool(it % 2), isref=bool(it % 3 == 0))


def reference_outputs(ref, c):
    w = ref.reverse_complement_trace(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"])
    return dict(acgt_sum=[int(x) for x in (w[0].astype(np.int64) * np.arange(1, w[0].shape[1] + 1)).sum(axis=1)], bcpos=[int(x) for x in w[1]], qual=[int(x) for x in w[2]],
                primary=w[3].decode("latin-1"), secondary=w[4].decode("latin-1"),
                byrow=ref.aligned_trace_by_row(c["rows"], c["row"], c["name"], c["fwd"], c["isref"]).decode("latin-1"))


if __name__ == "__main__":
    ref = loader.ref()
    assert ref is not None, "needs the reference build (oracle/_ref/libtracy_ref.so)"
    out = [reference_outputs(ref, c) for c in cases(41, 30)]
    with open(os.path.join(ROOT, "tests", "golden", "assemble_writers_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote assemble_writers_golden.json:", len(out), "cases")
