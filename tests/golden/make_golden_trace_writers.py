"""Golden outputs of the reference's per-trace writers (traceTxtOut src/abif.h:512-534, traceJsonOut src/json.h:108-117,
alignmentTracePadding + traceAlignJsonOut src/json.h:383-479, 197-217) through oracle/ref_bridge.cpp, for tests/test_writers.py.
Run in the build container (needs /root/reference): python tests/golden/make_golden_trace_writers.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402


def cases(seed, n, max_samples):
    """Seeded inputs shared by the generator and the differential test: traces, basecalls (some with repeated or out-of-range
    positions for the two plain writers), an alignment row 0 that holds the basecalls with random gap runs, a row 1."""
    rng = np.random.default_rng(seed)
    for it in range(n):
        ns = int(rng.integers(20, max_samples))
        nb = int(rng.integers(1, max(2, ns // 8)))
        acgt = rng.integers(0, 3000, (4, ns)).astype(np.int32)
        bcpos = np.sort(rng.choice(ns, nb, replace=False)).astype(np.int32)
        wellformed = not (it % 7 == 0 or it % 11 == 0)
        if it % 7 == 0 and nb > 2:
            bcpos[nb // 2] = bcpos[nb // 2 - 1]
        if it % 11 == 0:
            bcpos[-1] = ns + 3
        qual = rng.integers(0, 61, nb).astype(np.uint8)
        pri = bytes(rng.choice(list(b"ACGTN"), nb).astype(np.uint8))
        sec = bytes(np.where(rng.random(nb) < 0.7, np.frombuffer(pri, np.uint8), rng.choice(list(b"ACGTRYSWKMNX"), nb)).astype(np.uint8))
        row0 = bytearray()
        for ch in pri:
            if rng.random() < 0.08:
                row0 += b"-" * int(rng.integers(1, 4))
            row0.append(ch)
        if rng.random() < 0.5:
            row0 = bytearray(b"-" * int(rng.integers(1, 5))) + row0
        if rng.random() < 0.5:
            row0 += b"-" * int(rng.integers(1, 5))
        row1 = bytes(rng.choice(list(b"ACGT-"), len(row0)).astype(np.uint8))
        yield dict(acgt=acgt, bcpos=bcpos, qual=qual, pri=pri, sec=sec, tl=int(rng.integers(0, nb + 3)), tr=int(rng.integers(0, nb + 3)), row0=bytes(row0), row1=row1,
                   pos=int(rng.integers(0, 10 ** 6)), fwd=bool(it % 2), wellformed=wellformed)


def reference_outputs(ref, c):
    txt = ref.trace_outputs("txt", c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"], c["tl"], c["tr"])
    js = ref.trace_outputs("json", c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"])
    aj = b""
    if c["wellformed"]:      # the padding indexes bcPos by the number of bases seen so far: only defined for increasing, in-range calls
        aj = ref.trace_outputs("align_json", c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"], row0=c["row0"], row1=c["row1"], chr_name=b"chr7",
                               pos=c["pos"], forward=c["fwd"])
    return txt, js, aj


if __name__ == "__main__":
    ref = loader.ref()
    assert ref is not None, "needs the reference build (oracle/_ref/libtracy_ref.so)"
    out = {}
    n = 24
    for i, c in enumerate(cases(77, n, 160)):
        txt, js, aj = reference_outputs(ref, c)
        out[f"txt{i}"] = np.frombuffer(txt, np.uint8)
        out[f"json{i}"] = np.frombuffer(js, np.uint8)
        out[f"ajson{i}"] = np.frombuffer(aj, np.uint8)
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "trace_writers_golden.npz"), **out)
    print("wrote trace_writers_golden.npz,", n, "cases")
