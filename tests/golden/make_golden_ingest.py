"""Generates tests/golden/ingest_golden.npz from the REFERENCE ITSELF (oracle/_ref: unmodified src/abif.h readab, src/scf.h
readscf / traceFormat) on synthetic ABIF / SCF files written by tracy_b200.synth.

    python tests/golden/make_golden_ingest.py        (build container only: needs /root/reference)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tracy_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def chans(rng, ns, amp):
    x = np.arange(ns)
    out = []
    for k in range(4):
        c = rng.integers(0, 25, ns).astype(np.int64)
        for p in range(6 + 3 * k, ns, 12):
            c += (amp * rng.random() * np.exp(-0.5 * ((x - p) / 2.0) ** 2)).astype(np.int64)
        out.append(c)
    return out


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(606)
    files = []
    for i, (ns, nb, order, amp) in enumerate([(1500, 110, b"GATC", 1500), (2400, 190, b"ACGT", 900), (900, 60, b"TCGA", 30000), (5000, 400, b"GATC", 2000)]):
        ploc = np.sort(rng.choice(np.arange(5, ns - 5), nb, replace=False))
        bases = bytes(rng.choice(list(b"ACGTNRYK"), nb + (3 if i == 1 else 0), p=[.23, .23, .23, .23, .02, .02, .02, .02]).astype(np.uint8))
        b2 = bytes(rng.choice(list(b"ACGT"), nb).astype(np.uint8)) if i in (1, 3) else None
        qual = rng.integers(0, 62, nb + (5 if i == 3 else 0))
        extra = [(b"SMPL", 1, 18, 1, b"\x06sample"), (b"LANE", 1, 4, 2, b"\x00\x07")] if i == 0 else []
        files.append(synth.abif_bytes(chans(rng, ns, amp), order, ploc, bases, qual, b2, extra))
    files.append(synth.abif_bytes(chans(rng, 600, 800), b"GATC", [], b"", []))                       # "File lacks basecalls!"
    for ns, nb, amp in [(1800, 140, 1200), (700, 50, 32000), (3000, 260, 20000)]:                    # SCF 3.x incl. int16 wrap-around
        files.append(synth.scf_bytes(chans(rng, ns, amp), np.sort(rng.choice(np.arange(5, ns - 5), nb, replace=False))))
    files.append(synth.scf_bytes(chans(rng, 400, 900), np.arange(30) * 12 + 6, version=b"2.00"))      # too old: readscf returns false
    files.append(b">seq\nACGTACGT\n")                                                                # not a trace
    files.append(b"AB")
    d = dict(n=np.int64(len(files)))
    tmp = tempfile.mktemp()
    for i, f in enumerate(files):
        open(tmp, "wb").write(f)
        r = ref.read_trace(tmp)
        ns = [len(x) for x in r["samples"]]
        print(i, r["format"], r["ok"], ns, len(r["basecallpos"]), len(r["basecalls1"]), len(r["basecalls2"]), len(r["qual"]))
        d[f"file{i}"] = np.frombuffer(f, np.uint8)
        d[f"meta{i}"] = np.array([r["format"], int(r["ok"]), len(r["basecallpos"])] + ns, np.int64)
        d[f"samples{i}"] = np.concatenate(r["samples"]) if sum(ns) else np.zeros(0, np.int32)
        d[f"ploc{i}"], d[f"qual{i}"] = r["basecallpos"], r["qual"]
        d[f"b1_{i}"], d[f"b2_{i}"] = np.frombuffer(r["basecalls1"], np.uint8), np.frombuffer(r["basecalls2"], np.uint8)
    os.remove(tmp)
    np.savez_compressed(os.path.join(OUT, "ingest_golden.npz"), **d)


if __name__ == "__main__":
    main()
