"""Golden for the de novo assemble path at N = 64 traces: the reference's own revSeqBasedOnDist -> msa ->
consensus (oracle/_ref: unmodified src/msa.h) on overlapping synthetic traces, half of them reverse-complemented.
Run here (needs /root/reference through oracle/_ref); writes tests/golden/assemble64_golden.npz."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader          # noqa: E402
from tracy_b200 import synth       # noqa: E402

SC = (3, -5, -10, -4)
N, L, STEP = 64, 220, 40


def main():
    ref = loader.ref()
    assert ref is not None, "oracle/_ref is only built where /root/reference exists"
    rng = np.random.default_rng(64)
    contig = synth.random_seq(rng, STEP * N + L)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    profs, planted = [], []
    for i in range(N):
        s = synth.mutate_seq(rng, contig[STEP * i: STEP * i + L + 8], 0.01, 0.004)[:L - int(rng.integers(0, 30))]
        rc = bool(rng.integers(0, 2))
        planted.append(not rc)
        profs.append(synth.profile_from_seq(rng, s.translate(comp)[::-1] if rc else s, 0.3))
    t0 = time.time()
    fwd = ref.rev_seq_based_on_dist(profs, [True] * N, SC)
    oriented = [p if f else ref.revcomp_profile(p) for p, f in zip(profs, fwd)]
    r = ref.msa(oriented, SC, 0.05)
    print(f"reference: {time.time() - t0:.1f} s; flips kept {sum(1 for f in fwd if not f)} (planted {sum(1 for f in planted if not f)}), "
          f"{r['rows'].shape[1]} columns, consensus {len(r['cons'])} bp")
    d = {"n": np.int64(N), "fwd": np.array(fwd, np.uint8), "rows": r["rows"], "seqidx": r["seqidx"].astype(np.int64), "dist": r["dist"].astype(np.int64),
         "gapped": np.frombuffer(r["gapped"], np.uint8), "cons": np.frombuffer(r["cons"], np.uint8), "qual": np.frombuffer(r["qual"], np.uint8)}
    for k, p in enumerate(profs):
        d[f"p{k}"] = p
    out = os.path.join(ROOT, "tests", "golden", "assemble64_golden.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
