"""Generates tests/golden/consensus_golden.npz: the DP sequence of consensus() (reference src/consensus.h:499-556) composed
from the REFERENCE's own functions (oracle/_ref): createProfile, reverseComplementProfile, gotohScore, gotoh, and
pairwiseConsensus (:189-238) over each alignment.

    python tests/golden/make_golden_consensus.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from make_golden_align_genome import COMP, sanger  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SC = (3, -5, -10, -4)


def main():
    ref = loader.ref()
    assert ref is not None
    rng = np.random.default_rng(90210)
    genome = bytes(rng.choice(list(b"ACGT"), 3000).astype(np.uint8))
    d = dict(n=np.int64(8))
    for i in range(8):
        p1, L1 = int(rng.integers(0, 1000)), int(rng.integers(300, 600))
        p2, L2 = p1 + int(rng.integers(50, 250)), int(rng.integers(300, 600))
        s1, s2 = genome[p1:p1 + L1], bytearray(genome[p2:p2 + L2])
        for q in rng.integers(0, L2, 4):
            s2[q] = b"ACGT"[int(rng.integers(0, 4))]
        s2 = bytes(s2)
        if i % 2 == 0:
            s2 = s2.translate(COMP)[::-1]
        if i == 7:
            s2 = bytes(rng.choice(list(b"ACGT"), L2).astype(np.uint8))
        profs = []
        for s, (tl, trr) in ((s1, (20, 30)), (s2, (25, 15))):
            tr, pos = sanger(rng, s)
            bc = ref.basecall(tr, pos, 0.33)
            profs.append(ref.create_profile(tr, bc["bcPos"], bc["primary"], bc["secondary"], tl, trr))
        t1, f2 = profs
        r2 = ref.revcomp_profile(f2)
        gf, gr = ref.gotoh_score(t1, f2, 1, 1, SC), ref.gotoh_score(t1, r2, 1, 1, SC)
        fw = gf > gr
        score, r0, r1 = ref.gotoh(t1, f2 if fw else r2, 1, 1, SC)
        print(i, fw, gf, gr, score, len(r0))
        d[f"p1_{i}"], d[f"p2_{i}"] = t1, f2
        d[f"meta{i}"] = np.array([fw, score], np.int64)
        d[f"r0_{i}"], d[f"r1_{i}"] = np.frombuffer(r0, np.uint8), np.frombuffer(r1, np.uint8)
        # pairwiseConsensus (src/consensus.h:189-238) over that alignment, union / intersection x plain / IUPAC letters
        for un in (0, 1):
            for iu in (0, 1):
                cons, qual = ref.pairwise_consensus(r0, r1, t1, f2 if fw else r2, un, iu)
                d[f"cons{un}{iu}_{i}"], d[f"qual{un}{iu}_{i}"] = np.frombuffer(cons, np.uint8), qual
    np.savez_compressed(os.path.join(OUT, "consensus_golden.npz"), **d)


if __name__ == "__main__":
    main()
