"""BASELINE.json configs[3]: `tracy assemble` de novo over N overlapping 900 bp traces tiling a random contig with step 115
(>= 85 % neighbour overlap), half of them reverse-complemented: drivers.assemble_denovo end to end (orientation by
revSeqBasedOnDist, exclusion loop, all-pairs distance matrix, UPGMA, progressive alignment, consensus). Host clock; the
reference's own revSeqBasedOnDist + msa + consensus (oracle/_ref) on a smaller N for scale."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import DnaScore, drivers, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NREF = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L, STEP = 900, 115
rng = np.random.default_rng(46)
contig = synth.random_seq(rng, STEP * N + L)
comp = bytes.maketrans(b"ACGT", b"TGCA")


def traces(n):
    out = []
    for i in range(n):
        s = bytearray(contig[STEP * i: STEP * i + L])
        for q in rng.integers(0, L, 9):
            s[q] = b"ACGT"[int(rng.integers(0, 4))]
        s = bytes(s)
        out.append(synth.profile_from_seq(rng, s.translate(comp)[::-1] if i % 2 else s, 0.3))
    return out


ctx = tracy_b200.Context(0)
profs = traces(N)
drivers.assemble_denovo(ctx, profs[:16])                       # warm-up
st0 = ctx.stats()
t0 = time.perf_counter()
res = drivers.assemble_denovo(ctx, profs, fraction_called=0.01)   # the default 0.1 calls nothing when 512 traces tile a contig 8 deep
dt = time.perf_counter() - t0
cons = res["consensus"]
want = contig[: STEP * (N - 1) + L]
ident = max(sum(a == b for a, b in zip(cons, w)) for w in (want, want.translate(comp)[::-1])) / max(len(want), 1)
out = {"traces": N, "seconds": dt, "kept": len(res["kept"]), "consensus_len": len(cons), "contig_len": len(want), "identity_to_planted_contig": ident,
       "kernel_launches": ctx.stats()["kernel_launches"] - st0["kernel_launches"]}
from oracle import loader
ref = loader.ref()
if ref is not None and NREF > 1:
    sub = traces(NREF)
    t0 = time.perf_counter()
    fwd = [True] * NREF
    ref.rev_seq_based_on_dist(sub, fwd, (3, -5, -10, -4))
    dr = time.perf_counter() - t0
    out["cpu_baseline"] = {"traces": NREF, "seconds_revSeqBasedOnDist_only": dr, "kind": "reference", "cores": 1,
                           "note": "the orientation stage alone; its cost grows with N^2 per sweep"}
print(json.dumps(out))
