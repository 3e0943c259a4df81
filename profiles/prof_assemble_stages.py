"""Where drivers.assemble_denovo spends its time (BASELINE configs[3] shape): the stages of the de novo branch timed one by
one on the host clock, with the GPU calls each one makes."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import DnaScore, msa, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L, STEP = 900, 115
rng = np.random.default_rng(46)
contig = synth.random_seq(rng, STEP * N + L)
comp = bytes.maketrans(b"ACGT", b"TGCA")
profs = []
for i in range(N):
    s = bytearray(contig[STEP * i: STEP * i + L])
    for q in rng.integers(0, L, 9):
        s[q] = b"ACGT"[int(rng.integers(0, 4))]
    s = bytes(s)
    profs.append(synth.profile_from_seq(rng, s.translate(comp)[::-1] if i % 2 else s, 0.3))
ctx = tracy_b200.Context(0)
sc = DnaScore(3, -5, -10, -4)
msa.msa(ctx, [p.copy() for p in profs[:8]], sc)                 # warm-up
out = {"traces": N}


def stage(name, fn):
    l0 = ctx.stats()["kernel_launches"]; t0 = time.perf_counter()
    r = fn()
    out[name] = {"seconds": round(time.perf_counter() - t0, 3), "kernel_launches": ctx.stats()["kernel_launches"] - l0}
    return r


fwd = [True] * N
stage("revSeqBasedOnDist", lambda: msa.rev_seq_based_on_dist(ctx, profs, fwd, sc))
keep = stage("exclude_unmatched", lambda: msa.exclude_unmatched(ctx, profs, sc, 0.5))
kept = [profs[i] for i, k in enumerate(keep) if k]
d = stage("distance_matrix", lambda: msa.distance_matrix(ctx, kept, sc))
phylo, root = stage("upgma", lambda: msa.upgma(d, len(kept)))
rows, _, seqidx = stage("palign", lambda: msa.palign(ctx, kept, phylo, root, sc))
stage("consensus", lambda: msa.consensus(rows, 0.01, False))
out["msa_columns"] = int(rows.shape[1])
print(json.dumps(out))
