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
t_all = time.perf_counter()
T = stage("orientation_table", lambda: msa.orientation_table(ctx, profs, sc))
out["orientation_table"]["pairs"] = 4 * N * (N - 1)
out["orientation_table"]["kernel_ms"] = sum(ctx.last_kernel_ms()[k] for k in ("packed_ms", "general_ms"))
d, T, o = stage("revSeqBasedOnDist_replay", lambda: msa.rev_seq_based_on_dist(ctx, profs, fwd, sc, table=T, with_state=True))
keep = stage("exclude_unmatched", lambda: msa.exclude_unmatched(ctx, profs, sc, 0.5, dist=d))
idx = [i for i, k in enumerate(keep) if k]
kept = [profs[i] for i in idx]
dm = stage("distance_matrix_from_table", lambda: msa.oriented_distance(T, o, idx))
phylo, root = stage("upgma", lambda: msa.upgma(dm, len(kept)))
rows, _, seqidx = stage("palign", lambda: msa.palign(ctx, kept, phylo, root, sc))
stage("consensus", lambda: msa.consensus(rows, 0.01, False))
out["total_seconds"] = round(time.perf_counter() - t_all, 3)
out["total_kernel_launches"] = sum(v["kernel_launches"] for v in out.values() if isinstance(v, dict))
out["msa_columns"] = int(rows.shape[1])
out["traces_kept"] = len(kept)
out["flipped"] = int(sum(1 for f in fwd if not f))
print(json.dumps(out))
