"""palign of the configs[3] shape level by level: pairs, largest merge, GPU call wall time, kernel time, host merge time."""
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
profs = []
for i in range(N):
    s = bytearray(contig[STEP * i: STEP * i + L])
    for q in rng.integers(0, L, 9):
        s[q] = b"ACGT"[int(rng.integers(0, 4))]
    profs.append(synth.profile_from_seq(rng, bytes(s), 0.3))
ctx = tracy_b200.Context(0)
sc = DnaScore(3, -5, -10, -4)
msa.msa(ctx, [p.copy() for p in profs[:8]], sc)
d = msa.distance_matrix(ctx, profs, sc)
phylo, root = msa.upgma(d, N)
levels = []
real = ctx.gotoh


def timed(kind, A, B, *a, **k):
    t0 = time.perf_counter()
    r = real(kind, A, B, *a, **k)
    dt = time.perf_counter() - t0
    km = ctx.last_kernel_ms()
    levels.append({"pairs": len(A), "max_m": max(x.shape[1] for x in A), "max_n": max(x.shape[1] for x in B), "call_ms": round(dt * 1e3, 2),
                   "kernel_ms": round(km["packed_ms"] + km["general_ms"], 2), "big_pairs": int(ctx.last_big_pairs())})
    return r


ctx.gotoh = timed
import cProfile, pstats
pr = cProfile.Profile() if os.environ.get("PROFILE") else None
t0 = time.perf_counter()
if pr: pr.enable()
rows, _, _ = msa.palign(ctx, profs, phylo, root, sc)
if pr: pr.disable()
total = time.perf_counter() - t0
if pr: pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(14)
print(json.dumps({"traces": N, "palign_seconds": round(total, 3), "gpu_call_seconds": round(sum(l["call_ms"] for l in levels) / 1e3, 3),
                  "kernel_seconds": round(sum(l["kernel_ms"] for l in levels) / 1e3, 3), "columns": int(rows.shape[1]), "levels": levels}, indent=1))
