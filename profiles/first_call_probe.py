"""Where does the first large traceback call of a context spend its time? (palign's first level: 437 ms around a 1 ms kernel)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth

rng = np.random.default_rng(1)
ctx = tracy_b200.Context(0)
sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, True)
profs = [synth.profile_from_seq(rng, synth.random_seq(rng, 900), 0.3) for _ in range(400)]
out = []


def call(tag, n, tb):
    t0 = time.perf_counter()
    ctx.gotoh("pp", profs[:n], profs[n:2 * n], sc, ac, traceback=tb)
    out.append({"call": tag, "pairs": n, "traceback": tb, "ms": round((time.perf_counter() - t0) * 1e3, 2)})


call("warm-up, 4 pairs, traceback", 4, True)
call("4 pairs again", 4, True)
call("200 pairs, score only", 200, False)
call("200 pairs, traceback (first large)", 200, True)
call("200 pairs, traceback (second)", 200, True)
call("100 pairs, traceback", 100, True)
call("200 pairs, traceback (third)", 200, True)
from tracy_b200 import msa
t0 = time.perf_counter()
msa.distance_matrix(ctx, profs[:320], sc)
out.append({"call": "distance_matrix of 320 profiles (51 040 score fills)", "ms": round((time.perf_counter() - t0) * 1e3, 2)})
call("200 pairs, traceback, after the big score-only call", 200, True)
call("200 pairs, traceback, again", 200, True)
t0 = time.perf_counter()
msa.distance_matrix(ctx, profs[:400], sc)
out.append({"call": "distance_matrix of 400 profiles", "ms": round((time.perf_counter() - t0) * 1e3, 2)})
call("200 pairs, traceback, after the second big call", 200, True)
print(json.dumps(out, indent=1))
