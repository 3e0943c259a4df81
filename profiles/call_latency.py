"""Host-call latency of small batches (what single-file / CLI-style use and the sequential stages of assemble see)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth
ctx = tracy_b200.Context(0)
rng = np.random.default_rng(1)
out = {}
for name, kind, n, m, k, tb in (("ps_1pair_800x2000_tb", "ps", 1, 800, 2000, True), ("ps_1pair_800x2000_score", "ps", 1, 800, 2000, False),
                                ("pp_511pairs_900x900_score", "pp", 511, 900, 900, False), ("ss_2pairs_1000x4000_tb", "ss", 2, 1000, 4000, True)):
    if kind == "ps":
        a, b = synth.align_batch(n, m, k, seed=3)
        A, B = tracy_b200.uniform_profiles(a), tracy_b200.uniform_seqs(b)
    elif kind == "pp":
        ps = [synth.profile_from_seq(rng, synth.random_seq(rng, m), 0.3) for _ in range(8)]
        A = tracy_b200.pack_profiles([ps[i % 8] for i in range(n)]); B = tracy_b200.pack_profiles([ps[(i + 3) % 8] for i in range(n)])
    else:
        A = tracy_b200.pack_seqs([synth.random_seq(rng, m) for _ in range(n)]); B = tracy_b200.pack_seqs([synth.random_seq(rng, k) for _ in range(n)])
    ac = AlignConfig(True, kind == "pp")
    for _ in range(3):
        ctx.gotoh(kind, A, B, DnaScore(3, -5, -10, -4), ac, traceback=tb)
    t0 = time.perf_counter()
    R = 50
    for _ in range(R):
        ctx.gotoh(kind, A, B, DnaScore(3, -5, -10, -4), ac, traceback=tb)
    dt = (time.perf_counter() - t0) / R
    km = ctx.last_kernel_ms()
    out[name] = {"ms_per_call": dt * 1e3, "kernel_ms": km["packed_ms"] + km["general_ms"]}
print(json.dumps(out, indent=1))
