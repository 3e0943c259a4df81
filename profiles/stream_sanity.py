"""A streamed host batch of tiny pairs (16+ waves) next to the launch-per-chunk form: same outputs. Small enough to run under compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth

N, m, n = 45000, 48, 96
base_p, base_w = synth.align_batch(512, m, n, seed=5)
idx = (np.arange(N) * 7 + 1) % 512
prof, win = np.ascontiguousarray(base_p[idx]), np.ascontiguousarray(base_w[idx])
win[::101, 10] = ord("R")
ctx = tracy_b200.Context(0)
a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
os.environ["TRACY_B200_TRACE"] = "1"
os.environ["TRACY_B200_FORCE_STREAM"] = "1"
s0, o0, l0, r0, r1 = ctx.gotoh("ps", a1, a2, sc, ac, rows=True)
os.environ["TRACY_B200_NO_STREAM"] = "1"
s1, o1, l1, q0, q1 = ctx.gotoh("ps", a1, a2, sc, ac, rows=True)
mask = np.arange(o0.shape[1])[None, :] < l0[:, None]
ok = np.array_equal(s0, s1) and np.array_equal(l0, l1) and np.array_equal(o0 * mask, o1 * mask) and np.array_equal(r0 * mask, q0 * mask) and np.array_equal(r1 * mask, q1 * mask)
print("streamed == chunked:", ok, "pairs", N, "packed", ctx.last_packed_pairs())
sys.exit(0 if ok else 1)
