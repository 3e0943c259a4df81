"""Host-buffer call (what bench.py's e2e leg times) against the chunk size of the TB_MEM_HOST pipeline: 100 000 pairs of 1000 x 4000
from pinned buffers, score + 2-bit packed traceback back, wall clock per call. TRACY_B200_CHUNK fixes the chunk size (no ramps);
unset = the library's own schedule (ramp up, steady chunks in whole waves, ramp down)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth
P, m, n = 100000, 1000, 4000
ctx = tracy_b200.Context(0)
base_p, base_w = synth.align_batch(4096, m, n, seed=44)
h_prof = torch.empty((P, 6, m), dtype=torch.float32, pin_memory=True)
h_win = torch.empty((P, n), dtype=torch.uint8, pin_memory=True)
idx = np.arange(P) % 4096
for s in range(0, P, 4096):
    h_prof.numpy()[s: s + 4096] = base_p[idx[s: s + 4096]]
    h_win.numpy()[s: s + 4096] = base_w[idx[s: s + 4096]]
h_scores = torch.empty(P, dtype=torch.int32, pin_memory=True); h_len = torch.empty(P, dtype=torch.int32, pin_memory=True)
pstride = ((m + n + 3) // 4 + 15) // 16 * 16
h_pk = torch.empty((P, pstride), dtype=torch.uint8, pin_memory=True)
a1 = tracy_b200.uniform_profiles(h_prof.numpy(), trace_profiles=True); a2 = tracy_b200.uniform_seqs(h_win.numpy())
sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
out = {}
wave = 148 * 12
for tag, env in (("library schedule", {}), ("no ramp down", {"TRACY_B200_NO_RAMPDOWN": "1"}),
                 ("schedule, steady chunks of 10 waves", {"TRACY_B200_CHUNK_TARGET": "20000", "TRACY_B200_CHUNK_PARTS": "5"}),
                 ("schedule, steady chunks of 14 waves", {"TRACY_B200_CHUNK_TARGET": "32768", "TRACY_B200_CHUNK_PARTS": "4"}),
                 ("schedule, steady chunks of 21 waves", {"TRACY_B200_CHUNK_TARGET": "40000", "TRACY_B200_CHUNK_PARTS": "2"}),
                 ("schedule, steady chunks of 4 waves", {"TRACY_B200_CHUNK_TARGET": "7104", "TRACY_B200_CHUNK_PARTS": "16"}), ("3 waves", {"TRACY_B200_CHUNK": str(3 * wave)}), ("7 waves", {"TRACY_B200_CHUNK": str(7 * wave)}),
                 ("14 waves", {"TRACY_B200_CHUNK": str(14 * wave)}), ("28 waves", {"TRACY_B200_CHUNK": str(28 * wave)}), ("one chunk", {"TRACY_B200_CHUNK": str(P)})):
    for k, v in env.items():
        os.environ[k] = v
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.gotoh("ps", a1, a2, sc, ac, traceback=True, out=(h_scores.numpy(), h_pk.numpy(), h_len.numpy()), packed=True)
        torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
    k = ctx.last_kernel_ms()
    out[tag] = {"wall_ms": round(best, 2), "kernel_ms_sum": round(k["packed_ms"] + k["general_ms"], 2)}
    for k in env:
        del os.environ[k]
print(json.dumps(out, indent=1))
