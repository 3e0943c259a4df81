"""BASELINE.json configs[4]-style run from ONE process through the C ABI's multi-device handle (tb_multi / MultiContext): P pairs of
1000 x 4000 per device in pinned host buffers, one call spreads them over every visible GPU; score + 2-bit packed traceback back.
Prints one JSON object with the GCUPS of 1 .. N devices (host clock around the call: copies inside)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth
per_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ndev = torch.cuda.device_count()
m, n = 1000, 4000
P = per_dev * ndev
base_p, base_w = synth.align_batch(4096, m, n, seed=44)
h_prof = torch.empty((P, 6, m), dtype=torch.float32, pin_memory=True); h_win = torch.empty((P, n), dtype=torch.uint8, pin_memory=True)
idx = np.arange(P) % 4096
for s in range(0, P, 4096):
    h_prof.numpy()[s: s + 4096] = base_p[idx[s: s + 4096]]; h_win.numpy()[s: s + 4096] = base_w[idx[s: s + 4096]]
h_scores = torch.empty(P, dtype=torch.int32, pin_memory=True); h_len = torch.empty(P, dtype=torch.int32, pin_memory=True)
pstride = ((m + n + 3) // 4 + 15) // 16 * 16
h_pk = torch.empty((P, pstride), dtype=torch.uint8, pin_memory=True)
sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
out = {"pairs_per_device": per_dev, "visible_devices": ndev, "runs": []}
ref_scores = None
for k in sorted({1, 2, 4, ndev} & set(range(1, ndev + 1))):
    Pk = per_dev * k
    a1 = tracy_b200.uniform_profiles(h_prof.numpy()[:Pk], trace_profiles=True); a2 = tracy_b200.uniform_seqs(h_win.numpy()[:Pk])
    with tracy_b200.MultiContext(list(range(k))) as mc:
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            mc.gotoh("ps", a1, a2, sc, ac, traceback=True, out=(h_scores.numpy()[:Pk], h_pk.numpy()[:Pk], h_len.numpy()[:Pk]), packed=True)
            best = min(best, time.perf_counter() - t0)
        if ref_scores is None:
            ref_scores = h_scores.numpy()[:per_dev].copy()
        assert np.array_equal(h_scores.numpy()[:per_dev], ref_scores)                       # the first device's range gives the same scores at every k
        out["runs"].append({"devices": k, "pairs": Pk, "ms": best * 1e3, "gcups": Pk * m * n / best / 1e9, "ranges": mc.last_ranges})
print(json.dumps(out))
