"""Device-resident fill-only (gotohScore) against fill + traceback (gotoh) on the same pairs: what the traceback costs."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth
P, m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 53280, 1000, 4000
ctx = tracy_b200.Context(0)
dev = torch.device("cuda", 0)
prof, win = synth.align_batch(4096, m, n, seed=44)
idx = np.arange(P) % 4096
tp, tw = torch.from_numpy(prof).to(dev)[torch.from_numpy(idx).to(dev)].contiguous(), torch.from_numpy(win).to(dev)[torch.from_numpy(idx).to(dev)].contiguous()
aoff = (torch.arange(P, dtype=torch.int64) * 6 * m).to(dev); boff = (torch.arange(P, dtype=torch.int64) * n).to(dev)
alen = torch.full((P,), m, dtype=torch.int32, device=dev); blen = torch.full((P,), n, dtype=torch.int32, device=dev)
scores = torch.zeros(P, dtype=torch.int32, device=dev)
stride = 5008
ops = torch.zeros((P, stride), dtype=torch.uint8, device=dev); ol = torch.zeros(P, dtype=torch.int32, device=dev)
out = {"pairs": P}
for name, tb in (("score_only", False), ("traceback", True)):
    best = 1e9
    for _ in range(4):
        ctx.gotoh_device("ps", tp.data_ptr(), aoff.data_ptr(), alen.data_ptr(), tw.data_ptr(), boff.data_ptr(), blen.data_ptr(), P, scores.data_ptr(),
                         ops.data_ptr() if tb else None, stride if tb else 0, ol.data_ptr() if tb else None, DnaScore(3, -5, -10, -4), AlignConfig(True, False))
        best = min(best, ctx.last_call_ms())
    out[name] = {"ms": best, "gcups": P * m * n / best / 1e6}
print(json.dumps(out))
