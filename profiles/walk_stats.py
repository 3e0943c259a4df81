"""Traceback-round statistics of the packed kernel (profiling build with -DTB_WALK_STATS, loaded through TRACY_B200_LIB):
rounds per pair, clocks per phase. Same pairs as tb_cost_probe.py."""
import ctypes, json, os, sys
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
lib = ctypes.CDLL(os.environ["TRACY_B200_LIB"])
buf = (ctypes.c_ulonglong * 16)()
def run():
    ctx.gotoh_device("ps", tp.data_ptr(), aoff.data_ptr(), alen.data_ptr(), tw.data_ptr(), boff.data_ptr(), blen.data_ptr(), P, scores.data_ptr(),
                     ops.data_ptr(), stride, ol.data_ptr(), DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    return ctx.last_call_ms()
run(); lib.tb_debug_walk_stats(buf, 1)
ms = run(); lib.tb_debug_walk_stats(buf, 1)
v = list(buf)
names = ["pairs", "rounds", "horizontal_rounds", "rounds_64_iterations", "clk_plan_edges", "clk_tile", "clk_walk", "clk_total", "walk_iterations", "window_loads"]
out = {"ms": ms, "raw": dict(zip(names, v))}
pairs = max(v[0], 1)
out["per_pair"] = {k: x / pairs for k, x in zip(names[1:], v[1:10])}
out["per_round_clk"] = {k: x / max(v[1], 1) for k, x in zip(names[4:7], v[4:7])}
# rounds split by pair kind: ops length distribution as a proxy
L = ol.cpu().numpy(); sc = scores.cpu().numpy()
out["score_negative_fraction"] = float((sc < 0).mean())
print(json.dumps(out))
