"""Does PCIe traffic slow the packed kernel? The device-resident call (100 000 pairs of 1000 x 4000, checkpoint traceback) alone,
and with host<->device copies of the e2e leg's size running next to it on other streams (torch, pinned buffers)."""
import json, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth
P, m, n = 100000, 1000, 4000
ctx = tracy_b200.Context(0)
dev = torch.device("cuda", 0)
prof, win = synth.align_batch(4096, m, n, seed=44)
idx = torch.from_numpy(np.arange(P) % 4096).to(dev)
tp, tw = torch.from_numpy(prof).to(dev)[idx].contiguous(), torch.from_numpy(win).to(dev)[idx].contiguous()
aoff = (torch.arange(P, dtype=torch.int64) * 6 * m).to(dev); boff = (torch.arange(P, dtype=torch.int64) * n).to(dev)
alen = torch.full((P,), m, dtype=torch.int32, device=dev); blen = torch.full((P,), n, dtype=torch.int32, device=dev)
scores = torch.zeros(P, dtype=torch.int32, device=dev)
stride = 5008
ops = torch.zeros((P, stride), dtype=torch.uint8, device=dev); ol = torch.zeros(P, dtype=torch.int32, device=dev)
def step():
    ctx.gotoh_device("ps", tp.data_ptr(), aoff.data_ptr(), alen.data_ptr(), tw.data_ptr(), boff.data_ptr(), blen.data_ptr(), P, scores.data_ptr(),
                     ops.data_ptr(), stride, ol.data_ptr(), DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    return ctx.last_call_ms()
for _ in range(2): step()
out = {"alone_ms": min(step() for _ in range(3))}
hbuf = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True); dbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hbuf2 = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True); dbuf2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for tag, do_in, do_out in (("with_h2d", True, False), ("with_d2h", False, True), ("with_both", True, True)):
    stop = False
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    moved = [0]
    def pump():
        while not stop:
            if do_in:
                with torch.cuda.stream(s1): dbuf.copy_(hbuf, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2): hbuf2.copy_(dbuf2, non_blocking=True)
            s1.synchronize(); s2.synchronize(); moved[0] += (256 << 20) * (int(do_in) + int(do_out))
    th = threading.Thread(target=pump); th.start()
    time.sleep(0.05)
    m0 = moved[0]; t0 = time.perf_counter()
    ms = min(step() for _ in range(3))
    dt = time.perf_counter() - t0
    out[tag + "_ms"] = ms; out[tag + "_GBps"] = (moved[0] - m0) / dt / 1e9
    stop = True; th.join()
print(json.dumps(out))
