"""PCIe / pipeline probe for the TB_MEM_HOST path (run on the GPU box): raw pinned copy bandwidth, then e2e vs chunk size."""
import os, sys, time, json, subprocess
import torch
dev = torch.device("cuda", 0)
h = torch.empty(2_800_000_000, dtype=torch.uint8, pin_memory=True); h.fill_(1)
d = torch.empty_like(h, device=dev)
ho = torch.empty(500_000_000, dtype=torch.uint8, pin_memory=True)
do = torch.empty_like(ho, device=dev)
def t(fn, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
d.copy_(h, non_blocking=True); torch.cuda.synchronize()
s = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 2.8GB one copy: %.1f ms %.1f GB/s" % (s * 1e3, 2.8 / s))
def chunks():
    for i in range(8): d[i * 350_000_000:(i + 1) * 350_000_000].copy_(h[i * 350_000_000:(i + 1) * 350_000_000], non_blocking=True)
s = t(chunks); print("H2D 8x350MB: %.1f ms %.1f GB/s" % (s * 1e3, 2.8 / s))
s = t(lambda: ho.copy_(do, non_blocking=True)); print("D2H 0.5GB: %.1f ms %.1f GB/s" % (s * 1e3, 0.5 / s))
s2 = torch.cuda.Stream()
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
s = t(both); print("H2D 2.8GB + D2H 0.5GB concurrent: %.1f ms" % (s * 1e3))
del h, d, ho, do
for ch in sys.argv[1:]:
    env = dict(os.environ, TRACY_B200_CHUNK=ch)
    out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--steps", "3", "--warmup", "2"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    j = json.loads(out); print(ch, round(j["value"]), round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"], 1), flush=True)
