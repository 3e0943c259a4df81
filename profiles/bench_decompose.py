"""BASELINE.json configs[2] end to end through drivers.decompose_batch (the DP sequence of `tracy decompose`): synthetic
heterozygous traces against ~1 kb single-FASTA references, maxindel 30. Host clock around the whole driver (Python glue
included), with the GPU kernels' own time beside it, and the reference composed from its own functions on a sample."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import tracy_b200
from tracy_b200 import DnaScore, drivers, synth
from make_golden_drivers import het_trace, reference_decompose

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(45)
base = []
for i in range(64):
    nref = int(rng.integers(900, 1100))
    refseq = synth.random_seq(rng, nref)
    start, L, bp = int(rng.integers(20, 120)), int(rng.integers(600, 750)), int(rng.integers(200, 500))
    ins, dl = [(0, int(rng.integers(1, 26))), (int(rng.integers(1, 26)), 0)][i % 2]
    tr, pos, pri, sec = het_trace(rng, refseq, start, L, bp, ins, dl, 1, rc=bool(i % 2))
    base.append((tr, pos, pri, sec, refseq))
sel = [base[i % 64] for i in range(N)]
ctx = tracy_b200.Context(0)
args = ([x[0] for x in sel], [x[1] for x in sel], [x[2] for x in sel], [x[3] for x in sel], [x[4] for x in sel], DnaScore(3, -5, -10, -4), 50, 50, 30, 5)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    drivers.decompose_batch(ctx, *[a[:64] if isinstance(a, list) else a for a in args])     # warm-up
    st0 = ctx.stats(); t0 = time.perf_counter()
    res = drivers.decompose_batch(ctx, *args)
    dt = time.perf_counter() - t0
out = {"traces": N, "seconds": dt, "traces_per_s": N / dt, "decomposed": sum(r is not None for r in res), "kernel_launches": ctx.stats()["kernel_launches"] - st0["kernel_launches"]}
from oracle import loader
ref = loader.ref()
if ref is not None:
    ns = 16
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        for x in base[:ns]:
            reference_decompose(ref, x[0], x[1], x[2], x[3], x[4], 50, 50, 30, 5)
        dr = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": ns / dr, "unit": "traces/s", "cores": 1, "kind": "reference", "sample": f"{ns} traces, indigo()'s DP sequence composed from the reference's functions"}
print(json.dumps(out))
