"""`tracy assemble` files in -> files out for ONE job of N trace files tiling a contig (BASELINE.json configs[3] as a subcommand):
wall clock of subcommands.assemble, split into the call up to the last GPU stage and the writers."""
import cProfile, json, os, pstats, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracy_b200
from tracy_b200 import subcommands, synth
from subcmd_cases import sanger_file, COMP
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(46)
L, STEP = 900, 115
contig = synth.random_seq(rng, STEP * N + L)
ctx = tracy_b200.Context(0)
with tempfile.TemporaryDirectory() as d:
    paths = []
    for i in range(N):
        s = synth.mutate_seq(rng, contig[STEP * i: STEP * i + L], 0.005, 0.001)
        if i % 2:
            s = s.translate(COMP)[::-1]
        p = os.path.join(d, f"t{i}.ab1")
        open(p, "wb").write(sanger_file(rng, s, het=0.01))
        paths.append(p)
    subcommands.assemble(ctx, [(paths[:8], None, os.path.join(d, "warm"))], fraction_called=0.01)
    pr = cProfile.Profile()
    t0 = time.perf_counter(); pr.enable()
    rc = subcommands.assemble(ctx, [(paths, None, os.path.join(d, "out"))], fraction_called=0.01, workers=8)
    pr.disable(); dt = time.perf_counter() - t0
    sizes = {s: os.path.getsize(os.path.join(d, "out" + s)) for s in (".align.fa", ".json", ".vertical", ".cons.fa") if os.path.exists(os.path.join(d, "out" + s))}
    print(json.dumps({"traces": N, "rc": rc, "seconds": dt, "output_bytes": sizes}))
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
