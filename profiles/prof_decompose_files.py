"""cProfile of subcommands.decompose over N synthetic jobs (host side of the files-in -> files-out pipeline of tracy decompose)."""
import cProfile, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracy_b200
from tracy_b200 import subcommands
from subcmd_cases import make_decompose_jobs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = tracy_b200.Context(0)
with tempfile.TemporaryDirectory() as d:
    jobs, _ = make_decompose_jobs(d, n=N, seed=9)
    jobs = [j for j in jobs if os.path.exists(j[0]) and j[1].endswith(".fa")]
    subcommands.decompose(ctx, jobs[:32], maxindel=30, chunk=32)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    rc = subcommands.decompose(ctx, jobs, maxindel=30, chunk=512, workers=8)
    pr.disable()
    dt = time.perf_counter() - t0
    print("jobs", len(jobs), "ok", rc.count(0), "seconds", round(dt, 3), "traces/s", round(len(jobs) / dt, 1))
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
