"""cProfile of subcommands.align over N synthetic jobs (host side of the files-in -> files-out pipeline)."""
import cProfile, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracy_b200
from tracy_b200 import subcommands
from subcmd_cases import make_align_jobs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
WORKERS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = tracy_b200.Context(0)
with tempfile.TemporaryDirectory() as d:
    jobs, _ = make_align_jobs(d, n=N, seed=7)
    jobs = [j for j in jobs if os.path.exists(j[0]) and j[1].endswith(".fa") and "big" not in j[1] and "multi" not in j[1]][:N]
    subcommands.align(ctx, jobs[:32], chunk=32)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    rc = subcommands.align(ctx, jobs, chunk=256, workers=WORKERS)
    pr.disable()
    dt = time.perf_counter() - t0
    print("jobs", len(jobs), "ok", rc.count(0), "seconds", round(dt, 3), "traces/s", round(len(jobs) / dt, 1))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
