import cProfile, pstats, sys, io, contextlib, runpy
sys.argv = ["bench_decompose.py", "1000"]
pr = cProfile.Profile()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    try:
        runpy.run_path("profiles/bench_decompose.py", run_name="__main__")
    except SystemExit:
        pass
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
