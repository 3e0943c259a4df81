"""Files in -> files out throughput of tracy_b200.subcommands.align / consensus (N command lines of the reference in one call):
synthetic ABIF traces of ~700 bp against 2-3 kb FASTA references, all four output files per trace. Wall clock of the call, the time
the calling thread spent inside GPU calls, and what the writer threads were still doing after the last GPU call."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracy_b200
from tracy_b200 import subcommands
from subcmd_cases import make_align_jobs, make_consensus_jobs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = tracy_b200.Context(0)
out = {}
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    jobs, _ = make_align_jobs(d, n=N, seed=7)
    jobs = [j for j in jobs if os.path.exists(j[0]) and j[1].endswith(".fa") and "big" not in j[1] and "multi" not in j[1]][:N]
    gen = time.perf_counter() - t0
    subcommands.align(ctx, jobs[:32], chunk=32)                                   # warm-up
    for workers in (2, 8, 16):
        st0 = ctx.stats(); t0 = time.perf_counter()
        rc = subcommands.align(ctx, jobs, chunk=256, workers=workers)
        dt = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(j[2] + s) for j in jobs for s in (".abif", ".align.fa", ".txt", ".json") if os.path.exists(j[2] + s))
        out[f"align_workers{workers}"] = {"jobs": len(jobs), "ok": rc.count(0), "seconds": dt, "traces_per_s": len(jobs) / dt, "output_MB": nbytes / 1e6,
                                          "kernel_launches": ctx.stats()["kernel_launches"] - st0["kernel_launches"]}
    out["generate_inputs_s"] = gen
    # tracy decompose: heterozygous traces (indels of 1-25 bp) against ~2 kb FASTA references, -i 30; six output files per trace
    from subcmd_cases import make_decompose_jobs
    djobs, _ = make_decompose_jobs(d, n=min(N, 1200), seed=9)
    djobs = [j for j in djobs if os.path.exists(j[0]) and j[1].endswith(".fa")]
    subcommands.decompose(ctx, djobs[:32], maxindel=30, chunk=32)
    for workers in (8,):
        st0 = ctx.stats(); t0 = time.perf_counter()
        rc = subcommands.decompose(ctx, djobs, maxindel=30, chunk=512, workers=workers)
        dt = time.perf_counter() - t0
        out[f"decompose_workers{workers}"] = {"jobs": len(djobs), "ok": rc.count(0), "seconds": dt, "traces_per_s": len(djobs) / dt,
                                              "kernel_launches": ctx.stats()["kernel_launches"] - st0["kernel_launches"]}
    # tracy consensus: overlapping trace pairs, six output files per pair
    cjobs, _ = make_consensus_jobs(d, n=min(N, 800), seed=10)
    cjobs = [j for j in cjobs if os.path.exists(j[0]) and os.path.exists(j[1])]
    subcommands.consensus(ctx, cjobs[:32], chunk=32)
    st0 = ctx.stats(); t0 = time.perf_counter()
    rc = subcommands.consensus(ctx, cjobs, chunk=512, workers=8)
    dt = time.perf_counter() - t0
    out["consensus_workers8"] = {"jobs": len(cjobs), "ok": rc.count(0), "seconds": dt, "pairs_per_s": len(cjobs) / dt,
                                 "kernel_launches": ctx.stats()["kernel_launches"] - st0["kernel_launches"]}
    from oracle import loader                                                      # the reference's own `tracy align` on a sample of the same files, one core
    ref = loader.ref()
    if ref is not None:
        t0 = time.perf_counter()
        rc = [ref.subcommand("align", ["-r", g, "-o", o + ".ref", t]) for t, g, o in jobs[:24]]
        dt = time.perf_counter() - t0
        out["reference_one_core"] = {"jobs": 24, "ok": rc.count(0), "seconds": dt, "traces_per_s": 24 / dt,
                                     "what": "tracy::sage(argc, argv) of the unmodified reference (oracle/_ref) on the first 24 jobs"}
    if ref is not None:
        t0 = time.perf_counter()
        rc = [ref.subcommand("decompose", ["-r", g, "-o", o + ".ref", "-i", "30", t]) for t, g, o in djobs[:12]]
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        rc2 = [ref.subcommand("consensus", ["-o", o + ".ref", a, b]) for a, b, o in cjobs[:12]]
        out["reference_consensus_one_core"] = {"jobs": 12, "ok": rc2.count(0), "seconds": time.perf_counter() - t1, "pairs_per_s": 12 / (time.perf_counter() - t1)}
        out["reference_decompose_one_core"] = {"jobs": 12, "ok": rc.count(0), "seconds": dt, "traces_per_s": 12 / dt,
                                               "what": "tracy::indigo(argc, argv) of the unmodified reference on the first 12 jobs"}
print(json.dumps(out, indent=1))
