"""Throughput of the anchoring row (tb_index_build / tb_anchor) on one B200, beside the reference's own FM-index path
(oracle/_ref: unmodified src/fmindex.h over sdsl csa_wt) timed on a sample of the same traces on the host.
usage: python profiles/bench_anchor.py [--mbp 64] [--traces 100000] [--cpu-sample 200] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=int, default=64)
    ap.add_argument("--traces", type=int, default=100000)
    ap.add_argument("--len", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import tracy_b200
    rng = np.random.default_rng(46)
    n, nt, L = a.mbp * 1_000_000, a.traces, a.len
    text = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).astype(np.uint8)
    for c in range(1, 8):
        text[c * (n // 8)] = ord("\n")                             # 8 sequences
    text[-1] = ord("\n")
    pos = rng.integers(0, n - L - 2, nt)
    fw = rng.random(nt) < 0.5
    cons = np.empty((nt, L), np.uint8)
    for i in range(nt):
        cons[i] = text[pos[i]:pos[i] + L]
    cons[cons == ord("\n")] = ord("A")
    sub = rng.random((nt, L)) < 0.01                               # 1 % substitutions
    cons[sub] = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(sub.sum()))
    rc = np.frombuffer(cons.tobytes().translate(COMP), np.uint8).reshape(nt, L)[:, ::-1]
    cons = np.where(fw[:, None], cons, rc)
    arena = tracy_b200.uniform_seqs(np.ascontiguousarray(cons))

    ctx = tracy_b200.Context(0)
    t0 = time.perf_counter()
    idx = ctx.build_index(text)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    idx2 = ctx.build_index(text)
    build2_s = time.perf_counter() - t0
    idx2.close()
    r = ctx.anchor(idx, arena)                                     # warm-up
    ms, wall = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        r = ctx.anchor(idx, arena)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(ctx.last_anchor_ms())
    ok = float((r["anchored"] & (r["forward"] == fw) & (np.abs(r["bestpos"] - pos) <= 3)).mean())
    kq = nt * 2 * (L - 100)                                        # k-mer lookups per call (both strands)
    out = {"metric": "anchoring (scanSequence+findMaxFreq+orientation), traces/s", "text_mbp": a.mbp, "traces": nt, "trace_len": L,
           "index_build_s_first": build_s, "index_build_s": build2_s, "index_device_bytes": idx.device_bytes,
           "unique_kernel_ms": float(np.median(ms)), "e2e_ms_host_buffers": float(np.median(wall)),
           "traces_per_s_kernel": nt / (np.median(ms) * 1e-3), "traces_per_s_e2e": nt / (np.median(wall) * 1e-3),
           "kmer_lookups_per_s_kernel": kq / (np.median(ms) * 1e-3), "recovered_planted": ok,
           "decided_by_unique_pass": float((r["pass_"] == 1).mean())}
    if not a.no_cpu:
        from oracle import loader
        ref = loader.ref()
        if ref is not None:
            t0 = time.perf_counter()
            h = ref.fm_build(text.tobytes())
            out["cpu_fm_build_s"] = time.perf_counter() - t0
            ns = min(a.cpu_sample, nt)
            t0 = time.perf_counter()
            agree = 0
            for i in range(ns):
                g = ref.get_reference_slice(h, 1, cons[i].tobytes(), 50, 50, 15, 1000, 3, refslice=b"A")
                agree += int(g["ok"] == bool(r["anchored"][i]) and g["forward"] == bool(r["forward"][i]) and g["kmersupport"] == int(r["kmersupport"][i]))
            dt = time.perf_counter() - t0
            ref.fm_free(h)
            out["cpu_baseline"] = {"value": ns / dt, "unit": "traces/s", "cores": 1, "kind": "reference", "sample": f"{ns} of the same traces, getReferenceSlice over sdsl csa_wt<>",
                                   "agree_with_gpu": agree}
    print(json.dumps(out))
    idx.close()
    ctx.close()


if __name__ == "__main__":
    main()
