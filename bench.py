#!/usr/bin/env python
"""bench.py -- GCUPS of the batched Gotoh fill + traceback on synthetic 1 kb x 4 kb trace/window pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs P]

One "step" = one pass of the hot path (tb_gotoh_ps with traceback, reference src/gotoh.h:71-174) over one batch of
P pairs per GPU (BASELINE.json configs[1]: P = 100 000, m = 1000, n = 4000, scores 3/-5/-10/-4, AlignConfig<true,false>).
Prints ONE JSON line (rank 0):
  value        whole-job GCUPS (sum of m*n over all ranks' pairs / max-over-ranks device time), inputs resident in HBM
  e2e          same metric through the public host-buffer call (Context.gotoh / tb_gotoh_ps with TB_MEM_HOST): pinned
               host inputs -> H2D -> kernels -> D2H of scores + traceback strings, all inside the timed region
  roofline     dominant kernel: algorithmic bytes (0.5075 B/cell) / CUDA-event kernel time vs measured HBM peak
  cpu_baseline tracy's own gotoh() (oracle/_ref, unmodified reference headers) on this box's host cores, bounded sample
`--impl reference` times only that CPU path (rank 0; other ranks exit 0).
For N > 1 launch under torchrun (one rank per GPU); pairs are independent so ranks shard the batch with no data-path
collective; the timed step ends with the one all-gather of scores the north star names.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GCUPS (Gotoh DP cells/s), batched fill + traceback, 1 kb trace profiles x 4 kb reference windows"
SC = (3, -5, -10, -4)
HFREE, VFREE = 1, 0


def algorithmic_bytes_per_pair(m, n):
    """SURVEY.md section 8d / DESIGN.md: inputs + 4-bit pointer stream written once + pointers read on the walk + ops + score."""
    return 16 * m + n + ((m + 1) * (n + 1) + 1) // 2 + (m + n) // 2 + (m + n) + 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def gen_batch_into(prof_out, win_out, m, n, seed, chunk=5000):
    """Config-2 generator (tracy_b200/synth.py) written chunk-wise into preallocated (pinned) host arrays."""
    from tracy_b200 import synth
    P = prof_out.shape[0]
    for lo in range(0, P, chunk):
        hi = min(P, lo + chunk)
        p, w = synth.align_batch(hi - lo, m, n, seed=seed + 7919 * (lo // chunk))
        prof_out[lo:hi] = p
        win_out[lo:hi] = w


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---- the CPU reference arm ------------------------------------------------------------------------------------
def cpu_reference_gcups(m, n, pairs_per_thread, threads, repeats=1, seed=4242):
    """tracy's gotoh() (fill + bitsets + traceback + _createAlignment) on `threads` host threads, each running the
    reference single-threaded path over its own pairs (ctypes releases the GIL). Returns (gcups, seconds, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import loader
    from tracy_b200 import synth
    impl, kind = loader.ref(), "reference"
    if impl is None:
        impl, kind = loader.port(), "port"
    total = pairs_per_thread * threads
    prof, win = synth.align_batch(total, m, n, seed=seed)
    parts = [(np.ascontiguousarray(prof[t * pairs_per_thread:(t + 1) * pairs_per_thread]),
              bytes(np.ascontiguousarray(win[t * pairs_per_thread:(t + 1) * pairs_per_thread]).reshape(-1))) for t in range(threads)]

    def work(part):
        cells, _ = impl.bench_gotoh_ps(part[0], part[1], m, n, HFREE, VFREE, SC, True)
        return cells

    best = None
    with ThreadPoolExecutor(threads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            cells = sum(ex.map(work, parts))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return cells / best / 1e9, best, kind, total


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.cpu_threads or cores))
    ppt = args.cpu_pairs_per_thread
    for _ in range(args.warmup and 1):
        cpu_reference_gcups(args.m, args.n, 1, threads)
    vals, secs = [], []
    kind = total = None
    for _ in range(args.steps):
        g, s, kind, total = cpu_reference_gcups(args.m, args.n, ppt, threads)
        vals.append(g); secs.append(s)
    v = float(np.mean(vals))
    sample = f"{total} pairs of {args.m}x{args.n} per step ({ppt} per thread x {threads} threads), gotoh() with traceback + _createAlignment"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(secs) * 1e3), "higher_is_better": True, "scaling": "strong" if args.total_pairs else "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": threads, "kind": kind, "sample": sample, "cpu_model": cpu_model()},
        "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def workload_config(args, world):
    """The same dict for both arms at every N (the driver compares them): `world` is --gpus, not the rank count of the arm."""
    return {"workload": f"batched Gotoh fill+traceback (tb_gotoh_ps): {args.pairs} pairs/GPU of {args.m} bp trace profile x {args.n} bp reference window "
                        f"(BASELINE.json configs[1]), scores 3/-5/-10/-4, AlignConfig<true,false>",
            "pairs_per_gpu": args.pairs, "m": args.m, "n": args.n, "traceback": True, "outputs": "score + s/h/v string + both gapped alignment rows per pair",
            "sharding": f"{world} x independent pair ranges, no data-path collective"
            + ("; one all-gather of scores per step" if world > 1 else ""),
            "l2": "inputs (2.8 GB/GPU) and the pointer scratch (GBs) exceed the 126 MB L2; no explicit flush"}


# ---- the B200 arm ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """Pin this rank to the CPUs NVML names as local to its GPU BEFORE the pinned staging buffers are allocated, so that
    they land on the GPU's NUMA node (one process per GPU; 8 ranks copying 2.4 GB per step each otherwise share one node's
    memory controller). Returns the number of CPUs bound to, 0 when NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_b200(args, rank, world, local_rank):
    import torch
    import tracy_b200
    from tracy_b200 import AlignConfig, DnaScore
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tracy_b200 has no CPU path (use --impl reference for the CPU baseline)")
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else 0
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = tracy_b200.Context(local_rank)
    P, m, n = args.pairs, args.m, args.n
    stride = (m + n + 15) // 16 * 16
    sc, ac = DnaScore(*SC), AlignConfig(bool(HFREE), bool(VFREE))

    # pinned host staging (inputs and outputs of the e2e call)
    h_prof = torch.empty((P, 6, m), dtype=torch.float32, pin_memory=True)
    h_win = torch.empty((P, n), dtype=torch.uint8, pin_memory=True)
    h_scores = torch.empty(P, dtype=torch.int32, pin_memory=True)
    h_len = torch.empty(P, dtype=torch.int32, pin_memory=True)
    pstride = ((m + n + 3) // 4 + 15) // 16 * 16                      # 2-bit packed ops: the form that travels in the e2e leg
    h_pk = torch.empty((P, pstride), dtype=torch.uint8, pin_memory=True)
    h_row0 = torch.empty((P, stride), dtype=torch.uint8, pin_memory=True)
    h_row1 = torch.empty((P, stride), dtype=torch.uint8, pin_memory=True)
    t0 = time.perf_counter()
    gen_batch_into(h_prof.numpy(), h_win.numpy(), m, n, seed=44 + 1000003 * rank)
    gen_s = time.perf_counter() - t0
    a1 = tracy_b200.uniform_profiles(h_prof.numpy(), trace_profiles=True)   # the generator keeps createProfile's invariant: rows 4, 5 exactly zero
    a2 = tracy_b200.uniform_seqs(h_win.numpy())

    # device-resident copies for the `value` leg
    dev = torch.device("cuda", local_rank)
    d_prof, d_win = h_prof.to(dev), h_win.to(dev)
    d_aoff, d_alen = torch.from_numpy(a1.off).to(dev), torch.from_numpy(a1.len).to(dev)
    d_boff, d_blen = torch.from_numpy(a2.off).to(dev), torch.from_numpy(a2.len).to(dev)
    d_scores = torch.zeros(P, dtype=torch.int32, device=dev)
    d_ops = torch.zeros((P, stride), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(P, dtype=torch.int32, device=dev)
    counts = [P] * world
    g_scores = torch.empty(P * world, dtype=torch.int32, device=dev) if world > 1 else None
    bcast = None
    if world > 1:   # the one broadcast the north star names: the reference text, once, outside the steps; every rank indexes its copy
        try:
            from tracy_b200 import shard
            nref = 8_000_000
            text = torch.zeros(nref, dtype=torch.uint8, device=dev)
            if rank == 0:
                text.copy_(torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (nref,), device=dev)])
                text[-1] = 10
            idx = shard.broadcast_genome_and_index(ctx, text, src=0)
            probe = text[123456 + 1000 * rank: 123456 + 1000 * rank + 900].cpu().numpy().tobytes()
            hit = ctx.anchor(idx, [probe])
            bcast = {"text_bytes": nref, "index_device_bytes": int(idx.device_bytes),
                     "anchor_check": bool(hit["anchored"][0]) and int(hit["bestpos"][0]) == 123456 + 1000 * rank}
            idx.close()
            del text
        except Exception as e:   # the broadcast + index check must never take the bench down
            bcast = {"error": str(e)[:200]}

    def device_step():
        ctx.gotoh_device("ps", d_prof.data_ptr(), d_aoff.data_ptr(), d_alen.data_ptr(), d_win.data_ptr(), d_boff.data_ptr(), d_blen.data_ptr(), P,
                         d_scores.data_ptr(), d_ops.data_ptr(), stride, d_len.data_ptr(), sc, ac)
        ms = ctx.last_call_ms()
        k = ctx.last_kernel_ms()
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(g_scores, d_scores)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, k

    def host_step():     # what a drop-in gotoh() caller gets back: score, traceback (2 bits per op on the wire), both gapped rows
        ctx.gotoh("ps", a1, a2, sc, ac, traceback=True, out=(h_scores.numpy(), h_pk.numpy(), h_len.numpy()), rows=(h_row0.numpy(), h_row1.numpy()), packed=True)

    def host_step_ops_only():   # the compact form: score + packed traceback, no rows
        ctx.gotoh("ps", a1, a2, sc, ac, traceback=True, out=(h_scores.numpy(), h_pk.numpy(), h_len.numpy()), packed=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        device_step()
    st0 = ctx.stats()
    sampler = ClockSampler(local_rank)
    sync()
    sampler.start()
    w0 = time.perf_counter()
    dev_ms, kern = 0.0, {"packed_ms": 0.0, "general_ms": 0.0}
    for _ in range(args.steps):
        ms, k = device_step()
        dev_ms += ms
        for key in kern:
            kern[key] += k[key]
    sync()
    wall_ms = (time.perf_counter() - w0) * 1e3
    st1 = ctx.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    # ---- e2e: host buffers through the public call ----
    for _ in range(min(args.warmup, 2)):
        host_step()
    se0 = ctx.stats()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    se1 = ctx.stats()
    # the same call without the rows (score + packed ops only), one warm-up + min(steps, 2) steps
    host_step_ops_only()
    so0 = ctx.stats()
    sync()
    t0 = time.perf_counter()
    nops = min(args.steps, 2)
    for _ in range(nops):
        host_step_ops_only()
    sync()
    e2e_ops_ms = (time.perf_counter() - t0) * 1e3 / nops
    so1 = ctx.stats()
    host_step()                                                        # leave the rows of the full call in the buffers for the parity check

    # spot parity inside the bench: device leg and host leg agree, and a few pairs match the CPU oracle
    assert torch.equal(d_scores.cpu(), h_scores), "device-resident and host-buffer legs disagree"
    assert torch.equal(d_len.cpu(), h_len), "device-resident and host-buffer legs disagree on the traceback lengths"
    chk = {"pairs_checked_vs_oracle": 0}
    if rank == 0:
        from oracle import loader
        port = loader.port()
        pk_h, len_h, r0_h, r1_h = h_pk.numpy(), h_len.numpy(), h_row0.numpy(), h_row1.numpy()
        for i in range(0, P, max(1, P // 4))[:4]:
            ws, wops = port.gotoh_ps(h_prof.numpy()[i], bytes(h_win.numpy()[i]), HFREE, VFREE, SC)
            assert int(h_scores[i]) == ws and tracy_b200.unpack_ops(pk_h[i], len_h[i]) == wops, f"bench parity: pair {i} differs from the oracle"
            want_rows = tracy_b200.rows_from_ops("ps", h_prof.numpy()[i], bytes(h_win.numpy()[i]), wops)
            assert (bytes(r0_h[i, : len_h[i]]), bytes(r1_h[i, : len_h[i]])) == want_rows, f"bench parity: rows of pair {i} differ"
            chk["pairs_checked_vs_oracle"] += 1
    checksum = int(h_scores.numpy().astype(np.int64).sum())

    if world > 1:
        t = torch.tensor([dev_ms, wall_ms, e2e_ms, e2e_ops_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, e2e_ms, e2e_ops_ms = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    cells_step = float(P) * m * n * world
    value = cells_step * args.steps / (dev_ms * 1e-3) / 1e9
    e2e_val = cells_step * args.steps / (e2e_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    dom = "packed_ms" if kern["packed_ms"] >= kern["general_ms"] else "general_ms"
    dom_name = {"packed_ms": "gotoh_packed_kernel<traceback, vfree=0, classes=4>", "general_ms": "gotoh_general_kernel<PS,traceback>"}[dom]
    k_ms = kern[dom] / args.steps
    bytes_launch = float(algorithmic_bytes_per_pair(m, n)) * P
    achieved = bytes_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(float(P) * m * n),
            "peak_source": peak_src, "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": bytes_launch,
            "note": "integer DP: the binding limit is the ALU/FMA pipe issue rate (DESIGN.md section 4), HBM fraction is reported as the contract asks",
            "packed_pairs_last_step": ctx.last_packed_pairs()}
    prof_info = load_profile_info()
    if prof_info.get("warp_instructions_per_pair") and k_ms > 0:
        # the bound that actually binds: warp-instruction issue slots, 1 per clock per scheduler, 4 schedulers per SM
        ipp = float(prof_info["warp_instructions_per_pair"])
        ach = ipp * P / (k_ms * 1e-3)
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        roof["issue"] = {"bound": "alu+lsu issue (warp instructions / s over 148 SMs x 4 schedulers)", "achieved": ach, "peak": 148 * 4 * sm_mhz * 1e6,
                         "frac": ach / (148 * 4 * sm_mhz * 1e6), "peak_at_1965_mhz": 148 * 4 * 1965e6, "peak_at_1342_mhz": 148 * 4 * 1342e6,
                         "frac_at_1965_mhz": ach / (148 * 4 * 1965e6), "warp_instructions_per_pair": ipp,
                         "pipes_from_capture": prof_info.get("pipes"), "source": prof_info.get("capture")}
    out = {
        "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_pairs else "weak", "vs_baseline": None,
        "dtype": "u16x2",            # biased unsigned 16-bit DP cells, two per 32-bit register (int32 in the general kernel)
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_val, "unit": "GCUPS", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": world * (se1["h2d_bytes"] - se0["h2d_bytes"]) // args.steps,      # every rank moves the same amount
                "d2h_bytes_per_step": world * (se1["d2h_bytes"] - se0["d2h_bytes"]) // args.steps},
        "e2e_ops_only": {"value": float(P) * m * n * world / (e2e_ops_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ops_ms,
                         "h2d_bytes_per_step": world * (so1["h2d_bytes"] - so0["h2d_bytes"]) // nops, "d2h_bytes_per_step": world * (so1["d2h_bytes"] - so0["d2h_bytes"]) // nops,
                         "note": "same host-buffer call returning score + 2-bit packed traceback only (no gapped rows)"},
        "gpu_launches": launches, "roofline": roof, "clocks": clocks, "host_cpus_bound_per_rank": numa_cpus, "reference_broadcast": bcast, "wall_ms_per_step": wall_ms / args.steps,
        "kernel_ms": {k: v / args.steps for k, v in kern.items()}, "parity": dict(chk, score_checksum=checksum), "gen_seconds": gen_s,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, args.cpu_threads or cores))
        ppt = max(1, (args.cpu_sample_pairs + threads - 1) // threads)
        g, s, kind, total = cpu_reference_gcups(m, n, ppt, threads)
        g1, s1, _, total1 = cpu_reference_gcups(m, n, args.cpu_single_pairs, 1, seed=777)
        out["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": kind, "seconds": s, "cpu_model": cpu_model(),
                               "sample": f"{total} pairs of {m}x{n} ({ppt} per thread x {threads} threads), tracy gotoh() with traceback + _createAlignment",
                               "single_thread": {"value": g1, "unit": "GCUPS", "cores": 1, "seconds": s1,
                                                 "sample": f"{total1} pairs of {m}x{n} on one thread: the faithful single-threaded tracy path"}}
    print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def load_profile_info():
    """Per-pair instruction count and pipe utilisations of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def load_traffic(cells_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json holds bytes per DP cell of that capture; scaled here to this launch's cells), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dram_bytes_per_cell"]) * cells_per_launch
    except (OSError, KeyError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=100000, help="pairs per GPU per step")
    ap.add_argument("--m", type=int, default=1000)
    ap.add_argument("--n", type=int, default=4000)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-pairs-per-thread", type=int, default=8, help="--impl reference: pairs per thread per step")
    ap.add_argument("--cpu-sample-pairs", type=int, default=2048, help="cpu_baseline of the b200 arm: pairs timed over all host threads (BASELINE.md section 3)")
    ap.add_argument("--cpu-single-pairs", type=int, default=32, help="cpu_baseline.single_thread: pairs timed on one thread")
    ap.add_argument("--total-pairs", type=int, default=0, help="strong split: this many pairs over all GPUs (BASELINE.json configs[4]: 1 000 000 over 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.total_pairs:
        args.pairs = (args.total_pairs + max(args.gpus, 1) - 1) // max(args.gpus, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
