"""Throughput of the other entry points / BASELINE.json configs (not the headline bench): device time from the library's
CUDA events where available, otherwise host clock around the host-buffer call. Prints one JSON object."""
import json, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth, msa

ctx = tracy_b200.Context(0)
sc = DnaScore(3, -5, -10, -4)
out = {}

def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best

# config 2 shape, score only (gotohScore: orientation picks)
N, m, n = 20000, 1000, 4000
prof, win = synth.align_batch(N, m, n, seed=44)
a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
t = timed(lambda: ctx.gotoh("ps", a1, a2, sc, AlignConfig(True, False), traceback=False))
k = ctx.last_kernel_ms()
out["ps_score_only_1000x4000"] = {"pairs": N, "host_call_gcups": N * m * n / t / 1e9, "kernel_gcups": N * m * n / ((k["packed_ms"] + k["general_ms"]) * 1e-3) / 1e9}
# the same call with the batch in page-locked memory (Context.pinned_empty = tb_host_alloc)
pp_, pw_ = ctx.pinned_empty(prof.shape, np.float32), ctx.pinned_empty(win.shape, np.uint8)
pp_[:] = prof; pw_[:] = win
b1, b2 = tracy_b200.uniform_profiles(pp_), tracy_b200.uniform_seqs(pw_)
ps_ = ctx.pinned_empty((N,), np.int32)
t = timed(lambda: ctx.gotoh("ps", b1, b2, sc, AlignConfig(True, False), traceback=False, out=(ps_, None, None)))
out["ps_score_only_1000x4000"]["host_call_gcups_pinned"] = N * m * n / t / 1e9

# string x string with traceback (allele alignments of decompose)
N2 = 4000
seqs = [bytes(b"ACGT"[x] for x in np.random.default_rng(i).integers(0, 4, 1000)) for i in range(64)]
S1 = [seqs[i % 64] for i in range(N2)]
S2 = [bytes(win[i]) for i in range(N2)]
t = timed(lambda: ctx.gotoh("ss", S1, S2, sc, AlignConfig(True, False)))
k = ctx.last_kernel_ms()
out["ss_traceback_1000x4000"] = {"pairs": N2, "kernel_gcups": N2 * 1000 * 4000 / ((k["packed_ms"] + k["general_ms"]) * 1e-3) / 1e9, "packed_pairs": int(ctx.last_packed_pairs())}

# config 4: all-pairs profile x profile, score only (distance matrix) and with traceback
rng = np.random.default_rng(5)
NT, L = 256, 900
contig = synth.random_seq(rng, 115 * NT + L)
profs = [synth.profile_from_seq(rng, contig[115 * i: 115 * i + L], 0.3) for i in range(NT)]
ii, jj = np.triu_indices(NT, 1)
A = tracy_b200.pack_profiles([profs[i] for i in ii]); B = tracy_b200.pack_profiles([profs[j] for j in jj])
t = timed(lambda: ctx.gotoh("pp", A, B, sc, AlignConfig(True, True), traceback=False), reps=2)
k = ctx.last_kernel_ms()
out["pp_all_pairs_score_900x900"] = {"pairs": len(ii), "kernel_gcups": len(ii) * L * L / ((k["general_ms"] + k["packed_ms"]) * 1e-3) / 1e9, "host_call_s": t}
sub = 4096
A2 = tracy_b200.pack_profiles([profs[i] for i in ii[:sub]]); B2 = tracy_b200.pack_profiles([profs[j] for j in jj[:sub]])
t = timed(lambda: ctx.gotoh("pp", A2, B2, sc, AlignConfig(True, True)), reps=2)
k = ctx.last_kernel_ms()
out["pp_traceback_900x900"] = {"pairs": sub, "kernel_gcups": sub * L * L / ((k["general_ms"] + k["packed_ms"]) * 1e-3) / 1e9}

# config 3: decompose sweeps, 10k traces, maxindel 30
NTd = 10000
ref = [synth.random_seq(rng, 1100, b"ACGT-") for _ in range(64)]
pri = [synth.random_seq(rng, 1000, b"ACGTN") for _ in range(64)]
sec = [synth.random_seq(rng, 1000, b"ACGTRYSWKMN") for _ in range(64)]
R = [ref[i % 64] for i in range(NTd)]; P = [pri[i % 64] for i in range(NTd)]; S = [sec[i % 64] for i in range(NTd)]
t = timed(lambda: ctx.decompose_sweep(R, P, S, [950] * NTd, [300] * NTd, [310] * NTd, [30] * NTd, [30] * NTd), reps=2)
out["decompose_sweep_10k_traces_pm30"] = {"traces": NTd, "kernel_ms": ctx.last_kernel_ms()["sweep_ms"], "host_call_s": t,
                                          "compared_columns_per_s": NTd * 59 * 650 / (ctx.last_kernel_ms()["sweep_ms"] * 1e-3)}

# config 3 tail: allelicFraction (FP64 grid fit) for 2 000 heterozygous traces of 600 basecalls, ~150 differing positions each
NF = 2000
tr_l, pos_l, pri_l, sec_l = [], [], [], []
for i in range(64):
    nbc = 600
    ns = 12 * nbc + 30
    tr = rng.integers(0, 40, size=(4, ns)).astype(np.int32)
    pos = (12 * np.arange(nbc) + 8).astype(np.int32)
    pri = bytearray(rng.choice(list(b"ACGT"), nbc).astype(np.uint8).tobytes()); sec = bytearray(pri)
    for j in range(nbc):
        a = b"ACGT".index(pri[j]); h = int(rng.integers(500, 1500))
        if rng.random() < 0.3:
            b = (a + int(rng.integers(1, 4))) % 4
            sec[j] = b"ACGT"[b]; tr[a, pos[j]] += int(h * 0.6); tr[b, pos[j]] += int(h * 0.4)
        else:
            tr[a, pos[j]] += h
    tr_l.append(tr); pos_l.append(pos); pri_l.append(bytes(pri)); sec_l.append(bytes(sec))
sel = [i % 64 for i in range(NF)]
t = timed(lambda: ctx.allelic_fraction([tr_l[i] for i in sel], [pos_l[i] for i in sel], [pri_l[i] for i in sel], [sec_l[i] for i in sel], 50, 50), reps=2)
fm = C = None
import ctypes
v = ctypes.c_float()
ctx._lib.tb_ctx_last_fraction_ms(ctx._h, ctypes.byref(v))
out["allelic_fraction_2k_traces"] = {"traces": NF, "kernel_ms": v.value, "host_call_s": t, "traces_per_s_kernel": NF / (v.value * 1e-3),
                                     "grid_points_per_trace": 176851}

# ingest: 4 000 ABIF files of 12 k samples x 4 channels (host directory walk + GPU unpack), host clock around the call
chs = [rng.integers(0, 2000, 12000) for _ in range(4)]
f = synth.abif_bytes(chs, b"GATC", np.arange(990) * 12 + 20, synth.random_seq(rng, 990), rng.integers(0, 60, 990))
files = [f] * 4000
t = timed(lambda: ctx.read_traces(files), reps=2)
out["ingest_4k_abif_files"] = {"files": len(files), "file_bytes": len(f), "host_call_s": t, "files_per_s": len(files) / t,
                               "note": "Python list/array assembly included; the GPU unpack itself is copy-bound"}
print(json.dumps(out, indent=1))
