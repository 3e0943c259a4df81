"""Throughput of the profile x profile paths (B200, CUDA events of the library): the fp32x2 kernel in its two register
variants against the general int32 kernel on the all-pairs shape of assemble (900 x 900, score only and with traceback), and
single big pairs (band-pipelined over many warps) against the same pair on one warp. Prints one JSON object."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth

ctx = tracy_b200.Context(0)
sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, True)
out = {}
rng = np.random.default_rng(5)
NT, L = 320, 900
contig = synth.random_seq(rng, 115 * NT + L)
profs = [synth.profile_from_seq(rng, contig[115 * i: 115 * i + L], 0.3) for i in range(NT)]
ii, jj = np.triu_indices(NT, 1)
A = tracy_b200.pack_profiles([profs[i] for i in ii]); B = tracy_b200.pack_profiles([profs[j] for j in jj])
sub = 8192
A2 = tracy_b200.pack_profiles([profs[i] for i in ii[:sub]]); B2 = tracy_b200.pack_profiles([profs[j] for j in jj[:sub]])


def kernel_ms():
    k = ctx.last_kernel_ms()
    return k["packed_ms"] + k["general_ms"]


def run(tag, env):
    for k, v in env.items():
        os.environ[k] = v
    ref = None
    for tb, (a, b, n) in (("score", (A, B, len(ii))), ("traceback", (A2, B2, sub))):
        best = 1e9
        for _ in range(3):
            s, _, _ = ctx.gotoh("pp", a, b, sc, ac, traceback=(tb == "traceback"))
            best = min(best, kernel_ms())
        out[f"all_pairs_900x900_{tb}_{tag}"] = {"pairs": n, "kernel_ms": best, "gcups": n * L * L / (best * 1e-3) / 1e9,
                                                "fast_pairs": int(ctx.last_packed_pairs()), "checksum": int(np.asarray(s, np.int64).sum())}
    for k in env:
        del os.environ[k]


run("fp32x2_arr", {})
if os.environ.get("PP_PROBE_QUICK"):
    print(json.dumps(out, indent=1)); sys.exit(0)
run("fp32x2_sel", {"TRACY_B200_PP_VARIANT": "sel"})
run("fp32x2_arr_literal_only", {"TRACY_B200_PP_SCREEN": "0"})
run("fp32x2_sel_literal_only", {"TRACY_B200_PP_SCREEN": "0", "TRACY_B200_PP_VARIANT": "sel"})
run("general", {"TRACY_B200_NO_PPFAST": "1"})

# big pairs: MSA-like profiles of growing length, one pair per call, with traceback
for n in (4000, 12000, 36000):
    g = synth.random_seq(rng, 2 * n)
    a = synth.profile_from_seq(rng, synth.mutate_seq(rng, g[: n + 50], 0.02, 0.01)[:n], 0.3)
    b = synth.profile_from_seq(rng, synth.mutate_seq(rng, g[n // 2: n // 2 + n + 50], 0.02, 0.01)[:n], 0.3)
    for tag, env in (("pipelined", {}), ("one_warp", {"TRACY_B200_NO_BIG": "1"})):
        if tag == "one_warp" and n > 12000:
            continue
        for k, v in env.items():
            os.environ[k] = v
        best, wall = 1e9, 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            s, ops, ol = ctx.gotoh("pp", [a], [b], sc, ac)
            wall = min(wall, time.perf_counter() - t0)
            best = min(best, kernel_ms())
        out[f"big_pair_{n}x{n}_{tag}"] = {"kernel_ms": best, "host_call_ms": wall * 1e3, "gcups": n * n / (best * 1e-3) / 1e9, "score": int(s[0]),
                                          "ops_len": int(ol[0]), "big_pairs": int(ctx.last_big_pairs())}
        for k in env:
            del os.environ[k]
print(json.dumps(out, indent=1))
