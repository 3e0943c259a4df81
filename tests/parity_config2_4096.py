"""SURVEY section 8d at full size: BASELINE configs[1] (100 000 pairs of 1 kb x 4 kb, 3/-5/-10/-4, AlignConfig<true,false>) through the
public host-buffer call -- the streamed form: one launch, inputs gated chunk by chunk -- and a fixed 4 096-pair subset of it pair by
pair against tracy's own gotoh() (unmodified reference headers, oracle/_ref; all host threads): score and both gapped rows; plus
a hash of all 100 000 scores for the record. Test infrastructure (it runs the oracle), kept under tests/; run by hand: python tests/parity_config2_4096.py (2.8 GB of inputs, ~1 min)."""
import hashlib, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth
from oracle import loader

P, m, n, SUB = 100000, 1000, 4000, 4096
ctx = tracy_b200.Context(0)
prof = ctx.pinned_empty((P, 6, m), np.float32)
win = ctx.pinned_empty((P, n), np.uint8)
for lo in range(0, P, 5000):
    p, w = synth.align_batch(5000, m, n, seed=777 + lo)
    prof[lo:lo + 5000] = p; win[lo:lo + 5000] = w
a1, a2 = tracy_b200.uniform_profiles(prof, trace_profiles=True), tracy_b200.uniform_seqs(win)
sc = (3, -5, -10, -4)
k0 = ctx.stats()["kernel_launches"]
t0 = time.perf_counter()
s, ops, ol, r0, r1 = ctx.gotoh("ps", a1, a2, DnaScore(*sc), AlignConfig(True, False), rows=True)
dt = time.perf_counter() - t0
launches = ctx.stats()["kernel_launches"] - k0
ref = loader.ref()
idx = (np.arange(SUB) * (P // SUB) + 7) % P


def one(i):
    return ref.gotoh(prof[i], ref.onehot(bytes(win[i])), 1, 0, sc)


t1 = time.perf_counter()
with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
    want = list(ex.map(one, [int(i) for i in idx]))
cpu_s = time.perf_counter() - t1
bad = [int(i) for k, i in enumerate(idx) if (int(s[i]), bytes(r0[i, : ol[i]]), bytes(r1[i, : ol[i]])) != want[k]]
ops_bad = 0
for k, i in enumerate(idx[::16]):
    if tracy_b200.rows_from_ops("ps", prof[i], bytes(win[i]), bytes(ops[i, : ol[i]])) != (want[16 * k][1], want[16 * k][2]):
        ops_bad += 1
print(json.dumps({"pairs": P, "kernel_launches": launches, "packed_pairs": int(ctx.last_packed_pairs()), "host_call_s": round(dt, 3),
                  "subset_pairs_checked_against_reference_headers": SUB, "mismatches": len(bad), "first_mismatches": bad[:5],
                  "ops_strings_checked": len(idx[::16]), "ops_mismatches": ops_bad, "reference_cpu_s": round(cpu_s, 1), "threads": os.cpu_count(),
                  "sha256_of_all_scores": hashlib.sha256(np.ascontiguousarray(s, np.int32).tobytes()).hexdigest(),
                  "score_sum": int(np.asarray(s, np.int64).sum())}))
sys.exit(1 if bad or ops_bad else 0)
