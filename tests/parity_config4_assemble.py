"""The de novo branch of `tracy assemble` on the configs[3] shape (900 bp traces tiling a contig, step 115, every other one
reverse-complemented) at the largest size the reference finishes in minutes on one core (its revSeqBasedOnDist runs its fills one
after the other): N = 112 traces, about 44 000 fills. Reference: revSeqBasedOnDist -> msa -> consensus from the unmodified headers
(oracle/_ref). This package: drivers.assemble_denovo on the GPU. Compared: orientation vector, leaf order, every alignment row,
gapped consensus, consensus and quality strings. Test infrastructure (it runs the oracle), run by hand: python tests/parity_config4_assemble.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import DnaScore, drivers, msa, synth
from oracle import loader

N = int(sys.argv[1]) if len(sys.argv) > 1 else 112
L, STEP, SC = 900, 115, (3, -5, -10, -4)
rng = np.random.default_rng(146)
contig = synth.random_seq(rng, STEP * N + L)
comp = bytes.maketrans(b"ACGT", b"TGCA")
profs = []
for i in range(N):
    s = synth.mutate_seq(rng, contig[STEP * i: STEP * i + L + 16], 0.01, 0.004)[:L - int(rng.integers(0, 40))]
    profs.append(synth.profile_from_seq(rng, s.translate(comp)[::-1] if i % 2 else s, 0.3))
ctx = tracy_b200.Context(0)
msa.orientation_table(ctx, profs[:6], DnaScore(*SC))
t0 = time.perf_counter()
T = msa.orientation_table(ctx, [p.copy() for p in profs], DnaScore(*SC))
r = drivers.assemble_denovo(ctx, [p.copy() for p in profs], DnaScore(*SC), 0.5, 0.05, table=T)
gpu_s = time.perf_counter() - t0
big = int(ctx.last_big_pairs())
ref = loader.ref()
t1 = time.perf_counter()
fwd = ref.rev_seq_based_on_dist(profs, [True] * N, SC)
oriented = [p if f else ref.revcomp_profile(p) for p, f in zip(profs, fwd)]
g = ref.msa(oriented, SC, 0.05)
ref_s = time.perf_counter() - t1
same = {"forward": r["forward"] == [bool(x) for x in fwd], "kept_all": r["kept"] == list(range(N)),
        "leaf_order": list(r["seqidx"]) == [int(x) for x in g["seqidx"]], "rows": bool(np.array_equal(r["rows"], g["rows"])),
        "gapped_consensus": r["gapped"] == bytes(g["gapped"]), "consensus": r["consensus"] == bytes(g["cons"]), "quality": r["quality"] == bytes(g["qual"])}
print(json.dumps({"traces": N, "flipped": int(sum(1 for f in fwd if not f)), "msa_columns": int(g["rows"].shape[1]), "consensus_bp": len(g["cons"]),
                  "identical": same, "gpu_seconds": round(gpu_s, 3), "reference_seconds_one_core": round(ref_s, 1),
                  "last_call_pairs_spread_over_warps": big}))
sys.exit(0 if all(same.values()) else 1)
