"""BASELINE configs[3] at full size (512 trace profiles of 900 bp): the orientation table of `tracy assemble` -- 1 046 528 score fills
in ONE launch of the screened profile x profile kernel -- and a fixed sample of 65 536 of its entries against tracy's own gotohScore()
(unmodified reference headers, oracle/_ref, all host threads); 64 alignments of overlapping neighbours (score and both gapped rows)
against tracy's gotoh(); the table computed with the literal substitution score only (TRACY_B200_PP_SCREEN=0) must be identical
entry for entry. Test infrastructure (it runs the oracle), run by hand: python tests/parity_config4_512.py (about a minute)."""
import hashlib, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, msa, synth
from oracle import loader

N, L, STEP, SAMPLE = 512, 900, 115, 65536
rng = np.random.default_rng(46)
contig = synth.random_seq(rng, STEP * N + L)
comp = bytes.maketrans(b"ACGT", b"TGCA")
profs = []
for i in range(N):
    s = bytearray(contig[STEP * i: STEP * i + L])
    for q in rng.integers(0, L, 9):
        s[q] = b"ACGT"[int(rng.integers(0, 4))]
    s = bytes(s)
    profs.append(synth.profile_from_seq(rng, s.translate(comp)[::-1] if i % 2 else s, 0.3))
ctx = tracy_b200.Context(0)
sc = DnaScore(3, -5, -10, -4)
sct = (3, -5, -10, -4)
msa.orientation_table(ctx, profs[:8], sc)
t0 = time.perf_counter()
T = msa.orientation_table(ctx, profs, sc)
t_table = time.perf_counter() - t0
os.environ["TRACY_B200_PP_SCREEN"] = "0"
T_literal = msa.orientation_table(ctx, profs, sc)
del os.environ["TRACY_B200_PP_SCREEN"]
ref = loader.ref()
flip = [ref.revcomp_profile(p) for p in profs]
pick = np.random.default_rng(7)
# half of the sample among overlapping neighbours (large scores), half anywhere
cases = []
while len(cases) < SAMPLE:
    i = int(pick.integers(0, N))
    k = int(pick.integers(max(0, i - 7), min(N, i + 8))) if len(cases) % 2 == 0 else int(pick.integers(0, N))
    if i != k:
        cases.append((i, k, int(pick.integers(0, 2)), int(pick.integers(0, 2))))


def one(c):
    i, k, oi, ok = c
    return ref.gotoh_score(flip[i] if oi else profs[i], flip[k] if ok else profs[k], 1, 1, sct)


t1 = time.perf_counter()
with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
    want = list(ex.map(one, cases))
cpu_s = time.perf_counter() - t1
bad = [c for c, w in zip(cases, want) if int(T[c[0], c[1], c[2], c[3]]) != w]
# alignments of overlapping neighbours in matching orientation
pairs = [(i, i + 1) for i in range(0, 128, 2)]
A = [profs[i] if i % 2 == 0 else flip[i] for i, _ in pairs]
B = [profs[k] if k % 2 == 0 else flip[k] for _, k in pairs]
s, ops, ol, r0, r1 = ctx.gotoh("pp", A, B, sc, AlignConfig(True, True), rows=True)
with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
    wal = list(ex.map(lambda q: ref.gotoh(A[q], B[q], 1, 1, sct), range(len(pairs))))
abad = [q for q in range(len(pairs)) if (int(s[q]), bytes(r0[q, : ol[q]]), bytes(r1[q, : ol[q]])) != wal[q]]
print(json.dumps({"traces": N, "table_entries": int(4 * N * (N - 1)), "table_seconds": round(t_table, 3),
                  "entries_checked_against_reference_headers": SAMPLE, "mismatches": len(bad), "first_mismatches": bad[:5],
                  "screened_equals_literal_only_table": bool(np.array_equal(T, T_literal)),
                  "alignments_checked": len(pairs), "alignment_mismatches": len(abad), "reference_cpu_s": round(cpu_s, 1), "threads": os.cpu_count(),
                  "largest_checked_score": int(max(want)), "sha256_of_table": hashlib.sha256(np.ascontiguousarray(T, np.int64).tobytes()).hexdigest()}))
sys.exit(1 if bad or abad or not np.array_equal(T, T_literal) else 0)
