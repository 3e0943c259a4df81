"""GPU parity tests of the rows either side of the DP (a3 createProfile, a5 reverseComplementProfile,
a16/a17 the assemble glue on top of the batched profile x profile kernels) against the reference-generated goldens."""
import os

import numpy as np
import pytest

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, msa, synth
from test_glue import GOLD, SC, run_msa_case

pytestmark = pytest.mark.gpu


def test_create_profile_golden(ctx):
    n = int(GOLD["ncp"])
    tr = [GOLD[f"cp_tr{i}"] for i in range(n)]
    pos = [GOLD[f"cp_pos{i}"] for i in range(n)]
    pri = [bytes(GOLD[f"cp_pri{i}"]) for i in range(n)]
    sec = [bytes(GOLD[f"cp_sec{i}"]) for i in range(n)]
    tl = [int(GOLD[f"cp_trim{i}"][0]) for i in range(n)]
    trr = [int(GOLD[f"cp_trim{i}"][1]) for i in range(n)]
    out = ctx.create_profile(tr, pos, pri, sec, tl, trr)          # one ragged batch
    for i in range(n):
        want = GOLD[f"cp_out{i}"]
        assert out[i].shape == want.shape, i
        assert np.array_equal(out[i].view(np.uint32), want.view(np.uint32)), i      # bit-exact floats
    one = ctx.create_profile(tr[3:4], pos[3:4], pri[3:4], sec[3:4])                 # defaults: no trimming
    assert one[0].shape[1] == len(pos[3])


def test_revcomp_profile_golden(ctx):
    n = int(GOLD["ncp"])
    out = ctx.revcomp_profile([GOLD[f"cp_out{i}"] for i in range(n)])
    for i in range(n):
        assert np.array_equal(out[i].view(np.uint32), GOLD[f"cp_rc{i}"].view(np.uint32)), i
    twice = ctx.revcomp_profile(out)
    for i in range(n):
        assert np.array_equal(twice[i], GOLD[f"cp_out{i}"])                          # involution


def test_profile_feeds_the_dp_without_host_math(ctx, oracle_port):
    """createProfile output goes straight into tb_gotoh_ps: same score/traceback as the oracle run on the golden profile."""
    i = 4
    p = ctx.create_profile([GOLD[f"cp_tr{i}"]], [GOLD[f"cp_pos{i}"]], [bytes(GOLD[f"cp_pri{i}"])], [bytes(GOLD[f"cp_sec{i}"])])[0]
    ref = synth.random_seq(np.random.default_rng(3), 400)
    s, ops, ol = ctx.gotoh("ps", [p], [ref], DnaScore(*SC), AlignConfig(True, False))
    want = ctx.create_profile([GOLD[f"cp_tr{i}"]], [GOLD[f"cp_pos{i}"]], [bytes(GOLD[f"cp_pri{i}"])], [bytes(GOLD[f"cp_sec{i}"])], 0, 0)[0]
    assert (int(s[0]), bytes(ops[0, : ol[0]])) == oracle_port.gotoh_ps(want, ref, 1, 0, SC)


@pytest.mark.parametrize("idx", range(int(GOLD["nmsa"])))
def test_msa_pipeline_gpu(ctx, idx):
    run_msa_case(ctx, idx)


def test_assemble64_gpu(ctx):
    """SURVEY section 8d-4 at N = 64: orientation vector, distance matrix, leaf order, MSA rows and consensus equal the reference's."""
    from test_glue import run_assemble64
    run_assemble64(ctx)
    assert ctx.stats()["kernel_launches"] > 0


def test_exclude_unmatched(ctx, oracle_port):
    """The exclusion loop of assemble() (src/assemble.h:428-448): a stray trace is dropped, overlapping ones are kept;
    the batched rounds give the same booleans as the reference's first-hit scan done with the CPU oracle."""
    rng = np.random.default_rng(12)
    contig = synth.random_seq(rng, 400)
    profs = [synth.profile_from_seq(rng, contig[s: s + 160], 0.3) for s in (0, 60, 120, 200)]
    profs.append(synth.profile_from_seq(rng, synth.random_seq(rng, 150), 0.3))       # unrelated
    keep = msa.exclude_unmatched(ctx, profs, DnaScore(*SC), 0.5)
    want = []
    for i in range(len(profs)):
        hit = False
        for j in range(len(profs)):
            if i == j:
                continue
            gs, ops = oracle_port.gotoh_pp(profs[i], profs[j], 1, 1, SC)
            na = ops.count(b"s")
            thr = float(np.float32(np.float32(np.float32(na) * np.float32(0.5)) * np.float32(3)) + np.float32(np.float32(np.float32(na) * np.float32(0.5)) * np.float32(-5)))
            if na / profs[i].shape[1] > 0.1 and na > 25 and gs > thr:
                hit = True
                break
        want.append(hit)
    assert keep == want and want == [True, True, True, True, False]


def test_align_batch_matches_reference_sequence(ctx, oracle_port, oracle_ref):
    """drivers.align_batch = the DP sequence of sage() for FASTA references (src/sage.h:233-260, :311), all traces in three
    batched GPU calls. Checked against the same sequence composed from the CPU oracle one trace at a time (and, where the
    reference build is present, from tracy's own headers with its reverseComplementProfile(one-hot) formulation)."""
    from tracy_b200 import drivers
    rng = np.random.default_rng(2024)
    trimmed, full, refs = [], [], []
    for t in range(12):
        nref = int(rng.integers(400, 1500))
        ref = synth.random_seq(rng, nref, b"ACGTN" if t % 4 == 0 else b"ACGT")
        st, ln = int(rng.integers(0, nref - 300)), int(rng.integers(150, 300))
        core = synth.mutate_seq(rng, ref[st: st + ln].replace(b"N", b"A"), 0.02, 0.01)
        p = synth.profile_from_seq(rng, core, 0.3)
        if t % 2:                                              # trace sequenced from the other strand
            p = np.ascontiguousarray(p[[3, 2, 1, 0, 4, 5], ::-1])
        tl, tr_ = int(rng.integers(0, 20)), int(rng.integers(0, 20))
        full.append(p); trimmed.append(np.ascontiguousarray(p[:, tl: p.shape[1] - tr_])); refs.append(ref)
    got = drivers.align_batch(ctx, trimmed, full, refs, DnaScore(*SC), 50, 50)
    assert any(g["forward"] for g in got) and not all(g["forward"] for g in got)
    for i, g in enumerate(got):
        rc = drivers.reverse_complement_seq(refs[i])
        if oracle_ref is not None:                             # tracy's own formulation of the two orientation scores
            fwdp = oracle_ref.onehot(refs[i])
            gs_f = oracle_ref.gotoh_score(trimmed[i], fwdp, 1, 0, SC)
            gs_r = oracle_ref.gotoh_score(trimmed[i], oracle_ref.revcomp_profile(fwdp), 1, 0, SC)
        else:
            gs_f = oracle_port.gotoh_ps(trimmed[i], refs[i], 1, 0, SC)[0]
            gs_r = oracle_port.gotoh_ps(trimmed[i], rc, 1, 0, SC)[0]
        fw = gs_f > gs_r
        pref = refs[i] if fw else rc
        _, ops = oracle_port.gotoh_ps(trimmed[i], pref, 1, 0, SC)
        r0 = bytes(0x2D if o == ord("h") else 0x58 for o in ops); r1 = bytes(0x2D if o == ord("v") else 0x58 for o in ops)
        if oracle_ref is not None:
            sl, pos = oracle_ref.trim_reference_slice(r0, r1, pref, fw, 0, 50, 50)
        else:
            sl, pos = tracy_b200.trim_reference_slice(r0, r1, pref, fw, 0, 50, 50)
        sc2, ops2 = oracle_port.gotoh_ps(full[i], sl, 1, 0, SC)
        rows = oracle_port.rows_from_ops(full[i], oracle_port.onehot(sl), ops2)
        assert (g["forward"], g["refslice"], g["pos"], g["score"], g["row0"], g["row1"]) == (fw, sl, pos, sc2, rows[0], rows[1]), i


def test_assemble_denovo_recovers_contig(ctx):
    """drivers.assemble_denovo on overlapping traces of one contig, half of them reverse-complemented: the consensus
    equals the planted contig region covered by >= fraction_called of the traces, up to the global strand."""
    from tracy_b200 import drivers
    rng = np.random.default_rng(99)
    contig = synth.random_seq(rng, 700)
    profs = []
    for t in range(10):
        st = 50 * t
        p = synth.profile_from_seq(rng, contig[st: st + 250], 0.25)
        if t % 3 == 1:
            p = np.ascontiguousarray(p[[3, 2, 1, 0, 4, 5], ::-1])
        profs.append(p)
    profs.append(synth.profile_from_seq(rng, synth.random_seq(rng, 200), 0.25))    # a stray trace
    r = drivers.assemble_denovo(ctx, profs, DnaScore(*SC), 0.5, 0.1)
    assert 10 not in r["kept"] and len(r["kept"]) == 10
    cons = r["consensus"]
    assert len(cons) >= 650
    assert cons in contig or drivers.reverse_complement_seq(cons) in contig


def test_basecall_golden(ctx):
    """tb_basecall (device basecall(), reference src/abif.h:408-511) against the reference's calls on synthetic
    chromatograms: positions, primary / secondary / consensus strings, dropped empty windows; then straight into
    tb_create_profile."""
    Gb = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "basecall_golden.npz"))
    n = int(Gb["n"])
    by_ratio = {}
    for i in range(n):
        by_ratio.setdefault(float(Gb[f"ratio{i}"]), []).append(i)
    for ratio, idx in by_ratio.items():
        res = ctx.basecall([Gb[f"tr{i}"] for i in idx], [Gb[f"ploc{i}"] for i in idx], ratio)
        for i, r in zip(idx, res):
            assert np.array_equal(r["bcPos"], Gb[f"pos{i}"]), i
            for k in ("primary", "secondary", "consensus"):
                assert r[k] == bytes(Gb[f"{k}{i}"]), (i, k)
        profs = ctx.create_profile([Gb[f"tr{i}"] for i in idx], [r["bcPos"] for r in res], [r["primary"] for r in res], [r["secondary"] for r in res])
        for p, r in zip(profs, res):
            assert p.shape == (6, len(r["bcPos"])) and np.isfinite(p).all()
