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
