"""GPU parity tests (-m gpu) that sit ON the exactness bound of the packed 16x2 kernel (tracy_b200/csrc/gotoh_packed.cu, the
per-pair range check): pairs whose bias + upper bound lands a few units either side of the largest field value the kernel
allows, with contents that drive the DP to both extremes, in both traceback modes and every end-gap configuration -- each pair is
checked against the int32 oracle and against which kernel took it. Plus the shapes that used to fall off the packed path:
windows of tens of thousands of columns with free horizontal end gaps (a 50 kb FASTA reference, reference src/sage.h:227-230)."""
import numpy as np
import pytest

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth

pytestmark = pytest.mark.gpu

K_NEG, K_MAX, K_ROWS = 2048, 0x7BFF - 16, 1024        # kPkNeg, kPkMaxField, kPkRows


def headroom(m, n, sc, hfree, smin, smax):
    """kPkMaxField - (bias + ub) exactly as the kernel computes it: >= 0 -> the packed kernel takes the pair."""
    ma, mi, go, ge = sc
    goe = go + ge
    npass = (m + K_ROWS - 1) // K_ROWS
    far = npass * K_ROWS + 100 + (0 if hfree else n)
    lb = 2 * go + 2 * goe + far * ge - 16 + 72 * goe + 72 * min(smin, 0)
    bias = K_NEG + 64 - lb
    ub = max(smax, 0) * min(m, n) + 72 * max(smax, 0)
    return K_MAX - (bias + ub)


def onehot_profile(seq):
    p = np.zeros((6, len(seq)), np.float32)
    p[[b"ACGT".index(bytes([c])) for c in seq], np.arange(len(seq))] = 1.0
    return p


def shapes_at_offsets(sc, hfree, offsets):
    """(m, n) with headroom exactly `off` for each requested offset (one-hot traces: smin = mismatch, smax = match)."""
    out = {}
    for m in range(40, 400):
        for n in range(m, 9000):
            h = headroom(m, n, sc, hfree, sc[1], sc[0])
            if h in offsets and h not in out:
                out[h] = (m, n)
            if h < min(offsets) - 8:
                break
        if len(out) == len(offsets):
            break
    return out


@pytest.mark.parametrize("vfree", [0, 1])
def test_bound_straddled_global_rows(ctx, oracle_port, vfree, monkeypatch):
    """AlignConfig<false, V>: the bias grows with every column, so a sweep over n walks through the bound. Contents: all-match
    (fields climb to bias + match * min(m, n)), all-mismatch and a trace against an unrelated window (fields sink towards the
    all-gap corner values)."""
    sc = (3, -5, -10, -4)
    want = [16, 2, 1, 0, -1, -2]
    shp = shapes_at_offsets(sc, False, want)
    assert sorted(shp) == sorted(want), shp
    rng = np.random.default_rng(90 + vfree)
    for mode_env in (None, "flags"):
        if mode_env:
            monkeypatch.setenv("TRACY_B200_TB_MODE", mode_env)
        for off, (m, n) in sorted(shp.items()):
            seq = synth.random_seq(rng, n)
            cases = [(onehot_profile(b"A" * m), b"A" * n), (onehot_profile(b"A" * m), b"C" * n),
                     (onehot_profile(seq[:m]), seq), (onehot_profile(synth.random_seq(rng, m)), seq)]
            A, B = [c[0] for c in cases], [c[1] for c in cases]
            s, ops, ol = ctx.gotoh("ps", A, B, DnaScore(*sc), AlignConfig(False, bool(vfree)))
            took = ctx.last_packed_pairs()
            so, _, _ = ctx.gotoh("ps", A, B, DnaScore(*sc), AlignConfig(False, bool(vfree)), traceback=False)
            assert took == (len(A) if off >= 0 else 0), (off, m, n, took)
            for i in range(len(A)):
                ws, wops = oracle_port.gotoh_ps(A[i], B[i], 0, vfree, sc)
                assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops), (off, m, n, i, mode_env)
                assert int(so[i]) == ws
        if mode_env:
            monkeypatch.delenv("TRACY_B200_TB_MODE")


def test_bound_straddled_by_scores(ctx, oracle_port):
    """AlignConfig<true, false> (tracy align / decompose): the bias no longer depends on the window, so the bound is reached through
    the score values instead: match scores large enough that match * min(m, n) climbs to the limit."""
    rng = np.random.default_rng(95)
    m, n = 200, 900
    hits = {}
    for match in range(60, 160):                    # the largest match score the packed kernel still takes, and the next one
        sc = (match, -5, -10, -4)
        if headroom(m, n, sc, True, -5, match) >= 0:
            hits[True] = sc
        elif False not in hits:
            hits[False] = sc
    assert set(hits) == {True, False} and hits[False][0] == hits[True][0] + 1
    for packed, sc in hits.items():
        seq = synth.random_seq(rng, n)
        cases = [(onehot_profile(b"A" * m), b"A" * n), (onehot_profile(seq[300: 300 + m]), seq), (onehot_profile(b"C" * m), b"A" * n)]
        A, B = [c[0] for c in cases], [c[1] for c in cases]
        s, ops, ol = ctx.gotoh("ps", A, B, DnaScore(*sc), AlignConfig(True, False))
        assert ctx.last_packed_pairs() == (len(A) if packed else 0), (sc, ctx.last_packed_pairs())
        for i in range(len(A)):
            assert (int(s[i]), bytes(ops[i, : ol[i]])) == oracle_port.gotoh_ps(A[i], B[i], 1, 0, sc), (sc, i)


@pytest.mark.parametrize("n", [20000, 50000])
def test_long_windows_stay_packed(ctx, oracle_port, n):
    """1000 x 20 000 and 1000 x 50 000 profile x string pairs with free horizontal end gaps: all on the packed kernel, bit-exact."""
    rng = np.random.default_rng(n)
    N, m = 12, 1000
    A, B = [], []
    for i in range(N):
        g = synth.random_seq(rng, n)
        off = int(rng.integers(0, n - m - 10))
        t = synth.mutate_seq(rng, g[off: off + m + 20], 0.02, 0.01)[:m]
        if i % 3 == 2:
            t = t.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]
        A.append(synth.profile_from_seq(rng, t, 0.3)); B.append(g)
    sc = (3, -5, -10, -4)
    s, ops, ol = ctx.gotoh("ps", A, B, DnaScore(*sc), AlignConfig(True, False))
    assert ctx.last_packed_pairs() == N
    k = ctx.last_kernel_ms()
    so, _, _ = ctx.gotoh("ps", A, B, DnaScore(*sc), AlignConfig(True, False), traceback=False)
    assert ctx.last_packed_pairs() == N
    for i in range(0, N, 3 if n > 20000 else 2):
        ws, wops = oracle_port.gotoh_ps(A[i], B[i], 1, 0, sc)
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops), i
        assert int(so[i]) == ws
    assert np.array_equal(s, so)
    print(f"1000 x {n}: {N} pairs, packed kernel {k['packed_ms']:.2f} ms with traceback")
