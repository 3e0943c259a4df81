"""GPU parity tests (-m gpu) of the profile x profile kernel (tracy_b200/csrc/gotoh_pp.cu): the fp32x2 substitution score and
the band-pipelined big pairs (one pair spread over many warps), against the oracle. Bit-exact scores and s/h/v strings."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from tracy_b200 import AlignConfig, DnaScore, synth

pytestmark = pytest.mark.gpu
SC = (3, -5, -10, -4)


def msa_profile(rng, seq, depth_max=8):
    """Column-frequency profile of a stack of noisy copies of `seq` with gaps, as _createProfile(char MSA) makes them
    (reference src/align.h:138-180): counts / coverage, gap row filled, N row zero (assemble's rows hold A,C,G,T,'-' only)."""
    m = len(seq)
    code = np.array([b"ACGT".index(bytes([c])) for c in seq], np.int64)
    depth = rng.integers(1, depth_max + 1, m)
    p = np.zeros((6, m), np.float32)
    for j in range(m):
        probs = np.full(6, 0.02)
        probs[4] = 0.0
        probs[5] = 0.04
        probs[code[j]] = 0.0
        probs[code[j]] = 1.0 - probs.sum()
        cnt = rng.multinomial(depth[j], probs)
        p[:, j] = cnt.astype(np.float32) / np.float32(depth[j])
    return p


def related_pair(rng, m, n, style, overlap=0.7):
    """Two profiles over one random genome, overlapping like neighbouring contigs of an assembly (a suffix of a1 matches a
    prefix of a2), with substitutions and indels between them."""
    ov = int(min(m, n) * overlap)
    g = synth.random_seq(rng, m + n - ov + 64)
    s1 = synth.mutate_seq(rng, g[:m + 16], 0.02, 0.01)[:m]
    s2 = synth.mutate_seq(rng, g[m - ov: m - ov + n + 16], 0.02, 0.01)[:n]
    if style == "msa":
        return msa_profile(rng, s1), msa_profile(rng, s2)
    if style == "msaN":      # N mass in some columns: the 25-term sum
        a, b = msa_profile(rng, s1), msa_profile(rng, s2)
        for p in (a, b):
            j = rng.integers(0, p.shape[1], max(1, p.shape[1] // 50))
            p[4, j] = np.float32(0.25)
            p[:4, j] *= np.float32(0.75)
        return a, b
    return synth.profile_from_seq(rng, s1), synth.profile_from_seq(rng, s2)


def oracle_many(port, A, B, hf, vf, sc, threads=16):
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda ab: port.gotoh_pp(ab[0], ab[1], hf, vf, sc), zip(A, B)))


def check(ctx, port, A, B, hf, vf, sc=SC, want_big=None):
    s, ops, ol = ctx.gotoh("pp", A, B, DnaScore(*sc), AlignConfig(bool(hf), bool(vf)))
    nbig = ctx.last_big_pairs()
    so, _, _ = ctx.gotoh("pp", A, B, DnaScore(*sc), AlignConfig(bool(hf), bool(vf)), traceback=False)
    want = oracle_many(port, A, B, hf, vf, sc)
    for i, (ws, wops) in enumerate(want):
        assert int(s[i]) == ws, ("score", i, A[i].shape, B[i].shape)
        assert int(so[i]) == ws, ("score-only", i)
        assert bytes(ops[i, : ol[i]]) == wops, ("ops", i, A[i].shape, B[i].shape)
    if want_big is not None:
        assert nbig == want_big, (nbig, want_big)
    return nbig


def test_small_batch_all_configs(ctx, oracle_port):
    """Whole pairs on one warp each (1 to 3 bands), every end-gap configuration, trace / MSA / N-carrying profiles."""
    rng = np.random.default_rng(301)
    for cfg in range(4):
        hf, vf = cfg & 1, cfg >> 1
        A, B = [], []
        for it in range(40):
            m, n = int(rng.integers(1, 1400)), int(rng.integers(1, 1400))
            a, b = related_pair(rng, m, n, ["trace", "msa", "msaN"][it % 3])
            A.append(a); B.append(b)
        # 40 pairs fill well under half the machine: pairs of three bands or more become big pairs
        check(ctx, oracle_port, A, B, hf, vf)


def test_many_small_pairs_stay_whole(ctx, oracle_port):
    """A batch that fills the machine keeps one warp per pair (no big pairs below sixteen bands)."""
    rng = np.random.default_rng(302)
    base = [related_pair(rng, int(rng.integers(850, 950)), int(rng.integers(850, 950)), "trace") for _ in range(24)]
    reps = (2 * 148 * 12) // len(base) + 1
    A = [p[0] for p in base] * reps
    B = [p[1] for p in base] * reps
    s, _, _ = ctx.gotoh("pp", A, B, DnaScore(*SC), AlignConfig(True, True), traceback=False)
    assert ctx.last_big_pairs() == 0
    assert ctx.last_packed_pairs() == len(A)
    want = oracle_many(oracle_port, [p[0] for p in base], [p[1] for p in base], 1, 1, SC)
    for i in range(len(A)):
        assert int(s[i]) == want[i % len(base)][0]


@pytest.mark.parametrize("cfg", range(4))
def test_big_pairs_vs_oracle(ctx, oracle_port, cfg):
    """One pair over many warps: 2 k x 2 k .. 9 k x 7 k, all end-gap configurations, MSA-style profiles included."""
    rng = np.random.default_rng(310 + cfg)
    hf, vf = cfg & 1, cfg >> 1
    shapes = [(2048, 2048), (2100, 3000), (3500, 1700), (5000, 5200), (9000, 7000), (1537, 4100)]
    A, B = [], []
    for k, (m, n) in enumerate(shapes):
        a, b = related_pair(rng, m, n, ["msa", "trace", "msaN"][k % 3])
        A.append(a); B.append(b)
    nbig = check(ctx, oracle_port, A, B, hf, vf)
    assert nbig == len(shapes)


def test_big_and_small_mixed(ctx, oracle_port):
    """A progressive-alignment level: a few large merges next to many small ones in one call."""
    rng = np.random.default_rng(320)
    A, B = [], []
    for k in range(60):
        m, n = (int(rng.integers(3000, 6000)), int(rng.integers(3000, 6000))) if k % 15 == 0 else (int(rng.integers(200, 1200)), int(rng.integers(200, 1200)))
        a, b = related_pair(rng, m, n, "msa" if k % 2 else "trace")
        A.append(a); B.append(b)
    nbig = check(ctx, oracle_port, A, B, 1, 1)
    assert nbig >= 4


def test_huge_pair(ctx, oracle_port):
    """The shape of palign's last merges (reference src/msa.h:116): tens of thousands of columns on both sides."""
    rng = np.random.default_rng(330)
    a, b = related_pair(rng, 24000, 20000, "msa", overlap=0.5)
    check(ctx, oracle_port, [a], [b], 1, 1, want_big=1)


def test_pp_variants_agree(ctx, oracle_port, monkeypatch):
    """The select variant of the kernel and the general int32 kernel give the same answers as the default (register arrays)."""
    rng = np.random.default_rng(340)
    A, B = [], []
    for k in range(24):
        a, b = related_pair(rng, int(rng.integers(300, 2600)), int(rng.integers(300, 2600)), ["trace", "msa", "msaN"][k % 3])
        A.append(a); B.append(b)
    s0, o0, l0 = ctx.gotoh("pp", A, B, DnaScore(*SC), AlignConfig(True, True))
    for env in ("TRACY_B200_PP_VARIANT", "TRACY_B200_NO_PPFAST"):
        monkeypatch.setenv(env, "sel" if env.endswith("VARIANT") else "1")
        s1, o1, l1 = ctx.gotoh("pp", A, B, DnaScore(*SC), AlignConfig(True, True))
        monkeypatch.delenv(env)
        assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
        for i in range(len(A)):
            assert bytes(o0[i, : l0[i]]) == bytes(o1[i, : l1[i]])


# ---- the screened substitution score (gotoh_pp.cu header, 3.; the rule itself is attacked on the CPU in test_pp_screen.py) ----
def family_profile(rng, kind, n, with_n=False):
    from test_pp_screen import family
    p = np.zeros((6, n), np.float32)
    p[:5] = family(rng, kind, 5, n) if with_n else np.concatenate([family(rng, kind, 4, n), np.zeros((1, n), np.float32)])
    return p


@pytest.mark.parametrize("sc", [SC, (5, -4, -10, -1), (1, -1, -2, -1), (300, -500, -600, -100)])
def test_screen_value_families_vs_oracle(ctx, oracle_port, sc):
    """Columns whose float sums land ON integers (one-hot, dyadic), next to them (k/d fractions) or anywhere (blends), on both
    sides and mixed inside one profile: the short form, the literal fallback and the switch between them."""
    rng = np.random.default_rng(350 + sc[0])
    kinds = ["onehot", "dyadic", "fractions", "blend", "uniform", "tiny"]
    A, B = [], []
    for ka in kinds:
        for kb in kinds:
            m, n = int(rng.integers(120, 330)), int(rng.integers(120, 330))
            A.append(family_profile(rng, ka, m)); B.append(family_profile(rng, kb, n))
    for k in range(6):                                           # families interleaved column by column, some with N mass
        m, n = int(rng.integers(200, 700)), int(rng.integers(200, 700))
        a = np.concatenate([family_profile(rng, kinds[(k + i) % 5], 40, with_n=k % 2 == 1) for i in range(m // 40 + 1)], 1)[:, :m]
        b = np.concatenate([family_profile(rng, kinds[(k + 2 * i) % 5], 64) for i in range(n // 64 + 1)], 1)[:, :n]
        A.append(np.ascontiguousarray(a)); B.append(np.ascontiguousarray(b))
    check(ctx, oracle_port, A, B, 1, 1, sc=sc)
    check(ctx, oracle_port, A[:12], B[:12], 1, 0, sc=sc)


def test_screen_on_off_agree_at_scale(ctx, monkeypatch):
    """6 000 pairs of trace-like and MSA-like profiles: the screened kernel and the literal-only kernel (TRACY_B200_PP_SCREEN=0)
    return the same scores and the same s/h/v strings."""
    rng = np.random.default_rng(360)
    A, B = [], []
    for k in range(6000):
        a, b = related_pair(rng, int(rng.integers(150, 420)), int(rng.integers(150, 420)), ["trace", "trace", "msa", "msaN"][k % 4])
        A.append(a); B.append(b)
    s1, o1, l1 = ctx.gotoh("pp", A, B, DnaScore(*SC), AlignConfig(True, True))
    monkeypatch.setenv("TRACY_B200_PP_SCREEN", "0")
    s0, o0, l0 = ctx.gotoh("pp", A, B, DnaScore(*SC), AlignConfig(True, True))
    monkeypatch.delenv("TRACY_B200_PP_SCREEN")
    assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
    w = int(l0.max())
    mask = np.arange(w)[None, :] < l0[:, None]
    assert np.array_equal(np.where(mask, o0[:, :w], 0), np.where(mask, o1[:, :w], 0))
