"""GPU parity tests of the packed 16x2 kernel (gotoh_packed.cu): bit-exact against the oracle and against the general
int32 kernel (TRACY_B200_NO_PACKED=1 disables the packed launch), over the shapes that stress its layout: one and two
half-bands, several 1024-row passes, windows shorter than the 32-column half-band lag, all end-gap configurations,
profiles with exact ties and MSA-style columns, reference strings with N / lower case / foreign characters, and
pairs that must fall through to the general kernel (score range too wide for 15-bit fields)."""
import os

import numpy as np
import pytest

import tracy_b200
from oracle import loader
from tracy_b200 import AlignConfig, DnaScore, synth

pytestmark = pytest.mark.gpu


def _both(ctx, A, Bs, sc, ac, traceback=True):
    os.environ.pop("TRACY_B200_NO_PACKED", None)
    r1 = ctx.gotoh("ps", A, Bs, DnaScore(*sc), ac, traceback=traceback)
    packed = ctx.last_packed_pairs()
    os.environ["TRACY_B200_NO_PACKED"] = "1"
    try:
        r2 = ctx.gotoh("ps", A, Bs, DnaScore(*sc), ac, traceback=traceback)
        assert ctx.last_packed_pairs() == 0
    finally:
        os.environ.pop("TRACY_B200_NO_PACKED", None)
    assert np.array_equal(r1[0], r2[0])
    if traceback:
        assert np.array_equal(r1[2], r2[2])
        for i in range(len(A)):
            assert bytes(r1[1][i, : r1[2][i]]) == bytes(r2[1][i, : r2[2][i]]), i
    return r1, packed


SHAPES = [(1, 1), (1, 40), (5, 3), (16, 16), (17, 31), (33, 32), (100, 33), (511, 64), (512, 200), (513, 90), (600, 1),
          (1000, 31), (1024, 300), (1025, 77), (1500, 260), (2049, 40), (300, 1500), (1000, 1300)]


@pytest.mark.parametrize("hf,vf", [(1, 0), (0, 0), (1, 1), (0, 1)])
def test_packed_shapes_vs_oracle(ctx, oracle_port, hf, vf):
    rng = np.random.default_rng(100 + 2 * hf + vf)
    sc = (3, -5, -10, -4)
    A, Bs = [], []
    for k, (m, n) in enumerate(SHAPES):
        a = synth.random_profile(rng, m, ["trace", "ties", "msa"][k % 3])
        cons = bytes(b"ACGT"[int(x)] for x in np.argmax(a[:4], axis=0))
        b = synth.mutate_seq(rng, (cons * (n // m + 1))[:n], 0.05, 0.04) if k % 2 == 0 else synth.random_seq(rng, n, b"ACGTNacgtn")
        b = (b or b"A")
        A.append(a); Bs.append(b)
    (s, ops, ol), packed = _both(ctx, A, Bs, sc, AlignConfig(bool(hf), bool(vf)))
    assert packed == len(A), "every one of these pairs is inside the packed kernel's exact range"
    for i in range(len(A)):
        ws, wops = oracle_port.gotoh_ps(A[i], Bs[i], hf, vf, sc)
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops), (SHAPES[i], hf, vf)
    (s2, _, _), _ = _both(ctx, A, Bs, sc, AlignConfig(bool(hf), bool(vf)), traceback=False)
    assert np.array_equal(s, s2)


@pytest.mark.parametrize("sc", [(5, -4, -10, -1), (1, -1, -2, -1), (4, -4, -6, 0), (2, -7, 0, -3), (3, -5, -10, -4), (0, 0, 0, 0), (-1, -2, -3, -1)])
def test_packed_scores_fuzz(ctx, oracle_port, sc):
    rng = np.random.default_rng(abs(hash(sc)) % 1000)
    A, Bs = [], []
    for it in range(48):
        m = int(rng.integers(1, 1200 if it % 8 == 0 else 150))
        n = int(rng.integers(1, 900 if it % 8 == 0 else 200))
        A.append(synth.random_profile(rng, m, ["ties", "trace", "msa"][it % 3]))
        Bs.append(synth.random_seq(rng, n, b"ACGTN" if it % 5 == 0 else b"ACGT"))
    for hf, vf in ((1, 0), (1, 1)):
        (s, ops, ol), packed = _both(ctx, A, Bs, sc, AlignConfig(bool(hf), bool(vf)))
        assert packed > 0
        for i in range(len(A)):
            ws, wops = oracle_port.gotoh_ps(A[i], Bs[i], hf, vf, sc)
            assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops), (sc, i)


def test_packed_falls_through_when_range_too_wide(ctx, oracle_port):
    """Unnormalised profile values (|sub| large) and big gap scores do not fit 15-bit fields: the packed kernel must
    decline those pairs and the general kernel must finish them, in the same call."""
    rng = np.random.default_rng(9)
    A, Bs = [], []
    for it in range(12):
        a = synth.random_profile(rng, 200, "trace")
        if it % 2 == 0:
            a = a * np.float32(40.0)       # substitution scores up to +-200 -> 200 rows x 200 overflows 15 bits
        A.append(a); Bs.append(synth.random_seq(rng, 300))
    sc = (3, -5, -10, -4)
    (s, ops, ol), packed = _both(ctx, A, Bs, sc, AlignConfig(True, False))
    assert 0 < packed < len(A)
    for i in range(len(A)):
        ws, wops = oracle_port.gotoh_ps(A[i], Bs[i], 1, 0, sc)
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops), i
    # gap scores beyond the packed kernel's host-side limits: whole batch on the general kernel
    sc2 = (3, -5, -700, -4)
    s2, ops2, ol2 = ctx.gotoh("ps", A[1::2], Bs[1::2], DnaScore(*sc2), AlignConfig(True, False))
    assert ctx.last_packed_pairs() == 0
    for i, (a, b) in enumerate(zip(A[1::2], Bs[1::2])):
        assert (int(s2[i]), bytes(ops2[i, : ol2[i]])) == oracle_port.gotoh_ps(a, b, 1, 0, sc2)


def test_foreign_reference_characters_fall_through(ctx, oracle_port):
    """Reference characters outside ACGTN score 0 against every channel (src/align.h:121-136): the packed kernel's
    tables carry five classes only, so such pairs go to the general kernel; mixed batches stay correct."""
    rng = np.random.default_rng(31)
    A = [synth.random_profile(rng, 150 + 10 * i, "trace") for i in range(10)]
    Bs = [synth.random_seq(rng, 400, b"ACGT-RYK" if i % 2 else b"ACGTN") for i in range(10)]
    (s, ops, ol), packed = _both(ctx, A, Bs, (3, -5, -10, -4), AlignConfig(True, False))
    assert packed == 5
    for i in range(10):
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == oracle_port.gotoh_ps(A[i], Bs[i], 1, 0, (3, -5, -10, -4)), i


def test_checkpoint_and_flag_tracebacks_agree(ctx, oracle_port):
    """The packed kernel has two traceback implementations: checkpoints + tile recompute (default) and pointer flags for
    every cell (TRACY_B200_TB_MODE=flags). Same strings on shapes that stress the round planner: long free end-gap runs
    (horizontal rounds), long internal gaps (speculation misses), several passes, tiny windows, every end-gap config."""
    rng = np.random.default_rng(77)
    A, Bs = [], []
    for it in range(40):
        m = int(rng.integers(1, 2300 if it % 10 == 0 else 700))
        core = synth.random_seq(rng, m)
        a = synth.profile_from_seq(rng, core, 0.35)
        kind = it % 5
        if kind == 0:      # trace inside a long window: long horizontal runs on both ends
            b = synth.random_seq(rng, int(rng.integers(0, 2500))) + synth.mutate_seq(rng, core, 0.02, 0.01) + synth.random_seq(rng, int(rng.integers(0, 2500)))
        elif kind == 1:    # a big deletion / insertion in the middle: the diagonal speculation must miss and re-plan
            cut = m // 2
            b = synth.mutate_seq(rng, core[:cut], 0.02, 0.01) + synth.random_seq(rng, int(rng.integers(40, 400))) + synth.mutate_seq(rng, core[cut:], 0.02, 0.01)
        elif kind == 2:
            b = synth.mutate_seq(rng, core[: m // 3] + core[2 * m // 3:], 0.02, 0.01) or b"A"
        elif kind == 3:
            b = synth.random_seq(rng, int(rng.integers(1, 70)))
        else:
            b = synth.mutate_seq(rng, core, 0.1, 0.08) or b"A"
        A.append(a); Bs.append(b or b"A")
    for hf, vf in ((1, 0), (0, 0), (1, 1), (0, 1)):
        ac = AlignConfig(bool(hf), bool(vf))
        os.environ.pop("TRACY_B200_TB_MODE", None)
        s1, o1, l1 = ctx.gotoh("ps", A, Bs, DnaScore(3, -5, -10, -4), ac)
        packed = ctx.last_packed_pairs()
        assert packed >= len(A) - 4          # the longest traces against the longest windows exceed the 15-bit range
        os.environ["TRACY_B200_TB_MODE"] = "flags"
        try:
            s2, o2, l2 = ctx.gotoh("ps", A, Bs, DnaScore(3, -5, -10, -4), ac)
            assert ctx.last_packed_pairs() == packed
        finally:
            os.environ.pop("TRACY_B200_TB_MODE", None)
        assert np.array_equal(s1, s2) and np.array_equal(l1, l2)
        for i in range(len(A)):
            assert bytes(o1[i, : l1[i]]) == bytes(o2[i, : l2[i]]), (hf, vf, i)
        for i in range(0, len(A), 3):
            assert (int(s1[i]), bytes(o1[i, : l1[i]])) == oracle_port.gotoh_ps(A[i], Bs[i], hf, vf, (3, -5, -10, -4)), (hf, vf, i)


def test_packed_config2_full_shape(ctx, oracle_port):
    """BASELINE configs[1] shape through the packed kernel: 256 pairs of 1000 x 4000, 8 exhaustively vs the oracle."""
    m, n, N = 1000, 4000, 256
    prof, win = synth.align_batch(N, m, n, seed=45)
    a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
    s, ops, ol = ctx.gotoh("ps", a1, a2, DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    assert ctx.last_packed_pairs() == N
    for i in range(0, N, 32):
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, (3, -5, -10, -4)), i
    os.environ["TRACY_B200_NO_PACKED"] = "1"
    try:
        s2, ops2, ol2 = ctx.gotoh("ps", a1, a2, DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    finally:
        os.environ.pop("TRACY_B200_NO_PACKED", None)
    assert np.array_equal(s, s2) and np.array_equal(ol, ol2) and np.array_equal(ops, ops2)


def test_host_pipeline_chunks_lanes_and_stage2(ctx, oracle_port, monkeypatch):
    """The TB_MEM_HOST pipeline (chunks over three lanes, one kernel per chunk, stage 2 only for chunks whose windows hold N or
    foreign characters, 5-of-6 profile rows copied): many small chunks give the same results as one chunk and as the oracle."""
    rng = np.random.default_rng(2024)
    N, m, n = 5000, 96, 260
    prof, win = synth.align_batch(N, m, n, seed=9)
    prof = prof.copy(); win = win.copy()
    prof[:, 5, :] = rng.random((N, m), dtype=np.float32)          # the '-' row never enters _score (src/align.h:112-116)
    for i in rng.choice(N, 400, replace=False):
        win[i, rng.integers(0, n, 3)] = ord("N")                   # 5-class packed kernel
    for i in rng.choice(N, 60, replace=False):
        win[i, rng.integers(0, n)] = ord("R")                      # general kernel (scores 0 against everything)
    a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
    sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
    monkeypatch.delenv("TRACY_B200_CHUNK", raising=False)
    s0, o0, l0 = ctx.gotoh("ps", a1, a2, sc, ac)
    for chunk, lanes in (("333", "3"), ("1000", "2"), ("128", "4")):
        monkeypatch.setenv("TRACY_B200_CHUNK", chunk)
        monkeypatch.setenv("TRACY_B200_LANES", lanes)
        s1, o1, l1 = ctx.gotoh("ps", a1, a2, sc, ac)
        assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
        assert all(bytes(o0[i, : l0[i]]) == bytes(o1[i, : l1[i]]) for i in range(0, N, 7))
        s2 = ctx.gotoh("ps", a1, a2, sc, ac, traceback=False)[0]
        assert np.array_equal(s0, s2)
    monkeypatch.delenv("TRACY_B200_CHUNK"); monkeypatch.delenv("TRACY_B200_LANES")
    for i in list(range(0, N, 97)) + [int(x) for x in np.nonzero((win == ord("R")).any(1))[0][:8]]:
        ws, wops = oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, (3, -5, -10, -4))
        assert int(s0[i]) == ws and bytes(o0[i, : l0[i]]) == wops, i


def test_host_pipeline_ramp_schedule(ctx, oracle_port, monkeypatch):
    """A batch large enough for the full chunk schedule of the TB_MEM_HOST pipeline (ramp up 1-2-4 waves, steady chunks in whole waves,
    ramp down 4-2-1 waves with the remainder, inputs copied one chunk at a time): same scores, traceback lengths and rows as the plain
    schedule (TRACY_B200_NO_RAMPDOWN) and as fixed small chunks; pairs at the chunk boundaries against the oracle."""
    N, m, n = 60000, 300, 900
    base_p, base_w = synth.align_batch(2048, m, n, seed=31)
    idx = np.arange(N) % 2048
    idx[1::3] = (idx[1::3] * 7 + 3) % 2048                                 # neighbours differ
    prof, win = np.ascontiguousarray(base_p[idx]), np.ascontiguousarray(base_w[idx])
    a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
    sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
    for k in ("TRACY_B200_CHUNK", "TRACY_B200_LANES", "TRACY_B200_NO_RAMPDOWN"):
        monkeypatch.delenv(k, raising=False)
    s0, o0, l0, r0, r1 = ctx.gotoh("ps", a1, a2, sc, ac, rows=True)
    assert ctx.last_packed_pairs() == N
    monkeypatch.setenv("TRACY_B200_NO_RAMPDOWN", "1")
    s1, o1, l1, q0, q1 = ctx.gotoh("ps", a1, a2, sc, ac, rows=True)
    monkeypatch.delenv("TRACY_B200_NO_RAMPDOWN")
    monkeypatch.setenv("TRACY_B200_CHUNK", "7001")
    s2, o2, l2 = ctx.gotoh("ps", a1, a2, sc, ac)
    monkeypatch.delenv("TRACY_B200_CHUNK")
    assert np.array_equal(s0, s1) and np.array_equal(l0, l1) and np.array_equal(s0, s2) and np.array_equal(l0, l2)
    mask = np.arange(o0.shape[1])[None, :] < l0[:, None]
    assert np.array_equal(o0 * mask, o1 * mask) and np.array_equal(o0 * mask, o2[:, : o0.shape[1]] * mask)
    assert np.array_equal(r0 * mask, q0 * mask) and np.array_equal(r1 * mask, q1 * mask)
    wave = 148 * 12
    edges = sorted({0, N - 1, wave - 1, wave, 3 * wave - 1, 3 * wave, 7 * wave, N - wave, N - 3 * wave - 1, N - 7 * wave, N // 2, N - 500})
    for i in edges:
        ws, wops = oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, (3, -5, -10, -4))
        assert int(s0[i]) == ws and bytes(o0[i, : l0[i]]) == wops, i

def test_string_pairs_through_packed_kernel(ctx, oracle_port):
    """tb_gotoh_ss: upper-case ACGTN pairs run on the packed kernel (a1's characters as one-hot profile columns); lower case,
    IUPAC codes and gaps on either side are byte-compared by the general string kernel. Same results as the reference's
    string overload (src/align.h:96-101) either way."""
    rng = np.random.default_rng(321)
    a1, a2 = [], []
    for i in range(160):
        m, n = int(rng.integers(1, 700)), int(rng.integers(1, 1400))
        alpha1 = [b"ACGT", b"ACGTN", b"ACGTacgt", b"ACGTRY-", b"ACGT"][i % 5]
        alpha2 = [b"ACGT", b"ACGTN", b"ACGT", b"ACGT", b"ACGTNn-K"][i % 5]
        t = synth.random_seq(rng, n, alpha2)
        s = bytearray(t[: min(m, n)]) + bytearray(synth.random_seq(rng, max(0, m - n), alpha1))
        for q in rng.integers(0, len(s), max(1, len(s) // 12)):
            s[q] = alpha1[int(rng.integers(0, len(alpha1)))]
        a1.append(bytes(s)); a2.append(t)
    total_packed = 0
    for hf, vf, sc in ((1, 0, (3, -5, -10, -4)), (0, 0, (3, -5, -10, -4)), (1, 1, (5, -4, -10, -1)), (0, 1, (2, -3, -6, -2))):
        s, ops, ol = ctx.gotoh("ss", a1, a2, DnaScore(*sc), AlignConfig(bool(hf), bool(vf)))
        total_packed += ctx.last_packed_pairs()
        s2 = ctx.gotoh("ss", a1, a2, DnaScore(*sc), AlignConfig(bool(hf), bool(vf)), traceback=False)[0]
        assert np.array_equal(s, s2)
        for i in range(len(a1)):
            ws, wops = oracle_port.gotoh_ss(a1[i], a2[i], hf, vf, sc)
            assert int(s[i]) == ws and bytes(ops[i, : ol[i]]) == wops, (hf, vf, i)
    assert total_packed >= 4 * 60          # the ACGT / ACGTN pairs (2 of 5 alphabets) took the packed kernel


def test_spans_at_the_window_start(ctx, oracle_port):
    """The tile recompute runs a span's two 32-column halves side by side from two column checkpoints; spans that begin left
    of a block's first checkpoint start from column 0 and run all 64 columns in one field. Every start offset 0..71 of the
    trace inside the window (so that the diagonal crosses column 0..71 in the top blocks), windows that end right after the
    trace and windows with a long right flank (horizontal rounds that reach column 0), against the oracle."""
    rng = np.random.default_rng(2025)
    A, Bs = [], []
    for off in range(72):
        m = int(rng.integers(40, 700))
        core = synth.random_seq(rng, m)
        A.append(synth.profile_from_seq(rng, core, 0.3))
        right = int(rng.integers(0, 50)) if off % 3 else int(rng.integers(200, 2200))
        Bs.append(synth.random_seq(rng, off) + synth.mutate_seq(rng, core, 0.02, 0.02) + synth.random_seq(rng, right))
    for hf, vf in ((1, 0), (1, 1), (0, 0)):
        s, ops, ol = ctx.gotoh("ps", A, Bs, DnaScore(3, -5, -10, -4), AlignConfig(bool(hf), bool(vf)))
        assert ctx.last_packed_pairs() == len(A)
        for i in range(len(A)):
            assert (int(s[i]), bytes(ops[i, : ol[i]])) == oracle_port.gotoh_ps(A[i], Bs[i], hf, vf, (3, -5, -10, -4)), (hf, vf, i)


def test_config2_1024_pairs_vs_reference_headers(ctx, oracle_ref, oracle_port):
    """a fixed subset of BASELINE configs[1] (1 kb x 4 kb, scores 3/-5/-10/-4, AlignConfig<true,false>) pair by pair
    against tracy's own gotoh() -- the unmodified reference headers (oracle/_ref) where they travelled with the snapshot, else the C
    restatement -- on all host threads: score and both gapped rows (made on the device) of each of 1 024 pairs."""
    from concurrent.futures import ThreadPoolExecutor
    m, n, N = 1000, 4000, 1024
    prof, win = synth.align_batch(N, m, n, seed=4096)
    a1, a2 = tracy_b200.uniform_profiles(prof, trace_profiles=True), tracy_b200.uniform_seqs(win)
    sc = (3, -5, -10, -4)
    s, ops, ol, r0, r1 = ctx.gotoh("ps", a1, a2, DnaScore(*sc), AlignConfig(True, False), rows=True)
    assert ctx.last_packed_pairs() == N

    def one(i):
        if oracle_ref is not None:
            return oracle_ref.gotoh(prof[i], oracle_ref.onehot(bytes(win[i])), 1, 0, sc)
        ws, wops = oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, sc)
        return (ws,) + tracy_b200.rows_from_ops("ps", prof[i], bytes(win[i]), wops)
    with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
        want = list(ex.map(one, range(N)))
    bad = [i for i in range(N) if (int(s[i]), bytes(r0[i, : ol[i]]), bytes(r1[i, : ol[i]])) != want[i]]
    assert not bad, bad[:10]
    for i in range(0, N, 64):      # the s/h/v strings say the same as the rows
        assert tracy_b200.rows_from_ops("ps", prof[i], bytes(win[i]), bytes(ops[i, : ol[i]])) == (want[i][1], want[i][2])


def test_streamed_batch_equals_chunk_pipeline(ctx, oracle_port, monkeypatch):
    """The streamed form of a large host batch (ONE launch of the packed kernel, inputs gated chunk by chunk, results sent off as the
    chunks finish; capi.cu run_gotoh `streamed`) against the launch-per-chunk pipeline (TRACY_B200_NO_STREAM): scores, lengths,
    plain ops, 2-bit packed ops and both gapped rows; score only; and a batch in which every 97th window carries a character the packed
    kernel leaves to the second stage (which then runs on the resident batch)."""
    N, m, n = 60000, 300, 900
    base_p, base_w = synth.align_batch(2048, m, n, seed=77)
    idx = (np.arange(N) * 13 + 5) % 2048
    prof, win = np.ascontiguousarray(base_p[idx]), np.ascontiguousarray(base_w[idx])
    sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
    for k in ("TRACY_B200_CHUNK", "TRACY_B200_LANES", "TRACY_B200_NO_RAMPDOWN", "TRACY_B200_NO_STREAM"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("TRACY_B200_FORCE_STREAM", "1")      # pageable inputs arrive slowly: without it the context would fall back after the first call

    def both(fn):
        k0 = ctx.stats()["kernel_launches"]
        got = fn()
        k1 = ctx.stats()["kernel_launches"]
        monkeypatch.setenv("TRACY_B200_NO_STREAM", "1")
        want = fn()
        monkeypatch.delenv("TRACY_B200_NO_STREAM")
        assert k1 - k0 <= 3, "the streamed form is one launch (plus a second stage)"
        return got, want

    for variant in range(2):
        if variant == 1:
            win = win.copy()
            win[::97, 450] = ord("R")                                     # IUPAC code: not a packed-kernel pair
        a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
        (s0, o0, l0, r0, r1), (s1, o1, l1, q0, q1) = both(lambda: ctx.gotoh("ps", a1, a2, sc, ac, rows=True))
        assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
        mask = np.arange(o0.shape[1])[None, :] < l0[:, None]
        assert np.array_equal(o0 * mask, o1 * mask) and np.array_equal(r0 * mask, q0 * mask) and np.array_equal(r1 * mask, q1 * mask)
        (s2, p2, l2), (s3, p3, l3) = both(lambda: ctx.gotoh("ps", a1, a2, sc, ac, packed=True))
        assert np.array_equal(s0, s2) and np.array_equal(s0, s3) and np.array_equal(l0, l2) and np.array_equal(l0, l3)
        pm = np.arange(p2.shape[1])[None, :] < ((l0 + 3) // 4)[:, None]
        assert np.array_equal(p2 * pm, p3 * pm)
        (s4, _, _), (s5, _, _) = both(lambda: ctx.gotoh("ps", a1, a2, sc, ac, traceback=False))
        assert np.array_equal(s0, s4) and np.array_equal(s0, s5)
        for i in (0, 1, 97, 1775, 1776, 5328, N // 2, N - 1777, N - 1):
            ws, wops = oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, (3, -5, -10, -4))
            assert int(s0[i]) == ws and bytes(o0[i, : l0[i]]) == wops, (variant, i)


def test_streamed_batch_other_shapes(ctx, monkeypatch):
    """The streamed form on the other packed-kernel inputs: string x string pairs, the 4-row upload of trace profiles
    (TB_A1_TRACE_PROFILES), pinned arenas and result arrays, global alignment (no free end gaps) -- each against the launch-per-chunk
    pipeline."""
    N = 45000
    rng = np.random.default_rng(9)
    sc = DnaScore(3, -5, -10, -4)
    monkeypatch.setenv("TRACY_B200_FORCE_STREAM", "1")

    def both(fn):
        monkeypatch.delenv("TRACY_B200_NO_STREAM", raising=False)
        k0 = ctx.stats()["kernel_launches"]
        got = fn()
        k1 = ctx.stats()["kernel_launches"]
        monkeypatch.setenv("TRACY_B200_NO_STREAM", "1")
        want = fn()
        monkeypatch.delenv("TRACY_B200_NO_STREAM")
        assert k1 - k0 <= 3, "the streamed form is one launch (plus a second stage)"
        return got, want

    def same(got, want):
        (s0, o0, l0), (s1, o1, l1) = got[:3], want[:3]
        assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
        mask = np.arange(o0.shape[1])[None, :] < l0[:, None]
        assert np.array_equal(o0 * mask, o1 * mask)
        for r, q in zip(got[3:], want[3:]):
            assert np.array_equal(r * mask, q * mask)

    # string x string
    s1 = rng.choice(np.frombuffer(b"ACGT", np.uint8), (N, 40))
    s2 = rng.choice(np.frombuffer(b"ACGTN", np.uint8), (N, 70))
    a1, a2 = tracy_b200.uniform_seqs(s1), tracy_b200.uniform_seqs(s2)
    same(*both(lambda: ctx.gotoh("ss", a1, a2, sc, AlignConfig(False, False), rows=True)))
    assert ctx.last_packed_pairs() == N
    # trace profiles (rows 4, 5 zero): 4-row upload, pinned buffers on both sides, free horizontal end gaps
    base_p, base_w = synth.align_batch(1024, 60, 150, seed=3)
    idx = (np.arange(N) * 5 + 2) % 1024
    prof = ctx.pinned_empty((N, 6, 60), np.float32); prof[:] = base_p[idx]
    win = ctx.pinned_empty((N, 150), np.uint8); win[:] = base_w[idx]
    assert not prof[:, 4:].any()
    p1, p2 = tracy_b200.uniform_profiles(prof, trace_profiles=True), tracy_b200.uniform_seqs(win)
    same(*both(lambda: ctx.gotoh("ps", p1, p2, sc, AlignConfig(True, False), rows=True)))
    same(*both(lambda: ctx.gotoh("ps", p1, p2, sc, AlignConfig(False, False))))
