"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the golden vectors (produced by the
reference's own headers), against the oracle on seeded fuzz, and -- at BASELINE.json's full shapes -- through
size-independent properties. Bit-exact: scores, s/h/v strings and gapped rows."""
import numpy as np
import pytest

import tracy_b200
from conftest import load_gotoh_golden
from oracle import loader
from tracy_b200 import AlignConfig, DnaScore, synth

pytestmark = pytest.mark.gpu
GOLD = load_gotoh_golden()


def _run_one(ctx, kind, a, b, hf, vf, sc):
    s, ops, ol = ctx.gotoh(kind, [a], [b], DnaScore(*sc), AlignConfig(bool(hf), bool(vf)), traceback=True)
    s2, _, _ = ctx.gotoh(kind, [a], [b], DnaScore(*sc), AlignConfig(bool(hf), bool(vf)), traceback=False)
    assert s2[0] == s[0], "score-only kernel disagrees with the traceback kernel"
    return int(s[0]), bytes(ops[0, : ol[0]])


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_golden(ctx, idx):
    c = GOLD[idx]
    s, ops = _run_one(ctx, c["kind"], c["a"], c["b"], c["hf"], c["vf"], c["sc"])
    assert s == c["score"]
    assert ops == loader.ops_from_rows(c["row0"], c["row1"])
    assert tracy_b200.rows_from_ops(c["kind"], c["a"], c["b"], ops) == (c["row0"], c["row1"])


def test_golden_as_one_ragged_batch(ctx):
    """All golden ps cases with equal scoring in ONE call: ragged shapes, work queue, scratch sizing."""
    for kind in ("ps", "pp", "ss"):
        groups = {}
        for c in GOLD:
            if c["kind"] == kind:
                groups.setdefault((c["hf"], c["vf"], c["sc"]), []).append(c)
        for (hf, vf, sc), cs in groups.items():
            s, ops, ol = ctx.gotoh(kind, [c["a"] for c in cs], [c["b"] for c in cs], DnaScore(*sc), AlignConfig(bool(hf), bool(vf)))
            for i, c in enumerate(cs):
                assert s[i] == c["score"]
                assert bytes(ops[i, : ol[i]]) == loader.ops_from_rows(c["row0"], c["row1"])


@pytest.mark.parametrize("kind", ["ps", "pp", "ss"])
def test_fuzz_vs_oracle(ctx, oracle_port, kind):
    rng = np.random.default_rng({"ps": 11, "pp": 12, "ss": 13}[kind])
    scs = [(3, -5, -10, -4), (5, -4, -10, -1), (1, -1, -2, -1), (4, -4, -6, 0), (2, -7, 0, -3)]
    for rnd in range(8):
        hf, vf = rnd & 1, (rnd >> 1) & 1
        sc = scs[rnd % len(scs)]
        A, Bs = [], []
        for it in range(96):
            big = it % 16 == 0
            m = int(rng.integers(1, 1300 if big else 200))
            n = int(rng.integers(1, 700 if big else 260))
            if kind == "ss":
                a = synth.random_seq(rng, m, b"ACGTN")
                b = synth.mutate_seq(rng, (a * (n // m + 1))[:n], 0.1, 0.1) or b"A"
            else:
                a = synth.random_profile(rng, m, ["trace", "ties", "msa"][it % 3])
                b = synth.random_seq(rng, n, b"ACGTNn-acgtRY" if it % 4 == 0 else b"ACGT") if kind == "ps" else \
                    synth.random_profile(rng, n, ["ties", "msa", "trace"][it % 3])
            A.append(a); Bs.append(b)
        s, ops, ol = ctx.gotoh(kind, A, Bs, DnaScore(*sc), AlignConfig(bool(hf), bool(vf)))
        so, _, _ = ctx.gotoh(kind, A, Bs, DnaScore(*sc), AlignConfig(bool(hf), bool(vf)), traceback=False)
        fn = {"ps": oracle_port.gotoh_ps, "pp": oracle_port.gotoh_pp, "ss": oracle_port.gotoh_ss}[kind]
        for i in range(len(A)):
            ws, wops = fn(A[i], Bs[i], hf, vf, sc)
            assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops), (kind, rnd, i, len(wops))
            assert int(so[i]) == ws


def test_empty_and_degenerate(ctx, oracle_port):
    sc = DnaScore(3, -5, -10, -4)
    s, ops, ol = ctx.gotoh("ss", [], [], sc, AlignConfig(True, False))
    assert len(s) == 0
    cases = [(b"", b"ACG"), (b"AC", b""), (b"", b""), (b"A", b"A"), (b"A", b"C")]
    for hf in (0, 1):
        for vf in (0, 1):
            s, ops, ol = ctx.gotoh("ss", [a for a, _ in cases], [b for _, b in cases], sc, AlignConfig(bool(hf), bool(vf)))
            for i, (a, b) in enumerate(cases):
                assert (int(s[i]), bytes(ops[i, : ol[i]])) == oracle_port.gotoh_ss(a, b, hf, vf, tuple([3, -5, -10, -4]))


def test_invalid_arguments(ctx):
    with pytest.raises(tracy_b200.TracyError) as e:
        ctx.gotoh("ss", [b"ACGT" * 10], [b"ACGT" * 10], DnaScore(300000, -5, -10, -4))
    assert e.value.code == 4   # TB_ERR_UNSUPPORTED: would collide with the 1e6 sentinel
    with pytest.raises(ValueError):
        ctx.gotoh("ss", [b"A"], [b"A", b"C"])


def test_config2_shape_properties(ctx, oracle_port):
    """1 kb x 4 kb, 3/-5/-10/-4, <true,false>: 512 pairs on the GPU; 6 checked exhaustively against the oracle, all
    checked through invariants: the ops string consumes exactly m rows and n columns, re-scoring the alignment with
    the reference's scoring rules reproduces the reported score, and a second run is bit-identical."""
    m, n, N = 1000, 4000, 512
    prof, win = synth.align_batch(N, m, n, seed=44)
    sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
    a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
    s, ops, ol = ctx.gotoh("ps", a1, a2, sc, ac)
    s2, ops2, ol2 = ctx.gotoh("ps", a1, a2, sc, ac)
    assert np.array_equal(s, s2) and np.array_equal(ol, ol2) and np.array_equal(ops, ops2)
    so, _, _ = ctx.gotoh("ps", a1, a2, sc, ac, traceback=False)
    assert np.array_equal(s, so)
    for i in (0, 1, 2, 255, 300, 511):
        ws, wops = oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, (3, -5, -10, -4))
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops)
    fm, fmm = np.float32(3), np.float32(-5)
    for i in range(N):
        o = ops[i, : ol[i]]
        assert (o != ord("h")).sum() == m and (o != ord("v")).sum() == n
        # re-score: substitution scores from the exact float rule, affine gaps, free gaps in first/last row
        r = c = 0
        total = 0
        prev = None
        code = {65: 0, 67: 1, 71: 2, 84: 3}
        for x in o.tobytes():
            if x == 115:
                acc = np.float32(0)
                b = code[win[i, c]]
                for k in range(5):
                    acc = np.float32(acc + np.float32(prof[i, k, r] * (fm if k == b else fmm)))
                total += int(acc)
                r += 1; c += 1
            elif x == 104:
                free = r == 0 or r == m
                if not free:
                    total += -4 + (-10 if prev != 104 else 0)
                c += 1
            else:
                total += -4 + (-10 if prev != 118 else 0)
                r += 1
            prev = x
        assert total == int(s[i]), i


def test_multiband_tall_pair(ctx, oracle_port):
    """a1 longer than one 512/1024-row band, a2 short and long: exercises the band boundary rows."""
    rng = np.random.default_rng(5)
    for m, n in ((1500, 300), (2100, 64), (1025, 1025), (513, 2000)):
        a = synth.random_profile(rng, m, "trace")
        cons = bytes(b"ACGT"[int(k)] for k in np.argmax(a[:4], axis=0))
        b = synth.mutate_seq(rng, (cons * (n // m + 1))[:n], 0.05, 0.03) or b"A"
        for hf, vf in ((1, 0), (1, 1), (0, 0)):
            s, ops, ol = ctx.gotoh("ps", [a], [b], DnaScore(3, -5, -10, -4), AlignConfig(bool(hf), bool(vf)))
            assert (int(s[0]), bytes(ops[0, : ol[0]])) == oracle_port.gotoh_ps(a, b, hf, vf, (3, -5, -10, -4)), (m, n, hf, vf)


def test_device_resident_call(ctx, oracle_port):
    """TB_MEM_DEVICE: every pointer is an HBM pointer (torch is only the allocator here)."""
    import torch
    m, n, N = 300, 700, 64
    prof, win = synth.align_batch(N, m, n, seed=7)
    dev = torch.device("cuda", 0)
    tp = torch.from_numpy(prof).to(dev); tw = torch.from_numpy(win).to(dev)
    aoff = (torch.arange(N, dtype=torch.int64) * 6 * m).to(dev); boff = (torch.arange(N, dtype=torch.int64) * n).to(dev)
    alen = torch.full((N,), m, dtype=torch.int32, device=dev); blen = torch.full((N,), n, dtype=torch.int32, device=dev)
    scores = torch.zeros(N, dtype=torch.int32, device=dev)
    stride = 1008
    ops = torch.zeros((N, stride), dtype=torch.uint8, device=dev); ol = torch.zeros(N, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.gotoh_device("ps", tp.data_ptr(), aoff.data_ptr(), alen.data_ptr(), tw.data_ptr(), boff.data_ptr(), blen.data_ptr(), N,
                     scores.data_ptr(), ops.data_ptr(), stride, ol.data_ptr())
    s, o, l = scores.cpu().numpy(), ops.cpu().numpy(), ol.cpu().numpy()
    for i in range(0, N, 7):
        assert (int(s[i]), bytes(o[i, : l[i]])) == oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, (3, -5, -10, -4))
    ms = ctx.last_kernel_ms()
    assert ms["packed_ms"] + ms["general_ms"] > 0


def test_single_pair_mirrors(ctx, oracle_port):
    """gotohScore()/gotoh() with the reference's call shape (src/gotoh.h:12-14, :71-73)."""
    rng = np.random.default_rng(21)
    p = synth.random_profile(rng, 80, "trace")
    seq = synth.random_seq(rng, 200)
    ac, sc = AlignConfig(True, False), DnaScore(3, -5, -10, -4)
    ws, wops = oracle_port.gotoh_ps(p, seq, 1, 0, (3, -5, -10, -4))
    assert tracy_b200.gotohScore(p, seq, ac, sc, ctx=ctx) == ws
    s, rows = tracy_b200.gotoh(p, seq, ac, sc, ctx=ctx)
    assert s == ws and rows == oracle_port.rows_from_ops(p, oracle_port.onehot(seq), wops)
    s, rows = tracy_b200.gotoh("ACGTACGT", "ACGACGT", AlignConfig(), DnaScore(), ctx=ctx)
    ws, wops = oracle_port.gotoh_ss(b"ACGTACGT", b"ACGACGT", 0, 0, (5, -4, -10, -1))
    assert s == ws and rows == oracle_port.rows_from_ops(b"ACGTACGT", b"ACGACGT", wops)


def test_pinned_arrays_from_the_context(ctx, oracle_port):
    """Context.pinned_empty (tb_host_alloc / tb_host_free): batches and results in page-locked memory give the same answers; the block
    is handed back when the last view goes."""
    import gc
    prof, win = synth.align_batch(64, 120, 300, seed=3)
    pp, pw = ctx.pinned_empty(prof.shape, np.float32), ctx.pinned_empty(win.shape, np.uint8)
    pp[:] = prof
    pw[:] = win
    s_out = ctx.pinned_empty((64,), np.int32)
    want = ctx.gotoh("ps", tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win), DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    got = ctx.gotoh("ps", tracy_b200.uniform_profiles(pp), tracy_b200.uniform_seqs(pw), DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[2], got[2])
    s2 = ctx.gotoh("ps", tracy_b200.uniform_profiles(pp), tracy_b200.uniform_seqs(pw), DnaScore(3, -5, -10, -4), AlignConfig(True, False), traceback=False,
                   out=(s_out, None, None))[0]
    assert s2 is s_out and np.array_equal(s_out, want[0])
    ws, _ = oracle_port.gotoh_ps(prof[5], bytes(win[5]), 1, 0, (3, -5, -10, -4))
    assert int(s_out[5]) == ws
    view = pp[3]
    del pp
    gc.collect()
    assert float(view.sum()) == float(prof[3].sum())                  # the block lives as long as a view does
    del view, pw, s_out, s2
    gc.collect()
