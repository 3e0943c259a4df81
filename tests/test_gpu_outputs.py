"""GPU parity tests (-m gpu) of the output forms made on the device (tracy_b200/csrc/post_ops.cu): the gapped alignment rows
gotoh() leaves in `align` (reference src/align.h:196-293) and the 2-bit packed s/h/v strings, plus the 4-row profile upload."""
import numpy as np
import pytest

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth

pytestmark = pytest.mark.gpu
SC = (3, -5, -10, -4)


def _batch(rng, kind, n):
    A, B = [], []
    for it in range(n):
        m, k = int(rng.integers(1, 700)), int(rng.integers(1, 900))
        if kind == "ss":
            a = synth.random_seq(rng, m, b"ACGTN" if it % 3 else b"ACGTNacgt-RY")
            b = synth.mutate_seq(rng, (a * (k // m + 1))[:k], 0.1, 0.1) or b"A"
        else:
            a = synth.random_profile(rng, m, ["trace", "ties", "msa"][it % 3])
            b = synth.random_seq(rng, k, b"ACGTNn-acgtRY" if it % 4 == 0 else b"ACGT") if kind == "ps" else synth.random_profile(rng, k, ["ties", "msa", "trace"][it % 3])
        A.append(a); B.append(b)
    return A, B


@pytest.mark.parametrize("kind", ["ps", "pp", "ss"])
def test_rows_and_packed_ops_host_mode(ctx, oracle_port, kind):
    rng = np.random.default_rng({"ps": 51, "pp": 52, "ss": 53}[kind])
    for cfg in range(4):
        hf, vf = cfg & 1, cfg >> 1
        A, B = _batch(rng, kind, 48)
        s, ops, ol = ctx.gotoh(kind, A, B, DnaScore(*SC), AlignConfig(bool(hf), bool(vf)))
        s2, pk, ol2, r0, r1 = ctx.gotoh(kind, A, B, DnaScore(*SC), AlignConfig(bool(hf), bool(vf)), rows=True, packed=True)
        assert np.array_equal(s, s2) and np.array_equal(ol, ol2)
        assert pk.shape[1] < ops.shape[1] // 2
        for i in range(len(A)):
            o = bytes(ops[i, : ol[i]])
            assert tracy_b200.unpack_ops(pk[i], ol[i]) == o, (kind, cfg, i)
            w0, w1 = tracy_b200.rows_from_ops(kind, A[i], B[i], o)
            assert bytes(r0[i, : ol[i]]) == w0 and bytes(r1[i, : ol[i]]) == w1, (kind, cfg, i)
        # rows without ops: still a traceback call
        s3, none_ops, ol3, q0, q1 = ctx.gotoh(kind, A, B, DnaScore(*SC), AlignConfig(bool(hf), bool(vf)), traceback=False, rows=True)
        assert none_ops is None and np.array_equal(s3, s) and np.array_equal(ol3, ol)
        assert np.array_equal(q0, r0) and np.array_equal(q1, r1)
    # rows against the oracle's own _createAlignment for a few pairs (profiles and strings)
    if kind != "ps":
        fn = oracle_port.gotoh_pp if kind == "pp" else oracle_port.gotoh_ss
        for i in range(0, len(A), 7):
            ws, wops = fn(A[i], B[i], hf, vf, SC)
            assert oracle_port.rows_from_ops(A[i], B[i], wops) == (bytes(r0[i, : ol[i]]), bytes(r1[i, : ol[i]]))


def test_rows_device_mode(ctx):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    P, m, n = 300, 400, 1500
    prof, win = synth.align_batch(P, m, n, seed=9)
    a1, a2 = tracy_b200.uniform_profiles(prof), tracy_b200.uniform_seqs(win)
    stride = (m + n + 15) // 16 * 16
    pstride = ((m + n + 3) // 4 + 15) // 16 * 16
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    dp, dw, ao, al, bo, bl = d(prof), d(win), d(a1.off), d(a1.len), d(a2.off), d(a2.len)
    sc_t = torch.zeros(P, dtype=torch.int32, device=dev)
    pk = torch.zeros((P, pstride), dtype=torch.uint8, device=dev)
    ln = torch.zeros(P, dtype=torch.int32, device=dev)
    r0 = torch.zeros((P, stride), dtype=torch.uint8, device=dev)
    r1 = torch.zeros((P, stride), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.gotoh_device("ps", dp.data_ptr(), ao.data_ptr(), al.data_ptr(), dw.data_ptr(), bo.data_ptr(), bl.data_ptr(), P, sc_t.data_ptr(), pk.data_ptr(), pstride,
                     ln.data_ptr(), DnaScore(*SC), AlignConfig(True, False), row0=r0.data_ptr(), row1=r1.data_ptr(), rows_stride=stride, packed=True)
    s, ops, ol = ctx.gotoh("ps", a1, a2, DnaScore(*SC), AlignConfig(True, False))
    assert np.array_equal(sc_t.cpu().numpy(), s) and np.array_equal(ln.cpu().numpy(), ol)
    pkh, r0h, r1h = pk.cpu().numpy(), r0.cpu().numpy(), r1.cpu().numpy()
    for i in range(0, P, 11):
        o = bytes(ops[i, : ol[i]])
        assert tracy_b200.unpack_ops(pkh[i], ol[i]) == o
        assert (bytes(r0h[i, : ol[i]]), bytes(r1h[i, : ol[i]])) == tracy_b200.rows_from_ops("ps", prof[i], bytes(win[i]), o)


def test_four_row_upload(ctx, oracle_port, monkeypatch):
    """TB_A1_TRACE_PROFILES: uniform back-to-back trace profiles whose N and '-' rows are exact zeros travel as 4 rows of 6; results
    (rows included, which read the '-' row) equal the full upload's and the oracle's."""
    P, m, n = 1200, 500, 1200
    prof, win = synth.align_batch(P, m, n, seed=21)
    assert not prof[:, 4:, :].any()
    a1, a2 = tracy_b200.uniform_profiles(prof, trace_profiles=True), tracy_b200.uniform_seqs(win)
    h0 = ctx.stats()["h2d_bytes"]
    s, ops, ol, r0, r1 = ctx.gotoh("ps", a1, a2, DnaScore(*SC), AlignConfig(True, False), rows=True)
    sent4 = ctx.stats()["h2d_bytes"] - h0
    h0 = ctx.stats()["h2d_bytes"]
    s6, ops6, ol6, q0, q1 = ctx.gotoh("ps", tracy_b200.uniform_profiles(prof), a2, DnaScore(*SC), AlignConfig(True, False), rows=True)
    sent6 = ctx.stats()["h2d_bytes"] - h0
    assert sent6 - sent4 == P * 2 * m * 4
    assert np.array_equal(s, s6) and np.array_equal(ops, ops6) and np.array_equal(r0, q0) and np.array_equal(r1, q1)
    for i in range(0, P, 97):
        ws, wops = oracle_port.gotoh_ps(prof[i], bytes(win[i]), 1, 0, SC)
        assert (int(s[i]), bytes(ops[i, : ol[i]])) == (ws, wops)
    # a stale lane buffer must not leak into the cleared rows: first a batch with gap-row mass through the same lane, then the promise
    prof2 = prof.copy()
    prof2[:, 5, :] = 0.9
    ctx.gotoh("ps", tracy_b200.uniform_profiles(prof2), a2, DnaScore(*SC), AlignConfig(True, False), rows=True)
    s8, ops8, ol8, t0, t1 = ctx.gotoh("ps", a1, a2, DnaScore(*SC), AlignConfig(True, False), rows=True)
    assert np.array_equal(s8, s) and np.array_equal(t0, r0) and np.array_equal(t1, r1)


def test_odd_strides(ctx):
    """Caller-chosen strides that leave the strings unaligned (the C++ single-pair wrapper passes len1 + len2 + 16)."""
    rng = np.random.default_rng(77)
    A, B = _batch(rng, "ps", 20)
    s, ops, ol, r0, r1 = ctx.gotoh("ps", A, B, DnaScore(*SC), AlignConfig(True, False), rows=True)
    mx = int(max(a.shape[1] + len(b) for a, b in zip(A, B)))
    for stride in (mx, mx + 1, mx + 3):
        o2, l2 = np.zeros((len(A), stride), np.uint8), np.zeros(len(A), np.int32)
        q0, q1 = np.zeros((len(A), stride + 2), np.uint8), np.zeros((len(A), stride + 2), np.uint8)
        s2, _, _, _, _ = ctx.gotoh("ps", A, B, DnaScore(*SC), AlignConfig(True, False), out=(np.zeros(len(A), np.int32), o2, l2), rows=(q0, q1))
        assert np.array_equal(s2, s) and np.array_equal(l2, ol)
        for i in range(len(A)):
            assert bytes(o2[i, : ol[i]]) == bytes(ops[i, : ol[i]])
            assert bytes(q0[i, : ol[i]]) == bytes(r0[i, : ol[i]]) and bytes(q1[i, : ol[i]]) == bytes(r1[i, : ol[i]])
