"""GPU parity of tb_index_build / tb_anchor (anchor.cu) against the goldens generated from the reference's own
src/fmindex.h over sdsl's FM-index, against the C restatement on seeded inputs, and through size-independent properties
(planted positions recovered) at sizes the CPU oracle cannot reach."""
import numpy as np
import pytest

from test_anchor import COMP, api_revcomp, expected, load_anchor_golden

pytestmark = pytest.mark.gpu


def test_anchor_matches_reference_goldens(ctx):
    g = load_anchor_golden()
    idx = ctx.build_index(g["text"])
    assert idx.text_len == len(g["text"])
    for ci, (tl, tr, k, mi, ms) in enumerate(g["cfgs"]):
        r = ctx.anchor(idx, g["cons"], tl, tr, k, ms)
        for i, row in enumerate(g["rows"][ci]):
            ok, fw, ks, bp, ps = expected(row, ms)
            got = (bool(r["anchored"][i]), bool(r["forward"][i]), int(r["kmersupport"][i]), int(r["bestpos"][i]), int(r["pass_"][i]))
            assert got == (ok, fw, ks, bp, ps), (ci, i, got)
    idx.close()


def test_anchor_single_sequence_reference(ctx):
    g = load_anchor_golden()
    idx = ctx.build_index(g["seqs"][0])
    r = ctx.anchor(idx, g["cons"][:12], 50, 50, 15, 3)
    for i, row in enumerate(g["fasta_rows"]):
        assert (bool(r["anchored"][i]), bool(r["forward"][i]), int(r["kmersupport"][i])) == (bool(row[0]), bool(row[1]), int(row[2])), i
    idx.close()


def test_anchor_vs_port_seeded(ctx, oracle_port):
    rng = np.random.default_rng(77)
    seqs = [bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8)) for n in (7000, 5000)]
    seqs[1] = seqs[1][:1000] + seqs[0][2000:3500] + seqs[1][2500:]
    text = b"\n".join(seqs) + b"\n"
    cons = []
    for i in range(24):
        c = i % 2
        L = int(rng.integers(5, 900))
        p = int(rng.integers(0, len(seqs[c]) - L))
        t = bytearray(seqs[c][p:p + L])
        for q in rng.integers(0, L, max(1, L // 60)):
            t[q] = b"ACGTNRYK"[int(rng.integers(0, 8))]
        t = bytes(t)
        cons.append(t.translate(COMP)[::-1] if i % 3 == 0 else t)
    cons += [b"", b"A", b"ACGTTGCAAC"]
    idx = ctx.build_index(text)
    for tl, tr, k, ms in ((50, 50, 15, 3), (0, 0, 8, 1), (3, 200, 16, 5), (0, 10, 1, 3)):
        r = ctx.anchor(idx, cons, tl, tr, k, ms)
        for i, t in enumerate(cons):
            want = oracle_port.anchor(text, t, tl, tr, k, ms)
            got = (bool(r["anchored"][i]), bool(r["forward"][i]), int(r["kmersupport"][i]), int(r["bestpos"][i]), int(r["pass_"][i]))
            assert got == want, (tl, tr, k, ms, i, got, want)
    idx.close()


def test_anchor_long_consensus_uses_global_tables(ctx, oracle_port):
    """Scans longer than the shared-memory tables hold (> 4096 k-mers) take the global-table path."""
    rng = np.random.default_rng(5)
    seq = bytes(rng.choice(list(b"ACGT"), 12000).astype(np.uint8))
    text = seq + b"\n"
    cons = [seq[1000:7000], seq[500:6000].translate(COMP)[::-1], seq[100:900]]
    idx = ctx.build_index(text)
    r = ctx.anchor(idx, cons, 50, 50, 15, 3)
    for i, t in enumerate(cons):
        want = oracle_port.anchor(text, t, 50, 50, 15, 3)
        got = (bool(r["anchored"][i]), bool(r["forward"][i]), int(r["kmersupport"][i]), int(r["bestpos"][i]), int(r["pass_"][i]))
        assert got == want, (i, got, want)
    idx.close()


def test_anchor_planted_positions_large(ctx):
    """4 Mbp text, 20 000 error-free 1 kb traces: every trace anchors at its planted offset and strand."""
    rng = np.random.default_rng(11)
    n, nt, L = 4_000_000, 20000, 1000
    text = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).astype(np.uint8)
    text[1_000_000] = ord("\n")
    text[-1] = ord("\n")
    pos = rng.integers(0, n - L - 2, nt)
    pos[(pos <= 1_000_000) & (pos + L > 1_000_000)] = 2_000_000       # keep traces inside one sequence
    fw = rng.random(nt) < 0.5
    tb = text.tobytes()
    cons = [tb[p:p + L] if f else tb[p:p + L].translate(COMP)[::-1] for p, f in zip(pos, fw)]
    idx = ctx.build_index(text)
    r = ctx.anchor(idx, cons, 50, 50, 15, 3)
    assert r["anchored"].all()
    assert np.array_equal(r["forward"], fw)
    assert np.array_equal(r["bestpos"], pos)
    # a random 15-mer recurs somewhere in 4 Mbp with p ~ 0.4 %: a few of the 900 k-mers per trace are not unique
    assert (r["kmersupport"] <= L - 100).all() and (r["kmersupport"] >= L - 130).all() and (r["pass_"] == 1).all()
    assert ctx.last_anchor_ms() > 0
    idx.close()


def test_index_rejects_foreign_bytes(ctx):
    import tracy_b200
    with pytest.raises(tracy_b200.TracyError):
        ctx.build_index(b"ACGT*ACGT\n")
    with pytest.raises(tracy_b200.TracyError):
        idx = ctx.build_index(b"ACGTACGTACGTACGTACGTAAAC\n")
        try:
            ctx.anchor(idx, [b"ACGTACGTACGTACGTACGT"], 0, 0, 17, 3)      # kmer beyond the index depth
        finally:
            idx.close()


def test_align_genome_batch_golden(ctx):
    """samples -> basecall -> createProfile -> anchor -> gotoh -> trimReferenceSlice -> gotoh, all through the library,
    against sage() for an indexed genome composed from the reference's own functions (make_golden_align_genome.py)."""
    import os
    from conftest import ROOT
    from tracy_b200 import DnaScore, drivers
    G = np.load(os.path.join(ROOT, "tests", "golden", "align_genome_golden.npz"))
    n = int(G["n"])
    tl, trr, kmer, maxindel, ms = (int(x) for x in G["cfg"])
    text = bytes(G["text"])
    seqs = text[:-1].split(b"\n")
    traces, ploc = [G[f"tr{i}"] for i in range(n)], [G[f"ploc{i}"] for i in range(n)]
    bc = ctx.basecall(traces, ploc, 0.33)
    args = ([b["bcPos"] for b in bc], [b["primary"] for b in bc], [b["secondary"] for b in bc])
    full = ctx.create_profile(traces, *args, trim_left=0, trim_right=0)
    trimmed = ctx.create_profile(traces, *args, trim_left=tl, trim_right=trr)
    idx = ctx.build_index(text)
    res = drivers.align_genome_batch(ctx, idx, seqs, [b["consensus"] for b in bc], trimmed, full, DnaScore(3, -5, -10, -4), tl, trr, kmer, ms, maxindel)
    idx.close()
    for i, r in enumerate(res):
        if not int(G[f"ok{i}"]):
            assert r is None, i
            continue
        assert r is not None, i
        assert [int(r["forward"]), r["chr"], r["kmersupport"], r["pos"], r["score"]] == [int(x) for x in G[f"meta{i}"]], i
        for k in ("refslice", "row0", "row1"):
            assert r[k] == bytes(G[f"{k}{i}"]), (i, k)
