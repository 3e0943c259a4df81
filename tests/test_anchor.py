"""CPU tests of the anchoring row: the C restatement against the goldens generated from the reference's
own src/fmindex.h (tests/golden/make_golden_anchor.py), the live reference bridge when present, and the host-side slice
arithmetic tb_reference_slice (no GPU needed)."""
import os

import numpy as np
import pytest

from conftest import ROOT

COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def load_anchor_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "anchor_golden.npz"))
    g = dict(text=bytes(z["text"]), names=bytes(z["names"]).split(b"\n"), cfgs=[tuple(int(x) for x in c) for c in z["cfgs"]],
             cons=[bytes(z[f"cons{i}"]) for i in range(int(z["ntraces"]))], fasta_rows=z["fasta_rows"])
    g["rows"] = [z[f"rows{ci}"] for ci in range(len(g["cfgs"]))]
    g["slices"] = [bytes(z[f"slices{ci}"]).split(b"\n") for ci in range(len(g["cfgs"]))]
    g["seqs"] = g["text"][:-1].split(b"\n")
    return g


def expected(row, ms):
    """(anchored, forward, kmersupport, bestpos, pass) from a golden row, by the rule of src/fmindex.h:251-284."""
    for base, ps in ((0, 1), (6, 4)):
        _, gf, ff, _, gr, fr = (int(x) for x in row[base:base + 6])
        if ff >= ms and ff > 2 * fr:
            return True, True, ff, gf, ps
        if fr >= ms and fr > 2 * ff:
            return True, False, fr, gr, ps
    return False, True, 0, 0, 0


def test_golden_rows_are_self_consistent():
    g = load_anchor_golden()
    for cfg, rows in zip(g["cfgs"], g["rows"]):
        for r in rows:
            ok, fw, ks, _, _ = expected(r, cfg[4])
            assert (ok, fw, ks) == (bool(r[12]), bool(r[13]), int(r[14]))


def test_port_scan_and_anchor_match_reference_goldens(oracle_port):
    g = load_anchor_golden()
    for ci in (0, 1):
        tl, tr, k, mi, ms = g["cfgs"][ci]
        for i in list(range(0, 38, 3)) + [26, 27, 28]:
            t, row = g["cons"][i], g["rows"][ci][i]
            hf, gf, ff = oracle_port.scan_sequence(g["text"], t, tl, tr, k, True)
            assert (len(hf), gf, ff) == tuple(int(x) for x in row[0:3]), (ci, i)
            assert oracle_port.anchor(g["text"], t, tl, tr, k, ms) == expected(row, ms), (ci, i)


def test_port_matches_live_reference(oracle_port, oracle_ref):
    if oracle_ref is None:
        pytest.skip("reference bridge not built here")
    rng = np.random.default_rng(9)
    text = bytes(rng.choice(list(b"ACGT"), 6000).astype(np.uint8)) + b"\n" + bytes(rng.choice(list(b"ACGTN"), 3000).astype(np.uint8)) + b"\n"
    h = oracle_ref.fm_build(text)
    try:
        for _ in range(6):
            p, L = int(rng.integers(0, 5000)), int(rng.integers(30, 700))
            t = bytearray(text[p:p + L].replace(b"\n", b"A"))
            for q in rng.integers(0, len(t), 4):
                t[q] = b"ACGTRN"[int(rng.integers(0, 6))]
            t = bytes(t)
            for tl, tr, k, uniq in ((5, 9, 11, True), (0, 0, 6, False), (40, 3, 16, True)):
                a = oracle_ref.scan_sequence(h, t, tl, tr, k, uniq)
                b = oracle_port.scan_sequence(text, t, tl, tr, k, uniq)
                assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    finally:
        oracle_ref.fm_free(h)


def test_reference_slice_arithmetic_matches_get_reference_slice():
    """tb_reference_slice (host) + the inclusive faidx fetch reproduce rs.pos / rs.chr / rs.refslice of the goldens."""
    import tracy_b200
    from tracy_b200 import api
    g = load_anchor_golden()
    seqlen = [len(s) + 1 for s in g["seqs"]]                       # src/fmindex.h:247
    n = 0
    for ci, (tl, tr, k, mi, ms) in enumerate(g["cfgs"]):
        for i, row in enumerate(g["rows"][ci]):
            ok, fw, ks, bp, _ = expected(row, ms)
            if not ok:
                continue
            ri, chrpos, s0, s1 = api.reference_slice(bp, seqlen, len(g["cons"][i]), mi)
            assert ri == int(row[16]) and s0 == int(row[15]), (ci, i)
            seq = g["seqs"][ri]
            sl = seq[s0: min(s1, len(seq) - 1) + 1]
            if not fw:
                sl = api_revcomp(sl)
            assert sl == g["slices"][ci][i], (ci, i)
            n += 1
    assert n > 100


def api_revcomp(s):
    """reverseComplement(std::string&), reference src/fmindex.h:11-26 (non-ACGTN keep the forward character of the slot)."""
    out = bytearray(s)
    for i, ch in enumerate(reversed(s.upper())):
        if ch in b"ACGTN":
            out[i] = bytes([ch]).translate(COMP)[0]
    return bytes(out)
