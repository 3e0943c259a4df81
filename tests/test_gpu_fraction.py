"""GPU parity of tb_allelic_fraction (fraction.cu) against goldens generated from the reference's own allelicFraction
(src/decompose.h:412-617, tests/golden/make_golden_fraction.py): FP64, bit-exact, including the NaN case (a position with
no signal) and non-ACGT codes."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_allelic_fraction_golden(ctx):
    G = np.load(os.path.join(ROOT, "tests", "golden", "fraction_golden.npz"))
    groups = {}
    for i in range(int(G["n"])):
        groups.setdefault(tuple(int(x) for x in G[f"cfg{i}"]), []).append(i)
    for (tl, trr), idx in groups.items():
        out = ctx.allelic_fraction([G[f"tr{i}"] for i in idx], [G[f"pos{i}"] for i in idx], [bytes(G[f"pri{i}"]) for i in idx],
                                   [bytes(G[f"sec{i}"]) for i in idx], tl, trr)
        for row, i in zip(out, idx):
            assert row.tobytes() == G[f"out{i}"].tobytes(), (i, row, G[f"out{i}"])


def test_allelic_fraction_live_reference(ctx, oracle_ref):
    if oracle_ref is None:
        pytest.skip("reference bridge not present")
    rng = np.random.default_rng(8)
    tr_l, pos_l, pri_l, sec_l, want = [], [], [], [], []
    for _ in range(24):
        nbc = int(rng.integers(120, 700))
        ns = 12 * nbc + 30
        tr = rng.integers(0, 40, size=(4, ns)).astype(np.int32)
        pos = (12 * np.arange(nbc) + 8).astype(np.int32)
        pri = bytearray(rng.choice(list(b"ACGT"), nbc).astype(np.uint8).tobytes())
        sec = bytearray(pri)
        f = rng.random()
        for j in range(nbc):
            a = b"ACGT".index(pri[j])
            h = int(rng.integers(500, 1500))
            if rng.random() < 0.3:
                b = (a + int(rng.integers(1, 4))) % 4
                sec[j] = b"ACGTN"[b if rng.random() > 0.02 else 4]
                tr[a, pos[j]] += int(h * f); tr[b, pos[j]] += int(h * (1 - f))
            else:
                tr[a, pos[j]] += h
        tr_l.append(tr); pos_l.append(pos); pri_l.append(bytes(pri)); sec_l.append(bytes(sec))
        want.append(oracle_ref.allelic_fraction(tr, pos, bytes(pri), bytes(sec), 50, 50))
    out = ctx.allelic_fraction(tr_l, pos_l, pri_l, sec_l, 50, 50)
    assert out.tobytes() == np.array(want, np.float64).tobytes()


def test_allelic_fraction_perfect_fits_and_ties(ctx, oracle_ref):
    """The table pre-filter of fraction.cu must never drop a point that wins or ties: noise-free traces whose peak ratios sit exactly on
    grid values (several grid points reach the same, tiny or zero, sum of squares; (0.5, 0.5) ties with the start value), few and
    many differing positions, a third allele on the grid."""
    if oracle_ref is None:
        pytest.skip("reference bridge not present")
    rng = np.random.default_rng(81)
    tr_l, pos_l, pri_l, sec_l, want = [], [], [], [], []
    for it in range(20):
        nbc = int(rng.integers(60, 400))
        ns = 12 * nbc + 30
        tr = np.zeros((4, ns), np.int32)
        pos = (12 * np.arange(nbc) + 8).astype(np.int32)
        pri = bytearray(rng.choice(list(b"ACGT"), nbc).astype(np.uint8).tobytes())
        sec = bytearray(pri)
        pa, pb, pc = [(50, 50, 0), (60, 40, 0), (75, 25, 0), (33, 67, 0), (100, 0, 0), (48, 40, 12), (1, 99, 0), (20, 20, 60)][it % 8]
        ndiff = [1, 2, 17, 40, nbc // 2][it % 5]
        diff = set(int(x) for x in rng.choice(np.arange(nbc), size=ndiff, replace=False))
        for j in range(nbc):
            a = b"ACGT".index(pri[j])
            if j in diff:
                b = (a + int(rng.integers(1, 4))) % 4
                sec[j] = b"ACGT"[b]
                c3 = next(k for k in range(4) if k not in (a, b))
                tr[a, pos[j]] += pa * 10; tr[b, pos[j]] += pb * 10; tr[c3, pos[j]] += pc * 10
            else:
                tr[a, pos[j]] += 1000
        tr_l.append(tr); pos_l.append(pos); pri_l.append(bytes(pri)); sec_l.append(bytes(sec))
        want.append(oracle_ref.allelic_fraction(tr, pos, bytes(pri), bytes(sec), 10, 10))
    out = ctx.allelic_fraction(tr_l, pos_l, pri_l, sec_l, 10, 10)
    assert out.tobytes() == np.array(want, np.float64).tobytes(), (out, want)
