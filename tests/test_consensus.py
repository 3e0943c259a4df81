"""The consensus letters of `tracy consensus` (gtLetter / pairwiseConsensus, reference src/consensus.h:94-238): the Python statement
(tracy_b200/consensus.py) and the native function (tb_pairwise_consensus) against goldens made by the reference
(tests/golden/make_golden_consensus.py) and, where the reference build exists, against it directly."""
import os

import numpy as np
import pytest

from conftest import ROOT
from tracy_b200 import consensus as cons_mod
from tracy_b200 import msa

G = np.load(os.path.join(ROOT, "tests", "golden", "consensus_golden.npz"))


def _second(i):
    p2 = G[f"p2_{i}"]
    return p2 if int(G[f"meta{i}"][0]) else msa._revcomp(p2)


@pytest.mark.parametrize("fn", [cons_mod.pairwise_consensus, cons_mod.pairwise_consensus_native])
def test_pairwise_consensus_golden(fn):
    checked = 0
    for i in range(int(G["n"])):
        for un in (0, 1):
            for iu in (0, 1):
                c, q = fn(bytes(G[f"r0_{i}"]), bytes(G[f"r1_{i}"]), G[f"p1_{i}"], _second(i), un, iu)
                assert c == bytes(G[f"cons{un}{iu}_{i}"]) and np.array_equal(np.asarray(q, np.uint32), G[f"qual{un}{iu}_{i}"]), (i, un, iu)
                checked += 1
    assert checked == 32


def test_gt_letter_differential(oracle_ref):
    if oracle_ref is None:
        pytest.skip("reference build not present")
    rng = np.random.default_rng(1)
    for t in range(4000):
        kind = t % 5
        if kind == 0:
            cl = rng.random(6)
        elif kind == 1:
            cl = np.float64(np.float32(rng.random(6)) * np.float32(rng.integers(0, 2, 6)))
        elif kind == 2:
            cl = np.float64(rng.integers(0, 4, 6)) / 2
        elif kind == 3:
            cl = np.float64(np.float32(10.0 ** rng.uniform(-30, 0, 6)))
        else:
            cl = np.zeros(6)
            cl[rng.integers(0, 6)] = rng.random()
            cl[rng.integers(0, 6)] += rng.random() * 1e-3
        for iu in (0, 1):
            w = oracle_ref.gt_letter(cl, iu)
            assert (w[0].decode(), w[1]) == cons_mod.gt_letter(cl, iu), (cl, iu)


def test_native_pairwise_consensus_differential(oracle_ref):
    """Random alignments over random profiles (incl. N / gap weights, zero columns): native == Python == reference."""
    rng = np.random.default_rng(2)
    for it in range(40):
        m, n = int(rng.integers(1, 80)), int(rng.integers(1, 80))
        p1, p2 = rng.random((6, m)).astype(np.float32), rng.random((6, n)).astype(np.float32)
        if it % 4 == 0:
            p1[:, rng.integers(0, m)] = 0
            p2[4:, :] = 0
        row0, row1, i, j = bytearray(), bytearray(), 0, 0
        while i < m or j < n:
            u = rng.random()
            if i < m and j < n and u < 0.7:
                row0.append(65); row1.append(67); i += 1; j += 1
            elif i < m and (j >= n or u < 0.85):
                row0.append(65); row1.append(45); i += 1
            else:
                row0.append(45); row1.append(67); j += 1
        for un in (0, 1):
            for iu in (0, 1):
                a = cons_mod.pairwise_consensus(bytes(row0), bytes(row1), p1, p2, un, iu)
                b = cons_mod.pairwise_consensus_native(bytes(row0), bytes(row1), p1, p2, un, iu)
                assert a[0] == b[0] and np.array_equal(np.asarray(a[1], np.uint32), b[1]), (it, un, iu)
                if oracle_ref is not None:
                    c = oracle_ref.pairwise_consensus(bytes(row0), bytes(row1), p1, p2, un, iu)
                    assert c[0] == b[0] and np.array_equal(c[1], b[1]), (it, un, iu)
