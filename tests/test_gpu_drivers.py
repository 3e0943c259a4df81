"""GPU parity test of drivers.decompose_batch: the DP sequence of `tracy decompose` (reference src/indigo.h:190-388) for a
batch of traces, against tests/golden/drivers_golden.npz -- the same sequence composed one trace at a time from the
reference's own functions (tests/golden/make_golden_drivers.py)."""
import os

import numpy as np
import pytest

from tracy_b200 import DnaScore, drivers

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "drivers_golden.npz"))


def _cases():
    groups = {}
    for i in range(int(G["n"])):
        tl, trr, maxindel, madc, ok = (int(x) for x in G[f"cfg{i}"])
        groups.setdefault((tl, trr, maxindel, madc), []).append(i)
    return groups


def test_decompose_batch_golden(ctx, capsys):
    for (tl, trr, maxindel, madc), idx in _cases().items():
        res = drivers.decompose_batch(ctx, [G[f"tr{i}"] for i in idx], [G[f"pos{i}"] for i in idx], [bytes(G[f"pri{i}"]) for i in idx],
                                      [bytes(G[f"sec{i}"]) for i in idx], [bytes(G[f"ref{i}"]) for i in idx], DnaScore(3, -5, -10, -4),
                                      tl, trr, maxindel, madc)
        for i, r in zip(idx, res):
            if int(G[f"cfg{i}"][4]) == 0:
                assert r is None, i
                continue
            assert r is not None, i
            assert [int(r["forward"]), r["score"]] == [int(x) for x in G[f"fw{i}"]], i
            bp = r["breakpoint"]
            assert np.array_equal(np.array([bp["indelshift"], bp["traceleft"], bp["breakpoint"], bp["bestDiff"]], np.float64), G[f"bp{i}"]), i
            for k in ("refslice", "row0", "row1", "primary", "secondary", "secDecompose"):
                assert r[k] == bytes(G[f"{k}{i}"]), (i, k)
            assert np.array_equal(r["decomp"], G[f"decomp{i}"]), i
            assert np.array(r["allele_fractions"], np.float64).tobytes() == G[f"frac{i}"].tobytes(), (i, r["allele_fractions"], G[f"frac{i}"])
            for name in ("align1", "align2", "align3"):
                a = r[name]
                assert [a["score"], a["pos"]] == [int(x) for x in G[f"{name}_s{i}"]], (i, name)
                assert (a["row0"], a["row1"], a["refslice"]) == (bytes(G[f"{name}_r0{i}"]), bytes(G[f"{name}_r1{i}"]), bytes(G[f"{name}_sl{i}"])), (i, name)


def test_consensus_batch_golden(ctx):
    """drivers.consensus_batch against the DP sequence of consensus() composed from the reference's functions."""
    Gc = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "consensus_golden.npz"))
    n = int(Gc["n"])
    res = drivers.consensus_batch(ctx, [Gc[f"p1_{i}"] for i in range(n)], [Gc[f"p2_{i}"] for i in range(n)], DnaScore(3, -5, -10, -4),
                                  min_overlap=100, match_fraction=0.5)
    for i, r in enumerate(res):
        assert [int(r["forward"]), r["score"]] == [int(x) for x in Gc[f"meta{i}"]], i
        assert (r["row0"], r["row1"]) == (bytes(Gc[f"r0_{i}"]), bytes(Gc[f"r1_{i}"])), i
        a0, a1 = np.frombuffer(r["row0"], np.uint8), np.frombuffer(r["row1"], np.uint8)
        both = (a0 != 45) & (a1 != 45)
        na, nm = int(both.sum()), int((both & (a0 == a1)).sum())
        assert r["num_aligned"] == na and r["num_match"] == nm
        assert r["ok"] == (not (na < 100 or (nm / na if na else 0.0) < 0.5)), i          # src/consensus.h:546-549
    assert not res[-1]["ok"] and any(r["ok"] for r in res)
