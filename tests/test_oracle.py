"""CPU tests (-m "not gpu"): the plain-C oracle (oracle/gotoh_oracle.c) against
 (1) the committed golden vectors, which were produced by the reference's own unmodified headers
     (tests/golden/make_golden.py), and
 (2) the reference build itself (oracle/_ref), fuzzed, wherever that build is present.
Also the host-side decompose glue (tracy_b200/decompose.py) driven by the oracle's sweep."""
import numpy as np
import pytest

from conftest import load_decompose_golden, load_gotoh_golden
from oracle import loader
from tracy_b200 import decompose, synth

GOLD = load_gotoh_golden()


def _port_run(port, c):
    if c["kind"] == "ps":
        s, ops = port.gotoh_ps(c["a"], c["b"], c["hf"], c["vf"], c["sc"])
        rows = port.rows_from_ops(c["a"], port.onehot(c["b"]), ops)
    elif c["kind"] == "pp":
        s, ops = port.gotoh_pp(c["a"], c["b"], c["hf"], c["vf"], c["sc"])
        rows = port.rows_from_ops(c["a"], c["b"], ops)
    else:
        s, ops = port.gotoh_ss(c["a"], c["b"], c["hf"], c["vf"], c["sc"])
        rows = port.rows_from_ops(c["a"], c["b"], ops)
    return s, ops, rows


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_port_matches_golden(oracle_port, idx):
    c = GOLD[idx]
    s, ops, rows = _port_run(oracle_port, c)
    assert s == c["score"]
    assert ops == loader.ops_from_rows(c["row0"], c["row1"])
    assert rows == (c["row0"], c["row1"])


def test_golden_covers_edge_shapes():
    kinds = {c["kind"] for c in GOLD}
    assert kinds == {"ps", "pp", "ss"}
    assert {(c["hf"], c["vf"]) for c in GOLD} == {(0, 0), (0, 1), (1, 0), (1, 1)}
    assert any(len(c["row0"]) >= 4000 for c in GOLD)          # the config-2 shape
    assert any(c["kind"] == "ps" and b"N" in c["b"] for c in GOLD)


def test_port_vs_reference_fuzz(oracle_port, oracle_ref):
    if oracle_ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(99)
    scs = [(3, -5, -10, -4), (5, -4, -10, -1), (1, -1, -2, -1), (4, -4, -6, 0)]
    for it in range(400):
        m, n = int(rng.integers(1, 90)), int(rng.integers(1, 120))
        hf, vf = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        sc = scs[it % 4]
        a = synth.random_profile(rng, m, ["trace", "ties", "msa"][it % 3])
        if it % 3 == 0:
            b = synth.random_seq(rng, n, b"ACGTNn-acgtRY" if it % 2 else b"ACGT")
            s, ops = oracle_port.gotoh_ps(a, b, hf, vf, sc)
        elif it % 3 == 1:
            b = synth.random_profile(rng, n, ["ties", "msa", "trace"][it % 3])
            s, ops = oracle_port.gotoh_pp(a, b, hf, vf, sc)
        else:
            a = synth.random_seq(rng, m, b"ACGTN")
            b = synth.mutate_seq(rng, a, 0.2, 0.2) or b"C"
            s, ops = oracle_port.gotoh_ss(a, b, hf, vf, sc)
        rs, r0, r1 = oracle_ref.gotoh(a, b, hf, vf, sc)
        assert (s, ops) == (rs, loader.ops_from_rows(r0, r1)), (it, m, n, hf, vf)
        assert oracle_ref.gotoh_score(a, b, hf, vf, sc) == rs


def test_known_answers(oracle_port):
    """Hand-derivable cases for the recurrences and tie rules."""
    sc = (3, -5, -10, -4)
    # identical strings: all diagonal
    assert oracle_port.gotoh_ss(b"ACGT", b"ACGT", 0, 0, sc) == (12, b"ssss")
    # one deletion in a1, global: gap open+extend = -14, three matches
    s, ops = oracle_port.gotoh_ss(b"ACT", b"ACGT", 0, 0, sc)
    assert s == 9 - 14 and sorted(ops) == sorted(b"sshs") and ops.count(b"h") == 1
    # free horizontal end gaps: the trace floats inside the reference at no cost
    s, ops = oracle_port.gotoh_ss(b"CG", b"AACGTT", 1, 0, sc)
    assert s == 6 and ops == b"hhsshh"
    # without free end gaps the same input pays for them
    s2, _ = oracle_port.gotoh_ss(b"CG", b"AACGTT", 0, 0, sc)
    assert s2 < s
    # tie H vs V vs diagonal: horizontal wins (src/gotoh.h:134-135) -> with all-equal costs the walk prefers 'h'
    s, ops = oracle_port.gotoh_ss(b"A", b"C", 0, 0, (0, 0, 0, 0))
    assert s == 0 and ops == b"vh"   # end->start: 'h' first, then 'v'; reversed for start->end
    # empty inputs
    assert oracle_port.gotoh_ss(b"", b"ACG", 0, 0, sc) == (-10 - 12, b"hhh")
    assert oracle_port.gotoh_ss(b"", b"ACG", 1, 0, sc) == (0, b"hhh")
    assert oracle_port.gotoh_ss(b"AC", b"", 0, 1, sc) == (0, b"vv")


def test_all_quarter_columns(oracle_port):
    """An all-0.25 trace column (totalsig == 0 in createProfile, src/profile.h:38-39) scores trunc(0.25*(m+3mm))."""
    p = np.zeros((6, 3), np.float32)
    p[:4] = 0.25
    s, ops = oracle_port.gotoh_ps(p, b"ACG", 0, 0, (3, -5, -10, -4))
    assert ops == b"sss" and s == 3 * int(0.25 * 3 + 3 * 0.25 * -5)


def test_phase_ref_allele_table(oracle_port):
    alphabet = "ACGTNRYSWKM-X"
    for p in alphabet:
        for s in alphabet:
            for r in alphabet:
                want = decompose.phase_ref_allele(p, s, r)
                got = oracle_port.phase(p.encode(), s.encode(), r.encode()).decode()
                assert got == want, (p, s, r)


DGOLD = load_decompose_golden()


@pytest.mark.parametrize("idx", range(len(DGOLD)))
def test_decompose_glue_with_oracle_sweep(oracle_port, idx, capsys):
    """Host glue + oracle sweep == the reference's decomposeAlleles on the golden cases."""
    c = DGOLD[idx]

    def sweep(refrow, pri, sec, vi_end, ai, vi, ndel, nins, grid):
        return oracle_port.decompose_sweep(refrow, pri, sec, vi_end, ai, vi, ndel, nins, grid)

    pri, sec, dcp, info = decompose.decompose_alleles(c["row0"], c["row1"], c["pri"], c["sec"], c["trimL"], c["trimR"], c["maxindel"],
                                                      c["madc"], c["bp"], c["nref"], sweep)
    assert pri == c["pri_out"]
    assert sec == c["sec_out"]
    assert np.array_equal(dcp, c["dcp"])


def test_decompose_golden_modes_covered(oracle_port):
    modes = set()
    for c in DGOLD:
        sweep = lambda *a: oracle_port.decompose_sweep(*a)
        modes.add(decompose.decompose_alleles(c["row0"], c["row1"], c["pri"], c["sec"], c["trimL"], c["trimR"], c["maxindel"], c["madc"],
                                              c["bp"], c["nref"], sweep)[3]["mode"])
    assert {"del", "ins"} <= modes, modes


def test_decompose_alleles_batch_lockstep(oracle_port, capsys):
    """decompose_alleles_batch runs every trace's decomposeAlleles in lock-step with ONE sweep call per phase; with the
    sweeps served by the CPU oracle it must reproduce the golden outputs of all cases at once."""
    class Ctx:
        calls = 0

        def decompose_sweep(self, refrows, pris, secs, vi_end, ai, vi, ndel, nins, grid=False):
            Ctx.calls += 1
            S = int(max(1, max(ndel), max(nins)))
            n = len(refrows)
            fref, fins = np.zeros((n, S), np.int32), np.zeros((n, S), np.int32)
            g = np.zeros((n, S, S), np.int32) if grid else None
            for t in range(n):
                a, b, c = oracle_port.decompose_sweep(refrows[t], pris[t], secs[t], vi_end[t], ai[t], vi[t], ndel[t], nins[t], grid)
                fref[t, : ndel[t]] = a
                fins[t, : nins[t]] = b
                if grid:
                    g[t, : nins[t], : ndel[t]] = c
            return fref, fins, g
    items = [dict(row0=c["row0"], row1=c["row1"], primary=c["pri"], secondary=c["sec"], trim_left=c["trimL"], trim_right=c["trimR"],
                  maxindel=c["maxindel"], madc=c["madc"], breakpoint=c["bp"], refslice_len=c["nref"]) for c in DGOLD]
    res = decompose.decompose_alleles_batch(Ctx(), items)
    assert Ctx.calls <= 2                      # the indel sweep for everybody, the ins x del grid for those without a candidate
    for c, (pri, sec, dcp, info) in zip(DGOLD, res):
        assert pri == c["pri_out"] and sec == c["sec_out"] and np.array_equal(dcp, c["dcp"])
