"""CPU tests of the host logic either side of the DP kernels against golden vectors
produced by the reference's own headers (tests/golden/make_golden_glue.py). The msa.py glue is exercised with a stand-in
context whose gotoh()/revcomp_profile() are served by the CPU oracle, so tree building, row merging, consensus and the
orientation loop are checked here; the same pipeline runs against the CUDA kernels in tests/test_gpu_glue.py."""
import os

import numpy as np
import pytest

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, msa

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glue_golden.npz"))
SC = (3, -5, -10, -4)


class OracleContext:
    """Context stand-in for CPU tests: same call shapes as tracy_b200.Context, DP served by oracle/gotoh_oracle.c."""

    def __init__(self, port):
        self.port = port

    def gotoh(self, kind, a1, a2, sc=DnaScore(3, -5, -10, -4), ac=AlignConfig(True, False), traceback=True, out=None):
        assert kind == "pp"
        s4 = (sc.match, sc.mismatch, sc.go, sc.ge)

        def items(a):                                              # an Arena of offsets into one packed array, or a plain list
            if isinstance(a, tracy_b200.Arena):
                return [a.base[o: o + 6 * n_].reshape(6, n_) for o, n_ in zip(a.off, a.len)]
            return a
        a1, a2 = items(a1), items(a2)
        n = len(a1)
        scores = np.zeros(n, np.int32)
        stride = max(max((np.asarray(x).shape[1] + np.asarray(y).shape[1] for x, y in zip(a1, a2)), default=1), 1)
        ops = np.zeros((n, stride), np.uint8) if traceback else None
        ol = np.zeros(n, np.int32) if traceback else None
        for i, (x, y) in enumerate(zip(a1, a2)):
            if traceback:
                scores[i], o = self.port.gotoh_pp(x, y, int(ac.horizontal), int(ac.vertical), s4)
                ops[i, : len(o)] = np.frombuffer(o, np.uint8)
                ol[i] = len(o)
            else:
                scores[i] = self.port.gotoh_score_pp(x, y, int(ac.horizontal), int(ac.vertical), s4)
        return scores, ops, ol

    def revcomp_profile(self, profiles):
        return [self.port.revcomp_profile(p) for p in profiles]


def test_trim_reference_slice_golden():
    for i in range(int(GOLD["ntrim"])):
        fw, pos, tl, trr, npos = (int(x) for x in GOLD[f"tr_cfg{i}"])
        out, p2 = tracy_b200.trim_reference_slice(bytes(GOLD[f"tr_r0{i}"]), bytes(GOLD[f"tr_r1{i}"]), bytes(GOLD[f"tr_ref{i}"]), bool(fw), pos, tl, trr)
        assert out == bytes(GOLD[f"tr_out{i}"]) and p2 == npos, i


def test_find_breakpoint_golden():
    cases = [(GOLD[f"bp_p{i}"], GOLD[f"bp_out{i}"]) for i in range(int(GOLD["nbp"]))]
    cases += [(GOLD[f"cp_out{i}"], GOLD[f"cp_bp{i}"]) for i in range(int(GOLD["ncp"]))]
    for k, (p, want) in enumerate(cases):
        b = tracy_b200.find_breakpoint(p)
        got = np.array([b["indelshift"], b["traceleft"], b["breakpoint"], b["bestDiff"]], np.float64)
        assert np.array_equal(got, want), (k, got, want)


def test_profile_from_alignment_golden():
    for i in range(int(GOLD["nal"])):
        assert np.array_equal(msa.profile_from_alignment(GOLD[f"al_rows{i}"]), GOLD[f"al_out{i}"]), i


def test_upgma_quirks():
    """closestPair starts from -1 with a strict '>': negative scores never merge, ties take the first (i, j) in scan order,
    merged distances use C++ integer division (src/msa.h:44-70)."""
    d = np.zeros((4, 4), np.int64)
    d[0, 1], d[0, 2], d[0, 3], d[1, 2], d[1, 3], d[2, 3] = 7, 7, -3, 5, -9, -1
    p, root = msa.upgma(d, 4)
    assert (p[0, 0], p[1, 0], p[4, 1], p[4, 2]) == (4, 4, 0, 1)          # first maximum wins
    assert p[2, 0] == 5 and p[4, 0] == 5 and root == 5                   # (7 + 5) / 2 = 6 joins node 4 with 2
    assert p[3, 0] == -1                                                  # (-3 + -9)/2 = -6, (-6 + -1)/2 = -3 (truncation): never > -1
    p, root = msa.upgma(np.zeros((1, 1), np.int64), 1)
    assert root == 0 and (p[0] == -1).all()


@pytest.mark.parametrize("idx", range(int(GOLD["nmsa"])))
def test_msa_pipeline_host_logic(oracle_port, idx):
    """revSeqBasedOnDist -> msa -> consensus with the DP served by the CPU oracle: orientation vector, distance matrix,
    leaf order, every alignment row and the consensus strings equal the reference's."""
    run_msa_case(OracleContext(oracle_port), idx)


def run_msa_case(ctx, idx):
    n = int(GOLD[f"ms_n{idx}"])
    profs = [GOLD[f"ms_p{idx}_{k}"].copy() for k in range(n)]
    fwd = [True] * n
    msa.rev_seq_based_on_dist(ctx, profs, fwd, DnaScore(*SC))
    assert fwd == [bool(x) for x in GOLD[f"ms_fwd{idx}"]]
    rows, seqidx, dist = msa.msa(ctx, profs, DnaScore(*SC))
    assert np.array_equal(np.triu(dist, 1), np.triu(GOLD[f"ms_dist{idx}"], 1))
    assert list(seqidx) == list(GOLD[f"ms_idx{idx}"])
    assert np.array_equal(rows, GOLD[f"ms_rows{idx}"])
    gapped, cs, qs = msa.consensus(rows, 0.5, False)
    assert gapped == bytes(GOLD[f"ms_gapped{idx}"]) and cs == bytes(GOLD[f"ms_cons{idx}"]) and qs == bytes(GOLD[f"ms_qual{idx}"])


def run_assemble64(ctx):
    """N = 64 overlapping traces, 24 of them planted reverse-complemented (tests/golden/make_golden_assemble64.py: the reference's
    revSeqBasedOnDist -> msa -> consensus): orientation vector, distance matrix, leaf order (the guide tree's post-order), every
    alignment row and the consensus through drivers.assemble_denovo -- orientation table, score-ordered exclusion, table-fed msa."""
    from tracy_b200 import drivers
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assemble64_golden.npz"))
    n = int(g["n"])
    profs = [g[f"p{k}"].copy() for k in range(n)]
    T = msa.orientation_table(ctx, profs, DnaScore(*SC))
    r = drivers.assemble_denovo(ctx, profs, DnaScore(*SC), 0.5, 0.05, table=T)
    assert r["forward"] == [bool(x) for x in g["fwd"]]
    assert r["kept"] == list(range(n))
    assert list(r["seqidx"]) == list(g["seqidx"])
    assert np.array_equal(r["rows"], g["rows"])
    assert (r["gapped"], r["consensus"], r["quality"]) == (bytes(g["gapped"]), bytes(g["cons"]), bytes(g["qual"]))
    # the table-fed distance matrix is the one distanceMatrix() computes on the oriented traces
    fwd = [True] * n
    d, T, o = msa.rev_seq_based_on_dist(ctx, [p.copy() for p in profs], fwd, DnaScore(*SC), table=T, with_state=True)
    assert np.array_equal(np.triu(msa.oriented_distance(T, o), 1), np.triu(g["dist"], 1))


def test_assemble64_host_logic(oracle_port):
    run_assemble64(OracleContext(oracle_port))


def test_reverse_complement_seq(oracle_ref):
    """drivers.reverse_complement_seq against reverseComplement(std::string&) (src/fmindex.h:11-26), incl. its quirk for
    characters outside ACGTN; the three fixed cases were produced by the reference."""
    from tracy_b200 import drivers
    fixed = {b"ACGTNRacg": b"CGTTNACGT", b"": b"", b"AAxCC-GT": b"ACxGG-TT"}
    for k, v in fixed.items():
        assert drivers.reverse_complement_seq(k) == v
    if oracle_ref is None:
        pytest.skip("reference build not present")
    rng = np.random.default_rng(4)
    alphabet = np.frombuffer(b"ACGTNacgtnRYX-", np.uint8)
    for _ in range(100):
        s = bytes(alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 60)))])
        assert drivers.reverse_complement_seq(s) == oracle_ref.reverse_complement(s), s


ASM_REF = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assemble_reference_golden.npz"))


def run_assemble_reference_case(ctx, it):
    from tracy_b200 import drivers
    G = ASM_REF
    profs = [G[f"p{it}_{i}"] for i in range(int(G[f"n{it}"]))]
    got = drivers.assemble_reference(ctx, profs, bytes(G[f"ref{it}"]), DnaScore(*SC), 0.5, 0.5, bool(G[f"inc{it}"]))
    assert got["idx"] == [int(x) for x in G[f"idx{it}"]] and got["forward"] == [bool(x) for x in G[f"fwd{it}"]]
    assert np.array_equal(got["rows"], G[f"rows{it}"])
    assert got["gapped"] == bytes(G[f"gapped{it}"]) and got["consensus"] == bytes(G[f"cons{it}"]) and got["quality"] == bytes(G[f"qual{it}"])
    assert sorted(got["idx"] + got["excluded"]) == list(range(len(profs)))


@pytest.mark.parametrize("it", range(8))
def test_assemble_reference_host_logic(oracle_port, it):
    """The reference-guided branch of assemble() (orientation scores in one batch, threshold, ranking, iterative alignment against the
    growing alignment profile, consensus) with the DP served by the CPU oracle, against the sequence composed from the reference's
    functions (tests/golden/make_golden_assemble_reference.py)."""
    run_assemble_reference_case(OracleContext(oracle_port), it)


def test_onehot_profile_all_bytes(oracle_ref):
    """_createProfile(std::string) (src/align.h:119-136) for every byte value: A, C, G, T, N in either case and '-' have a row."""
    from tracy_b200 import msa as _msa
    p = _msa.onehot_profile(bytes(range(1, 256)))
    assert p.sum() == 11 and p[5, 0x2D - 1] == 1 and p[4, ord("n") - 1] == 1 and p[:, ord("x") - 1].sum() == 0
    if oracle_ref is not None:
        assert np.array_equal(oracle_ref.onehot(bytes(range(1, 256))), p)
