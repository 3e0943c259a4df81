"""Text writers (tracy_b200/writers.py) against outputs of the reference's own plotAlignment / writeDecomposition
(tests/golden/make_golden_writers.py), byte for byte."""
import os

import numpy as np

from conftest import ROOT
from tracy_b200 import writers


def test_plot_alignment_and_decomposition_match_reference():
    G = np.load(os.path.join(ROOT, "tests", "golden", "writers_golden.npz"))
    for i in range(int(G["n"])):
        pos, rl, fw, score, key, ll = (int(x) for x in G[f"cfg{i}"])
        got = writers.plot_alignment(bytes(G[f"r0_{i}"]), bytes(G[f"r1_{i}"]), bytes(G[f"chr{i}"]), pos, rl, bool(fw), score, key, tuple(G[f"a1a2_{i}"]), ll)
        assert got.encode("latin-1") == bytes(G[f"txt{i}"]), i
    assert writers.write_decomposition(G["decomp"]).encode() == bytes(G["decomp_txt"])


def test_align_fasta_layout():
    # reference src/sage.h:326-339: ">" stem, row 0, ">" chr " (forward)" | " (reverse)", row 1
    assert writers.align_fasta("trace1", b"AC-GT", b"ACTGT", b"chr2", True) == ">trace1\nAC-GT\n>chr2 (forward)\nACTGT\n"
    assert writers.align_fasta("t", b"A", b"A", "x", False) == ">t\nA\n>x (reverse)\nA\n"


def _trace_writer_cases(seed, n, max_samples):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_trace_writers", os.path.join(ROOT, "tests", "golden", "make_golden_trace_writers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.cases(seed, n, max_samples)


def _mine(c):
    txt = writers.trace_txt(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"], c["tl"], c["tr"]).encode("latin-1")
    js = writers.trace_json(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"]).encode("latin-1")
    aj = b""
    if c["wellformed"]:
        aj = writers.trace_align_json(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"], c["row0"], c["row1"], b"chr7", c["pos"], c["fwd"]).encode("latin-1")
    return txt, js, aj


def test_trace_writers_match_reference_goldens():
    """P.abif (traceTxtOut), the basecall JSON (traceJsonOut) and P.json of `tracy align` (alignmentTracePadding +
    traceAlignJsonOut) byte for byte against outputs of the reference (tests/golden/make_golden_trace_writers.py)."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "trace_writers_golden.npz"))
    _, cs = _trace_writer_cases(77, int(G["n"]), 160)
    for i, c in enumerate(cs):
        txt, js, aj = _mine(c)
        assert txt == bytes(G[f"txt{i}"]), i
        assert js == bytes(G[f"json{i}"]), i
        assert aj == bytes(G[f"ajson{i}"]), i


def test_trace_writers_differential(oracle_ref):
    """More seeded cases against the reference build itself, where it exists."""
    import pytest
    if oracle_ref is None:
        pytest.skip("reference build not present")
    mod, cs = _trace_writer_cases(5, 150, 500)
    for i, c in enumerate(cs):
        assert _mine(c) == mod.reference_outputs(oracle_ref, c), i
