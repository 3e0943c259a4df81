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
