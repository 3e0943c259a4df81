"""tracy_b200/trim.py (penalty track, estimated base qualities, trimming heuristics) against the reference's findBestTraceSection /
estimateQualities (src/abif.h:164-253) and trimTrace (src/trim.h:35-99): goldens made by the reference and, where its build exists, a
larger differential run."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from tracy_b200 import trim

_spec = importlib.util.spec_from_file_location("make_golden_trim", os.path.join(ROOT, "tests", "golden", "make_golden_trim.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)


def _mine(c):
    q = trim.estimate_qualities(c["bcpos"], c["sec"])
    _, best, _ = trim.find_best_trace_section(c["bcpos"], c["sec"])
    left, right = trim.trim_trace(c["bcpos"], c["sec"], c["stringency"])
    kept = trim.trim_basecalls(c["nsamples"], c["bcpos"], q, c["sec"], c["sec"], c["sec"], c["tl"], c["tr"])
    return dict(qual=[int(x) for x in q], best=best, left=left, right=right, kept_bcpos=[int(x) for x in kept["bcpos"]], kept_primary=kept["primary"])


def test_trim_goldens():
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "trim_golden.json")))
    for i, c in enumerate(_gen.cases(31, len(want))):
        assert _mine(c) == want[i], i


def test_qualities_shape_of_the_scale():
    """A clean trace with regular peaks and a noisy tail: qualities stay within 0..60 and fall to 0 in the tail; the first window
    counts the distance from sample 0 to the first peak, so the first basecalls score lower too (and a few of them are trimmed);
    trimming with a finite stringency cuts the whole noisy tail."""
    n = 400
    bcpos = np.arange(n, dtype=np.int32) * 12 + 20
    sec = b"ACGT" * 85 + b"RYNKS" * 12
    q = trim.estimate_qualities(bcpos, sec)
    assert q.min() >= 0 and q.max() <= 60 and q[10:300].min() == 60 and q[-40:].max() == 0 and q[0] < 60
    left, right = trim.trim_trace(bcpos, sec, 2.0)
    assert left <= 20 and 60 <= right <= 80


def test_trim_differential(oracle_ref):
    if oracle_ref is None:
        pytest.skip("reference build not present")
    for i, c in enumerate(_gen.cases(7, 300)):
        assert _mine(c) == _gen.reference_outputs(oracle_ref, c), i


def test_nearest_snp_goldens(oracle_ref):
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "nearest_snp_golden.json")))
    for i, c in enumerate(_gen.snp_cases(32, len(want))):
        assert trim.nearest_snp(*c) == want[i], i
    if oracle_ref is not None:
        for i, c in enumerate(_gen.snp_cases(33, 1500)):
            assert trim.nearest_snp(*c) == oracle_ref.nearest_snp(*c), i
