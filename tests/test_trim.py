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


def test_native_trace_quality_goldens():
    """tb_trace_quality (csrc/trimq.cu: estimateQualities + trimTrace for one trace, host code) on the reference's goldens."""
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "trim_golden.json")))
    for i, c in enumerate(_gen.cases(31, len(want))):
        q, lr = trim.trace_quality(c["bcpos"], c["sec"], c["stringency"])
        assert [int(x) for x in q] == want[i]["qual"] and lr == (want[i]["left"], want[i]["right"]), i


def test_native_and_python_trace_quality_where_positions_wrap(oracle_ref):
    """Basecall positions that do not increase make the reference's uint32 peak distances wrap, its (int32_t) cast saturate and its
    int32 penalties overflow: the native function and the Python statement follow it there too (reference build present: compared
    with it; otherwise with each other)."""
    rng = np.random.default_rng(5)
    for it in range(400):
        n = int(rng.integers(2, 30)) if it % 5 == 0 else int(rng.integers(30, 900))
        pos = np.cumsum(rng.integers(1, 25, n)).astype(np.int32)
        if it % 3 == 0 and n > 5:
            pos[n // 2] = pos[n // 2 - 1] - int(rng.integers(0, 30))
        sec = bytes(rng.choice(list(b"ACGTNRYKM"), n, p=[.2, .2, .2, .2, .05, .04, .04, .04, .03]).astype(np.uint8))
        if it % 13 == 1:
            sec = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
        st = float([1, 2, 4, 9, 0.5, 3.3][it % 6])
        q2, l2 = trim.trace_quality(pos, sec, st)
        q1, l1 = trim.estimate_qualities(pos, sec), trim.trim_trace(pos, sec, st)
        assert np.array_equal(q1, q2) and tuple(l1) == tuple(l2), it
        if oracle_ref is not None:
            qr, _ = oracle_ref.estimate_qualities(pos, sec, sec)
            assert np.array_equal(qr, q2) and tuple(oracle_ref.trim_trace(pos, sec, st)) == tuple(l2), it
