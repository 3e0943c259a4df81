"""tracy_b200/variants.py against the reference's callVariants / insertVariant / variantType (src/variants.h:34-138): committed
goldens made by the reference (tests/golden/make_golden_variants.py) and, where the reference build exists, a larger
differential run."""
import importlib.util
import json
import os

import pytest

from conftest import ROOT
from tracy_b200 import variants

_spec = importlib.util.spec_from_file_location("make_golden_variants", os.path.join(ROOT, "tests", "golden", "make_golden_variants.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)


def _mine(als):
    var = []
    for r0, r1, c, p in als:
        variants.call_variants(r0, r1, c, p, var)
    return [[v["pos"], v["basenum"], v["gt"], v["chr"], v["ref"], v["alt"], variants.variant_type(v["ref"], v["alt"])] for v in var]


def test_call_variants_goldens():
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "variants_golden.json")))
    for i, als in enumerate(_gen.cases(11, len(want))):
        assert _mine(als) == want[i], i


def test_call_variants_edge_cases():
    assert _mine([(b"", b"", b"c", 5)]) == []                                   # empty alignment
    assert _mine([(b"---", b"ACG", b"c", 5)]) == []                             # the allele has no base at all
    assert _mine([(b"AC", b"AG", b"c", 0)]) == [[2, 2, 1, "c", "G", "C", "SNV"]]
    assert _mine([(b"A-C", b"AGC", b"c", 9)]) == [[10, 1, 1, "c", "AG", "A", "Deletion"]]   # basenum = bases of the allele seen when the run ends
    assert _mine([(b"AGC", b"A-C", b"c", 9)]) == [[10, 2, 1, "c", "A", "AG", "Insertion"]]
    assert _mine([(b"AG", b"A-", b"c", 9)]) == []                               # an insertion still open at the last base is dropped
    assert _mine([(b"GAC", b"GNC", b"c", 3)]) == []                             # a reference allele with N is not recorded
    two = [(b"AC", b"AG", b"c", 0)] * 2
    assert _mine(two) == [[2, 2, 2, "c", "G", "C", "SNV"]]                      # the same call from both alleles: gt 2


def test_call_variants_differential(oracle_ref):
    if oracle_ref is None:
        pytest.skip("reference build not present")
    for i, als in enumerate(_gen.cases(5, 300)):
        assert _mine(als) == [list(x) for x in oracle_ref.call_variants(als)], i
