"""The reference-guided branch of `tracy assemble` (drivers.assemble_reference) on the B200 against goldens composed from the
reference's functions (tests/golden/make_golden_assemble_reference.py). The same cases run on the CPU with the oracle as DP
provider in tests/test_glue.py. (This file sorts last on purpose: it was added after the round's GPU budget was spent.)"""
import pytest

from test_glue import run_assemble_reference_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("it", range(8))
def test_assemble_reference_gpu(ctx, it):
    run_assemble_reference_case(ctx, it)
