"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; with no GPU the entry points fail loudly (no CPU fallback); the product never touches oracle/."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from tracy_b200 import build, capi
    build.build()
    return capi.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tracy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from tracy_b200 import capi
    decl = _declared_symbols()
    assert decl, "no declarations parsed from include/tracy_b200.h"
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/tracy_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == decl


def test_strerror_and_version(lib):
    assert lib.tb_strerror(0) == b"ok"
    assert b"sm_100a" in lib.tb_version()
    assert lib.tb_strerror(4).startswith(b"unsupported")


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    h = C.c_void_p()
    rc = lib.tb_ctx_create(C.byref(h), 0)
    assert rc == 2 and not h.value          # TB_ERR_CUDA, no context
    import tracy_b200
    with pytest.raises(tracy_b200.TracyError):
        tracy_b200.Context(0)
    # every compute entry point refuses a null context instead of computing anything
    assert lib.tb_gotoh_ps(None, None, tracy_b200.DnaScore().c(), tracy_b200.AlignConfig().c(), None) == 1
    assert lib.tb_decompose_sweep(None, None, None) == 1


def test_rows_from_ops_host_helper(lib, oracle_port):
    """tb_rows_from_ops is host code (O(L) formatting, reference src/align.h:196-293): check it against the oracle."""
    import numpy as np
    import tracy_b200
    from tracy_b200 import synth
    rng = np.random.default_rng(3)
    for it in range(30):
        m, n = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        p1 = synth.random_profile(rng, m, ["trace", "ties", "msa"][it % 3])
        seq = synth.random_seq(rng, n, b"ACGTNn-acgtRY")
        s, ops = oracle_port.gotoh_ps(p1, seq, 1, 0, (3, -5, -10, -4))
        assert tracy_b200.rows_from_ops("ps", p1, seq, ops) == oracle_port.rows_from_ops(p1, oracle_port.onehot(seq), ops)
        p2 = synth.random_profile(rng, n, "msa")
        s, ops = oracle_port.gotoh_pp(p1, p2, 1, 1, (3, -5, -10, -4))
        assert tracy_b200.rows_from_ops("pp", p1, p2, ops) == oracle_port.rows_from_ops(p1, p2, ops)
        a, b = synth.random_seq(rng, m), synth.random_seq(rng, n)
        s, ops = oracle_port.gotoh_ss(a, b, 0, 0, (3, -5, -10, -4))
        assert tracy_b200.rows_from_ops("ss", a, b, ops) == oracle_port.rows_from_ops(a, b, ops)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tracy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libgotoh_oracle" not in text and "libtracy_ref" not in text, f


def test_multi_partition_balances_cost(lib):
    """tb_multi_partition (host): contiguous pair ranges of equal DP cost for the devices of a tb_multi."""
    import numpy as np
    import ctypes as C
    from tracy_b200 import capi
    L = capi.lib()
    L.tb_multi_partition.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
    rng = np.random.default_rng(1)
    l1 = rng.integers(10, 2000, 1000).astype(np.int32)
    l2 = rng.integers(10, 6000, 1000).astype(np.int32)
    first = (C.c_size_t * 9)()
    assert L.tb_multi_partition(l1.ctypes.data, l2.ctypes.data, 1000, 8, first) == 0
    f = [int(x) for x in first]
    cost = (l1.astype(np.float64) + 1) * (l2 + 1)
    parts = [cost[a:b].sum() for a, b in zip(f, f[1:])]
    assert f[0] == 0 and f[-1] == 1000 and max(parts) / (cost.sum() / 8) < 1.15, parts
