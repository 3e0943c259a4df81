import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_b200():
    """True when the C ABI can open a context (a B200 is visible). No torch, no compute."""
    try:
        import tracy_b200
        c = tracy_b200.Context(0)
        c.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _have_b200():
        return
    skip = pytest.mark.skip(reason="no B200 visible: tracy_b200 has no CPU path (run with -m 'not gpu' here)")
    for it in gpu_items:
        it.add_marker(skip)


def load_gotoh_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "gotoh_golden.npz"))
    cases = []
    for i in range(int(z["n"])):
        kind = bytes(z[f"kind{i}"]).decode()
        a, b = z[f"a{i}"], z[f"b{i}"]
        if kind == "ss":
            a = bytes(a)
        if kind in ("ss", "ps"):
            b = bytes(b)
        hf, vf, ma, mi, go, ge, score = (int(x) for x in z[f"cfg{i}"])
        cases.append(dict(kind=kind, a=a, b=b, hf=hf, vf=vf, sc=(ma, mi, go, ge), score=score,
                          row0=bytes(z[f"row0_{i}"]), row1=bytes(z[f"row1_{i}"])))
    return cases


def load_decompose_golden(name="decompose_golden.npz"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    cases = []
    for i in range(int(z["n"])):
        trimL, trimR, maxindel, madc, bp, nref = (int(x) for x in z[f"cfg{i}"])
        cases.append(dict(row0=bytes(z[f"row0{i}"]), row1=bytes(z[f"row1{i}"]), pri=bytes(z[f"pri{i}"]), sec=bytes(z[f"sec{i}"]),
                          pri_out=bytes(z[f"pri_out{i}"]), sec_out=bytes(z[f"sec_out{i}"]), dcp=z[f"dcp{i}"],
                          trimL=trimL, trimR=trimR, maxindel=maxindel, madc=madc, bp=bp, nref=nref))
    return cases


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import loader
    return loader.port()


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle import loader
    return loader.ref()


@pytest.fixture(scope="session")
def ctx():
    import tracy_b200
    c = tracy_b200.Context(0)
    yield c
    c.close()
