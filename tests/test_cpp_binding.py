"""The C++ host layer include/tracy_b200.hpp: (CPU) it compiles on its own against the C ABI and links against the library;
(GPU) the drop-in check -- one translation unit with the UNMODIFIED reference headers and tracy_b200.hpp, reference calls on
the CPU against tracy_b200 calls on the B200 with the same argument objects (tests/cpp/dropin.cpp, built into
oracle/_ref/dropin_test by oracle/Makefile in the build container)."""
import os
import subprocess

import pytest

from conftest import ROOT


def test_header_compiles_and_links_standalone(tmp_path):
    from tracy_b200 import capi
    capi.lib()                                                     # the library must exist (no compute calls here)
    exe = str(tmp_path / "header_only")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "header_only.cpp"),
           "-o", exe, "-L", os.path.join(ROOT, "tracy_b200"), "-ltracy_b200", "-Wl,-rpath," + os.path.join(ROOT, "tracy_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_host_logic_against_reference(tmp_path):
    """The C++ glue above the batched DP calls (revSeqBasedOnDist with grouped trials, matchingTraces, msa = UPGMA + level-wise
    progressive alignment, assembleDenovo, assembleReference) with a CPU double of the context served by the reference's own gotohScore / gotoh, against the
    reference's functions: tests/cpp/hostlogic.cpp, compiled with the unmodified reference headers where they exist."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("the reference headers are only present in the build container")
    from tracy_b200 import capi
    capi.lib()
    exe = str(tmp_path / "hostlogic")
    cmd = ["g++", "-std=c++17", "-O2", "-fno-tree-vectorize", "-DNDEBUG", "-w", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", ref, "-I", ref + "/htslib",
           "-I", ref + "/xxsds/include", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "hostlogic.cpp"), "-o", exe,
           "-L", os.path.join(ROOT, "tracy_b200"), "-ltracy_b200", "-Wl,-rpath," + os.path.join(ROOT, "tracy_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "89 checks, 0 mismatches" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
def test_dropin_against_unmodified_reference_headers():
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_test is built where /root/reference exists and travels with the snapshot")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "0 mismatches" in r.stdout
