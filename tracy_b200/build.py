"""In-tree build of libtracy_b200.so (nvcc, sm_100a only). No JIT cache: the .so sits next to this file so it
travels with the repository snapshot to the GPU box."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtracy_b200.so")
SOURCES = ["capi.cu", "gotoh_general.cu", "gotoh_packed.cu", "gotoh_pp.cu", "sweep.cu", "profile_ops.cu", "anchor.cu", "fraction.cu", "trace_io.cu", "multi.cu", "writers.cu", "trimq.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "tracy_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "--threads", "4",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA library when sources are newer than the .so. Needs nvcc (no GPU needed)."""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
