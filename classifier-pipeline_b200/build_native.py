"""Build ``libcptrack.so`` in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcptrack.so")
SOURCES = ["cptrack.cu", "extract_kernel.cu"]
HEADERS = ["cptrack_kernels.cuh", os.path.join("..", "..", "include", "cptrack.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))]
    cmd = [
        nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB,
    ] + (["-Xptxas", "-v"] if verbose else []) + srcs
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose="-v" in sys.argv))
