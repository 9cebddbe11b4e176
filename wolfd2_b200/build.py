"""Build libwolfd2_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwolfd2_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # parity build: no FMA contraction, so stencil results are bit-identical to a non-FMA CPU
    # build of the reference (gfortran -O2 on baseline x86-64 emits no FMA); IEEE div/sqrt.
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-ldl",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) + \
        [os.path.join(HERE, "..", "include", "wolfd2_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: an A/B variant of the same sources (e.g. extra=["-DSF_CPT=2"]), loaded with WOLFD2_B200_LIB=<out>."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(r.stderr)
    return out or LIB


HOST_BIN = os.path.join(HERE, "host", "wolfd2_host")


def build_host():
    """The compiled host driver (g++, links the in-tree library; no CUDA headers needed)."""
    src = os.path.join(HERE, "host", "wolfd2_host.cpp")
    if os.path.exists(HOST_BIN) and os.path.getmtime(HOST_BIN) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return HOST_BIN
    cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_BIN, src, "-L" + HERE, "-lwolfd2_b200", "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed for the host driver")
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
