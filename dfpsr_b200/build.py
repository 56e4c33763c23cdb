"""Builds dfpsr_b200/libdfpsr_b200.so from csrc/*.cu with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdfpsr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the reference is built without FMA contraction; bit-exact parity needs the same rounding
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-ldl",
]


def sources():
    # *.cpp: host-only arithmetic (SSE vector extensions the CUDA front end does not parse), handed to the host compiler as is
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cpp")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    stamp = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(HERE, "..", "include", "dfpsr_b200.h")]
    return any(os.path.getmtime(d) > stamp for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("DFPSR_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DRASTER_MIN_BLOCKS=6
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
