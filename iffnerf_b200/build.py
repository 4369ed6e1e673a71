"""Build recipe for libtvm_b200.so (the C-ABI library of include/tvm_b200.h).

    python -m iffnerf_b200.build [-v]

nvcc cross-compiles for sm_100a without a GPU; the .so is written in-tree
(iffnerf_b200/libtvm_b200.so, git-ignored) so it travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtvm_b200.so")
SOURCES = ["pack.cu", "march.cu", "march_bwd.cu", "shade.cu", "shade_bwd.cu", "shade_tc.cu", "shade_tc3.cu", "shade_ref.cu", "query.cu", "raygen.cu", "gridops.cu", "collective.cu"]
BENCH_LIB = os.path.join(HERE, "libtvm_bench.so")      # measurement aids (include/tvm_bench.h), not part of the product library
BENCH_SOURCES = ["microbench.cu"]
HEADERS = ["tvm_math.cuh", "tvm_common.cuh", "tvm_gather.cuh", "tvm_warp.cuh", "tvm_tc.cuh", os.path.join("..", "..", "include", "tvm_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--expt-relaxed-constexpr",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libtvm_b200.so cannot be built")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + BENCH_SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(verbose=False, force=False, defines=(), out=None):
    """defines/out: build a tuning variant (e.g. defines=["TVM_MARCH_MIN_BLOCKS=6"]) next to the default library."""
    if out is None and os.environ.get("TVM_B200_LIB"):
        return os.environ["TVM_B200_LIB"]
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] + \
        ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtvm_b200.so")
    if out == LIB:
        res = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-o", BENCH_LIB] + [os.path.join(CSRC, s) for s in BENCH_SOURCES],
                             capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building libtvm_bench.so")
    return out


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force=True)
    print(LIB)
