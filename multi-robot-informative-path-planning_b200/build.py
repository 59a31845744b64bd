"""In-tree build of the sm_100a C-ABI library (libipp_b200.so) and the fp64 peak micro-benchmark.

nvcc cross-compiles without a GPU; the built artefacts are git-ignored but travel with the repo
snapshot to the GPU box.  Usage: python build.py [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libipp_b200.so")
PEAK = os.path.join(HERE, "fp64_peak")
SOURCES = ["ipp_linalg.cu", "ipp_posterior.cu", "ipp_branch.cu", "ipp_field.cu", "ipp_estimation.cu", "ipp_tree_host.cu"]
# every header under csrc/ (a change in any of them rebuilds) + the public C header
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "ipp_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
         "--expt-relaxed-constexpr", "--extended-lambda"]


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _stale(target, srcs):
    stamp = target + ".stamp"
    want = _digest(srcs)
    if os.path.exists(target) and os.path.exists(stamp) and open(stamp).read() == want:
        return None
    return want


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    want = _stale(LIB, deps)
    if force or want:
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs
        subprocess.run(cmd, check=True, cwd=CSRC)
        open(LIB + ".stamp", "w").write(want or _digest(deps))
    peak_src = os.path.join(CSRC, "fp64_peak.cu")
    want = _stale(PEAK, [peak_src])
    if force or want:
        subprocess.run([NVCC] + FLAGS[:5] + ["-o", PEAK, peak_src], check=True, cwd=CSRC)
        open(PEAK + ".stamp", "w").write(want or _digest([peak_src]))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
