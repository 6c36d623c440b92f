"""Builds chord_detection_b200/lib/libchordb200.so in-tree with nvcc for sm_100a (no GPU needed)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libchordb200.so")
SOURCES = ["api.cu", "he.cu", "esacf.cu", "iterf0.cu", "prime.cu", "ingest.cu", "chroma.cu", "comm.cu"]
# esacf.cu: the Levenberg-Marquardt fit and the lfilter recurrences mirror host (scipy) arithmetic,
# which has no fused multiply-add; the DFT loops there call fma() explicitly.
NO_FMAD = {"esacf.cu", "chroma.cu"}  # chroma.cu: exact decimal rounding, no contraction
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "--fmad=true",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "chordb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    for s in SOURCES:
        o = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        flags = list(NVCC_FLAGS)
        if s in NO_FMAD:
            flags[flags.index("--fmad=true")] = "--fmad=false"
        cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
        objs.append(o)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
