"""Compile libsfb200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

``python -m starfish_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["cov_build.cu", "chol.cu", "ozaki.cu", "upstream.cu", "band.cu", "capi.cu"]
HEADERS = ["sfb_internal.cuh", os.path.join("..", "..", "include", "sfb200.h")]
LIB = os.path.join(HERE, "libsfb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libsfb200.so cannot be built")
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, experiments=False):
    """Build the shared library if it is missing or older than its sources. Returns its path.

    ``experiments=True`` defines SFB_EXPERIMENTS: the A/B environment knobs (SFB_DEBUG_MODE, SFB_OUTER_TILES)
    exist only in such a build; the shipped library reads no environment variables."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS]
    out = LIB
    if experiments:   # separate file: load it with SFB200_LIB=<path> (starfish_b200/_lib.py)
        cmd += ["-DSFB_EXPERIMENTS"]
        out = os.path.join(HERE, "libsfb200_exp.so")
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--experiments" in sys.argv, verbose="-v" in sys.argv,
                experiments="--experiments" in sys.argv))
