"""Build libagnpy_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libagnpy_b200.so"
LIB_CHECKED = PKG / "libagnpy_b200_checked.so"   # -DAGNPY_B200_CHECKS: index checks in the kernels
SOURCES = ["runtime.cu", "synch_ssc.cu", "ec.cu", "tau.cu", "bb.cu", "targets.cu", "integrate.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # numpy-identical rounding; FMAs only where written as fma()
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build(target=LIB):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "agnpy_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force=False, verbose=False, checked=False):
    """checked=True builds the variant with the kernels' AGB_CHECK index checks compiled in (a debugging
    aid, loaded instead of the product library when AGNPY_B200_CHECKED=1)"""
    target = LIB_CHECKED if checked else LIB
    if not force and not needs_build(target):
        return target
    flags = NVCC_FLAGS + (["-DAGNPY_B200_CHECKS"] if checked else [])
    cmd = [_nvcc(), *flags, "-o", str(target)] + [str(CSRC / s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libagnpy_b200.so")
    if verbose:
        print(res.stdout + res.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked="--checked" in sys.argv))
