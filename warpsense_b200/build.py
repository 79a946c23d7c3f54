"""Builds libwarpsense_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libwarpsense_b200.so")
SOURCES = ["capi.cu", "update_tsdf.cu", "registration.cu", "map_ops.cu", "preprocess.cu", "voxelgrid.cu", "hdf5_export.cu", "hdf5_import.cu"]
HEADERS = ["ws_common.cuh", "ws_internal.h", "march_math.cuh", os.path.join("..", "..", "include", "warpsense_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # the FP64/FP32 pose update must round like the host code
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, extra_flags=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> warpsense_b200/libwarpsense_b200.so"""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-ccbin", "/usr/bin/g++"] + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    build_library(force=True, verbose=True, extra_flags=tuple(sys.argv[1:]))
