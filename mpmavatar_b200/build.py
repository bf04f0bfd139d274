"""In-tree build of libmpm_b200.so (nvcc cross-compiles sm_100a without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libmpm_b200.so")
SOURCES = ["mpm_b200.cu", "mpm_kernels.cuh", "mpm_device.cuh", "mpm_mesh.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Wno-deprecated-declarations"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES] + [os.path.join(_HERE, "..", "include", "mpm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "mpm_b200.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
