"""In-tree build of libpsydiff_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("PSY_NVCC", "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
SOURCES = ["psy_capi.cu", "psy_compile.cpp", "psy_sweep.cpp", "psy_primitives.cpp", "psy_codegen.cpp"]
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC,-pthread"]


def build_library(force: bool = False) -> str:
    out = os.path.join(_HERE, "libpsydiff_b200.so")
    srcs = [os.path.join(_HERE, "csrc", s) for s in SOURCES if os.path.exists(os.path.join(_HERE, "csrc", s))]
    deps = srcs + [os.path.join(_HERE, "csrc", h) for h in ("psy_launch.h", "psy_reduce.cuh", "psy_k2.cuh", "psy_internal.h")] + \
        [os.path.join(_HERE, "..", "include", "psydiff_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = [NVCC, "-shared"] + FLAGS + ["-o", out] + srcs + ["-ldl"]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build_library(force=True))
