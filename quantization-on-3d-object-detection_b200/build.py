"""Build the C-ABI shared library (hand-written sm_100a CUDA kernels) in-tree with nvcc.

`python build.py` or `build.build()`; the product is qlidar/libqlidar_b200.so next to the Python mirror, so it
travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "qlidar", "libqlidar_b200.so")
SOURCES = ["voxelize.cu", "rulebook.cu", "spconv_mma.cu", "spconv_warp.cu", "elementwise.cu", "bev.cu", "smoothquant.cu", "centerhead.cu"]
HEADERS = ["ql_common.cuh", "ql_scan.cuh", os.path.join("..", "..", "include", "qlidar.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ablate: bool = False) -> str:
    """ablate=True builds the TEST-TIME variant libqlidar_b200_ablate.so (-DQL_SPCONV_ABLATE: the conv kernel's stages can be
    switched off one by one, tools/conv_sweep.py); the product library never contains those switches."""
    if ablate:
        out = OUT.replace(".so", "_ablate.so")
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        r = subprocess.run([nvcc] + NVCC_FLAGS + ["-DQL_SPCONV_ABLATE", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building the ablation variant")
        return out
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libqlidar_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ablate="--ablate" in sys.argv))
