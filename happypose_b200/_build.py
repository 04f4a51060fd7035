"""Builds libhpb200.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  `python -m happypose_b200._build`
rebuilds it; __graft_entry__.build() calls build().
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libhpb200.so")
SOURCES = ["hpb_api.cu", "hpb_raster.cu", "hpb_crop.cu", "hpb_pose.cu", "hpb_topk.cu", "hpb_icp.cu", "hpb_prologue.cu", "hpb_maxpool_tma.cu", "hpb_crop_tma.cu", "hpb_stem_tc.cu", "hpb_conv3x3_tc.cu"]

# -fmad=false: every fused multiply-add in the kernels is an explicit fmaf(), so the rasteriser's arithmetic is
# bit-identical to the CPU oracle the parity tests compare against.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhpb200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "hpb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    # the image exports CC=/opt/gcc/bin/gcc; let nvcc pick the system g++ it was validated with
    env = dict(os.environ)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
