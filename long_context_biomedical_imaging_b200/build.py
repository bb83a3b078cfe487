"""In-tree build of the C-ABI shared library (nvcc cross-compiles sm_100a without a GPU).

    python -m long_context_biomedical_imaging_b200.build [--force]

Produces long_context_biomedical_imaging_b200/csrc/liblcbi_b200.so. The .so is git-ignored but travels to
the GPU box with the tree snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "liblcbi_b200.so")
SOURCES = ["capi.cu", "dense_attn_fwd.cu", "dense_attn_bwd.cu", "window_attn.cu", "window_attn_small.cu", "window_attn_tc.cu", "patch_embed.cu", "patch_embed_mma.cu", "patch_embed_tc.cu",
           "attn_merge.cu", "layer_norm.cu", "row_copy.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build liblcbi_b200.so")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    sources = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "lcbi_b200.h"))
    objs = []
    nvcc = _nvcc()
    log = []
    for src in sources:
        obj = src[:-3] + ".o"
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(r.stderr)
            if verbose:
                print(" ".join(cmd))
                print(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(CSRC, "build.log"), "a") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True)
    print("built", path)
