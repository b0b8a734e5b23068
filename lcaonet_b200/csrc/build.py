"""Compile the CUDA sources in this directory into liblcao_b200.so (in-tree, sm_100a only).

    python -m lcaonet_b200.csrc.build [--force]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "liblcao_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["abi.cu", "graph_index.cu", "neighbor.cu", "geom_basis.cu", "edge_ops.cu", "pair_table.cu", "table_norm.cu", "threebody.cu", "threebody_mma.cu", "threebody_staged.cu", "gemm_simt.cu", "gemm_tc.cu",
           "linear.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-I" + os.path.join(ROOT, "include"), "-I" + HERE]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the lcao_b200 CUDA library cannot be built")
    return exe


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))) + ["../../include/lcao_b200.h"]
    for f in names:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    # (the include paths are absolute and differ between checkouts of the same tree — e.g. the GPU box's scratch copy —
    #  so they stay out of the digest: it identifies the SOURCES and the code-generation flags)
    h.update(" ".join(f for f in NVCC_FLAGS if not f.startswith("-I")).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ_DIR, "stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
