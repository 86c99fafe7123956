"""Compiles hgsfusion_b200/csrc/*.cu for sm_100a into hgsfusion_b200/libhgsfusion_b200.so (in-tree).

nvcc cross-compiles without a GPU.  -fmad=false: every FMA in the kernels is an explicit fmaf();
nothing else may be contracted, the reference's CPU rounding order is part of the contract.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libhgsfusion_b200.so")
SOURCES = ["pillar_path.cu", "contract_ops.cu", "pillarnet_ops.cu", "train_ops.cu", "hybrid_points.cu", "subm_conv.cu", "abi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]


def _newest_source() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files += [os.path.join(os.path.dirname(HERE), "include", h) for h in ("hgsfusion_b200.h", "hgsfusion_b200_debug.h")]
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("HGSF_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
