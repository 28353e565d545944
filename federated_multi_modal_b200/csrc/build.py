"""Builds libmfk.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link).

    python -m federated_multi_modal_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.environ.get("MFK_LIB_OUT") or os.path.join(PKG, "libmfk.so")
SOURCES = ["mfk_gemm.cu", "mfk_attn.cu", "mfk_elem.cu", "mfk_head.cu", "mfk_fed.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-cudart", "static"]


def _deps():
    hdrs = [os.path.join(HERE, "mfk_common.cuh"), os.path.join(os.path.dirname(PKG), "include", "mfk.h")]
    return [os.path.join(HERE, s) for s in SOURCES] + hdrs


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("MFK_DEFS", "").split()  # experiment builds (tools/ab_step.sh), e.g. -DMFK_NO_GTRACE
    if os.environ.get("MFK_LIB_OUT"):
        objdir = objdir + "_" + os.path.basename(LIB).replace(".", "_")
        os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    tmp = f"{LIB}.tmp.{os.getpid()}"  # link beside the target, then rename: a concurrent loader never sees a partial file
    cmd = [NVCC, "-shared", "-o", tmp, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
