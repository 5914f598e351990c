"""Builds setok_b200/libsetok_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m setok_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsetok_b200.so")
SOURCES = ["core.cu", "gemm_tcgen05.cu", "attention.cu", "attention_tcgen05.cu", "attention_fullrow.cu", "rowops.cu", "dpc.cu", "dpc_fused.cu", "splice.cu", "preprocess.cu", "train.cu", "api.cu"]
HEADERS = ["common.cuh", "rowops.cuh", os.path.join("..", "..", "include", "setok_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale() and not os.environ.get("SETOK_BUILD_OUT"):
        return LIB
    nvcc = _nvcc()
    extra = os.environ.get("SETOK_NVCC_EXTRA", "").split()     # e.g. -DSETOK_ATTN_TRACE for tools/attn_timeline.py
    out_lib = os.environ.get("SETOK_BUILD_OUT") or LIB          # A/B and trace builds go elsewhere (load with SETOK_B200_LIB)
    objdir = os.path.join(HERE, "build" if out_lib == LIB else "build_" + os.path.basename(out_lib).replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"---- nvcc {s} ----\n{out}\n")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", out_lib, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return out_lib


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
