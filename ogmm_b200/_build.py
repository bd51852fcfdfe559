"""Build the CUDA library in-tree: ``ogmm_b200/libogmm_b200.so`` (sm_100a only).

    python -m ogmm_b200._build          # rebuild if any source is newer than the library
    python -m ogmm_b200._build --force

nvcc cross-compiles without a GPU.  The library is a plain C-ABI shared object (no torch, no
pybind): see include/ogmm_b200.h.  It links the CUDA runtime statically, so it shares the driver's
primary context (and therefore device pointers and streams) with PyTorch in the same process.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libogmm_b200.so")
SOURCES = ["capi.cu", "knn.cu", "knn_sweep.cu", "knn_select.cu", "knn_tiles.cu", "knn_wide.cu", "knn_wide2.cu", "edge_conv.cu", "cluster.cu", "cluster_dsmem.cu", "sinkhorn.cu", "moments.cu", "moments_tc.cu", "moments_tma.cu", "moments_bwd.cu", "procrustes.cu", "procrustes_bwd.cu", "deepgmr_bwd.cu"]
# translation units whose kernels launch a follow-up grid from the device (CUDA dynamic parallelism): relocatable
# device code + the device runtime at link time
RDC_SOURCES = {"cluster.cu", "cluster_dsmem.cu", "sinkhorn.cu"}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC or put it on PATH)")
    return cand


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "ogmm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ogmm_b200.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(newest_header, os.path.getmtime(os.path.join(CSRC, src)))):
            objs.append(obj)          # up to date
            continue
        cmd = [nvcc, *NVCC_FLAGS, *(["-rdc=true"] if src in RDC_SOURCES else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(obj)
    tmp = LIB + ".tmp"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudadevrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
