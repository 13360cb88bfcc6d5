"""Build libsmalltts_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python -m smalltts_b200.build [--force]

The library is plain CUDA C++ behind the C ABI of include/smalltts_b200.h; it links the CUDA runtime
statically and resolves cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, so it needs no libcuda at
link time and cross-compiles on a machine without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmalltts_b200.so")
# the parity build: the same sources with fp16 (11-bit significand) instead of bf16 GEMM / attention operands (csrc/op16.cuh)
LIB_TIGHT = os.path.join(HERE, "libsmalltts_b200_tight.so")
SOURCES = ["gemm.cu", "kernels.cu", "convnext_fused.cu", "ffn_fused.cu", "dit_chain.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("STTS_EXTRA_NVCC_FLAGS", "").split()  # e.g. -DSTTS_FUSED_TRACE for tools/trace_fused.py


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "smalltts_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, tight: bool = False) -> str:
    """Compile the library (tight=True: the fp16-operand parity build) unless it is newer than its sources."""
    lib = LIB_TIGHT if tight else LIB
    if not force and not _stale(lib):
        return lib
    obj_dir = os.path.join(CSRC, "tight") if tight else CSRC
    os.makedirs(obj_dir, exist_ok=True)
    extra = ["-DSTTS_OPERAND_F16"] if tight else []
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    cmd = [_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(obj_dir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tight=True))
