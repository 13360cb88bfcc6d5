#!/usr/bin/env python
"""Build a compile-time variant of the library next to the regular one:  python tools/build_variant.py NAME -DFLAG ...
-> smalltts_b200/variants/libsmalltts_b200_NAME.so (load it with STTS_LIB_PATH=...).  For A/B timing only."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smalltts_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(b.HERE, "variants")
os.makedirs(os.path.join(out_dir, name), exist_ok=True)
objs = []
for src in b.SOURCES:
    obj = os.path.join(out_dir, name, src.replace(".cu", ".o"))
    subprocess.run([b._nvcc(), *[f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c", os.path.join(b.CSRC, src), "-o", obj],
                   check=True)
    objs.append(obj)
lib = os.path.join(out_dir, f"libsmalltts_b200_{name}.so")
subprocess.run([b._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], check=True)
print(lib)
