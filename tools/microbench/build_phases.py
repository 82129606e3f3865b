#!/usr/bin/env python
"""Phase-timer build of the library (same sources, -DTN_PHASES): tools/microbench/lib_phases/libtnac4o_b200.so.
Run here (nvcc cross-compiles); the .so travels to the GPU box with the snapshot."""
import os
import subprocess
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tnac4o_b200'))
import build as b  # noqa: E402
out = os.path.join(HERE, 'lib_phases')
os.makedirs(out, exist_ok=True)
objs = []
for src in b.SOURCES:
    o = os.path.join(out, src.replace('.cu', '.o'))
    subprocess.run([b.NVCC] + b.FLAGS + ['-DTN_PHASES', '-c', os.path.join(b.CSRC, src), '-o', o], check=True)
    objs.append(o)
lib = os.path.join(out, 'libtnac4o_b200.so')
subprocess.run([b.NVCC, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'], check=True)
for o in objs:
    os.remove(o)
print('built', lib)
