#!/usr/bin/env python
"""Cycle breakdown of the QR panel column loop (phase-timer build) and old/new Jacobi SVD timings.
    TNAC4O_B200_LIB=tools/microbench/lib_phases/libtnac4o_b200.so python tools/microbench/phases.py"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tnac4o_b200 import ops  # noqa: E402
from tnac4o_b200._native import lib  # noqa: E402

dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
names = ['load', 'dots', 'butterfly+stage', 'syncthreads1', 'dsmem stores', 'cluster.sync', 'params+T', 'syncthreads2',
         'apply+rotate', 'store']
if hasattr(lib, 'tn_debug_phases'):
    lib.tn_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
    buf = (ctypes.c_longlong * 16)()
    for m, n in [(8192, 512), (2048, 512), (512, 32)]:
        A = torch.from_numpy(rng.standard_normal((m, n))).to(dev)
        ops.qr_pos(A.clone())
        lib.tn_debug_phases(buf, 1)
        ops.qr_pos(A.clone())
        lib.tn_debug_phases(buf, 1)
        cols = min(m, n)
        print('qr %d x %d: cycles per column by phase (thread 0 of CTA 0)' % (m, n))
        for k, nm in enumerate(names):
            print('   %-18s %9.1f' % (nm, buf[k] / cols))
        print('   %-18s %9.1f' % ('total', sum(buf[:10]) / cols), flush=True)


def timeit(f, reps=10):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for k, want, decades in [(32, False, 25), (64, True, 25), (128, True, 25), (128, False, 25), (256, True, 25), (512, True, 40),
                         (512, True, 25), (512, True, 3)]:
    U, _ = np.linalg.qr(rng.standard_normal((k, k)))
    V, _ = np.linalg.qr(rng.standard_normal((k, k)))
    C = torch.from_numpy(np.ascontiguousarray(np.triu((U * np.logspace(0, -decades, k)) @ V.T))).to(dev)
    t = timeit(lambda: ops.svd(C, want_vectors=want))
    print('svd %4d vectors=%-5s decades=%2d: %9.1f us  sweeps %d  (TN_SVD_V1=%s)' % (
        k, want, decades, t, ops.last_svd_sweeps, os.environ.get('TN_SVD_V1', '0')), flush=True)
