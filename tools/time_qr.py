#!/usr/bin/env python
"""CUDA-event timing of tn_qr_pos / tn_svd for a few shapes (no profiler): python tools/time_qr.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tnac4o_b200 import ops
dev = torch.device('cuda', 0)
torch.cuda.set_stream(torch.cuda.Stream())      # a non-default stream, as the solver threads use
rng = np.random.default_rng(0)
def timeit(f, reps=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.launch_count()
    e0.record()
    for _ in range(reps): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps, (ops.launch_count() - l0) / reps
for m, n in [(8192, 16), (8192, 32), (8192, 128), (8192, 512), (2048, 512), (512, 512), (512, 16), (512, 32), (512, 128), (2048, 128)]:
    A = torch.from_numpy(rng.standard_normal((m, n))).to(dev)
    buf = torch.empty_like(A)
    def f():
        buf.copy_(A)
        ops.qr_pos(buf)
    t, nl = timeit(f)
    print('qr %5d x %4d : %9.1f us  %6.1f launches  (%.2f us per column)' % (m, n, t, nl, t / min(m, n)), flush=True)
for k, want in [(32, False), (128, False), (64, True), (128, True), (256, True), (512, True)]:
    U, _ = np.linalg.qr(rng.standard_normal((k, k))); V, _ = np.linalg.qr(rng.standard_normal((k, k)))
    C = torch.from_numpy(np.ascontiguousarray(np.triu((U * np.logspace(0, -25, k)) @ V.T))).to(dev)
    t, nl = timeit(lambda: ops.svd(C, want_vectors=want), reps=10)
    print('svd %4d (vectors=%s) graded: %9.1f us  %6.1f launches' % (k, want, t, nl), flush=True)
