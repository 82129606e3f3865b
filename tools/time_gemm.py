#!/usr/bin/env python
"""CUDA-event timing of tn_gemm for the large boundary-MPS products: python tools/time_gemm.py   (TN_GEMM_TMA=0: cp.async path)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tnac4o_b200 import ops
dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
for M, N, K in [(8192, 512, 512), (512, 8192, 512), (8192, 8192, 512), (16384, 2048, 2048)]:
    A = torch.from_numpy(rng.standard_normal((M, K))).to(dev)
    B = torch.from_numpy(rng.standard_normal((K, N))).to(dev)
    out = torch.empty((M, N), dtype=torch.float64, device=dev)
    for _ in range(3):
        ops.gemm(A, B, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ops.gemm(A, B, out=out)
    e1.record(); e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / reps
    err = float((out - A @ B).abs().max())
    print('gemm %6d x %5d x %5d  TMA=%s : %8.1f us  %6.2f TFLOP/s  max err vs torch %.2e' % (
        M, N, K, os.environ.get('TN_GEMM_TMA', '1'), t * 1e6, 2.0 * M * N * K / t / 1e12, err), flush=True)
