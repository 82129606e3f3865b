#!/usr/bin/env python
"""Run a few representative primitive calls once (for ncu launch lists): python tools/run_one.py qr|svd|gemm|row"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from tnac4o_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'qr'
dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
if what == 'qr':
    A = up(rng.standard_normal((8192, 512)))
    for _ in range(2):
        ops.qr_pos(A.clone())
elif what == 'svd':
    U, _ = np.linalg.qr(rng.standard_normal((512, 512)))
    V, _ = np.linalg.qr(rng.standard_normal((512, 512)))
    C = up(np.triu((U * np.logspace(0, -40, 512)) @ V.T))
    for _ in range(2):
        ops.svd(C, want_vectors=True)
elif what == 'gemm':
    A, B = up(rng.standard_normal((8192, 512))), up(rng.standard_normal((512, 512)))
    for _ in range(3):
        ops.gemm(A, B)
    A, B = up(rng.standard_normal((128, 512))), up(rng.standard_normal((512, 8192)))
    for _ in range(3):
        ops.gemm(A, B)
elif what == 'search':
    from conftest import droplet_couplings
    import tnac4o_b200
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=8, Ny=8, Nc=8, J=droplet_couplings(512), beta=3)
    ins.search_ground_state(M=1024, relative_P_cutoff=1e-8, Dmax=16)
    print(ins.energy, ins.stats)
elif what == 'gibbs':
    # the branch-parallel kernels at config-5 size: right environments (grouped DMMA GEMM) and marginals for 10^5 samples.
    # The boundary MPS is built first; ncu --profile-from-start off captures only what follows cudaProfilerStart.
    from conftest import droplet_couplings
    import tnac4o_b200
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=16, Ny=16, Nc=8, J=droplet_couplings(2048), beta=3)
    ins._setup_rhoT(Dmax=32)
    built = ins._setup_rhoT
    ins._setup_rhoT = lambda **kw: None
    np.random.seed(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ins.Ny = 1                                  # one lattice row is enough for the capture
    ins.order = np.arange(16)
    ins.gibbs_sampling(M=100000, Dmax=32)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(ins.stats)
torch.cuda.synchronize()
print('done', what)
