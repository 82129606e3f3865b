#!/usr/bin/env python
"""Throughput of B concurrent ground-state searches on one GPU (one host thread + one stream each)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
import tnac4o_b200  # noqa: E402
from tnac4o_b200 import ops, parallel  # noqa: E402

C = dict(bench.CFG)
if os.environ.get('PROBE_L') == '512':
    C.update(L=512, Nx=8, Ny=8)
    bench.CFG.update(L=512, Nx=8, Ny=8)
for B in [int(x) for x in sys.argv[1:]] or [1, 2, 4]:
    inss = [tnac4o_b200.tnac4o(mode='Ising', Nx=C['Nx'], Ny=C['Ny'], Nc=C['Nc'], J=bench.instance_couplings(i), beta=C['beta'])
            for i in range(B)]
    for ins in inss:
        ins._site_tables()
    job = lambda ins: (lambda: ins.search_ground_state(M=C['M'], relative_P_cutoff=C['relative_P_cutoff'], Dmax=C['Dmax']))
    if not os.environ.get('PROBE_NOWARM'):
        parallel.run_concurrently([job(i) for i in inss])      # warm-up
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    t0 = time.perf_counter()
    parallel.run_concurrently([job(i) for i in inss])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('B=%d  wall %.3f s  -> %.3f s/instance  launches %d  E0=%.6f' % (B, dt, dt / B, ops.launch_count() - l0, inss[0].energy[0]), flush=True)
    del inss
    torch.cuda.empty_cache()
