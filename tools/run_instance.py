#!/usr/bin/env python
"""Solve bench instance(s) <ids...> one after another (debug helper): python tools/run_instance.py 3 4"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
import tnac4o_b200  # noqa: E402
C = bench.CFG
for i in [int(x) for x in sys.argv[1:]]:
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=C['Nx'], Ny=C['Ny'], Nc=C['Nc'], J=bench.instance_couplings(i), beta=C['beta'])
    ins.search_ground_state(M=C['M'], relative_P_cutoff=C['relative_P_cutoff'], Dmax=C['Dmax'])
    torch.cuda.synchronize()
    print('instance', i, 'E', ins.energy[0], 'deg', ins.degeneracy, 'logP', ins.probability[0], 'neg', ins.negative_probability,
          'stats', ins.stats, flush=True)
