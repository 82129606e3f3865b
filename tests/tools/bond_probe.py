#!/usr/bin/env python
"""Bond dimensions and discarded weights of every boundary-MPS row, GPU path against the oracle (L=128 / 512):
    python tests/tools/bond_probe.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from conftest import SHAPES, droplet_couplings  # noqa: E402
import tnac4o_b200  # noqa: E402
from oracle import RefSolver  # noqa: E402

for L, D in ((128, 8), (128, 48), (512, 16), (512, 32)):
    Nx, Ny = SHAPES[L]
    J = droplet_couplings(L)
    ref = RefSolver(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=3)
    ref._setup_rhoT(Dmax=D)
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=3)
    ins.build_rhoT0 = True
    ins._setup_rhoT(Dmax=D)
    same = 0
    for ny in range(Ny):
        a = [ins.rhoT[ny].A[0].shape[0]] + [t.shape[2] for t in ins.rhoT[ny].A]
        b = [ref.rhoT[ny].A[0].shape[0]] + [t.shape[2] for t in ref.rhoT[ny].A]
        same += (a == b)
        if a != b:
            print('  L=%d D=%d row %d: gpu %s\n                     ref %s' % (L, D, ny, a, b))
    dd = max(abs(ins.rhoT_discarded[ny] - ref.rhoT_discarded[ny]) / max(ref.rhoT_discarded[ny], 1e-300) for ny in range(Ny))
    do = max(abs(ins.rhoT_overlap[ny] / ref.rhoT_overlap[ny] - 1) for ny in range(Ny))
    print('L=%d Dmax=%d: rows with identical bond dimensions %d / %d, max rel diff discarded %.2e, overlap %.2e' % (L, D, same, Ny, dd, do))
