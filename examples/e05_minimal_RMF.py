#!/usr/bin/env python
"""Minimal Random-Markov-Field example on a 5 x 3 grid (examples/e05_minimal_RMF.py of the reference)."""
import argparse

import numpy as np

from _common import setup_logging


def toy_model():
    Nx, Ny = 5, 3
    fun = {1: np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]]),      # penalty for neighbouring variables that differ
           2: np.array([-1.5, 0, 1.5]), 3: np.array([1.25, 0, -1.25])}
    fac = {}
    for ny in range(Ny):
        for nx in range(Nx):
            fac[(ny, nx)] = 3 if ny == 1 else 2
            if nx + 1 < Nx:
                fac[(ny, nx, ny, nx + 1)] = 1
            if ny + 1 < Ny:
                fac[(ny, nx, ny + 1, nx)] = 1
    return {'fun': fun, 'fac': fac, 'N': np.zeros((Ny, Nx), dtype=int) + 3, 'Nx': Nx, 'Ny': Ny}


if __name__ == '__main__':
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument('-r', type=int, default=0)
    p.add_argument('-D', type=int, default=32)
    p.add_argument('-M', type=int, default=2 ** 10)
    p.add_argument('-P', type=float, default=1e-12)
    p.add_argument('-dE', type=float, default=3.1)
    p.add_argument('-hd', type=int, default=0)
    p.add_argument('-max_st', type=int, default=2 ** 20)
    p.add_argument('-ee', type=int, default=1, choices=[1, 2, 3])
    p.add_argument('-pre', dest='pre', action='store_true')
    p.set_defaults(pre=False)
    args = p.parse_args()
    setup_logging()
    import tnac4o_b200 as tnac4o
    from tnac4o_b200 import drivers
    J = toy_model()
    ins = drivers.minimal_RMF(J, J['Nx'], J['Ny'], rot=args.r, D=args.D, M=args.M, relative_P_cutoff=args.P,
                              excitations_encoding=args.ee, dE=args.dE, hd=args.hd, max_states=args.max_st, precondition=args.pre)
    ins.show_solution(state=False)
    print('Number of states :', len(ins.energy))
    print('Max energy difference against energy_RMF :', np.max(np.abs(tnac4o.energy_RMF(J, ins.states) - ins.energy)))
