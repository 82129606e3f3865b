#!/usr/bin/env python
"""Full-size runs of BASELINE configs 4 and 5 on one GPU (timings + result dumps for offline comparison with the
reference fixtures):  python tools/probe_configs.py [m4096] [gibbs256] [gibbs1e5]"""
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from conftest import droplet_couplings  # noqa: E402
import tnac4o_b200  # noqa: E402

OUT = os.path.join(ROOT, 'gpurun_out')
os.makedirs(OUT, exist_ok=True)
J = droplet_couplings(2048)
what = sys.argv[1:] or ['m4096', 'gibbs256', 'gibbs1e5']


def new(beta):
    return tnac4o_b200.tnac4o(mode='Ising', Nx=16, Ny=16, Nc=8, J=J, beta=beta)


if 'm4096' in what:
    for rep in range(2):
        ins = new(3)
        t0 = time.time()
        ins.search_ground_state(M=2 ** 12, relative_P_cutoff=1e-8, Dmax=32)
        torch.cuda.synchronize()
        print('config4 M=4096 rep %d: wall %.3f s  rhoT %.3f  search %.3f  marginals %d -> %.3e /s  E=%.13f deg=%d logP=%.10f '
              'disc=%.6f neg=%.3e' % (rep, time.time() - t0, ins.stats['seconds_rhoT'], ins.stats['seconds_search'],
                                      ins.stats['marginals'], ins.stats['marginals'] / ins.stats['seconds_search'],
                                      ins.energy[0], ins.degeneracy, ins.probability[0], ins.discarded_probability,
                                      ins.negative_probability), flush=True)
    np.savez_compressed(os.path.join(OUT, 'gpu_l2048_m4096.npz'), energy=ins.energy, states=ins.states,
                        probability=ins.probability, degeneracy=ins.degeneracy)

if 'gibbs256' in what:
    ins = new(1)
    np.random.seed(1)
    ins.gibbs_sampling(M=256, Dmax=32)
    E = tnac4o_b200.energy_Jij(J, ins.binary_states())
    print('config5 M=256: rhoT %.3f sampling %.3f  <E>=%.6f  max|E - energy_Jij|=%.2e neg=%.3e' % (
        ins.stats['seconds_rhoT'], ins.stats['seconds_search'], ins.energy.mean(), np.max(np.abs(E - ins.energy)),
        ins.negative_probability), flush=True)
    np.savez_compressed(os.path.join(OUT, 'gpu_gibbs_l2048.npz'), energy=ins.energy, states=ins.states.astype(np.int16))

if 'gibbs1e5' in what:
    for rep in range(2):
        ins = new(1)
        np.random.seed(1)
        t0 = time.time()
        ins.gibbs_sampling(M=100000, Dmax=32)
        torch.cuda.synchronize()
        n = ins.stats.get('marginals', 0)
        print('config5 M=1e5 rep %d: wall %.3f s  rhoT %.3f  sampling %.3f  marginals %d -> %.3e /s  <E>=%.4f neg=%.3e  mem %.1f GB'
              % (rep, time.time() - t0, ins.stats['seconds_rhoT'], ins.stats['seconds_search'], n,
                 n / ins.stats['seconds_search'], ins.energy.mean(), ins.negative_probability,
                 torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
    sub = slice(0, 100000, 50)
    E = tnac4o_b200.energy_Jij(J, ins.binary_states()[sub])
    print('   max|E - energy_Jij| over 2000 samples = %.2e, unique states %d' % (
        np.max(np.abs(E - ins.energy[sub])), len(np.unique(ins.states[sub], axis=0))), flush=True)
