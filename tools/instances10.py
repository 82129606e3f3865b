#!/usr/bin/env python
"""Droplet instances 001-010 of every lattice size (the 10-instance mean) and the synthetic families
B (uniform couplings) and C (J124) at L=2048: found energy against groundstates_otn2d.txt, seconds per instance one at a
time and with the ten running concurrently.      python tools/instances10.py [D] [precondition 0/1]  -> JSON on stdout"""
import json
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
from conftest import SHAPES, droplet_couplings10, droplet_golden10  # noqa: E402
import tnac4o_b200  # noqa: E402
from tnac4o_b200 import parallel  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
PRE = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device('cuda', 0)


def solve(J, L, beta=3.0, M=2 ** 10):
    Nx, Ny = SHAPES[L]
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=beta, device=dev)
    if PRE:
        ins.precondition(mode='balancing')
    ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
    return ins


def family(kind, seed):
    """synthetic couplings on the pattern of chimera2048 #1: B uniform J in [-1, 1] and fields in [-0.2, 0.2] rounded to 1/75,
    C couplers uniform in {+-1, +-2, +-4} without fields"""
    rng = np.random.default_rng(seed)
    out = []
    for i, j, _ in droplet_couplings10(2048, 1):
        if kind == 'B':
            v = rng.uniform(-0.2, 0.2) if i == j else rng.uniform(-1, 1)
            out.append([i, j, round(v * 75) / 75])
        elif i != j:
            out.append([i, j, float(rng.choice([-4, -2, -1, 1, 2, 4]))])
    return out


res = {'Dmax': D, 'precondition': PRE, 'M': 1024, 'beta': 3.0}
solve(droplet_couplings10(128, 1), 128)                  # warm-up
for L in (128, 512, 1152, 2048):
    Js = [droplet_couplings10(L, k) for k in range(1, 11)]
    hits, t_each = [], []
    for k, J in enumerate(Js, 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ins = solve(J, L)
        torch.cuda.synchronize()
        t_each.append(time.perf_counter() - t0)
        e, bits = droplet_golden10(L, k)
        hits.append({'instance': k, 'energy': float(ins.energy[0]), 'golden': e, 'delta': float(ins.energy[0] - e),
                     'state_equals_golden': None if bits is None else bool(np.array_equal(ins.binary_states()[0], bits)),
                     'degeneracy': int(ins.degeneracy)})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    parallel.run_concurrently([(lambda J=J: solve(J, L)) for J in Js], device=dev)
    torch.cuda.synchronize()
    t_conc = (time.perf_counter() - t0) / 10
    res['L%d' % L] = {'mean_seconds_one_at_a_time': float(np.mean(t_each)), 'seconds_per_instance_10_concurrent': t_conc,
                      'golden_energy_reached': int(sum(abs(h['delta']) < 1e-5 for h in hits)),
                      'below_golden': int(sum(h['delta'] < -1e-5 for h in hits)), 'instances': hits}
    print('L=%d: mean %.3f s, concurrent %.3f s/instance, golden energy reached on %d / 10' % (
        L, np.mean(t_each), t_conc, res['L%d' % L]['golden_energy_reached']), file=sys.stderr, flush=True)
for kind in ('B', 'C'):
    Js = [family(kind, s) for s in range(4)]
    beta = 3.0 if kind == 'B' else 0.75
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = parallel.run_concurrently([(lambda J=J: solve(J, 2048, beta=beta)) for J in Js], device=dev)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / len(Js)
    ok = [bool(abs(tnac4o_b200.energy_Jij(J, x.binary_states()[:1])[0] - x.energy[0]) < 1e-6) for J, x in zip(Js, out)]
    res['family_%s_L2048' % kind] = {'beta': beta, 'seconds_per_instance_4_concurrent': t, 'energies': [float(x.energy[0]) for x in out],
                                     'degeneracies': [int(x.degeneracy) for x in out], 'energy_self_consistent': ok}
    print('family %s: %.3f s/instance' % (kind, t), file=sys.stderr, flush=True)
print(json.dumps(res))
