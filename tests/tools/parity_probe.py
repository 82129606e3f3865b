#!/usr/bin/env python
"""Measured parity margins on the GPU box (what the assertions of tests/test_solver_gpu.py are set from):
    python tests/tools/parity_probe.py"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from conftest import SHAPES, droplet_couplings, golden  # noqa: E402
import tnac4o_b200  # noqa: E402
from oracle import RefSolver  # noqa: E402


def make(J, L, beta=3, cls=tnac4o_b200.tnac4o):
    Nx, Ny = SHAPES[L]
    return cls(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=beta)


# ---- config 2: L = 512, probability and per-site marginals against the oracle trace
J = droplet_couplings(512)
z = golden('ref_l512.npz')
ref = make(J, 512, cls=RefSolver)
trace = {}
ref.trace = lambda kind, **kw: trace.setdefault((kind, kw['ny'], kw['nx']), kw)
t0 = time.time()
ref.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
print('oracle L=512 %.1f s; oracle vs fixture log2P rel %.2e' % (time.time() - t0, abs(ref.probability[0] / z['gs_probability'][0] - 1)))
ins = make(J, 512)
ins.native_search = False
seen = {}
orig = ins._site_marginals


def spy(ws, br, RRat, ny, nx, want_P=False):
    P = orig(ws, br, RRat, ny, nx, want_P=True)
    seen[(ny, nx)] = (P.cpu().numpy(), br.vind[:br.n].cpu().numpy().copy())
    return None


ins._site_marginals = spy
ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
worst, frac = 0.0, 1.0
for (ny, nx), (P, vind) in seen.items():
    r = trace[('marginals', ny, nx)]
    rows = {tuple(v): i for i, v in enumerate(r['vind'].view(np.uint8).tolist())}
    common = [(i, rows[tuple(v)]) for i, v in enumerate(vind.tolist()) if tuple(v) in rows]
    frac = min(frac, len(common) / max(1, len(rows)))
    a, b = zip(*common)
    worst = max(worst, float(np.max(np.abs(P[list(a)] - r['P'][list(b)]))))
print('L=512: max |dP| over all sites %.3e, min common-branch fraction %.4f' % (worst, frac))
print('L=512: log2P rel err vs fixture %.3e, vs oracle here %.3e; E equal %s; states equal %s' % (
    abs(ins.probability[0] / z['gs_probability'][0] - 1), abs(ins.probability[0] / ref.probability[0] - 1),
    ins.energy[0] == ref.energy[0], np.array_equal(ins.states, ref.states)))

# ---- synthetic family A at L = 512 (the kind of instance that fills the bench batches)
rng = np.random.default_rng(7)
vals = np.array([v for _, _, v in J])
Js = [[i, j, float(v)] for (i, j, _), v in zip(J, rng.permutation(vals))]
ref = make(Js, 512, cls=RefSolver)
ref.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
ins = make(Js, 512)
ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
print('synthetic L=512: E equal %s (%.10f), states equal %s, deg %d/%d, log2P rel %.3e' % (
    ins.energy[0] == ref.energy[0], ins.energy[0], np.array_equal(ins.states, ref.states), ins.degeneracy, ref.degeneracy,
    abs(ins.probability[0] / ref.probability[0] - 1)))

# ---- config 4: which member of the degenerate pair
for M, fx in ((2 ** 10, 'ref_l2048.npz'), (2 ** 12, 'ref_l2048_m4096.npz')):
    z = golden(fx)
    J2 = droplet_couplings(2048)
    ins = make(J2, 2048)
    ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=32)
    bits = ins.binary_states()[0]
    print('config 4 M=%d: hamming to the reference state %d, E diff %.3e, log2P rel %.3e, deg %d' % (
        M, int(np.sum(bits != z['gs_bits'][0])), ins.energy[0] - z['gs_energy'][0], abs(ins.probability[0] / z['gs_probability'][0] - 1),
        ins.degeneracy))

# ---- Gibbs: identical samples
z = golden('ref_small.npz')
ins = make(droplet_couplings(128), 128, beta=1)
np.random.seed(1)
ins.gibbs_sampling(M=128, Dmax=16)
print('Gibbs L=128: identical samples %d / 128' % int(np.all(ins.states == z['gibbs_states'], axis=1).sum()))
z = golden('ref_gibbs_l2048.npz')
ins = make(droplet_couplings(2048), 2048, beta=1)
np.random.seed(1)
ins.gibbs_sampling(M=256, Dmax=32)
same = np.all(ins.states == z['states'], axis=1)
print('Gibbs L=2048: identical samples %d / 256; max |dE| on those %.3e' % (int(same.sum()), float(np.max(np.abs(ins.energy[same] - z['energy'][same])))))
