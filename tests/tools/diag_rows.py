#!/usr/bin/env python
"""Row-by-row comparison of the boundary-MPS build, GPU path vs numpy oracle (run on the GPU box):
   sweeps and Schmidt-spectrum changes of every variational_compress call, bond dimensions, fidelity of rhoT[ny].
   python tests/tools/diag_rows.py [L] [Dmax]"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import SHAPES, droplet_couplings  # noqa: E402
import tnac4o_b200  # noqa: E402
from tnac4o_b200 import mps as gmps  # noqa: E402
from oracle import RefSolver  # noqa: E402
import oracle.mps_ref as omps  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
Nx, Ny = SHAPES[L]
J = droplet_couplings(L)

glog, olog = [], []


def g_var(self, phi, tol=None, max_sweeps=1, verbose=False):
    overlap = self.setup_RL_mix(phi)
    sweeps, diff, hist = 0, 1., []
    while diff > tol:
        if sweeps >= max_sweeps:
            break
        for n in range(self.L - 1, 0, -1):
            self.optimise_site(phi, n); self.orth_right(n); self.update_S(); self.update_RR_mix(phi, n)
        dmax = torch.zeros(1, dtype=torch.float64, device=self.device)
        for n in range(self.L):
            T1 = self.optimise_site(phi, n); self.orth_left(n)
            dmax = torch.maximum(dmax, self.update_S()); self.update_RL_mix(phi, n, T1)
        diff = float(dmax.item()); overlap = self.R[-1]; sweeps += 1; hist.append(diff)
    glog.append((max_sweeps, hist, list(self.D)))
    return float(overlap.item())


def o_var(self, phi, tol=None, max_sweeps=1):
    for n in range(self.L):
        self.push_left_env(phi, n)
    overlap = self.R[-1]
    sweeps, diff, hist = 0, 1.0, []
    while diff > tol:
        if sweeps >= max_sweeps:
            break
        for n in range(self.L - 1, 0, -1):
            self._fit_site(phi, n); self.orth_right(n); self.refresh_schmidt(); self.push_right_env(phi, n)
        diff = 0.0
        for n in range(self.L):
            self._fit_site(phi, n); self.orth_left(n)
            diff = np.maximum(diff, self.refresh_schmidt()); self.push_left_env(phi, n)
        overlap = self.R[-1]; sweeps += 1; hist.append(float(diff))
    olog.append((max_sweeps, hist, list(self.D)))
    return overlap


gmps.MPS.variational_compress = g_var
omps.RefMPS.variational_compress = o_var
ins = tnac4o_b200.tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=3)
ins.build_rhoT0 = True
ins.native_rows = False
ref = RefSolver(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=3)
ins._setup_rhoT(Dmax=D)
ref._setup_rhoT(Dmax=D)


def ov(x, y):
    E = np.ones((1, 1))
    for p, q in zip(x, y):
        E = np.einsum('ab,apc,bpd->cd', E, p, q)
        E /= np.max(np.abs(E))
    return E


def fidelity(x, y):
    def lognorm(a, b):
        E = np.ones((1, 1)); s = 0.0
        for p, q in zip(a, b):
            E = np.einsum('ab,apc,bpd->cd', E, p, q)
            m = np.max(np.abs(E)); E /= m; s += np.log(m)
        return s + np.log(abs(E.item())), np.sign(E.item())
    xy, sg = lognorm(x, y); xx, _ = lognorm(x, x); yy, _ = lognorm(y, y)
    return sg * np.exp(xy - 0.5 * (xx + yy))


k = 0
for ny in range(Ny - 1, -1, -1):
    a = [t.cpu().numpy() for t in ins.rhoT[ny].A]
    b = ref.rhoT[ny].A
    print('row %2d  1-fidelity(gpu,oracle) = %.3e   discarded gpu %.3e oracle %.3e   overlap gpu %.12f oracle %.12f' % (
        ny, 1 - fidelity(a, b), ins.rhoT_discarded[ny], ref.rhoT_discarded[ny], ins.rhoT_overlap[ny], ref.rhoT_overlap[ny]))
    for tag in ('4D-stage', 'final'):
        g, o = glog[k], olog[k]
        k += 1
        print('        %-8s gpu sweeps %d diffs %s | oracle sweeps %d diffs %s' % (
            tag, len(g[1]), ['%.2e' % d for d in g[1]], len(o[1]), ['%.2e' % d for d in o[1]]))
        if g[2] != o[2]:
            print('        bonds gpu', g[2], '\n        bonds ora', o[2])
