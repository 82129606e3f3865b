#!/usr/bin/env python
"""examples/e06 over the first J124 C=8 instances: python tools/j124_sweep.py [n_instances] [D] [concurrent 0/1]
prints energy / degeneracy per rotation, the selected pair and the line of results_C8_J124.txt"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
from tnac4o_b200 import drivers  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8
conc = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_j124_sweep.npz'))
ok = 0
t_all = time.time()
for k in range(1, n + 1):
    J = [[int(i) - 1, int(j) - 1, float(v)] for i, j, v in zip(z['J_%03d_i' % k], z['J_%03d_j' % k], z['J_%03d_v' % k])]
    t0 = time.time()
    E, deg, per = drivers.search_gs_degeneracy(J, 8, 8, Nc=8, beta=0.75, D=D, M=2 ** 12, relative_P_cutoff=1e-8,
                                                precondition=True, concurrent=conc)
    want = z['results'][k - 1]
    good = (round(E) == int(want[1])) and (deg == int(want[2]))
    ok += good
    print('%03d  E=%.6f deg=%d  want %d %d  %s  %.1fs  per-rotation %s' % (k, E, deg, want[1], want[2], 'OK' if good else 'MISMATCH',
                                                                      time.time() - t0, per), flush=True)
print('%d / %d lines reproduced with D=%d in %.1f s' % (ok, n, D, time.time() - t_all))
