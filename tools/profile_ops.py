#!/usr/bin/env python
"""Per-primitive device time of one ground-state search (CUDA events around every ops.* call).
    python tools/profile_ops.py [L] [Dmax] [M]      -> table on stdout (run on the GPU box)"""
import os
import sys
import time
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import SHAPES, droplet_couplings  # noqa: E402
import tnac4o_b200  # noqa: E402
from tnac4o_b200 import ops, mps, solver  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
M = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
log = defaultdict(list)


def wrap(mod, name, key):
    raw = getattr(mod, name)

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        t0 = time.perf_counter()
        e0.record()
        r = raw(*a, **k)
        e1.record()
        log[key(a, k)].append((e0, e1, time.perf_counter() - t0, ops.launch_count() - l0))
        return r
    setattr(mod, name, timed)


def shape_key(prefix):
    return lambda a, k: '%s %s' % (prefix, 'x'.join(str(int(s)) for s in a[0].shape))


wrap(ops, 'qr_pos', shape_key('qr'))
wrap(ops, 'svd', lambda a, k: 'svd%s %s' % ('' if k.get('want_vectors', True) else '_S', 'x'.join(str(int(s)) for s in a[0].shape)))
wrap(ops, 'gemm', lambda a, k: 'gemm')
wrap(ops, 'transpose', lambda a, k: 'transpose')
wrap(ops, 'mpo_apply', lambda a, k: 'mpo_apply')
wrap(ops, 'pow2_scale_', lambda a, k: 'pow2_scale')
wrap(ops, 'truncation_rank', lambda a, k: 'truncation_rank')
wrap(ops, 'diff_norm', lambda a, k: 'diff_norm')

Nx, Ny = SHAPES[L]
ins = tnac4o_b200.tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=droplet_couplings(L), beta=float(os.environ.get('TN_PROFILE_BETA', '3')))
ins.native_rows = False          # same kernel sequence as the native row driver, but every primitive call is visible here
ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)       # warm-up (not instrumented meaningfully)
log.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
rows = []
for key, v in log.items():
    dev = sum(a.elapsed_time(b) for a, b, _, _ in v) * 1e-3
    host = sum(h for _, _, h, _ in v)
    rows.append((dev, key, len(v), host, sum(n for _, _, _, n in v)))
rows.sort(reverse=True)
print('wall %.3f s  rhoT %.3f  search %.3f   E=%.10f' % (wall, ins.stats['seconds_rhoT'], ins.stats['seconds_search'], ins.energy[0]))
print('%-28s %8s %10s %10s %10s %9s' % ('op', 'calls', 'dev s', 'host s', 'launches', 'us/call'))
tot = 0
for dev, key, n, host, nl in rows[:40]:
    print('%-28s %8d %10.4f %10.4f %10d %9.1f' % (key, n, dev, host, nl, 1e6 * dev / n))
print('sum of device time over ops: %.3f s' % sum(r[0] for r in rows))
agg = defaultdict(float)
for dev, key, n, host, nl in rows:
    agg[key.split()[0]] += dev
print({k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
