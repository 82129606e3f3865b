import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def droplet_couplings(L, k=1):
    """coupling list of a droplet instance exactly as examples/e01 prepares it (load -> 0-based -> round to 1/75)"""
    z = golden('instances.npz')
    i, j, v = z['J_%d_%03d_i' % (L, k)], z['J_%d_%03d_j' % (L, k)], z['J_%d_%03d_v' % (L, k)]
    dJ = float(1 / 75)
    return [[int(a) - 1, int(b) - 1, round(float(c) / dJ) * dJ] for a, b, c in zip(i, j, v)]


def droplet_golden(L, k=1):
    z = golden('instances.npz')
    return float(z['gs_%d_%03d_energy' % (L, k)]), z['gs_%d_%03d_bits' % (L, k)]


def droplet_couplings10(L, k):
    """instances 001-010 of every size (instances10.npz stores 75 J as integers); identical floats to droplet_couplings"""
    z = golden('instances10.npz')
    dJ = float(1 / 75)
    return [[int(a) - 1, int(b) - 1, int(c) * dJ] for a, b, c in zip(z['J_%d_%03d_i' % (L, k)], z['J_%d_%03d_j' % (L, k)],
                                                                  z['J_%d_%03d_v75' % (L, k)])]


def droplet_golden10(L, k):
    z = golden('instances10.npz')
    key = 'gs_%d_%03d_bits' % (L, k)
    return float(z['gs_%d_%03d_energy' % (L, k)]), (z[key] if key in z.files else None)


SHAPES = {128: (4, 4), 512: (8, 8), 1152: (12, 12), 2048: (16, 16)}


@pytest.fixture(scope='session')
def J128():
    return droplet_couplings(128, 1)
