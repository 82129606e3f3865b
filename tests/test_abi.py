"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol include/tnac4o_b200.h
declares, the product fails loudly without a GPU, and the product never imports the oracle."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'tnac4o_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import tnac4o_b200._native as nat
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(nat.lib, name), name
    assert sorted(nat.EXPORTED) == declared
    assert nat.lib.tn_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import tnac4o_b200
    from conftest import droplet_couplings
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=droplet_couplings(128), beta=3)
    with pytest.raises(RuntimeError):
        ins.search_ground_state(M=16, Dmax=4)
    with pytest.raises(RuntimeError):
        tnac4o_b200.energy_Jij(droplet_couplings(128), [[1] * 128])


def test_product_does_not_import_oracle():
    code = "import sys; import tnac4o_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, '-c', code], check=True, cwd=ROOT)
    # nor does anything under include/, tools/ or examples/ (checker scripts that execute the oracle live under tests/tools/)
    for top in ('tnac4o_b200', 'include', 'tools', 'examples'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    assert 'oracle' not in open(os.path.join(dirpath, f)).read().replace('no oracle', ''), f


def test_host_model_tables_match_oracle(J128=None):
    """host-side prep of the product (tnac4o_b200/model.py) against the oracle's restatement, bit for bit"""
    import numpy as np
    from conftest import droplet_couplings
    from oracle import RefSolver
    from tnac4o_b200.model import IsingLattice, upper_triangular
    J = droplet_couplings(128)
    ref = RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=3)
    lat = IsingLattice(upper_triangular(J, 128), 4, 4, 8)
    for ny in range(4):
        for nx in range(4):
            for a, b in zip(lat.energy_tables(ny, nx), ref.energy_tables(ny, nx)):
                assert np.array_equal(a, b)
            Wc, dm, rm = lat.boltzmann(ny, nx, 3, ref.Xu[ny][nx], ref.Xl[ny][nx], ref.Xr[ny][nx], ref.Xd[ny][nx])
            Wr, dr, rr = ref.site_weights(ny, nx)
            assert np.array_equal(Wc, Wr) and np.array_equal(dm, dr) and np.array_equal(rm, rr)
            assert np.array_equal(lat.traced(Wc, dm, rm, 2 ** lat.sd[ny][nx], 2 ** lat.sr[ny][nx]), ref.traced_mpo(ny, nx))


def test_public_names_of_the_reference_package():
    """tnac4o/__init__.py:1-2 of the reference: the names a user script imports"""
    import numpy as np
    import tnac4o_b200
    for name in ('tnac4o', 'load', 'load_Jij', 'round_Jij', 'minus_Jij', 'Jij_f2p', 'energy_Jij', 'energy_RMF'):
        assert hasattr(tnac4o_b200, name), name
    # energy_RMF on the model of examples/e05_minimal_RMF.py (2 x 2 corner of it): one site table + one bond table
    J = {'fun': {1: np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]]), 2: np.array([-1.5, 0, 1.5])},
         'fac': {(0, 0, 0, 1): 1, (0, 0, 1, 0): 1, (0, 0): 2, (1, 1): 2}, 'N': np.zeros((2, 2), dtype=int) + 3, 'Nx': 2, 'Ny': 2}
    states = np.array([[0, 0, 0, 0], [0, 1, 2, 2], [2, 2, 0, 1]])
    assert np.allclose(tnac4o_b200.energy_RMF(J, states), [-3.0, 2.0, 2.5])
    assert abs(tnac4o_b200.round_Jij([[0, 1, 0.33]], 1 / 75)[0][2] - 25 / 75) < 1e-15
    assert tnac4o_b200.Jij_f2p([[1, 2, 0.5]]) == [[0, 1, 0.5]] and tnac4o_b200.minus_Jij([[0, 1, 0.5]]) == [[0, 1, -0.5]]


def test_rotate_graph_host_logic_matches_oracle():
    """rotate_graph (tnac4o.py:290-340) is host logic: lattice maps, rotated couplings and the per-cell divisions of the
    product equal the oracle's for every quarter turn, on a non-square lattice too"""
    import numpy as np
    import tnac4o_b200
    from conftest import droplet_couplings
    from oracle import RefSolver
    J = droplet_couplings(128)
    for shape in ((4, 4), (8, 2)):
        for rot in (1, 2, 3, 5):
            a = tnac4o_b200.tnac4o(mode='Ising', Nx=shape[0], Ny=shape[1], Nc=8, J=J, beta=3)
            b = RefSolver(mode='Ising', Nx=shape[0], Ny=shape[1], Nc=8, J=J, beta=3)
            a.rotate_graph(rot)
            b.rotate_graph(rot)
            assert (a.Nx, a.Ny, a.rotation) == (b.Nx, b.Ny, b.rotation)
            assert np.array_equal(a.order, b.order) and np.array_equal(a.order_i, b.order_i)
            assert (abs(a.J - b.J)).nnz == 0
            for ny in range(a.Ny):
                for nx in range(a.Nx):
                    assert np.array_equal(a.ind[ny][nx], b.ind[ny][nx])
                    assert np.array_equal(a.id[ny][nx], b.id[ny][nx]) and np.array_equal(a.ir[ny][nx], b.ir[ny][nx])
            assert np.array_equal(a.sd, b.sd) and np.array_equal(a.sr, b.sr)
            offs = a._key_offsets(a.Ny - 1, a.Nx - 1)                  # merge-key layout stays within 128 bits
            assert len(offs) == a.Nx + 1 and int(offs[-1]) < 128


def test_rotate_then_add_noise_matches_reference_fixture():
    """add_noise (tnac4o.py:917-941) draws one uniform per stored coupling in storage order: with the same global-RNG seed
    the couplings equal the reference's bit for bit, also after quarter turns (which rebuild the sparse matrix here) and on
    a non-square lattice; fixture written by the reference (tests/golden/make_golden.py noise)"""
    import numpy as np
    import tnac4o_b200
    from conftest import droplet_couplings, golden
    z = golden('ref_noise.npz')
    J = droplet_couplings(128)
    for shape in ((4, 4), (8, 2)):
        for rot in (0, 1, 2, 3):
            a = tnac4o_b200.tnac4o(mode='Ising', Nx=shape[0], Ny=shape[1], Nc=8, J=J, beta=3)
            if rot:
                a.rotate_graph(rot)
            np.random.seed(7)
            a.add_noise(amplitude=1e-7)
            c = a.J.tocoo()
            k = np.lexsort((c.col, c.row))
            tag = 'n_%dx%d_r%d_' % (shape[0], shape[1], rot)
            assert np.array_equal(c.row[k], z[tag + 'row']) and np.array_equal(c.col[k], z[tag + 'col'])
            assert np.array_equal(c.data[k], z[tag + 'val'])
            assert np.array_equal(a.order, z[tag + 'order'])


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every function include/tnac4o_b200.h declares to the reference interface it replaces"""
    import re
    from conftest import ROOT
    hdr = open(os.path.join(ROOT, 'include', 'tnac4o_b200.h')).read()
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    declared = set(re.findall(r'\b(tn_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) > 40
    assert sorted(n for n in declared if n not in doc) == []
