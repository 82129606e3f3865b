"""CPU tests of mode='RMF' host logic (model tables, rotation, noise, adjacency helpers) against the unmodified reference
when it is mounted (build container), and self-consistency checks that run anywhere."""
import os
import sys

import numpy as np
import pytest

REF = '/root/reference'


def rmf_model():
    """the toy model of examples/e05_minimal_RMF.py:31-52"""
    Nx, Ny = 5, 3
    N = np.zeros((3, 5), dtype=int) + 3
    fun = {1: np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]]), 2: np.array([-1.5, 0, 1.5]), 3: np.array([1.25, 0, -1.25])}
    fac = {}
    for ny in range(Ny):
        for nx in range(Nx - 1):
            fac[(ny, nx, ny, nx + 1)] = 1
    for ny in range(Ny - 1):
        for nx in range(Nx):
            fac[(ny, nx, ny + 1, nx)] = 1
    for ny in range(Ny):
        for nx in range(Nx):
            fac[(ny, nx)] = 3 if ny == 1 else 2
    return {'fun': fun, 'fac': fac, 'N': N, 'Nx': Nx, 'Ny': Ny}


def dense_from_tables(lat, ny, nx, beta, Xu, Xl, Xr, Xd):
    E0, E1, E4, dmap, rmap = lat.exponents(ny, nx, beta)
    N, L1, L4 = len(E0), E1.shape[1], E4.shape[1]
    L2, L3 = int(lat.ld[ny, nx]), int(lat.lr[ny, nx])
    W = np.zeros((N, L1, L2, L3, L4))
    for s in range(N):
        w = np.exp((E0[s] + E1[s][:, None]) + E4[s][None, :])
        w = w * Xu[None, :L4] * Xl[:L1, None] * Xr[rmap[s]] * Xd[dmap[s]]
        W[s, :, dmap[s], rmap[s], :] = w
    return W


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference is mounted in the build container only')
@pytest.mark.parametrize('rot', [0, 1, 2, 3])
def test_rmf_tables_match_reference(rot):
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import tnac4o as ref
    import tnac4o_b200
    a = ref.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
    b = tnac4o_b200.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
    if rot:
        a.rotate_graph(rot=rot)
        b.rotate_graph(rot=rot)
    np.random.seed(5); a.add_noise(1e-3)
    np.random.seed(5); b.add_noise(1e-3)
    assert (a.Nx, a.Ny, a.rotation) == (b.Nx, b.Ny, b.rotation)
    assert np.array_equal(a.order, b.order) and np.array_equal(a.order_i, b.order_i)
    rng = np.random.default_rng(rot)
    for X in ('Xu', 'Xl', 'Xr', 'Xd'):
        g = rng.uniform(0.5, 2.0, size=getattr(a, X).shape)
        setattr(a, X, g.copy()); setattr(b, X, g.copy())
    for ny in range(a.Ny):
        for nx in range(a.Nx):
            want = a._peps_tensor(ny, nx)
            got = dense_from_tables(b.lat, ny, nx, b.beta, b.Xu[ny, nx], b.Xl[ny, nx], b.Xr[ny, nx], b.Xd[ny, nx])
            assert want.shape == got.shape
            np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)
            # energy tables reproduce _update_Eng for every (state, left, up) combination, bit for bit
            Es, Esl, Esu = b.lat.energy_tables(ny, nx)
            nst = a.Nx * a.Ny
            N = int(a.N[ny][nx])
            for s in range(N):
                for l in range(Esl.shape[1]):
                    for u in range(Esu.shape[1]):
                        st = np.zeros((1, nst), dtype=np.int8)
                        st[0, ny * a.Nx + nx] = s
                        if nx > 0:
                            st[0, ny * a.Nx + nx - 1] = l
                        if ny > 0:
                            st[0, (ny - 1) * a.Nx + nx] = u
                        mine = Es[s]
                        if nx > 0:
                            mine = mine + Esl[s, l]
                        if ny > 0:
                            mine = mine + Esu[s, u]
                        assert a._update_Eng(st, ny, nx)[0] == mine
            assert b._key_offsets(ny, nx)[-1] <= 128


def test_rmf_rotation_is_a_permutation_and_energy_rmf():
    import tnac4o_b200
    b = tnac4o_b200.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
    b.rotate_graph(rot=4)
    assert (b.Nx, b.Ny) == (5, 3) and np.array_equal(b.order, np.arange(15))
    assert b.J['fac'] == rmf_model()['fac']
    states = np.zeros((2, 15), dtype=np.int8)
    states[1, 7] = 2
    E = tnac4o_b200.energy_RMF(rmf_model(), states)
    assert E[0] == -8.75 and E[1] == -8.75 - 2.5 + 4


def test_rmf_adjacency_helpers():
    from tnac4o_b200.droplets import AdjacencyDroplets
    book = AdjacencyDroplets(2, 'RMF')
    book.set_grid(5, 3)
    one = (np.array([0, 1, 6]), np.array([1, 2, 3]))
    assert book.connected(*one) and not book.connected(np.array([0, 2]), np.array([1, 1]))
    assert book.overlap(one, (np.array([11]), np.array([1]))) and not book.overlap(one, (np.array([13]), np.array([1])))
    assert book.hamming(np.array([1, 2, 3])) == 4
    assert book.hamming_between(one, (np.array([1, 6, 7]), np.array([2, 1, 1]))) == 3
