"""CPU tests of the product's host-side droplet structure for excitations_encoding = 2, 3 (tnac4o_b200/droplets.py).

The class receives, per site, exactly what the device search hands it (old branch of every winner; old branch, dE and
XOR difference of every merged-away branch).  Here the same records come from the oracle's search loop (trace hook), so
the whole host logic -- connectivity, overlap, shape dictionary, hierarchy, enumeration, rotation back to the model's
order, save / load -- is checked against the reference fixtures without a GPU."""
import os
import warnings

import numpy as np
import pytest

from conftest import droplet_couplings, golden
from oracle import RefSolver
from tnac4o_b200.droplets import AdjacencyDroplets

warnings.filterwarnings('ignore')


def run_with_book(J, shape, ee, rot, hd, M, D, dE):
    ins = RefSolver(mode='Ising', Nx=shape[0], Ny=shape[1], Nc=8, J=J, beta=3)
    if rot:
        ins.rotate_graph(rot)
    book = AdjacencyDroplets(ee)
    book.set_adjacency(ins.J, [ins.ind[ny][nx] for ny in range(ins.Ny) for nx in range(ins.Nx)])

    def hook(kind, **kw):
        if kind == 'droplets':
            book.site_update(kw['winner_parent'], kw['merged'], dE, hd)
            book.end_of_site()
        elif kind == 'row_end':
            book.end_of_row()
    ins.trace = hook
    ins.search_low_energy_spectrum(excitations_encoding=ee, M=M, relative_P_cutoff=1e-8, Dmax=D, max_dEng=dE, lim_hd=hd)
    book.finish(ins.order_i, hd)
    book.set_adjacency(ins.J0, [ins.ind0[ny][nx] for ny in range(ins.Ny_model) for nx in range(ins.Nx_model)])
    return ins, book


def decode(book, ground, E0, dE):
    Eng, flip = book.unpack(dE, 2 ** 20)
    states = np.repeat(ground[None, :], len(Eng), axis=0)
    for i, keys in enumerate(flip):
        for k in keys:
            dpos, dstate = book.d[k]
            states[i, dpos] = np.bitwise_xor(states[i, dpos], dstate)
    order = np.lexsort(states.T[::-1])
    return states[order], (Eng + E0)[order]


@pytest.mark.parametrize('ee,rot,hd', [(2, 0, 0), (2, 2, 0), (3, 0, 0), (3, 3, 0), (2, 0, 4), (3, 0, 4), (2, 1, 0), (3, 1, 0)])
def test_book_reproduces_reference_spectrum(J128, ee, rot, hd):
    z = golden('ref_encodings.npz')
    tag = 'ee%d_r%d_hd%d' % (ee, rot, hd)
    ins, book = run_with_book(J128, (4, 4), ee, rot, hd, 1024, 16, 1.0)
    assert len(book.d) == int(z[tag + '_n_shapes']) and len(book.el) == int(z[tag + '_n_first_layer'])
    np.testing.assert_allclose(sorted(e[0][0] for e in book.el), z[tag + '_first_layer_dE'], atol=1e-10)
    states, energy = decode(book, ins.states[0], ins.energy[0], 1.0)
    assert np.array_equal(states, z[tag + '_states'])
    np.testing.assert_allclose(energy, z[tag + '_energy'], atol=1e-10)
    # the oracle's own structure is the same, shape by shape
    assert sorted((tuple(p), tuple(int(x) for x in s)) for p, s in book.d.values()) == \
        sorted((tuple(p), tuple(int(x) for x in s)) for p, s in ins.d.values())


@pytest.mark.parametrize('ee', [2, 3])
def test_book_L512_hierarchy(ee):
    z = golden('ref_encodings.npz')
    ins, book = run_with_book(droplet_couplings(512), (8, 8), ee, 0, 0, 256, 8, 0.5)
    states, energy = decode(book, ins.states[0], ins.energy[0], 0.5)
    assert len(energy) == 302 and np.array_equal(states, z['L512_ee%d_states' % ee])
    np.testing.assert_allclose(energy, z['L512_ee%d_energy' % ee], atol=1e-10)


def test_geometry_primitives():
    """a 2 x 1 lattice of 2-spin cells: spins 0-1 | 2-3, couplings 0-1, 1-2 (cells touch through spins 1 and 2)"""
    import scipy.sparse
    J = scipy.sparse.lil_matrix((4, 4))
    J[0, 1] = J[1, 2] = 1.0
    book = AdjacencyDroplets(2)
    book.set_adjacency(J.tocsr(), [[0, 1], [2, 3]])
    one = lambda pos, pat: (np.array(pos), np.array(pat, dtype=np.int8))
    assert list(book.spins(*one([0, 1], [3, 1]))) == [0, 1, 2]
    assert book.connected(*one([0], [3])) and book.connected(*one([0, 1], [2, 1]))
    assert not book.connected(*one([0, 1], [1, 1]))              # spins 0 and 2 are not coupled
    assert not book.connected(*one([0, 1], [2, 2]))              # spin 3 is isolated
    assert book.overlap(one([0], [2]), one([1], [1])) and not book.overlap(one([0], [1]), one([1], [1]))
    pos, pat = book.combine(one([0, 1], [3, 1]), one([1], [1]))
    assert list(pos) == [0] and list(pat) == [3]
    pos, pat = book.combine(one([0], [1]), one([0, 1], [2, -128]))
    assert list(pos) == [0, 1] and list(pat) == [3, -128]
    assert book.hamming_between(one([0, 1], [3, 1]), one([1], [3])) == 3
    k = book.key_of(*one([0, 1], [3, 1]))
    assert book.key_of(*one([0, 1], [3, 1])) == k and book.key_of(*one([0], [3])) != k


def test_solver_object_decodes_saved_structure(J128, tmp_path):
    """save() / load() carry the adjacency for encodings 2 and 3 (tnac4o.py:61-72, 226-231) and _exc_unpack of a loaded
    object enumerates the same spectrum (host-only part of decode_low_energy_states)"""
    import tnac4o_b200
    z = golden('ref_encodings.npz')
    ins, book = run_with_book(J128, (4, 4), 3, 1, 0, 1024, 16, 1.0)
    sol = tnac4o_b200.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J128, beta=3)
    sol.excitations_encoding = 3
    sol.d, sol.invd, sol.el, sol.free_d, sol.adj = book.d, book.invd, book.el, book.free_d, book.adj
    sol.energy, sol.states = ins.energy, ins.states
    fn = os.path.join(tmp_path, 'spectrum.npy')
    sol.save(fn)
    back = tnac4o_b200.load(fn)
    assert back.excitations_encoding == 3 and back.adj.shape == (128, 128)
    Eng, flip = back._exc_unpack(max_dEng=1.0, max_states=2 ** 20)
    np.testing.assert_allclose(np.sort(Eng + back.energy[0]), np.sort(z['ee3_r1_hd0_energy']), atol=1e-10)
