"""GPU tests of mode='RMF' (Random Markov Field, tnac4o.py:160-163 and every `elif self.mode == 'RMF'` branch) against
fixtures written by the unmodified reference (tests/golden/make_golden.py rmf): the known answers of
examples/test_examples.py:107-136 (26 states below dE = 3.1 for all three encodings) plus ground-state search and Gibbs."""
import numpy as np
import pytest

from conftest import golden
from test_rmf_host import rmf_model

pytestmark = pytest.mark.gpu


def sorted_rows(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize('ee,rot', [(1, 0), (1, 1), (2, 2), (3, 3)])
def test_rmf_low_energy_spectrum(ee, rot):
    import tnac4o_b200
    z = golden('ref_rmf.npz')
    tag = 'ee%d_r%d' % (ee, rot)
    ins = tnac4o_b200.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
    if rot:
        ins.rotate_graph(rot=rot)
    if ee > 1:
        np.random.seed(7)
        ins.add_noise(amplitude=1e-7)
    ins.search_low_energy_spectrum(excitations_encoding=ee, M=1024, relative_P_cutoff=1e-12, Dmax=32, max_dEng=3.1, lim_hd=0)
    assert ins.energy[0] == z[tag + '_gs_energy'][0]
    assert np.array_equal(ins.states[0], z[tag + '_gs_states'][0])
    assert abs(ins.probability[0] - z[tag + '_gs_probability'][0]) <= 1e-8 * abs(z[tag + '_gs_probability'][0]) + 1e-12
    assert len(ins.d) == int(z[tag + '_n_shapes'])
    ins.decode_low_energy_states(max_dEng=3.1, max_states=100)
    assert len(ins.energy) == 26                                   # test_examples.py:110
    np.testing.assert_allclose(np.sort(ins.energy), np.sort(z[tag + '_energy']), rtol=0, atol=1e-12)
    assert np.array_equal(sorted_rows(ins.states), sorted_rows(z[tag + '_states']))
    check = tnac4o_b200.energy_RMF(rmf_model(), ins.states)
    np.testing.assert_allclose(np.sort(check), np.sort(z[tag + '_energy_check']), atol=1e-12)
    assert np.max(np.abs(check - ins.energy)) < 1e-4               # test_examples.py:127-136


@pytest.mark.parametrize('rot', [0, 1])
def test_rmf_ground_state(rot):
    import tnac4o_b200
    z = golden('ref_rmf.npz')
    tag = 'gs_r%d' % rot
    ins = tnac4o_b200.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
    if rot:
        ins.rotate_graph(rot=rot)
    ins.search_ground_state(M=64, relative_P_cutoff=1e-12, Dmax=32)
    n = len(z[tag + '_energy'])
    assert len(ins.energy) == n and np.array_equal(ins.energy, z[tag + '_energy'])
    assert np.array_equal(ins.states, z[tag + '_states']) and ins.degeneracy == int(z[tag + '_degeneracy'])
    np.testing.assert_allclose(ins.probability, z[tag + '_probability'], rtol=1e-8, atol=1e-12)
    assert abs(ins.discarded_probability - float(z[tag + '_discarded'])) < 1e-6 * abs(float(z[tag + '_discarded']))


def test_rmf_gibbs():
    import tnac4o_b200
    z = golden('ref_rmf.npz')
    ins = tnac4o_b200.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=1)
    np.random.seed(3)
    ins.gibbs_sampling(M=64, Dmax=32)
    assert np.array_equal(ins.states, z['gibbs_states']) and np.array_equal(ins.energy, z['gibbs_energy'])
    np.testing.assert_allclose(tnac4o_b200.energy_RMF(rmf_model(), ins.states), ins.energy, atol=1e-12)
