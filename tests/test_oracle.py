"""CPU tests: the numpy oracle against the reference's golden vectors (tests/golden, generated from the
unmodified reference by tests/golden/make_golden.py) and against the known answers of
examples/test_examples.py:27, 62 and the groundstates_otn2d.txt files."""
import warnings

import numpy as np
import pytest

from conftest import SHAPES, droplet_couplings, droplet_golden, golden
from oracle import RefSolver, RefMPS, ref_nfactor, ref_qr, ref_svd
from oracle.auxx_ref import energy_ising_dense, energy_ising_sparse

warnings.filterwarnings('ignore')


def run_gs(J, rot, pre, D, M, beta=3, trace=None):
    ins = RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=beta)
    if rot:
        ins.rotate_graph(rot)
    if pre:
        ins.precondition(mode='balancing')
    ins.trace = trace
    ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
    return ins


@pytest.mark.parametrize('rot,pre,D,M', [(0, False, 8, 256), (3, False, 8, 256), (0, True, 8, 256), (0, False, 48, 1024)])
def test_ground_state_matches_reference_fixture(J128, rot, pre, D, M):
    z = golden('ref_small.npz')
    tag = 'gs_r%d_p%d_D%d_M%d' % (rot, pre, D, M)
    ins = run_gs(J128, rot, pre, D, M)
    assert abs(ins.energy[0] - (-210.93333333)) < 1e-5                  # examples/test_examples.py:27
    assert abs(ins.energy[0] - z[tag + '_energy'][0]) < 1e-10
    assert np.array_equal(ins.states, z[tag + '_states'])
    assert np.array_equal(ins.binary_states(), z[tag + '_bits'])
    assert int(ins.degeneracy) == int(z[tag + '_degeneracy'])
    np.testing.assert_allclose(ins.probability, z[tag + '_probability'], rtol=1e-8)
    np.testing.assert_allclose(ins.discarded_probability, z[tag + '_discarded'], rtol=1e-8)
    np.testing.assert_allclose(ins.negative_probability, z[tag + '_negative'], atol=1e-10)
    e_file, bits_file = droplet_golden(128, 1)                              # groundstates_otn2d.txt:1
    assert abs(ins.energy[0] - e_file) < 1e-5
    assert np.array_equal(ins.binary_states()[0], bits_file)


def test_marginal_trace_matches_reference_fixture(J128):
    z = golden('ref_small.npz')
    tag = 'gs_r0_p0_D8_M256'
    seen = []
    run_gs(J128, 0, False, 8, 256, trace=lambda kind, **kw: seen.append(kw['P']) if kind == 'marginals' else None)
    P = np.concatenate(seen, axis=0)
    assert P.shape[0] == int(z[tag + '_trace_calls'])
    stride = int(z[tag + '_trace_stride'])
    assert np.max(np.abs(P[::stride] - z[tag + '_trace_P'])) <= 1e-8       # normalised 256-vectors, absolute


@pytest.mark.parametrize('rot', [0, 1])
def test_spectrum_and_decode(J128, rot):
    z = golden('ref_small.npz')
    ins = RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=J128, beta=3)
    if rot:
        ins.rotate_graph(rot)
    ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0)
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    assert len(ins.energy) == 31                                            # examples/test_examples.py:62
    order = np.lexsort(ins.states.T[::-1])
    assert np.array_equal(ins.states[order], z['sp_r%d_states' % rot])
    np.testing.assert_allclose(ins.energy[order], z['sp_r%d_energy' % rot], atol=1e-10)
    eJ = energy_ising_dense(J128, ins.binary_states())
    assert np.max(np.abs(eJ - ins.energy)) < 1e-4
    assert np.max(np.abs(energy_ising_sparse(J128, ins.binary_states()) - eJ)) < 1e-9


def test_gibbs_matches_reference_fixture(J128):
    z = golden('ref_small.npz')
    ins = RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=J128, beta=1)
    np.random.seed(1)
    ins.gibbs_sampling(M=128, Dmax=16)
    assert np.array_equal(ins.states, z['gibbs_states'])
    np.testing.assert_allclose(ins.energy, z['gibbs_energy'], atol=1e-10)
    eJ = energy_ising_dense(J128, ins.binary_states())
    assert np.max(np.abs(eJ - ins.energy)) < 1e-6                          # examples/test_examples.py:56


def test_linear_algebra_wrappers():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((40, 12))
    Q, R = ref_qr(A)
    assert np.all(np.diag(R) >= 0) and np.allclose(Q @ R, A) and np.allclose(Q.T @ Q, np.eye(12))
    U, S, V = ref_svd(A)
    assert np.allclose((U * S) @ V, A) and np.all(np.diff(S) <= 0)
    assert ref_nfactor(np.array([0.3, -5.0])) == 4.0 and ref_nfactor(np.array([1.0])) == 1.0
    assert ref_nfactor(np.zeros(3)) == 2.0 ** -1023


def test_compress_keeps_the_state():
    """compressing with a generous bond must not change the state: overlap of normalised states = 1"""
    rng = np.random.default_rng(1)
    psi = RefMPS(6, d=1)
    W = [rng.random((1 if n == 0 else 3, 1, 1 if n == 5 else 3, 4)) for n in range(6)]
    psi.apply_mpo(W, conj=True)
    full = psi.copy()
    full.canonise_right()
    ov = psi.compress(Dmax=64, tolS=1e-16, tolV=1e-10, max_sweeps=4)
    assert abs(ov - 1.0) < 1e-10
    assert max(psi.discarded) < 1e-12


def test_full_size_fixtures_are_consistent_with_the_couplings():
    """configs 4 and 5 at L=2048 (fixtures from the unmodified reference, too slow to regenerate in the CPU suite):
    the stored states reproduce the stored energies under the oracle's energy_Jij restatement, the M=2^12 search finds
    the same ground state as M=2^10 and as the groundstates_otn2d.txt line, and the Gibbs samples are at beta=1 energies"""
    J = droplet_couplings(2048)
    z4, z4b, z5 = golden('ref_l2048_m4096.npz'), golden('ref_l2048.npz'), golden('ref_gibbs_l2048.npz')
    assert list(z4['params']) == [2048, 32, 4096] and list(z5['params']) == [2048, 32, 256]
    e_file, bits_file = droplet_golden(2048, 1)
    assert abs(float(z4['gs_energy'][0]) - e_file) < 1e-5 and float(z4['gs_energy'][0]) == float(z4b['gs_energy'][0])
    assert int(z4['gs_degeneracy']) == 2
    assert abs(energy_ising_sparse(J, z4['gs_bits'])[0] - float(z4['gs_energy'][0])) < 1e-6
    assert int(z4['marginals']) > 3 * int(z4b['marginals'])                 # four times the branches, saturated
    assert np.max(np.abs(z5['energy_Jij'] - z5['energy'])) < 1e-6           # examples/test_examples.py:56
    # samples -> spins (1 - bit of the cell state, tnac4o.py:279-285) -> the same energies
    bits = 1 - ((z5['states'][:, :, None].astype(np.int64) >> np.arange(8)) & 1)
    assert np.max(np.abs(energy_ising_sparse(J, bits.reshape(256, 2048)) - z5['energy'])) < 1e-6


ENCODING_CASES = [(2, 0, 0), (2, 2, 0), (3, 0, 0), (3, 3, 0), (2, 0, 4), (3, 0, 4), (2, 1, 0), (3, 1, 0)]


@pytest.mark.parametrize('ee,rot,hd', ENCODING_CASES)
def test_adjacency_encodings_match_reference_fixture(J128, ee, rot, hd):
    """excitations_encoding = 2, 3 (tnac4o.py:943-1358; examples/test_examples.py:59-104 expects 31 states below dE = 1
    for every encoding and rotation): stored structure sizes and the decoded spectrum, state by state"""
    z = golden('ref_encodings.npz')
    tag = 'ee%d_r%d_hd%d' % (ee, rot, hd)
    ins = RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=J128, beta=3)
    if rot:
        ins.rotate_graph(rot)
    ins.search_low_energy_spectrum(excitations_encoding=ee, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0, lim_hd=hd)
    assert len(ins.d) == int(z[tag + '_n_shapes']) and len(ins.el) == int(z[tag + '_n_first_layer'])
    np.testing.assert_allclose(sorted(e[0][0] for e in ins.el), z[tag + '_first_layer_dE'], atol=1e-10)
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    if hd == 0:
        assert len(ins.energy) == 31
    order = np.lexsort(ins.states.T[::-1])
    assert np.array_equal(ins.states[order], z[tag + '_states'])
    np.testing.assert_allclose(ins.energy[order], z[tag + '_energy'], atol=1e-10)
    assert np.max(np.abs(energy_ising_sparse(J128, ins.binary_states()) - ins.energy)) < 1e-4


@pytest.mark.parametrize('ee', [2, 3])
def test_adjacency_encodings_L512(ee):
    """several layers of the droplet hierarchy: L = 512, dE <= 0.5 -> 302 states for every encoding"""
    z = golden('ref_encodings.npz')
    ins = RefSolver(mode='Ising', Nx=8, Ny=8, Nc=8, J=droplet_couplings(512), beta=3)
    ins.search_low_energy_spectrum(excitations_encoding=ee, M=256, relative_P_cutoff=1e-8, Dmax=8, max_dEng=0.5)
    assert len(ins.d) == int(z['L512_ee%d_n_shapes' % ee])
    ins.decode_low_energy_states(max_dEng=0.5, max_states=2 ** 20)
    order = np.lexsort(ins.states.T[::-1])
    assert np.array_equal(ins.states[order], z['L512_ee%d_states' % ee])
    assert np.array_equal(ins.states[order], z['L512_ee1_states'])           # the encodings agree on the spectrum
    np.testing.assert_allclose(ins.energy[order], z['L512_ee%d_energy' % ee], atol=1e-10)


@pytest.mark.parametrize('k', [1, 2, 3])
def test_max_energy_known_answers(k):
    """max_energy_otn2d.txt (the known-answer file for the sign-flipped problem, minus_Jij): the oracle on -J reaches the
    file's energy and state, and equals the reference's own run on the same couplings (ref_max_energy.npz)"""
    import tnac4o_b200
    z = golden('ref_max_energy.npz')
    J = tnac4o_b200.minus_Jij(droplet_couplings(128, k))
    ins = run_gs(J, 0, False, 48, 1024)
    tag = 'max_%03d' % k
    assert abs(-ins.energy[0] - float(z['file_%03d_energy' % k])) < 1e-5
    assert np.array_equal(ins.binary_states()[0], z['file_%03d_bits' % k])
    assert abs(ins.energy[0] - z[tag + '_energy'][0]) < 1e-10
    assert np.array_equal(ins.states, z[tag + '_states'])
    np.testing.assert_allclose(ins.probability, z[tag + '_probability'], rtol=1e-8)
    np.testing.assert_allclose(ins.discarded_probability, z[tag + '_discarded'], rtol=1e-8)
    # the sign-flipped couplings evaluated on the file's state give minus the file's energy
    assert abs(energy_ising_sparse(J, z['file_%03d_bits' % k][None, :])[0] + float(z['file_%03d_energy' % k])) < 1e-5
