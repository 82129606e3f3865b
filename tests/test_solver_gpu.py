"""GPU parity tests of the phases and of the whole path against the oracle and the reference fixtures."""
import warnings

import numpy as np
import pytest
import torch

from conftest import SHAPES, droplet_couplings, droplet_golden, golden

pytestmark = pytest.mark.gpu
warnings.filterwarnings('ignore')


def make(J, L=128, beta=3, rot=0, pre=False, cls=None):
    import tnac4o_b200
    from oracle import RefSolver
    Nx, Ny = SHAPES[L]
    ins = (cls or tnac4o_b200.tnac4o)(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=beta)
    if rot:
        ins.rotate_graph(rot)
    if pre:
        ins.precondition(mode='balancing')
    return ins


def upload_mps(ref_mps, dev):
    """oracle MPS -> device MPS (used to test the search phase in isolation from the boundary-MPS phase)"""
    from tnac4o_b200 import mps
    psi = mps.MPS(d=1, L=ref_mps.L, Dmax=1, initial='X', device=dev)
    psi.A = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in ref_mps.A]
    psi.D = [psi.A[0].shape[0]] + [a.shape[2] for a in psi.A]
    return psi


@pytest.mark.parametrize('M,cut', [(64, 1e-8), (256, 1e-8), (1024, 0.0)])
def test_search_phase_with_oracle_environments(J128, M, cut):
    """same rhoT tensors on both sides: marginals to 1e-8 absolute, branch sets / energies / states exact"""
    from oracle import RefSolver
    ref = make(J128, cls=RefSolver)
    trace = {}
    ref.trace = lambda kind, **kw: trace.setdefault((kind, kw['ny'], kw['nx']), kw)
    ref.search_ground_state(M=M, relative_P_cutoff=cut, Dmax=8)
    ins = make(J128)
    ins.native_search = False          # the spy below hooks the Python site loop
    dev = ins._dev()
    ins._setup_rhoT = lambda **kw: setattr(ins, 'rhoT', [upload_mps(p, dev) for p in ref.rhoT])
    seen = {}
    orig = ins._site_marginals

    def spy(ws, br, RRat, ny, nx, want_P=False):
        P = orig(ws, br, RRat, ny, nx, want_P=True)
        seen[(ny, nx)] = (P.cpu().numpy(), br.vind[:br.n].cpu().numpy().copy())
        return None
    ins._site_marginals = spy
    ins.search_ground_state(M=M, relative_P_cutoff=cut, Dmax=8)
    for (ny, nx), (P, vind) in seen.items():
        r = trace[('marginals', ny, nx)]
        rows = {tuple(v): i for i, v in enumerate(r['vind'].view(np.uint8).tolist())}
        idx = [rows[tuple(v)] for v in vind.tolist()]                    # same branch set, any order
        assert len(idx) == len(rows)
        assert np.max(np.abs(P - r['P'][idx])) <= 1e-8
    assert ins.energy[0] == ref.energy[0]                                   # bit-exact (same table look-ups, same order)
    assert np.array_equal(ins.states, ref.states)
    assert int(ins.degeneracy) == int(ref.degeneracy)
    np.testing.assert_allclose(ins.probability, ref.probability, rtol=1e-8)
    np.testing.assert_allclose(ins.discarded_probability, ref.discarded_probability, rtol=1e-8)


def test_boundary_mps_against_oracle(J128):
    """gauge-invariant comparison of rhoT: normalised overlaps with the oracle's MPS, row by row"""
    from oracle import RefSolver
    ref = make(J128, cls=RefSolver)
    ref._setup_rhoT(Dmax=8)
    ins = make(J128)
    ins.build_rhoT0 = True
    ins._setup_rhoT(Dmax=8)
    for ny in range(4):
        a = [t.cpu().numpy() for t in ins.rhoT[ny].A]
        b = ref.rhoT[ny].A

        def ov(x, y):
            E = np.ones((1, 1))
            for p, q in zip(x, y):
                E = np.einsum('ab,apc,bpd->cd', E, p, q)
            return E.item()
        assert abs(ov(a, b) / np.sqrt(ov(a, a) * ov(b, b)) - 1) < 1e-10
        assert abs(ins.rhoT_overlap[ny] / ref.rhoT_overlap[ny] - 1) < 1e-8


@pytest.mark.parametrize('rot,pre,D,M', [(0, False, 8, 256), (3, False, 8, 256), (0, False, 48, 1024), (0, True, 8, 256)])
def test_ground_state_end_to_end(J128, rot, pre, D, M):
    z = golden('ref_small.npz')
    tag = 'gs_r%d_p%d_D%d_M%d' % (rot, pre, D, M)
    ins = make(J128, rot=rot, pre=pre)
    ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
    assert abs(ins.energy[0] - z[tag + '_energy'][0]) < 1e-10
    assert np.array_equal(ins.states, z[tag + '_states'])
    assert np.array_equal(ins.binary_states(), z[tag + '_bits'])
    assert int(ins.degeneracy) == int(z[tag + '_degeneracy'])
    np.testing.assert_allclose(ins.probability, z[tag + '_probability'], rtol=1e-7)
    e_file, bits_file = droplet_golden(128, 1)
    assert abs(ins.energy[0] - e_file) < 1e-5 and np.array_equal(ins.binary_states()[0], bits_file)


def test_marginals_end_to_end_against_reference_trace(J128):
    """full GPU path (own rhoT) against the reference's marginals: max |dP| <= 1e-8 on normalised vectors"""
    from oracle import RefSolver
    ref = make(J128, cls=RefSolver)
    trace = {}
    ref.trace = lambda kind, **kw: trace.setdefault((kind, kw['ny'], kw['nx']), kw)
    ref.search_ground_state(M=256, relative_P_cutoff=1e-8, Dmax=8)
    ins = make(J128)
    ins.native_search = False          # the spy below hooks the Python site loop
    seen = {}
    orig = ins._site_marginals

    def spy(ws, br, RRat, ny, nx, want_P=False):
        P = orig(ws, br, RRat, ny, nx, want_P=True)
        seen[(ny, nx)] = (P.cpu().numpy(), br.vind[:br.n].cpu().numpy().copy())
        return None
    ins._site_marginals = spy
    ins.search_ground_state(M=256, relative_P_cutoff=1e-8, Dmax=8)
    worst = 0.0
    for (ny, nx), (P, vind) in seen.items():
        r = trace[('marginals', ny, nx)]
        rows = {tuple(v): i for i, v in enumerate(r['vind'].view(np.uint8).tolist())}
        common = [(i, rows[tuple(v)]) for i, v in enumerate(vind.tolist()) if tuple(v) in rows]
        assert len(common) >= 0.99 * len(rows)                              # borderline members may differ
        a, b = zip(*common)
        worst = max(worst, np.max(np.abs(P[list(a)] - r['P'][list(b)])))
    assert worst <= 1e-8
    assert ins.energy[0] == ref.energy[0] and np.array_equal(ins.states, ref.states)


def test_config2_L512(J128):
    """BASELINE config 2 against the reference fixture AND, site by site, against the oracle's trace of every conditional
    marginal: max |dP| <= 1e-8 on the normalised 256-vectors (measured 2.6e-9, tests/tools/parity_probe.py), accumulated
    log2 P to 1e-10 relative (measured 1.5e-13)"""
    from oracle import RefSolver
    z = golden('ref_l512.npz')
    J = droplet_couplings(512)
    ref = make(J, L=512, cls=RefSolver)
    trace = {}
    ref.trace = lambda kind, **kw: trace.setdefault((kind, kw['ny'], kw['nx']), kw)
    ref.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
    ins = make(J, L=512)
    ins.native_search = False          # the spy below hooks the Python site loop
    seen = {}
    orig = ins._site_marginals

    def spy(ws, br, RRat, ny, nx, want_P=False):
        P = orig(ws, br, RRat, ny, nx, want_P=True)
        seen[(ny, nx)] = (P.cpu().numpy(), br.vind[:br.n].cpu().numpy().copy())
        return None
    ins._site_marginals = spy
    ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
    worst = 0.0
    assert len(seen) == 64
    for (ny, nx), (P, vind) in seen.items():
        r = trace[('marginals', ny, nx)]
        rows = {tuple(v): i for i, v in enumerate(r['vind'].view(np.uint8).tolist())}
        common = [(i, rows[tuple(v)]) for i, v in enumerate(vind.tolist()) if tuple(v) in rows]
        assert len(common) >= 0.9 * len(rows)                               # borderline members of the top-M cut may differ
        a, b = zip(*common)
        worst = max(worst, np.max(np.abs(P[list(a)] - r['P'][list(b)])))
    assert worst <= 1e-8
    assert ins.energy[0] == ref.energy[0] and np.array_equal(ins.states, ref.states)
    assert abs(ins.energy[0] - z['gs_energy'][0]) < 1e-9
    assert np.array_equal(ins.states, z['gs_states'])
    e_file, bits_file = droplet_golden(512, 1)
    assert abs(ins.energy[0] - e_file) < 1e-5 and np.array_equal(ins.binary_states()[0], bits_file)
    np.testing.assert_allclose(ins.probability, z['gs_probability'], rtol=1e-10)
    np.testing.assert_allclose(ins.probability, ref.probability, rtol=1e-10)
    # native driver = the instrumented Python loop
    nat = make(J, L=512)
    nat.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
    assert np.array_equal(nat.energy, ins.energy) and np.array_equal(nat.states, ins.states)
    assert np.array_equal(nat.probability, ins.probability)


def test_synthetic_family_A_instance_against_oracle():
    """the kind of instance that fills the bench batches (file values permuted over the same coupling pattern,
    SURVEY.md section 8d family A) at L=512: energy and state bit-exact, log2 P to 1e-10 against the oracle"""
    from oracle import RefSolver
    J = droplet_couplings(512)
    rng = np.random.default_rng(7)
    vals = np.array([v for _, _, v in J])
    Js = [[i, j, float(v)] for (i, j, _), v in zip(J, rng.permutation(vals))]
    ref = make(Js, L=512, cls=RefSolver)
    ref.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
    ins = make(Js, L=512)
    ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=16)
    assert ins.energy[0] == ref.energy[0] and np.array_equal(ins.states, ref.states)
    assert int(ins.degeneracy) == int(ref.degeneracy)
    np.testing.assert_allclose(ins.probability, ref.probability, rtol=1e-10)


def test_gibbs_sampling(J128):
    import tnac4o_b200
    z = golden('ref_small.npz')
    ins = make(J128, beta=1)
    np.random.seed(1)
    ins.gibbs_sampling(M=128, Dmax=16)
    assert ins.states.shape == (128, 16)
    E = tnac4o_b200.energy_Jij(J128, ins.binary_states())
    assert np.max(np.abs(E - ins.energy)) < 1e-6                            # examples/test_examples.py:56
    assert np.array_equal(ins.states, z['gibbs_states'])                    # same uniforms => the same 128 samples
    assert np.max(np.abs(ins.energy - z['gibbs_energy'])) < 1e-10


@pytest.mark.parametrize('rot', [0, 1])
def test_spectrum_and_decode(J128, rot):
    import tnac4o_b200
    z = golden('ref_small.npz')
    ins = make(J128, rot=rot)
    ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0)
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    assert len(ins.energy) == 31                                            # examples/test_examples.py:62
    order = np.lexsort(ins.states.T[::-1])
    assert np.array_equal(ins.states[order], z['sp_r%d_states' % rot])
    np.testing.assert_allclose(ins.energy[order], z['sp_r%d_energy' % rot], atol=1e-10)
    E = tnac4o_b200.energy_Jij(J128, ins.binary_states())
    assert np.max(np.abs(E - ins.energy)) < 1e-4


def test_config4_L2048_M1024():
    """BASELINE config 4 (M = 2^10 variant): golden energy bit-level, degeneracy 2, state in the degenerate set"""
    import tnac4o_b200
    z = golden('ref_l2048.npz')
    J = droplet_couplings(2048)
    ins = make(J, L=2048)
    ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=32)
    assert abs(ins.energy[0] - z['gs_energy'][0]) < 1e-9
    assert int(ins.degeneracy) == int(z['gs_degeneracy']) == 2
    e_file, bits_file = droplet_golden(2048, 1)
    assert abs(ins.energy[0] - e_file) < 1e-5
    bits = ins.binary_states()[0]
    # The ground state is two-fold degenerate: the two members differ in 3 spins and have the SAME
    # float64 energy here and in the reference (E diff 0.0, tests/tools/parity_probe.py).  Which one survives is decided
    # where the two branches merge: the reference takes np.argmin over the group in the order its unstable argsort left
    # them (tnac4o.py:482, 497), this package the first minimum in the total order of the candidate ids.  At M = 2^10 the
    # reference returns one member, this package the other (the one groundstates_otn2d.txt lists, which the reference
    # itself returns at M = 2^12 -- test below, where the states are identical); either is accepted here, the energy
    # is exact and the degeneracy is counted.
    d_ref, d_file = int(np.sum(bits != z['gs_bits'][0])), int(np.sum(bits != bits_file))
    assert min(d_ref, d_file) == 0 and max(d_ref, d_file) == 3, (d_ref, d_file)
    E = tnac4o_b200.energy_Jij(J, ins.binary_states())
    assert abs(E[0] - ins.energy[0]) < 1e-6
    # L=2048 truncations are ill-conditioned: the reference against itself (gesdd vs gesvd, or two CPUs) moves rhoT by
    # 1 - fidelity ~ 4e-12 and the accumulated log2 P by ~1e-6 relative
    np.testing.assert_allclose(ins.probability, z['gs_probability'], rtol=5e-6)


def test_config4_L2048_M4096():
    """BASELINE config 4 as quoted (M = 2^12, Dmax = 32) against the reference fixture ref_l2048_m4096.npz"""
    import tnac4o_b200
    z = golden('ref_l2048_m4096.npz')
    J = droplet_couplings(2048)
    ins = make(J, L=2048)
    ins.search_ground_state(M=2 ** 12, relative_P_cutoff=1e-8, Dmax=32)
    assert abs(ins.energy[0] - z['gs_energy'][0]) < 1e-9
    assert int(ins.degeneracy) == int(z['gs_degeneracy']) == 2
    bits = ins.binary_states()[0]
    e_file, bits_file = droplet_golden(2048, 1)
    assert np.array_equal(bits, z['gs_bits'][0]) and np.array_equal(bits, bits_file)      # the reference's state, bit for bit
    assert abs(tnac4o_b200.energy_Jij(J, ins.binary_states())[0] - ins.energy[0]) < 1e-6
    np.testing.assert_allclose(ins.probability, z['gs_probability'], rtol=5e-6)
    assert abs(ins.discarded_probability - float(z['gs_discarded'])) < 1e-3
    # number of conditional marginals evaluated = number of live branches summed over the sites (borderline
    # candidates at the 1e-8 cut-off may differ by a few)
    assert abs(ins.stats['marginals'] - int(z['marginals'])) <= 64


def test_config5_gibbs_L2048_against_reference_samples():
    """BASELINE config 5 at the sample count the reference finishes on a CPU (256): beta = 1, np.random.seed(1);
    same uniforms => the same samples, state by state, and bit-identical energies"""
    z = golden('ref_gibbs_l2048.npz')
    J = droplet_couplings(2048)
    ins = make(J, L=2048, beta=1)
    np.random.seed(1)
    ins.gibbs_sampling(M=256, Dmax=32)
    assert np.array_equal(ins.states, z['states'])           # all 256 samples identical, state by state
    assert np.max(np.abs(ins.energy - z['energy'])) < 1e-9
    assert ins.negative_probability >= float(z['negative']) - 1e-9


def test_j124_degeneracy_counting():
    """J124 C8 #1 (examples/test_examples.py:139-147): E = -2309 exactly, degeneracy 1152 -- exercises the merge rule"""
    z = golden('instances.npz')
    J = [[int(a) - 1, int(b) - 1, float(c)] for a, b, c in zip(z['J124_C8_001_i'], z['J124_C8_001_j'], z['J124_C8_001_v'])]
    import tnac4o_b200
    ins = tnac4o_b200.tnac4o(mode='Ising', Nx=8, Ny=8, Nc=8, J=J, beta=0.75)
    ins.precondition(mode='balancing')
    ins.search_ground_state(M=2 ** 12, relative_P_cutoff=1e-8, Dmax=8)
    assert abs(ins.energy[0] - (-2309)) < 1e-12
    assert int(ins.degeneracy) == 1152
    ref = golden('ref_j124.npz')
    assert abs(ins.energy[0] - ref['gs_energy'][0]) < 1e-12 and int(ref['gs_degeneracy']) == 1152
    E = tnac4o_b200.energy_Jij(J, ins.binary_states())
    assert abs(E[0] + 2309) < 1e-9


def test_config3_L1152_spectrum_and_decode():
    """BASELINE config 3: low-energy spectrum of L=1152 #1 (ee=1, dE=1, Dmax=32), decoded: 545 966 states in the
    reference run; energies of all decoded states re-checked with the CSR energy kernel"""
    import tnac4o_b200
    z = golden('ref_l1152.npz')
    J = droplet_couplings(1152)
    ins = make(J, L=1152)
    ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=32, max_dEng=1.0)
    assert abs(ins.energy[0] - z['gs_energy'][0]) < 1e-9
    assert int(ins.degeneracy) == int(z['gs_degeneracy'])
    e0 = ins.energy[0]
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    n_ref = int(z['n_states'])
    assert abs(len(ins.energy) - n_ref) <= 0.001 * n_ref              # borderline branches may differ
    vals, counts = np.unique(np.round((ins.energy - e0) * 75).astype(np.int64), return_counts=True)
    ref_levels = dict(zip(z['level_75dE'].tolist(), z['level_count'].tolist()))
    low = [(v, c) for v, c in zip(vals.tolist(), counts.tolist()) if v <= 40]
    assert all(ref_levels.get(v) == c for v, c in low), (low[:10], [(k, ref_levels[k]) for k in sorted(ref_levels)[:10]])
    E = tnac4o_b200.energy_Jij(J, ins.binary_states())
    assert np.max(np.abs(E - ins.energy)) < 1e-6
    assert len(np.unique(ins.states, axis=0)) == len(ins.states)        # all decoded states are distinct


def test_config3_decode_of_the_reference_file_gives_the_reference_state_set():
    """examples/e04 on the file the reference's e03 -s wrote for config 3 (L=1152, dE=1): the device enumeration must return
    exactly the reference's 545 966 states -- the SET (sha256 of the sorted rows) and every energy bit for bit -- and do it
    at least 20x faster than the reference's 18.9 s Python loop (tnac4o.py:2295-2335)"""
    import hashlib
    import os
    import time
    import tnac4o_b200
    from conftest import GOLDEN
    z = golden('ref_l1152_decoded.npz')
    ins = tnac4o_b200.load(os.path.join(GOLDEN, 'ref_saved_spectrum_l1152.npy'))
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)          # warm-up (allocations)
    ins = tnac4o_b200.load(os.path.join(GOLDEN, 'ref_saved_spectrum_l1152.npy'))
    t0 = time.time()
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    seconds = time.time() - t0
    assert len(ins.energy) == int(z['n_states']) == 545966
    assert np.all(np.diff(ins.energy) >= 0)
    assert np.array_equal(ins.energy, z['energies_sorted'])
    st = np.ascontiguousarray(ins.states)
    st = st[np.lexsort(st.T[::-1])]
    assert np.array_equal(st[:64], z['states_head'])
    assert np.array_equal(np.frombuffer(hashlib.sha256(st.tobytes()).digest(), dtype=np.uint8), z['states_sha256'])
    assert seconds * 20 < float(z['seconds_decode']), seconds
    # the top-K cut keeps the lowest energies (ties at the cut are arbitrary in the reference as well)
    cut = tnac4o_b200.load(os.path.join(GOLDEN, 'ref_saved_spectrum_l1152.npy'))
    cut.decode_low_energy_states(max_dEng=1.0, max_states=1000)
    assert np.array_equal(cut.energy, z['energies_sorted'][:1000])


def test_config3_save_load_decode_round_trip(tmp_path):
    """e03 -s then e04 on this package's own file at config 3: search, save, load, decode = decode of the live object"""
    import tnac4o_b200
    J = droplet_couplings(1152)
    ins = make(J, L=1152)
    ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=32, max_dEng=1.0)
    fn = str(tmp_path / 'l1152.npy')
    ins.save(fn)
    back = tnac4o_b200.load(fn)
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    back.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    assert np.array_equal(ins.energy, back.energy) and np.array_equal(ins.states, back.states)
    z = golden('ref_l1152_decoded.npz')
    n_ref = int(z['n_states'])
    assert abs(len(back.energy) - n_ref) <= 0.001 * n_ref
    head = min(len(back.energy), n_ref, 4096)
    np.testing.assert_allclose(back.energy[:head], z['energies_sorted'][:head], rtol=0, atol=1e-9)


def test_config5_gibbs_L2048_reduced():
    """BASELINE config 5 at reduced sample count: beta = 1, L = 2048; sampled energies equal energy_Jij of the states"""
    import tnac4o_b200
    J = droplet_couplings(2048)
    ins = make(J, L=2048, beta=1)
    np.random.seed(1)
    ins.gibbs_sampling(M=2000, Dmax=32)
    assert ins.states.shape == (2000, 256) and ins.negative_probability > -1e-6
    E = tnac4o_b200.energy_Jij(J, ins.binary_states())
    assert np.max(np.abs(E - ins.energy)) < 1e-6
    assert len(np.unique(ins.states, axis=0)) > 1900
    assert -3200 < ins.energy.mean() < -2900                               # reference run: <E> = -3038.85 at M = 10^3


@pytest.mark.parametrize('L,D', [(128, 8), (512, 16)])
def test_native_row_driver_equals_python_mps_methods(L, D):
    """csrc/mps_native.cu issues the same kernel sequence as tnac4o_b200/mps.py (the mirror of the reference's MPS
    class): the boundary MPS of every row must come out bit-identical"""
    J = droplet_couplings(L)
    a = make(J, L=L)
    b = make(J, L=L)
    b.native_rows = False
    a.build_rhoT0 = b.build_rhoT0 = True
    a._setup_rhoT(Dmax=D)
    b._setup_rhoT(Dmax=D)
    for ny in range(a.Ny + 1):
        assert a.rhoT[ny].D == b.rhoT[ny].D
        for x, y in zip(a.rhoT[ny].A, b.rhoT[ny].A):
            assert torch.equal(x, y)
        if ny < a.Ny:
            assert a.rhoT_overlap[ny] == b.rhoT_overlap[ny]
            assert a.rhoT_discarded[ny] == b.rhoT_discarded[ny]
    a._setup_rhoB(Dmax=D)
    b._setup_rhoB(Dmax=D)
    for ny in range(a.Ny + 1):
        for x, y in zip(a.rhoB[ny].A, b.rhoB[ny].A):
            assert torch.equal(x, y)


def test_native_search_driver_equals_python_loop(J128):
    """csrc/search_native.cu against the Python row / site loop: identical kernels in identical order -> identical results"""
    for M, cut in ((64, 1e-8), (1024, 0.0)):
        a, b = make(J128), make(J128)
        b.native_search = False
        a.search_ground_state(M=M, relative_P_cutoff=cut, Dmax=8)
        b.search_ground_state(M=M, relative_P_cutoff=cut, Dmax=8)
        assert np.array_equal(a.energy, b.energy) and np.array_equal(a.states, b.states)
        assert np.array_equal(a.probability, b.probability) and a.degeneracy == b.degeneracy
        assert a.discarded_probability == b.discarded_probability and a.negative_probability == b.negative_probability
        assert a.stats['marginals'] == b.stats['marginals']


@pytest.mark.parametrize('L,D', [(128, 32), (512, 32)])
def test_ten_instances_reach_the_golden_energies(L, D):
    """droplet instances 001-010 against groundstates_otn2d.txt: energy to the file's print
    precision, and the file's state wherever the ground state is not degenerate"""
    import tnac4o_b200
    from conftest import droplet_couplings10, droplet_golden10
    for k in range(1, 11):
        J = droplet_couplings10(L, k)
        ins = make(J, L=L)
        ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=D)
        e_file, bits_file = droplet_golden10(L, k)
        assert abs(ins.energy[0] - e_file) < 1e-5, (L, k, ins.energy[0], e_file)
        assert abs(tnac4o_b200.energy_Jij(J, ins.binary_states()[:1])[0] - ins.energy[0]) < 1e-6
        if int(ins.degeneracy) == 1:
            assert np.array_equal(ins.binary_states()[0], bits_file), (L, k)


@pytest.mark.parametrize('L,D', [(128, 8), (512, 16)])
def test_boundary_mps_bond_dimensions_equal_the_oracle(L, D):
    """the rank decisions of truncateC (keep = #{S > S0 max(eps, tol)}, mps.py:805-806) on the real centre matrices: every
    bond dimension of every row equals the oracle's (LAPACK gesdd), the overlaps agree to 1e-12 (measured 2e-15,
    profiles/r2c_bond_probe.txt) -- so the deflation of svd.cu below 1e-2 eps ||C|| changes no kept rank here"""
    from oracle import RefSolver
    J = droplet_couplings(L)
    ref = make(J, L=L, cls=RefSolver)
    ref._setup_rhoT(Dmax=D)
    ins = make(J, L=L)
    ins.build_rhoT0 = True
    ins._setup_rhoT(Dmax=D)
    for ny in range(ins.Ny):
        a = [ins.rhoT[ny].A[0].shape[0]] + [t.shape[2] for t in ins.rhoT[ny].A]
        b = [ref.rhoT[ny].A[0].shape[0]] + [t.shape[2] for t in ref.rhoT[ny].A]
        assert a == b, (ny, a, b)
        assert abs(ins.rhoT_overlap[ny] / ref.rhoT_overlap[ny] - 1) < 1e-12
