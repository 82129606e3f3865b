#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container, where the reference is mounted read-only at
/root/reference (it does not exist on the GPU box -- tests read the committed .npz files).

    PYTHONDONTWRITEBYTECODE=1 OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 \
        python tests/golden/make_golden.py instances small l512 l2048 l1152

Fixtures
  instances.npz       coupling lists of droplet instances 001-003 (L=128), 001 (L=512, 1152, 2048), J124 C8 #1,
                      with the matching lines of groundstates_otn2d.txt / results_*.txt
  instances10.npz     droplet instances 001-010 of L = 128, 512, 1152, 2048 (couplings as 75 J, int16) + golden energies / states
  ref_small.npz       L=128 #1: search_ground_state / spectrum / Gibbs outputs for several (rot, precondition, D)
                      plus per-site traces (marginals of every branch, branch records after merge + top-M)
  ref_l512.npz        config 2: L=512 #1, M=2^10, Dmax=16
  ref_l2048.npz       config 4 (M=2^10 variant): L=2048 #1, Dmax=32
  ref_l2048_m4096.npz config 4 as quoted: L=2048 #1, M=2^12, Dmax=32
  ref_gibbs_l2048.npz config 5 at M=256 samples: L=2048 #1, beta=1, seed 1
  ref_encodings.npz   excitations_encoding = 2, 3 (adjacency-based droplets): L=128 spectra for several rotations / lim_hd, L=512
  ref_saved_spectrum_ee{1,2}.npy   files written by the reference's save() (pickled dict), read back by tnac4o_b200.load
  ref_l1152.npz       config 3: L=1152 #1 spectrum (ee=1, dE=1) -> number of decoded states, energies
  ref_saved_spectrum_l1152.npy + ref_l1152_decoded.npz   config 3 as e03 -s / e04 run it: the reference's saved file and the
                      decoded state set (sha256 of the sorted rows), all 545 966 energies, level counts
  ref_j124_sweep.npz  examples/e06 on J124 C=8 instances 1-20: per rotation energy / degeneracy and the selected pair,
                      with the couplings and the lines of results_*J124*.txt
  ref_rmf.npz         examples/e05 (RMF toy 5x3): spectra for the three encodings (test_examples.py:107-136)
  ref_max_energy.npz  known answers of max_energy_otn2d.txt (L=128 #1-3) and the reference's search on minus_Jij(J) for them
  ref_noise.npz       host logic: couplings of L=128 #1 (and of a 8x2 lattice) after rotate_graph(rot) and add_noise with
                      np.random.seed(7), as sorted (row, col, value) triplets, plus the cell order
"""
import os
import sys
import time
import logging
import warnings

import numpy as np

if not hasattr(np, 'int'):       # the reference still uses np.int on the ee=3 path (tnac4o.py:2213)
    np.int = int
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True
warnings.filterwarnings('ignore')
logging.disable(logging.CRITICAL)
import tnac4o as ref   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
INST = '/root/reference/instances'
SHAPES = {128: (4, 4), 512: (8, 8), 1152: (12, 12), 2048: (16, 16)}


def raw_couplings(L, k):
    return np.loadtxt('%s/Chimera_droplet_instances/chimera%d_spinglass_power/%03d.txt' % (INST, L, k))


def droplet_J(L, k):
    fn = '%s/Chimera_droplet_instances/chimera%d_spinglass_power/%03d.txt' % (INST, L, k)
    return ref.round_Jij(ref.Jij_f2p(ref.load_Jij(fn)), 1 / 75)


def golden_line(L, k):
    fn = '%s/Chimera_droplet_instances/chimera%d_spinglass_power/groundstates_otn2d.txt' % (INST, L)
    with open(fn) as f:
        for line in f:
            name, rest = line.split(':')
            if name.strip() == '%03d.txt' % k:
                vals = rest.split()
                return float(vals[0]), np.array(vals[1:], dtype=np.int8)
    raise KeyError(k)


def make_instances():
    out = {}
    for L, ks in ((128, (1, 2, 3)), (512, (1,)), (1152, (1,)), (2048, (1,))):
        for k in ks:
            raw = raw_couplings(L, k)
            out['J_%d_%03d_i' % (L, k)] = raw[:, 0].astype(np.int32)
            out['J_%d_%03d_j' % (L, k)] = raw[:, 1].astype(np.int32)
            out['J_%d_%03d_v' % (L, k)] = raw[:, 2].astype(np.float64)
            e, bits = golden_line(L, k)
            out['gs_%d_%03d_energy' % (L, k)] = np.float64(e)
            out['gs_%d_%03d_bits' % (L, k)] = bits
    raw = np.loadtxt('%s/Chimera_J124/C=8_J124/001.txt' % INST)
    out['J124_C8_001_i'] = raw[:, 0].astype(np.int32)
    out['J124_C8_001_j'] = raw[:, 1].astype(np.int32)
    out['J124_C8_001_v'] = raw[:, 2].astype(np.float64)
    out['J124_C8_001_energy_deg'] = np.array([-2309, 1152], dtype=np.int64)    # results file / test_examples.py:142
    np.savez_compressed(os.path.join(HERE, 'instances.npz'), **out)


def make_instances10():
    """droplet instances 001-010 of every lattice size with their lines of groundstates_otn2d.txt (the
    10-instance mean); couplings as 75 * J (odd integers, int16) to keep the file small"""
    out = {}
    for L in (128, 512, 1152, 2048):
        for k in range(1, 11):
            raw = raw_couplings(L, k)
            out['J_%d_%03d_i' % (L, k)] = raw[:, 0].astype(np.int16)
            out['J_%d_%03d_j' % (L, k)] = raw[:, 1].astype(np.int16)
            v75 = np.round(raw[:, 2] * 75)
            assert np.max(np.abs(v75 / 75 - raw[:, 2])) < 1e-5
            out['J_%d_%03d_v75' % (L, k)] = v75.astype(np.int16)
            e, bits = golden_line(L, k)
            out['gs_%d_%03d_energy' % (L, k)] = np.float64(e)
            if len(bits):
                out['gs_%d_%03d_bits' % (L, k)] = bits
    np.savez_compressed(os.path.join(HERE, 'instances10.npz'), **out)


def result_fields(ins, tag, out):
    out[tag + '_energy'] = np.asarray(ins.energy)
    out[tag + '_states'] = np.asarray(ins.states)
    out[tag + '_probability'] = np.asarray(ins.probability)
    out[tag + '_degeneracy'] = np.int64(ins.degeneracy)
    out[tag + '_discarded'] = np.float64(ins.discarded_probability)
    out[tag + '_negative'] = np.float64(ins.negative_probability)
    out[tag + '_bits'] = ins.binary_states()


def make_small():
    out = {}
    J = droplet_J(128, 1)
    for rot, pre, D, M in ((0, False, 8, 256), (3, False, 8, 256), (0, True, 8, 256), (1, True, 16, 1024), (0, False, 48, 1024)):
        ins = ref.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=3)
        if rot:
            ins.rotate_graph(rot)
        if pre:
            ins.precondition(mode='balancing')
        tag = 'gs_r%d_p%d_D%d_M%d' % (rot, pre, D, M)
        trace = []
        if (rot, pre, D) == (0, False, 8):
            orig = ins._calculate_Pn

            def spy(A, RL, AT, RR, _o=orig):
                P, flag = _o(A, RL, AT, RR)
                trace.append((P.copy(), flag))
                return P, flag
            ins._calculate_Pn = spy
        ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
        result_fields(ins, tag, out)
        if trace:
            # every 11th marginal (call order = site-major, branch-minor) keeps the fixture small
            out[tag + '_trace_stride'] = np.int64(11)
            out[tag + '_trace_calls'] = np.int64(len(trace))
            out[tag + '_trace_P'] = np.array([t[0] for t in trace[::11]])
            out[tag + '_trace_flag'] = np.array([t[1] for t in trace], dtype=np.float64)
            out[tag + '_rhoT_overlap'] = np.array(ins.rhoT_overlap, dtype=np.float64)
            out[tag + '_rhoT_discarded'] = np.array(ins.rhoT_discarded, dtype=np.float64)
    # spectrum + decode (test_examples.py:59-104 expects 31 states below dE = 1)
    for rot in (0, 1):
        ins = ref.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=3)
        if rot:
            ins.rotate_graph(rot)
        ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0)
        ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
        tag = 'sp_r%d' % rot
        order = np.lexsort(ins.states.T[::-1])
        out[tag + '_energy'] = ins.energy[order]
        out[tag + '_states'] = ins.states[order]
        out[tag + '_energy_Jij'] = ref.energy_Jij(J, ins.binary_states())[order]
    # Gibbs sampling with the global numpy stream (test_examples.py:36-56)
    ins = ref.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=1)
    np.random.seed(1)
    ins.gibbs_sampling(M=128, Dmax=16)
    out['gibbs_energy'] = ins.energy
    out['gibbs_states'] = ins.states.astype(np.int16)
    out['gibbs_negative'] = np.float64(ins.negative_probability)
    # J124 C8 #1 is the strongest exact known answer (degeneracy counting), but takes minutes: see `j124`
    np.savez_compressed(os.path.join(HERE, 'ref_small.npz'), **out)


def make_big(L, D, M, name):
    Nx, Ny = SHAPES[L]
    ins = ref.tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=droplet_J(L, 1), beta=3)
    calls = [0]
    orig = ins._calculate_Pn

    def spy(A, RL, AT, RR):
        calls[0] += 1
        return orig(A, RL, AT, RR)
    ins._calculate_Pn = spy
    t0 = time.time()
    ins._setup_rhoT(Dmax=D)
    t_rho = time.time() - t0
    setup = ins._setup_rhoT
    ins._setup_rhoT = lambda **kw: None          # already built; time the search phase separately
    t0 = time.time()
    ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
    t_search = time.time() - t0
    ins._setup_rhoT = setup
    out = {}
    result_fields(ins, 'gs', out)
    out['seconds_rhoT'], out['seconds_search'] = np.float64(t_rho), np.float64(t_search)
    out['marginals'] = np.int64(calls[0])
    out['rhoT_overlap'] = np.array(ins.rhoT_overlap, dtype=np.float64)
    out['rhoT_discarded'] = np.array(ins.rhoT_discarded, dtype=np.float64)
    out['params'] = np.array([L, D, M], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name), **out)


def make_gibbs_l2048(M=256):
    """config 5 at a CPU-feasible sample count: L=2048, beta=1, np.random.seed(1), Dmax=32"""
    ins = ref.tnac4o(mode='Ising', Nx=16, Ny=16, Nc=8, J=droplet_J(2048, 1), beta=1)
    np.random.seed(1)
    t0 = time.time()
    ins.gibbs_sampling(M=M, Dmax=32)
    out = {'energy': ins.energy, 'states': ins.states.astype(np.int16), 'negative': np.float64(ins.negative_probability),
           'seconds': np.float64(time.time() - t0), 'params': np.array([2048, 32, M], dtype=np.int64),
           'energy_Jij': ref.energy_Jij(droplet_J(2048, 1), ins.binary_states())}
    np.savez_compressed(os.path.join(HERE, 'ref_gibbs_l2048.npz'), **out)


def make_encodings():
    """adjacency-based droplet encodings (excitations_encoding = 2, 3; examples/test_examples.py:59-104): decoded
    spectra below dE = 1 for several rotations, a Hamming-distance limited run, and the sizes of the stored structures"""
    out = {}
    J = droplet_J(128, 1)
    for ee, rot, hd in ((2, 0, 0), (2, 2, 0), (3, 0, 0), (3, 3, 0), (2, 0, 4), (3, 0, 4), (2, 1, 0), (3, 1, 0)):
        ins = ref.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=3)
        if rot:
            ins.rotate_graph(rot)
        ins.search_low_energy_spectrum(excitations_encoding=ee, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0, lim_hd=hd)
        tag = 'ee%d_r%d_hd%d' % (ee, rot, hd)
        out[tag + '_n_shapes'] = np.int64(len(ins.d))
        out[tag + '_n_first_layer'] = np.int64(len(ins.el))
        out[tag + '_first_layer_dE'] = np.array(sorted(e[0][0] for e in ins.el))
        ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
        order = np.lexsort(ins.states.T[::-1])
        out[tag + '_energy'] = ins.energy[order]
        out[tag + '_states'] = ins.states[order]
        out[tag + '_energy_Jij'] = ref.energy_Jij(J, ins.binary_states())[order]
    # a wider window on a larger lattice: L = 512, dE <= 0.5 (hundreds of states, several layers of the hierarchy)
    J = droplet_J(512, 1)
    for ee in (1, 2, 3):
        ins = ref.tnac4o(mode='Ising', Nx=8, Ny=8, Nc=8, J=J, beta=3)
        ins.search_low_energy_spectrum(excitations_encoding=ee, M=256, relative_P_cutoff=1e-8, Dmax=8, max_dEng=0.5)
        tag = 'L512_ee%d' % ee
        out[tag + '_n_shapes'] = np.int64(len(ins.d))
        ins.decode_low_energy_states(max_dEng=0.5, max_states=2 ** 20)
        order = np.lexsort(ins.states.T[::-1])
        out[tag + '_energy'] = ins.energy[order]
        out[tag + '_states'] = ins.states[order]
    np.savez_compressed(os.path.join(HERE, 'ref_encodings.npz'), **out)


def make_saved_files():
    """files written by the reference's own save() (what e03 -s leaves for e04): L=128 spectra for encodings 1 and 2"""
    J = droplet_J(128, 1)
    for ee in (1, 2):
        ins = ref.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=3)
        ins.rotate_graph(1)
        ins.search_low_energy_spectrum(excitations_encoding=ee, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0)
        ins.save(os.path.join(HERE, 'ref_saved_spectrum_ee%d.npy' % ee))


def make_l1152():
    ins = ref.tnac4o(mode='Ising', Nx=12, Ny=12, Nc=8, J=droplet_J(1152, 1), beta=3)
    t0 = time.time()
    ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=32, max_dEng=1.0)
    t_search = time.time() - t0
    out = {}
    result_fields(ins, 'gs', out)
    out['n_shapes'] = np.int64(len(ins.d))
    t0 = time.time()
    ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    out['seconds_search'], out['seconds_decode'] = np.float64(t_search), np.float64(time.time() - t0)
    out['n_states'] = np.int64(len(ins.energy))
    out['energies_sorted_head'] = np.sort(ins.energy)[:4096]
    vals, counts = np.unique(np.round((ins.energy - ins.energy[0]) * 75).astype(np.int64), return_counts=True)
    out['level_75dE'], out['level_count'] = vals, counts
    np.savez_compressed(os.path.join(HERE, 'ref_l1152.npz'), **out)


def make_l1152_saved():
    """config 3 end to end as examples/e03 -s + e04 run it: search, save (the file is committed: 54 kB), load, decode;
    the fixture pins the decoded state SET (sha256 of the lexicographically sorted rows) and the whole spectrum"""
    import hashlib
    ins = ref.tnac4o(mode='Ising', Nx=12, Ny=12, Nc=8, J=droplet_J(1152, 1), beta=3)
    ins.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=1e-8, Dmax=32, max_dEng=1.0)
    fn = os.path.join(HERE, 'ref_saved_spectrum_l1152.npy')
    ins.save(fn)
    back = ref.load(fn)
    t0 = time.time()
    back.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    out = {'seconds_decode': np.float64(time.time() - t0), 'n_states': np.int64(len(back.energy))}
    st = np.ascontiguousarray(back.states.astype(np.int8))
    st = st[np.lexsort(st.T[::-1])]
    out['states_sha256'] = np.frombuffer(hashlib.sha256(st.tobytes()).digest(), dtype=np.uint8)
    out['energies_sorted'] = np.sort(back.energy)
    vals, counts = np.unique(np.round((back.energy - back.energy.min()) * 75).astype(np.int64), return_counts=True)
    out['level_75dE'], out['level_count'] = vals, counts
    out['states_head'] = st[:64]
    np.savez_compressed(os.path.join(HERE, 'ref_l1152_decoded.npz'), **out)


def make_j124_sweep(n_inst=20):
    """J124 C=8 instances 1..n_inst: coupling lists and the lines of results_C8_J124.txt (energy, degeneracy), the
    known answers of examples/e06 (4 rotations, lowest energy, largest degeneracy among the rotations reaching it)"""
    out = {}
    res = np.loadtxt('%s/Chimera_J124/C=8_J124/results_C8_J124.txt' % INST, dtype=np.int64)
    out['results'] = res[:n_inst]
    for k in range(1, n_inst + 1):
        raw = np.loadtxt('%s/Chimera_J124/C=8_J124/%03d.txt' % (INST, k))
        out['J_%03d_i' % k] = raw[:, 0].astype(np.int16)
        out['J_%03d_j' % k] = raw[:, 1].astype(np.int16)
        out['J_%03d_v' % k] = raw[:, 2].astype(np.int8)
        assert np.array_equal(out['J_%03d_v' % k], raw[:, 2])
    np.savez_compressed(os.path.join(HERE, 'ref_j124_sweep.npz'), **out)


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


def make_rmf():
    """examples/e05 / test_examples.py:107-136: RMF toy, dE = 3.1, the three encodings under different rotations, plus a
    ground-state search and a Gibbs run; add_noise consumes the global RNG, seeded here and in the tests"""
    out = {}
    for ee, rot in ((1, 0), (1, 1), (2, 2), (3, 3)):
        ins = ref.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
        if rot:
            ins.rotate_graph(rot=rot)
        if ee > 1:
            np.random.seed(7)
            ins.add_noise(amplitude=1e-7)
        ins.search_low_energy_spectrum(excitations_encoding=ee, M=1024, relative_P_cutoff=1e-12, Dmax=32, max_dEng=3.1, lim_hd=0)
        tag = 'ee%d_r%d' % (ee, rot)
        out[tag + '_gs_energy'], out[tag + '_gs_states'] = ins.energy.copy(), ins.states.copy()
        out[tag + '_gs_probability'] = np.asarray(ins.probability)
        out[tag + '_n_shapes'] = np.int64(len(ins.d))
        ins.decode_low_energy_states(max_dEng=3.1, max_states=100)
        out[tag + '_energy'], out[tag + '_states'] = ins.energy.copy(), ins.states.copy()
        out[tag + '_energy_check'] = ref.energy_RMF(rmf_model(), ins.states)
    for rot in (0, 1):
        ins = ref.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=4)
        if rot:
            ins.rotate_graph(rot=rot)
        ins.search_ground_state(M=64, relative_P_cutoff=1e-12, Dmax=32)
        tag = 'gs_r%d' % rot
        out[tag + '_energy'], out[tag + '_states'] = ins.energy.copy(), ins.states.copy()
        out[tag + '_probability'], out[tag + '_degeneracy'] = np.asarray(ins.probability), np.int64(ins.degeneracy)
        out[tag + '_discarded'] = np.float64(ins.discarded_probability)
    ins = ref.tnac4o(mode='RMF', Nx=5, Ny=3, J=rmf_model(), beta=1)
    np.random.seed(3)
    ins.gibbs_sampling(M=64, Dmax=32)
    out['gibbs_energy'], out['gibbs_states'] = ins.energy.copy(), ins.states.copy()
    np.savez_compressed(os.path.join(HERE, 'ref_rmf.npz'), **out)


def make_j124():
    raw = np.loadtxt('%s/Chimera_J124/C=8_J124/001.txt' % INST)
    J = [[int(r[0]) - 1, int(r[1]) - 1, float(r[2])] for r in raw]
    out = {}
    ins = ref.tnac4o(mode='Ising', Nx=8, Ny=8, Nc=8, J=J, beta=0.75)
    ins.precondition(mode='balancing')
    ins.search_ground_state(M=2 ** 12, relative_P_cutoff=1e-8, Dmax=8)
    result_fields(ins, 'gs', out)
    np.savez_compressed(os.path.join(HERE, 'ref_j124.npz'), **out)


def make_max_energy():
    out = {}
    fn = '%s/Chimera_droplet_instances/chimera128_spinglass_power/max_energy_otn2d.txt' % INST
    lines = {l.split(':')[0].strip(): l.split(':')[1].split() for l in open(fn)}
    for k in (1, 2, 3):
        vals = lines['%03d.txt' % k]
        out['file_%03d_energy' % k], out['file_%03d_bits' % k] = float(vals[0]), np.array(vals[1:], dtype=np.int8)
        ins = ref.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=ref.minus_Jij(droplet_J(128, k)), beta=3)
        ins.search_ground_state(M=2 ** 10, relative_P_cutoff=1e-8, Dmax=48)
        result_fields(ins, 'max_%03d' % k, out)
        out['max_%03d_bits' % k] = ins.binary_states()[0].astype(np.int8)
    np.savez_compressed(os.path.join(HERE, 'ref_max_energy.npz'), **out)


def make_noise():
    out = {}
    for shape in ((4, 4), (8, 2)):
        for rot in (0, 1, 2, 3):
            ins = ref.tnac4o(mode='Ising', Nx=shape[0], Ny=shape[1], Nc=8, J=droplet_J(128, 1), beta=3)
            if rot:
                ins.rotate_graph(rot=rot)
            np.random.seed(7)
            ins.add_noise(amplitude=1e-7)
            c = ins.J.tocoo()
            k = np.lexsort((c.col, c.row))
            tag = 'n_%dx%d_r%d_' % (shape[0], shape[1], rot)
            out[tag + 'row'], out[tag + 'col'], out[tag + 'val'] = c.row[k].astype(np.int32), c.col[k].astype(np.int32), c.data[k]
            out[tag + 'order'] = np.asarray(ins.order)
    np.savez_compressed(os.path.join(HERE, 'ref_noise.npz'), **out)


if __name__ == '__main__':
    for what in sys.argv[1:]:
        t0 = time.time()
        {'instances': make_instances, 'instances10': make_instances10, 'small': make_small,
         'l512': lambda: make_big(512, 16, 1024, 'ref_l512.npz'),
         'l2048': lambda: make_big(2048, 32, 1024, 'ref_l2048.npz'),
         'l2048m4096': lambda: make_big(2048, 32, 4096, 'ref_l2048_m4096.npz'),
         'gibbs2048': make_gibbs_l2048, 'encodings': make_encodings, 'saved': make_saved_files,
         'l1152': make_l1152, 'l1152saved': make_l1152_saved, 'j124': make_j124, 'j124sweep': make_j124_sweep,
         'rmf': make_rmf, 'noise': make_noise, 'maxenergy': make_max_energy}[what]()
        print(what, 'done in %.1f s' % (time.time() - t0), flush=True)
