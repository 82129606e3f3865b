"""On-disk format: files written by the reference's save() are read by tnac4o_b200.load(), and
files written by tnac4o_b200's save() are read -- and decoded -- by the reference's own load() / e04 path.

The second direction imports the reference and therefore only runs where it is mounted (the build container)."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, golden

warnings.filterwarnings('ignore')
REF = '/root/reference'


@pytest.mark.parametrize('ee', [1, 2])
def test_load_reads_files_written_by_the_reference(ee):
    """tnac4o.save (tnac4o.py:200-233) -> tnac4o_b200.load: every field, and the host half of decode_low_energy_states
    (the enumeration of droplet combinations) on the loaded structure gives the reference's spectrum"""
    import tnac4o_b200
    z = golden('ref_small.npz') if ee == 1 else golden('ref_encodings.npz')
    ins = tnac4o_b200.load(os.path.join(GOLDEN, 'ref_saved_spectrum_ee%d.npy' % ee))
    assert (ins.mode, ins.Nx, ins.Ny, ins.Nc, ins.beta) == ('Ising', 4, 4, 8, 3)
    assert ins.excitations_encoding == ee and ins.states.shape == (ins.energy.shape[0], 16) and ins.states.dtype == np.int8
    assert len(ins.ind0) == 4 and len(ins.ind0[0]) == 4 and len(ins.d) > 0 and ins.free_d >= len(ins.d)
    if ee == 2:
        assert ins.adj.shape == (128, 128)
    else:
        return      # encoding 1 is enumerated on the device: tests/test_decode_host.py (algorithm model) + the GPU tests
    Eng, flip = ins._exc_unpack(max_dEng=1.0, max_states=2 ** 20)
    assert len(Eng) == 31
    want = z['sp_r1_energy'] if ee == 1 else z['ee2_r1_hd0_energy']
    np.testing.assert_allclose(np.sort(Eng + ins.energy[0]), np.sort(want), atol=1e-10)
    # the flips reproduce the stored states (host XOR here; the device kernel tn_apply_droplets in the GPU tests)
    states = np.repeat(ins.states[:1], len(Eng), axis=0)
    for i, keys in enumerate(flip):
        for k in keys:
            dpos, dstate = ins.d[k]
            states[i, dpos] = np.bitwise_xor(states[i, dpos], dstate)
    want_states = z['sp_r1_states'] if ee == 1 else z['ee2_r1_hd0_states']
    assert np.array_equal(states[np.lexsort(states.T[::-1])], want_states)


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference is mounted in the build container only')
@pytest.mark.parametrize('ee', [1, 2, 3])
def test_reference_decodes_files_written_by_this_package(J128, ee, tmp_path):
    """tnac4o_b200 save() -> the reference's load() + decode_low_energy_states (examples/e04): 31 states, same spectrum.
    The spectrum structure comes from the oracle's search (the device search needs a GPU; its structure is compared with
    the oracle's in the GPU tests), the file is written by the product's save()."""
    import tnac4o_b200
    from oracle import RefSolver
    if not hasattr(np, 'int'):
        np.int = int                                   # the reference still uses np.int on the encoding-3 path
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import tnac4o as reference
    src = RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=J128, beta=3)
    src.rotate_graph(1)
    src.search_low_energy_spectrum(excitations_encoding=ee, M=1024, relative_P_cutoff=1e-8, Dmax=16, max_dEng=1.0)
    sol = tnac4o_b200.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J128, beta=3)
    sol.rotate_graph(1)
    sol.excitations_encoding = ee
    for name in ('energy', 'states', 'probability', 'degeneracy', 'discarded_probability', 'negative_probability',
                 'd', 'invd', 'el', 'free_d'):
        setattr(sol, name, getattr(src, name))
    if ee > 1:
        sol.adj = src.adj
    fn = os.path.join(tmp_path, 'from_b200_ee%d.npy' % ee)
    sol.save(fn)
    back = reference.load(fn)
    back.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
    z = golden('ref_small.npz') if ee == 1 else golden('ref_encodings.npz')
    tag = 'sp_r1' if ee == 1 else 'ee%d_r1_hd0' % ee
    order = np.lexsort(back.states.T[::-1])
    assert len(back.energy) == 31 and np.array_equal(back.states[order], z[tag + '_states'])
    np.testing.assert_allclose(back.energy[order], z[tag + '_energy'], atol=1e-10)
    assert np.array_equal(back.binary_states(), sol_bits(tnac4o_b200, back, J128))


def sol_bits(pkg, ref_ins, J):
    """binary_states of the product for the same states (tnac4o.py:261-286)"""
    sol = pkg.tnac4o(mode='Ising', Nx=4, Ny=4, Nc=8, J=J, beta=3)
    sol.states = ref_ins.states
    return sol.binary_states()
