"""CPU tests of the host half of the device decode (tnac4o_b200/csrc/decode.cu): the tree flattening of
solver._flatten_tree and -- on the flattened arrays -- a plain-Python model of the level-synchronous algorithm the kernels
implement (immutable records with `from` / `below` links, lazy pops, frontier waves), checked against the reference's
decoded spectrum (fixtures written by the reference's own save() / decode_low_energy_states)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden


def model_decode(dE, key, first, last, child_ptr, child_idx, nsites, max_dEng, max_states):
    """what decode_count/emit/pop kernels compute, one combination at a time"""
    E, rec, frame = [0.0], [0], [0]
    rec_node, rec_from, rec_below = [0], [0], [0]
    ends = set(int(x) for x in last[1:])
    for nn in range(nsites - 1, -1, -1):
        if nn not in ends:
            continue
        for i in range(len(E)):                       # lazy pops of all sites above nn
            f = frame[i]
            while first[rec_node[f]] >= nn + 1:
                f = rec_below[f]
            frame[i] = f
        lo, hi = 0, len(E)
        while hi > lo:
            new = []
            for i in range(lo, hi):
                node = rec_node[frame[i]]
                for k in range(child_ptr[node], child_ptr[node + 1]):
                    ch = child_idx[k]
                    if last[ch] == nn and E[i] + dE[ch] <= max_dEng:
                        new.append((E[i] + dE[ch], ch, rec[i], frame[i]))
                    elif last[ch] > nn:
                        break
            for e, ch, fr, below in new:
                r = len(rec_node)
                rec_node.append(ch); rec_from.append(fr); rec_below.append(below)
                E.append(e); rec.append(r); frame.append(r)
            lo, hi = hi, len(E)
        if len(E) > max_states:
            keep = sorted(range(len(E)), key=lambda i: (E[i], i))[:max_states]
            E, rec, frame = [E[i] for i in keep], [rec[i] for i in keep], [frame[i] for i in keep]
    order = sorted(range(len(E)), key=lambda i: (E[i], i))
    flips = []
    for i in order:
        r, ks = rec[i], []
        while r > 0:
            ks.append(int(key[rec_node[r]]))
            r = rec_from[r]
        flips.append(ks)
    return np.array([E[i] for i in order]), flips


def apply_flips(ins, slot_keys, flips):
    states = np.repeat(ins.states[:1], len(flips), axis=0)
    for i, ks in enumerate(flips):
        for k in ks:
            dpos, dstate = ins.d[slot_keys[k]]
            states[i, dpos] = np.bitwise_xor(states[i, dpos], dstate)
    return states


def test_flattened_tree_model_reproduces_reference_spectrum_l128():
    import tnac4o_b200
    z = golden('ref_small.npz')
    ins = tnac4o_b200.load(os.path.join(GOLDEN, 'ref_saved_spectrum_ee1.npy'))
    slot, drop_ptr, drop_pos, drop_xor = ins._droplet_csr_host()
    keys = sorted(slot, key=slot.get)
    dE, key, first, last, cp, ci = ins._flatten_tree(slot)
    assert cp[0] == 0 and cp[-1] == len(ci) == len(dE) - 1 and first[0] == -1 and last[0] == 15
    assert np.all(first[1:] <= last[1:]) and key.max() < len(keys)
    Eng, flips = model_decode(dE, key, first, last, cp, ci, 16, 1.0, 2 ** 20)
    assert len(Eng) == 31 and np.all(np.diff(Eng) >= 0)
    np.testing.assert_allclose(Eng + ins.energy[0], np.sort(z['sp_r1_energy']), atol=1e-10)
    states = apply_flips(ins, keys, flips)
    assert np.array_equal(states[np.lexsort(states.T[::-1])], z['sp_r1_states'])
    # the top-K cut keeps the lowest energies
    Eng5, _ = model_decode(dE, key, first, last, cp, ci, 16, 1.0, 5)
    np.testing.assert_array_equal(Eng5, Eng[:5])


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, 'ref_saved_spectrum_l1152.npy')), reason='fixture missing')
def test_flattened_tree_of_config3():
    """config 3: the reference's saved L=1152 file flattens to the tree the survey measured (109 shapes, 541 nodes)"""
    import tnac4o_b200
    ins = tnac4o_b200.load(os.path.join(GOLDEN, 'ref_saved_spectrum_l1152.npy'))
    slot, drop_ptr, drop_pos, drop_xor = ins._droplet_csr_host()
    dE, key, first, last, cp, ci = ins._flatten_tree(slot)
    assert len(slot) == int(golden('ref_l1152.npz')['n_shapes'])
    assert len(dE) - 1 == cp[-1] and last.max() <= 143 and drop_ptr[-1] == len(drop_pos)
