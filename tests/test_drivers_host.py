"""CPU tests of the host-only parts of tnac4o_b200/drivers.py (text outputs of examples/e02 and e06)."""
import numpy as np


class FakeSolution:
    L = 6
    energy = np.array([-1.5, 0.25])

    def binary_states(self):
        return np.array([[1, 0, 1, 1, 0, 2], [0, 0, 0, 1, 1, 1]], dtype=np.int8)


def test_states_txt_matches_e02_format(tmp_path):
    from tnac4o_b200 import drivers
    fn = str(tmp_path / 'gibbs.txt')
    drivers.write_states_txt(FakeSolution(), fn)
    lines = open(fn).read().strip().split('\n')
    assert lines[0].startswith('# One line per state; First column is the energy')
    assert lines[1] == '-1.500000 1 0 1 1 0 2' and lines[2] == '0.250000 0 0 0 1 1 1'


def test_gs_degeneracy_txt(tmp_path):
    from tnac4o_b200 import drivers
    fn = str(tmp_path / 'J124.txt')
    drivers.write_gs_degeneracy_txt(fn, -2309.0, 1152)
    assert open(fn).read().split('\n')[:3] == ['# Energy and degeneracy', '-2309', '1152']
