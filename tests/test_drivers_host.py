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


class _Found:
    def __init__(self, e, d):
        self.energy, self.degeneracy = np.array([e]), d


# rotation -> (energy, degeneracy) a fake search returns: rotations 1 and 3 reach the lowest energy, 3 counts more states
_FAKE = {0: (-10.0, 7), 1: (-12.0, 3), 2: (-11.0, 9), 3: (-12.0, 5)}


def _fake_search_gs(J, Nx, Ny, rot=0, **kw):
    return _Found(*_FAKE[rot])


def test_e06_selection_rule_single_process(monkeypatch):
    """examples/e06:98-110: lowest energy over the four rotations, then the largest degeneracy among those reaching it"""
    from tnac4o_b200 import drivers
    monkeypatch.setattr(drivers, 'search_gs', _fake_search_gs)
    E, deg, per = drivers.search_gs_degeneracy(None, 8, 8, concurrent=False)
    assert (E, deg) == (-12.0, 5)
    assert per == [(r,) + _FAKE[r] for r in range(4)]


def _e06_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tnac4o_b200 import drivers
    seen = []

    def search(J, Nx, Ny, rot=0, **kw):
        seen.append(rot)
        return _Found(*_FAKE[rot])

    drivers.search_gs = search
    E, deg, per = drivers.search_gs_degeneracy(None, 8, 8, concurrent=False)
    out.put((rank, seen, E, deg, per))
    dist.barrier()
    dist.destroy_process_group()


def test_e06_rotations_spread_over_two_ranks():
    """world-size-2 gloo: rank r searches the rotations r, r + 2; every rank ends with all four pairs and the same answer"""
    import torch.multiprocessing as mp
    from test_parallel import _free_port
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_e06_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 2] and got[1][1] == [1, 3]
    for _, _, E, deg, per in got:
        assert (E, deg) == (-12.0, 5) and per == [(r,) + _FAKE[r] for r in range(4)]
