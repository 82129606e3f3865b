"""Multi-process GPU tests of the sharded paths: the branch batch of one search spread over
two ranks, and Gibbs samples spread over two ranks, must reproduce the single-rank results bit for bit.

With two or more GPUs the ranks use one GPU each and NCCL; on a single-GPU box both ranks share cuda:0 and the
(tiny) exchange buffers are staged through the host over gloo -- the kernels, the partition and the replicated
selection are the same code either way."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import SHAPES, droplet_couplings, golden

pytestmark = pytest.mark.gpu
warnings.filterwarnings('ignore')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fields(ins):
    return {'energy': np.asarray(ins.energy).copy(), 'states': np.asarray(ins.states).copy(),
            'probability': np.asarray(ins.probability).copy(), 'degeneracy': int(ins.degeneracy),
            'discarded': float(ins.discarded_probability), 'negative': float(ins.negative_probability)}


def _worker(rank, world, port, nccl, case, out):
    warnings.filterwarnings('ignore')
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank if nccl else 0)
    torch.cuda.set_device(dev)
    if nccl:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import tnac4o_b200
    from tnac4o_b200 import parallel
    L, beta, M, D, kind = case
    Nx, Ny = SHAPES[L]
    J = droplet_couplings(L, 1)
    new = lambda: tnac4o_b200.tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=8, J=J, beta=beta, device=dev)
    res = {}
    if kind == 'search':
        sh = parallel.BranchShards()
        ins = new()
        ins.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D, shards=sh)
        res['sharded'] = _fields(ins)
        res['marginals'] = (ins.stats['marginals'], ins.stats['marginals_this_rank'])
        res['backend_host_staged'] = sh.host
        one = new()
        one.native_search = False
        one.search_ground_state(M=M, relative_P_cutoff=1e-8, Dmax=D)
        res['single'] = _fields(one)
        res['marginals_single'] = one.stats['marginals']
    elif kind == 'spectrum':
        sh = parallel.BranchShards()
        ins = new()
        ins.search_low_energy_spectrum(excitations_encoding=1, M=M, relative_P_cutoff=1e-8, Dmax=D, max_dEng=1.0, shards=sh)
        ins.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
        order = np.lexsort(ins.states.T[::-1])
        res['sharded'] = {'energy': ins.energy[order], 'states': ins.states[order]}
    elif kind == 'gibbs':
        ins = new()
        np.random.seed(1)
        ins.gibbs_sampling(M=M, Dmax=D, shard=(rank, world))
        E, S = parallel.gather_samples(ins.energy, ins.states)
        res['sharded'] = {'energy': E, 'states': S}
        one = new()
        np.random.seed(1)
        one.gibbs_sampling(M=M, Dmax=D)
        res['single'] = {'energy': one.energy, 'states': one.states}
    torch.cuda.synchronize(dev)
    out.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _run(case, world=2):
    nccl = torch.cuda.device_count() >= world
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nccl, case, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


def _same(a, b):
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], (k, a[k], b[k])


@pytest.mark.parametrize('case', [(128, 3, 256, 8, 'search'), (512, 3, 1024, 16, 'search')])
def test_branch_sharded_search_equals_single_rank(case):
    got = _run(case)
    for r in (0, 1):
        _same(got[r]['single'], got[r]['sharded'])           # bit-identical to the unsharded site loop on the same GPU
    _same(got[0]['sharded'], got[1]['sharded'])              # and the replicas agree
    tot, mine = zip(*(got[r]['marginals'] for r in (0, 1)))
    assert tot[0] == tot[1] == sum(mine) == got[0]['marginals_single']
    assert 0 < mine[0] and 0 < mine[1]
    if case[0] == 512:
        z = golden('ref_l512.npz')
        assert abs(got[0]['sharded']['energy'][0] - float(z['gs_energy'][0])) < 1e-10
        assert np.array_equal(got[0]['sharded']['states'][0], z['gs_states'][0])


def test_branch_sharded_spectrum_matches_reference_fixture():
    got = _run((128, 3, 1024, 16, 'spectrum'))
    z = golden('ref_small.npz')
    for r in (0, 1):
        assert np.array_equal(got[r]['sharded']['states'], z['sp_r0_states'])
        np.testing.assert_allclose(got[r]['sharded']['energy'], z['sp_r0_energy'], rtol=0, atol=1e-10)


def test_sample_sharded_gibbs_equals_single_rank():
    got = _run((128, 1, 128, 16, 'gibbs'))
    for r in (0, 1):
        _same(got[r]['single'], got[r]['sharded'])
    z = golden('ref_small.npz')
    same = np.all(got[0]['sharded']['states'] == z['gibbs_states'], axis=1)
    assert same.mean() >= 0.98                                              # as tests/test_solver_gpu.py::test_gibbs_sampling
