"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: partition, timing reduction, sample gathering and the
sliced uniform stream that makes sharded Gibbs sampling reproduce the single-GPU draw order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tnac4o_b200 import parallel
    M = 11
    lo, hi = parallel.partition(M, world, rank)
    np.random.seed(7)
    stream = parallel.UniformStream(M, rank, world)
    draws = np.stack([stream.draw() for _ in range(3)])                  # three "sites"
    energy = draws.sum(0)
    states = np.repeat(np.arange(lo, hi)[:, None], 4, axis=1)
    E, S = parallel.gather_samples(energy, states)
    t = parallel.max_over_ranks(1.0 + rank)
    tot = parallel.sum_over_ranks([hi - lo, 1])
    if rank == 0:
        out.put((E, S, t, tot))
    dist.destroy_process_group()


def test_two_rank_gibbs_plumbing_matches_single_rank():
    from tnac4o_b200 import parallel
    assert [parallel.partition(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    E, S, t, tot = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.random.seed(7)
    ref = np.stack([np.random.rand(11) for _ in range(3)]).sum(0)        # what one rank with all samples draws
    assert np.array_equal(E, ref)
    assert np.array_equal(S[:, 0], np.arange(11))
    assert t == 2.0 and tot == [11.0, 2.0]
