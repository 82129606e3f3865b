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


def _shard_worker(rank, world, port, out):
    """one 'site' of the sharded branch-and-bound with host tensors: each rank fills the rows of its slice, the
    gather + max reduce must leave every rank with what a single rank would hold"""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tnac4o_b200 import parallel
    sh = parallel.BranchShards()
    res = []
    for B in (1, 2, 5, 64, 1023):                           # fewer branches than ranks, ragged and even splits
        rng = np.random.default_rng(B)
        full = rng.standard_normal((B, 256))                # what one rank with all branches computes
        lo, hi = sh.slice(B)
        assert (hi - lo) <= sh.chunk(B) and sh.padded(B) >= B
        buf = torch.full((B + world, 256), np.nan, dtype=torch.float64)
        buf[lo:hi] = torch.from_numpy(full[lo:hi])
        sh.allgather_rows(buf, B)
        # best candidate: per-rank maximum in the order-preserving unsigned encoding, stored in an int64 tensor
        def ordered(x):
            u = np.array([x], dtype=np.float64).view(np.uint64)[0]
            u = (~u) if (u >> np.uint64(63)) else (u | np.uint64(1 << 63))
            return np.array([u], dtype=np.uint64).view(np.int64)
        local = full[lo:hi].max() if hi > lo else None
        bits = torch.from_numpy(ordered(local) if local is not None else np.zeros(1, dtype=np.int64)).clone()
        sh.allreduce_max_ordered_(bits)
        gmin = torch.tensor([full[lo:hi].min() if hi > lo else 1.0], dtype=torch.float64)
        sh.allreduce_min_(gmin)
        cnt = sh.allreduce_sum_(torch.tensor([float(hi - lo)], dtype=torch.float64))
        res.append((B, np.array_equal(buf[:B].numpy(), full), int(bits.item()) == int(ordered(full.max())[0]),
                    float(gmin.item()) == min(full.min(), 1.0), cnt.item() == B))
    t = torch.arange(6, dtype=torch.float64) * (rank + 1)
    sh.broadcast_(t, src=0)
    same = (sh.same_everywhere([(1, 2), (3,)]), sh.same_everywhere(rank))
    if rank == 1:
        out.put((res, t.tolist(), same))
    dist.destroy_process_group()


def test_two_rank_branch_shards_site_exchange():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res, t, same = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [1, 2, 5, 64, 1023]
    assert all(all(r[1:]) for r in res), res
    assert t == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    assert same == (True, False)
