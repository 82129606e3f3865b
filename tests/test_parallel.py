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
        # survivors-only exchange of the sharded selection (solver._select_sharded): every rank keeps the candidates of its
        # slice above a common threshold; counts, padded ids and padded values are all-gathered
        thr = full.max() - 1.0
        flat = full.reshape(-1)
        mine = np.nonzero(flat[lo * 256:hi * 256] > thr)[0] + lo * 256
        counts = sh.allgather_small(torch.tensor([len(mine)], dtype=torch.int64)).tolist()
        kmax = max(counts)
        ids = torch.zeros(kmax, dtype=torch.int32)
        vals = torch.zeros(kmax, dtype=torch.float64)
        ids[:len(mine)] = torch.from_numpy(mine.astype(np.int32))
        vals[:len(mine)] = torch.from_numpy(flat[mine])
        all_i, all_v = sh.allgather_small(ids).view(world, kmax), sh.allgather_small(vals).view(world, kmax)
        got_i = np.concatenate([all_i[r, :n].numpy() for r, n in enumerate(counts)])
        got_v = np.concatenate([all_v[r, :n].numpy() for r, n in enumerate(counts)])
        want = np.nonzero(flat > thr)[0]
        surv_ok = np.array_equal(np.sort(got_i), want) and np.array_equal(got_v, flat[got_i])
        res.append((B, np.array_equal(buf[:B].numpy(), full), int(bits.item()) == int(ordered(full.max())[0]),
                    float(gmin.item()) == min(full.min(), 1.0), cnt.item() == B, surv_ok))
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
