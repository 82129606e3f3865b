"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL).

Three ways the contraction path spreads over GPUs:
  * independent instances / rotations: replicas, no data-path collective (bench.py --gpus N);
  * Gibbs samples: every rank draws the same uniforms and keeps its slice (:class:`UniformStream`), results are
    gathered once at the end (:func:`gather_samples`);
  * the branch batch of ONE search: :class:`BranchShards` -- every rank evaluates the right environments and the
    conditional marginals of its slice of the live branches, the candidate log-probabilities are all-gathered, and
    the selection (cut-off, merge, global top-M) runs replicated and deterministic on every rank."""
import numpy as np
import torch
import torch.distributed as dist


def partition(n, world, rank):
    """contiguous, balanced slice [lo, hi) of n units owned by `rank`"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def max_over_ranks(seconds, device=None):
    """timing rule of the bench contract: the slowest rank defines the step"""
    rank, world = world_info()
    if world == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device=None):
    rank, world = world_info()
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def gather_samples(energy, states):
    """concatenate per-rank Gibbs samples in rank order (every rank receives the full arrays); the data travel as
    tensors (NCCL: device buffers), cell states as int32"""
    rank, world = world_info()
    if world == 1:
        return energy, states
    energy, states = np.asarray(energy, dtype=np.float64), np.asarray(states)
    nccl = dist.get_backend() == 'nccl'
    dev = torch.device('cuda', torch.cuda.current_device()) if nccl else torch.device('cpu')
    counts = [None] * world
    dist.all_gather_object(counts, int(len(energy)))
    cap = max(counts)
    e = torch.zeros(cap, dtype=torch.float64, device=dev)
    e[:len(energy)] = torch.from_numpy(energy).to(dev)
    st = torch.zeros((cap, states.shape[1]), dtype=torch.int32, device=dev)
    st[:len(energy)] = torch.from_numpy(states.astype(np.int32)).to(dev)
    E = torch.empty((world * cap,), dtype=torch.float64, device=dev)
    S = torch.empty((world * cap, states.shape[1]), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(E, e)
    dist.all_gather_into_tensor(S, st)
    E, S = E.cpu().numpy(), S.cpu().numpy()
    keep = np.concatenate([np.arange(r * cap, r * cap + c) for r, c in enumerate(counts)])
    return E[keep], S[keep].astype(states.dtype)


_SIGN = -2 ** 63


class BranchShards:
    """Slice of the live branches owned by this rank plus the collectives of one site of the branch-and-bound.

    Branch records (vind, states, Eng, prob, deg, RL) and the boundary MPS are replicated; the work that scales with
    the number of branches -- right environments (tn_rr_level) and conditional marginals (tn_gemm + tn_marginals) --
    is sharded by contiguous, equally sized chunks of rows (the last chunk may be short or empty).  Buffers that are
    gathered must have room for ``padded(B)`` rows."""

    def __init__(self, group=None, stage_through_host=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # NCCL moves device buffers directly; any other backend (gloo in the single-GPU tests) goes through the host
        self.host = (dist.get_backend(group) != 'nccl') if stage_through_host is None else bool(stage_through_host)
        self.bytes_gathered = 0

    def chunk(self, B):
        return -(-int(B) // self.world)

    def padded(self, B):
        return self.chunk(B) * self.world

    def slice(self, B):
        c = self.chunk(B)
        lo = min(self.rank * c, int(B))
        return lo, min(lo + c, int(B))

    def allgather_rows(self, buf, B):
        """buf (rows >= padded(B), ...) holds this rank's rows [lo, hi); afterwards every rank holds rows [0, B)"""
        c = self.chunk(B)
        full = buf[:c * self.world]
        mine = full[self.rank * c:(self.rank + 1) * c]
        self.bytes_gathered += full.numel() * full.element_size()
        if self.host and full.is_cuda:
            out = torch.empty(full.shape, dtype=full.dtype)
            dist.all_gather_into_tensor(out, mine.cpu().contiguous(), group=self.group)
            full.copy_(out)
        else:
            dist.all_gather_into_tensor(full, mine.clone(), group=self.group)
        return buf

    def allgather_small(self, t):
        """all-gather of a small 1-d tensor (same length on every rank) -> 1-d tensor of world * len, rank-major"""
        self.bytes_gathered += t.numel() * t.element_size() * self.world
        if self.host and t.is_cuda:
            out = torch.empty(self.world * t.numel(), dtype=t.dtype)
            dist.all_gather_into_tensor(out, t.cpu().contiguous(), group=self.group)
            return out.to(t.device)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def _allreduce(self, t, op):
        if self.host and t.is_cuda:
            h = t.cpu()
            dist.all_reduce(h, op=op, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    def allreduce_max_ordered_(self, bits):
        """max over ranks of an int64 tensor that holds order-preserving *unsigned* encodings (common.cuh:
        ordered_bits); flipping the top bit turns the unsigned order into int64 order"""
        t = torch.bitwise_xor(bits, _SIGN)
        self._allreduce(t, dist.ReduceOp.MAX)
        bits.copy_(torch.bitwise_xor(t, _SIGN))
        return bits

    def allreduce_min_(self, t):
        return self._allreduce(t, dist.ReduceOp.MIN)

    def allreduce_sum_(self, t):
        return self._allreduce(t, dist.ReduceOp.SUM)

    def broadcast_(self, t, src=0):
        if self.host and t.is_cuda:
            h = t.cpu()
            dist.broadcast(h, src=src, group=self.group)
            t.copy_(h)
        else:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def same_everywhere(self, obj):
        """True when every rank passes an equal (picklable) object"""
        parts = [None] * self.world
        dist.all_gather_object(parts, obj, group=self.group)
        return all(p == parts[0] for p in parts)


class UniformStream:
    """The reference draws np.random.rand(M) once per site from the global numpy stream (tnac4o.py:617).  With the
    samples sharded over ranks every rank draws the same M numbers (same seed) and keeps its slice, so the union of
    the ranks' samples equals the single-GPU run."""

    def __init__(self, M, rank=0, world=1):
        self.M = M
        self.lo, self.hi = partition(M, world, rank)

    def draw(self):
        return np.random.rand(self.M)[self.lo:self.hi]


class StreamPool:
    """Persistent worker threads, each bound to its own CUDA stream (and therefore to its own library context and its
    own allocator pool): independent solver jobs submitted to the pool overlap on one GPU.

    The boundary-MPS build of a single instance is a latency-bound chain of small kernels (cluster QR panels and
    Jacobi rounds occupy 8 of 148 SMs), so independent instances / rotations / beta steps overlap almost for free."""

    def __init__(self, workers, device=None):
        import queue
        import threading
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.workers = workers
        self._q = queue.Queue()
        self._threads = [threading.Thread(target=self._loop, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    def _loop(self):
        with torch.cuda.device(self.device):
            stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(stream):
                while True:
                    item = self._q.get()
                    if item is None:
                        from ._native import Context
                        stream.synchronize()
                        Context.release_thread()        # the thread's library contexts (scratch, pinned buffer, pool)
                        return
                    job, slot, results, errors, done = item
                    try:
                        results[slot] = job()
                        stream.synchronize()
                    except BaseException as e:      # noqa: BLE001  (re-raised by run() in the caller's thread)
                        errors[slot] = e
                    done.release()

    def run(self, jobs):
        """run zero-argument callables concurrently; returns their results in order, re-raises the first exception"""
        import threading
        results, errors = [None] * len(jobs), [None] * len(jobs)
        done = threading.Semaphore(0)
        for i, job in enumerate(jobs):
            self._q.put((job, i, results, errors, done))
        for _ in jobs:
            done.acquire()
        for e in errors:
            if e is not None:
                raise e
        return results

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()


_POOLS = {}


def run_concurrently(jobs, device=None):
    """Run independent solver jobs on ONE GPU at the same time through a cached :class:`StreamPool` with one worker
    per job.  ``jobs`` is a list of zero-argument callables; returns their results in order."""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    key = (dev.index, len(jobs))
    if key not in _POOLS:
        _POOLS[key] = StreamPool(len(jobs), dev)
    return _POOLS[key].run(jobs)
