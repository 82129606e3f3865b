#!/usr/bin/env python
"""Headline benchmark: L=2048 chimera ground-state search, seconds per instance (+ branch-marginals/s).

    python bench.py --gpus N --steps K --warmup W            # the sm_100a path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm (numpy port) on the host cores

A step is one pass of the hot path over one batch of `--batch` independent instances per GPU: for each instance the
boundary-MPS build (_setup_rhoT) + branch-and-bound (search_ground_state) at M = 2^10, Dmax = 32, beta = 3,
relative_P_cutoff = 1e-8, no preconditioning (BASELINE.json config 4, the M = 2^10 variant the north star quotes its
target on).  The instances of a batch run concurrently, one host thread and one CUDA stream each: a single
boundary-MPS build is a latency-bound chain that keeps 8 of 148 SMs busy, so instance-level concurrency is how one
GPU is filled.  Instance 0 of rank 0 is droplet instance 001 (checked against the reference's golden energy); all
others are synthetic instances on the same coupling pattern (weak scaling: independent instances, no data-path
collective -- DESIGN.md section 6).  `value` is whole-job seconds per instance; the single-instance latency is
reported next to it.
"""
import argparse
import json
import os

# BLAS / OpenMP thread pools are sized when the libraries load: pin them BEFORE numpy is imported, here and (through the
# inherited environment) in every spawned worker of the CPU legs.  One BLAS thread per process is the fastest setting
# for this workload (L=512 takes 43.7 s with 1 thread, 78.8 s with 8); parallelism of the CPU
# arm comes from independent single-thread workers, one per host core.
for _v in ('OPENBLAS_NUM_THREADS', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
    os.environ[_v] = '1'
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')     # one hardware queue per concurrent solver stream
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CFG = dict(L=2048, Nx=16, Ny=16, Nc=8, beta=3.0, M=2 ** 10, Dmax=32, relative_P_cutoff=1e-8)
METRIC = 'L=2048 chimera ground-state search, seconds per instance'
UNIT = 's/instance'
NCU_TRAFFIC_LARGE_GEMM = {'bytes': 35686912 + 315136,
                          'note': 'dram__bytes_read.sum + dram__bytes_write.sum of one gemm_tma_kernel launch (the 8192x512x512 attach GEMM) '
                                  'from this round\'s ncu --set full capture of this build (profiles/r2b_ncu_summary.txt): 35.7 MB read + '
                                  '0.3 MB written against 69 MB algorithmic operand + result bytes -- the 33.5 MB result stays in the '
                                  '126 MB L2 for its consumer.  A profiler cannot run inside the timed region; the other primitives of the '
                                  'roofline move < 2 MB per launch (QR panel: 1.1 MB read, Jacobi: 4.2 MB)'}


def instance_couplings(rank):
    """rank 0: droplet instance 001 as e01 prepares it; rank r > 0: same coupling pattern, values redrawn from the
    file's value set with seed r (synthetic family A)"""
    from conftest import droplet_couplings
    J = droplet_couplings(CFG['L'], 1)
    if rank == 0:
        return J
    rng = np.random.default_rng(rank)
    vals = np.array([v for _, _, v in J])
    return [[i, j, float(v)] for (i, j, _), v in zip(J, rng.permutation(vals))]


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------- GPU arm
def fp64_peak(torch, dev):
    """FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so it is measured here the way the file's
    bf16 number was: cuBLAS DGEMM 8192^3 through torch.matmul, best of 5, CUDA events."""
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = float('inf')
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e12


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def extra_configs(torch, dist, tnac4o_b200, parallel, dev, rank, world):
    """one run each of config 4 (M = 2^12) and config 5 (10^5 samples at beta = 1) spread over all ranks; device-timed,
    max over ranks"""
    J = instance_couplings(0)                      # the same instance on every rank
    new = lambda beta: tnac4o_b200.tnac4o(mode='Ising', Nx=CFG['Nx'], Ny=CFG['Ny'], Nc=CFG['Nc'], J=J, beta=beta, device=dev)

    def timed(fn):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record(); e1.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3, device=dev)

    out = {}
    shards = parallel.BranchShards() if world > 1 else None
    ins = new(CFG['beta'])
    ins._site_tables()
    search = lambda: ins.search_ground_state(M=2 ** 12, relative_P_cutoff=CFG['relative_P_cutoff'], Dmax=CFG['Dmax'], shards=shards)
    search()                                       # warm-up (allocator, NCCL channels)
    t = timed(search)
    s_search = parallel.max_over_ranks(ins.stats['seconds_search'], device=dev)
    out['config4_M4096'] = {'workload': 'e01 ground-state search L=2048, M=2^12, Dmax=32, branch batch sharded over %d GPU(s)' % world,
                            'seconds_per_instance': t, 'seconds_rhoT_replicated': ins.stats['seconds_rhoT'],
                            'seconds_search': s_search, 'branch_marginals': int(ins.stats['marginals']),
                            'branch_marginals_per_s': ins.stats['marginals'] / s_search,
                            'bytes_allgathered_per_rank': int(ins.stats.get('bytes_gathered', 0)),
                            'energy': float(ins.energy[0]), 'degeneracy': int(ins.degeneracy)}
    M5 = 100000
    gib = new(1.0)
    gib._site_tables()

    def sample():
        np.random.seed(1)
        gib.gibbs_sampling(M=M5, Dmax=CFG['Dmax'], shard=(rank, world))
    t = timed(sample)
    s_samp = parallel.max_over_ranks(gib.stats['seconds_search'], device=dev)
    E, S = parallel.gather_samples(gib.energy, gib.states)
    out['config5_gibbs'] = {'workload': 'e02 Gibbs sampling L=2048, beta=1, %d samples over %d GPU(s), Dmax=32' % (M5, world),
                            'seconds_total': t, 'seconds_rhoT_replicated': gib.stats['seconds_rhoT'], 'seconds_sampling': s_samp,
                            'samples_per_s': M5 / s_samp, 'branch_marginals_per_s': M5 * CFG['Nx'] * CFG['Ny'] / s_samp,
                            'mean_energy': float(np.mean(E)), 'samples_gathered': int(len(E))}
    # config 3: low-energy spectrum of L=1152 #1 (ee=1, dE=1, Dmax=32) and its decode; one rank (the search is not sharded
    # here), the others wait at the next barrier
    if rank == 0:
        import time as _t
        from conftest import droplet_couplings, golden
        sp = tnac4o_b200.tnac4o(mode='Ising', Nx=12, Ny=12, Nc=8, J=droplet_couplings(1152), beta=CFG['beta'], device=dev)
        torch.cuda.synchronize(dev)
        t0 = _t.perf_counter()
        sp.search_low_energy_spectrum(excitations_encoding=1, M=1024, relative_P_cutoff=CFG['relative_P_cutoff'], Dmax=CFG['Dmax'],
                                      max_dEng=1.0)
        torch.cuda.synchronize(dev)
        t_search = _t.perf_counter() - t0
        pairs, nodes = sp.stats.get('droplet_pairs'), sp.stats.get('droplet_nodes')
        sp.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)          # warm-up of the decode buffers
        sp2 = tnac4o_b200.tnac4o(mode='Ising', Nx=12, Ny=12, Nc=8, J=droplet_couplings(1152), beta=CFG['beta'], device=dev)
        sp2.excitations_encoding, sp2.d, sp2.invd, sp2.el, sp2.free_d = 1, sp.d, sp.invd, sp.el, sp.free_d
        sp2.energy, sp2.states = sp.energy[:1].copy(), sp.states[:1].copy()      # sorted ascending: entry 0 is the ground state
        t0 = _t.perf_counter()
        sp2.decode_low_energy_states(max_dEng=1.0, max_states=2 ** 20)
        t_decode = _t.perf_counter() - t0
        z = golden('ref_l1152.npz')
        out['config3_spectrum'] = {'workload': 'e03 + e04: low-energy spectrum L=1152 #1 (ee=1, dE=1, M=2^10, Dmax=32) and decode, one GPU',
                                   'seconds_search': t_search, 'seconds_decode': t_decode, 'decoded_states': int(len(sp2.energy)),
                                   'decoded_states_per_s': len(sp2.energy) / t_decode, 'droplet_pairs': pairs, 'tree_nodes_created': nodes,
                                   'droplet_shapes': len(sp.d),
                                   'reference_seconds_search_build_container': float(z['seconds_search']),
                                   'reference_seconds_decode_build_container': float(z['seconds_decode']),
                                   'reference_decoded_states': int(z['n_states'])}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the sm_100a path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import tnac4o_b200
    from tnac4o_b200 import ops, mps

    from tnac4o_b200 import parallel
    B = args.batch
    # One host thread + one CUDA stream per concurrent instance.  Waiting threads sleep (cudaDeviceScheduleBlockingSync)
    # instead of spinning, so the number of instances per GPU is not tied to the number of host cores: a thread needs
    # ~0.3 s of CPU per instance for its ~10^5 launches.  Cap: 6 threads per core.
    cores = len(os.sched_getaffinity(0))
    from tnac4o_b200._native import lib as _lib, check as _check
    WAIT_MODE = int(os.environ.get('TN_WAIT_MODE', '1'))      # 1 = sleeping waits, 2 = yield-spinning waits
    _check(_lib.tn_set_blocking_sync(WAIT_MODE))
    _check(_lib.tn_set_throughput_mode(0 if os.environ.get('TN_THROUGHPUT') == '0' else 1))
    if world * B > 6 * cores:
        B = max(2, 6 * cores // world)
    J = instance_couplings(rank * B)
    t_prep = time.time()
    inss = [tnac4o_b200.tnac4o(mode='Ising', Nx=CFG['Nx'], Ny=CFG['Ny'], Nc=CFG['Nc'], J=instance_couplings(rank * B + i),
                               beta=CFG['beta'], device=dev) for i in range(B)]
    t_prep = (time.time() - t_prep) / B
    ins = inss[0]
    for x in inss:
        x._site_tables()                                  # inputs resident in HBM before the timed region
    torch.cuda.synchronize(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def solve(x):
        return lambda: x.search_ground_state(M=CFG['M'], relative_P_cutoff=CFG['relative_P_cutoff'], Dmax=CFG['Dmax'])

    def step(e2e, which=None):
        which = inss if which is None else which
        if e2e:
            for x in which:
                x.drop_device_tables()                    # pinned host tables -> HBM inside the timed region
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if len(which) == 1:
            solve(which[0])()
        else:
            parallel.run_concurrently([solve(x) for x in which], device=dev)
        torch.cuda.synchronize(dev)
        e1.record(); e1.synchronize()
        st = {k: float(np.sum([x.stats[k] for x in which])) for k in ('marginals', 'seconds_rhoT', 'seconds_search')}
        return e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0, st

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step(False)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = ops.launch_count(dev)
    cpu0 = time.process_time()
    t_begin = time.perf_counter()
    dev_s, stats = [], []
    for _ in range(args.steps):
        d, _, st = step(False)
        dev_s.append(d); stats.append(st)
    barrier()
    total = time.perf_counter() - t_begin
    cpu_per_instance = (time.process_time() - cpu0) / (args.steps * B)      # host CPU seconds this rank spent per instance
    launches = ops.launch_count(dev) - launches0
    clk = clocks.stop()
    lite = bool(os.environ.get('TN_BENCH_LITE'))          # profiler runs: timed region only
    if lite:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': total / (args.steps * B * world), 'unit': UNIT, 'n_gpus': world,
                              'steps': args.steps, 'warmup': args.warmup, 'lite': True, 'gpu_launches': int(launches),
                              'host_cpu_seconds_per_instance': cpu_per_instance}))
        if world > 1:
            dist.destroy_process_group()
        return
    # end-to-end through the public API: host tables uploaded and results read back inside the timed region
    step(True)                                            # one untimed pass of the e2e path (allocator warm-up)
    barrier()
    e2e_steps = max(1, min(args.steps, 5))                # the e2e path is the same work plus the copies: a few steps pin it
    t_begin = time.perf_counter()
    for _ in range(e2e_steps):
        step(True)
    barrier()
    total_e2e = time.perf_counter() - t_begin
    h2d = B * ins._host_tables().nbytes
    d2h = B * (ins.energy.nbytes + ins.states.nbytes + ins.probability.nbytes + 3 * 8)
    # single-instance latency (one stream), two extra steps outside the timed region; the host spins while it waits here
    # (a sleeping thread adds ~0.4 ms per read-back, 0.7 s per instance, which only matters when nothing else runs)
    _check(_lib.tn_set_blocking_sync(0))
    _check(_lib.tn_set_throughput_mode(0))
    lat_runs = [step(False, [ins]) for _ in range(2)]
    lat = min(r[0] for r in lat_runs)
    lat_stats = lat_runs[-1][2]

    def roofline_pass():
        """one extra single-instance step on the NATIVE path (the path the timed region runs) with the library's own
        per-primitive timers on: CUDA events on the launching stream around every primitive call of csrc/mps_native.cu
        and csrc/search_native.cu, each tagged with its algorithmic flops / bytes"""
        from tnac4o_b200._native import Context
        peak = fp64_peak(torch, dev)
        hbm_peak = measured_peaks().get('hbm_gbs', 6533.2)
        ctx = Context.get(dev)
        ctx.profile(True)
        t_pass = step(False, [ins])[0]
        prof = ctx.profile_read()
        ctx.profile(False)
        per = {}
        for name, r in prof.items():
            sec = r['seconds']
            per[name] = {'seconds': round(sec, 5), 'calls': int(r['calls']), 'gflop': round(r['flops'] / 1e9, 3),
                         'tflops': (r['flops'] / sec / 1e12) if sec > 0 and r['flops'] else None,
                         'frac_fp64_tensor_peak': (r['flops'] / sec / 1e12 / peak) if sec > 0 and r['flops'] else None,
                         'gbytes': round(r['bytes'] / 1e9, 3),
                         'frac_hbm_peak': (r['bytes'] / sec / 1e9 / hbm_peak) if sec > 0 and r['bytes'] else None}
        contraction = ('gemm', 'qr', 'svd', 'right_env')
        flops = sum(prof[k]['flops'] for k in contraction)
        secs = sum(prof[k]['seconds'] for k in contraction)
        allsec = sum(r['seconds'] for r in prof.values())
        dominant = max(prof, key=lambda k: prof[k]['seconds'])
        return {'bound': 'tensor', 'kernel': 'boundary-MPS contraction kernels of one instance on the native path: DMMA GEMM (gemm_kernel), '
                                             'cluster Householder QR (qr_panel_reg_kernel + DMMA updates), cluster Jacobi SVD, right environments',
                'achieved': flops / secs / 1e12 if secs else None, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': (flops / secs / 1e12 / peak) if secs else None,
                'traffic': NCU_TRAFFIC_LARGE_GEMM['bytes'], 'traffic_note': NCU_TRAFFIC_LARGE_GEMM['note'],
                'algorithmic_gflop_per_instance': round(flops / 1e9, 2), 'device_seconds_contraction': secs,
                'device_seconds_all_primitives': allsec, 'dominant_primitive_by_device_time': dominant,
                'per_primitive': per, 'instrumented_instance_seconds': t_pass,
                'hbm_peak_gbs': hbm_peak,
                'note': 'achieved / frac = algorithmic flops of the reference schedule that were executed (QR 4mn^2 - 4n^3/3 incl. forming Q, '
                        'SVD 22k^3 with vectors / 8k^3/3 without, GEMM 2MNK, right environments) x instances x steps / wall time of the timed '
                        'region (all concurrent instances, all GPUs); single_instance and per_primitive = the same flops / device time of '
                        'the primitives of ONE instance alone on one stream (library timers); QR panels and Jacobi rounds run on the non-tensor FP64 pipe (measured ~15 FMA/clk/SM, '
                        'profiles/r2a_qr_panel_phase_cycles_and_svd_timings.txt) and are latency-bound, the GEMMs run on DMMA',
                'peak_source': 'measured here: torch.matmul f64 8192^3 (cuBLAS DGEMM), best of 5 -- MEASURED_PEAKS.json has no FP64 entry',
                'measured_in': 'one extra single-instance step on the native path right after the timed region, library timers on '
                               '(tn_profile: CUDA events on the launching stream around every primitive call)'}

    roofline = roofline_pass() if rank == 0 else None      # before the one-off heavy runs below: same thermal state as the timed region
    # (the one-off runs below are single-stream as well: the host keeps spinning)
    # ---- BASELINE configs 4 and 5 as quoted (outside the timed region): ONE search at M = 2^12 with its branch batch
    # sharded over all ranks (NCCL all-gather of the candidate log-probabilities per site), and
    # 10^5 Gibbs samples at beta = 1 sharded over the ranks (no collective until the final gather)
    extra = None if args.no_extra else extra_configs(torch, dist, tnac4o_b200, parallel, dev, rank, world)

    if world > 1:
        t = torch.tensor([total, total_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, total_e2e = t.tolist()
        cnt = torch.tensor([launches, sum(s['marginals'] for s in stats), sum(s['seconds_search'] for s in stats) / B],
                           dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches_all, marg_all, search_s_all = cnt.tolist()
    else:
        launches_all, marg_all = launches, sum(s['marginals'] for s in stats)
        search_s_all = sum(s['seconds_search'] for s in stats) / B
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    instances = args.steps * world * B
    frac_search = (sum(s['seconds_search'] for s in stats) /
                   max(1e-30, sum(s['seconds_rhoT'] + s['seconds_search'] for s in stats)))
    value = total / instances
    # ---- parity guard on the timed workload: golden energy of instance 001 (groundstates_otn2d.txt:1)
    from conftest import droplet_golden
    e_file, _ = droplet_golden(CFG['L'], 1)
    parity_ok = bool(abs(ins.energy[0] - e_file) < 1e-5)
    # ... and on every synthetic instance of the batch: the search's own energy against the independent CSR energy kernel
    # on the returned configuration (auxx.energy_Jij), i.e. energy bookkeeping and state decoding agree on all of them
    synth_ok = True
    for i, x in enumerate(inss):
        Ji = instance_couplings(rank * B + i)
        synth_ok = synth_ok and bool(abs(tnac4o_b200.energy_Jij(Ji, x.binary_states()[:1])[0] - x.energy[0]) < 1e-6)
    if roofline is not None and roofline.get('algorithmic_gflop_per_instance'):
        whole = roofline['algorithmic_gflop_per_instance'] * 1e9 * instances / total / 1e12
        # top level = the TIMED REGION: algorithmic flops of all instances of all steps / wall time of the timed region,
        # against the FP64 tensor peak of the GPUs used; the single-instance figures of the instrumented pass stay beside it
        roofline['single_instance'] = {'achieved': roofline['achieved'], 'frac': roofline['frac'],
                                       'device_seconds_contraction': roofline['device_seconds_contraction']}
        roofline['achieved'] = whole
        roofline['frac'] = whole / (roofline['peak'] * world)
        roofline['peak_all_gpus'] = roofline['peak'] * world
        roofline['whole_step_tflops'] = whole
        roofline['whole_step_frac_fp64_tensor_peak'] = whole / (roofline['peak'] * world)

    # ---- CPU baseline: the numpy port of the reference on a bounded sample
    cpu = cpu_sample(J)
    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f64', 'data': 'droplet instance 001 (rank 0) + synthetic couplings on the same chimera pattern (other ranks)',
           'config': {'workload': 'e01 ground-state search L=2048 (16x16x8 chimera), M=2^10, Dmax=32, beta=3, P_cutoff=1e-8, no preconditioning',
                      'batch_per_gpu': B, 'batch_requested': args.batch, 'host_cores': cores,
                      'concurrency': 'one host thread + one CUDA stream per instance, blocking host waits',
                      'l2': 'flushed between steps (256 MiB write)', 'parity_energy_matches_golden': parity_ok,
                      'parity_synthetic_instances_self_consistent': synth_ok},
           'latency_seconds_single_instance': lat,
           'seconds_rhoT_per_instance_under_concurrency': float(np.mean([s['seconds_rhoT'] for s in stats])) / B,
           'seconds_search_per_instance_under_concurrency': float(np.mean([s['seconds_search'] for s in stats])) / B,
           'branch_marginals_per_s': marg_all / (total * frac_search) if frac_search else None,
           'branch_marginals_per_s_single_stream': lat_stats['marginals'] / lat_stats['seconds_search'],
           'device_seconds_per_step_rank0': float(np.mean(dev_s)),
           'host_model_prep_seconds': t_prep, 'host_cpu_seconds_per_instance_rank0': cpu_per_instance,
           'e2e': {'value': total_e2e / (e2e_steps * world * B), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                   'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps},
           'gpu_launches': int(launches_all), 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu}
    if extra is not None:
        out['other_configs'] = extra
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- CPU arm
def _sample_once(J, cols=None):
    """Bounded sample of the workload for the numpy port: a 4-row slab of the instance (same Nx, M, Dmax), which
    contains every kind of row of the 16-row lattice.  Returns timings that extrapolate to the whole instance."""
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import RefSolver
    rows = 4
    Nx_full, Nc = CFG['Nx'], CFG['Nc']
    Nx = cols or Nx_full
    # spins of the slab: lattice rows 0..rows-1, columns 0..Nx-1, re-indexed to the narrower lattice
    def remap(i):
        cell, m = divmod(i, Nc)
        ny, nx = divmod(cell, Nx_full)
        return None if (ny >= rows or nx >= Nx) else (ny * Nx + nx) * Nc + m
    Jsub = [[remap(i), remap(j), v] for i, j, v in J if remap(i) is not None and remap(j) is not None]
    ins = RefSolver(mode='Ising', Nx=Nx, Ny=rows, Nc=Nc, J=Jsub, beta=CFG['beta'])
    t_rows = []
    from oracle.mps_ref import RefMPS
    ins.rhoT = [None] * (rows + 1)
    ins.rhoT_overlap, ins.rhoT_discarded = [1] * (rows + 1), [0] * (rows + 1)
    ins.rhoT[-1] = RefMPS(Nx, d=1)
    for ny in range(rows - 1, -1, -1):
        t0 = time.perf_counter()
        W = [ins.traced_mpo(ny, nx) for nx in range(Nx)]
        psi = ins.rhoT[ny + 1].copy()
        psi.apply_mpo(W, conj=True)
        psi.compress(CFG['Dmax'], 1e-16, 1e-10, 20, True)
        ins.rhoT[ny] = psi
        t_rows.append(time.perf_counter() - t0)
    # search: first lattice row only (branch count saturates at M within the first sites)
    ins._setup_rhoT = lambda *a, **k: None
    t0 = time.perf_counter()
    count = search_rows(ins, 1)
    t_search = time.perf_counter() - t0
    f = Nx_full / Nx            # a narrower slab is extrapolated linearly in the number of columns
    return {'t_rows': [t * f for t in t_rows], 't_search_row': t_search * f, 'marginals_row': count * f}


def search_rows(ins, nrows):
    """the oracle's branch-and-bound restricted to the first nrows lattice rows (timing sample)"""
    full, order = ins.Ny, ins.order
    ins.Ny, ins.order = nrows, np.arange(nrows * ins.Nx)
    try:
        ins.search_ground_state(M=CFG['M'], relative_P_cutoff=CFG['relative_P_cutoff'], Dmax=CFG['Dmax'])
    finally:
        ins.Ny, ins.order = full, order
    return ins.marginals_evaluated


def extrapolate(s):
    Ny = CFG['Ny']
    bottom, second, steady, top = s['t_rows']          # lattice rows Ny-1 (bond 16), Ny-2 (256), interior (512), 0 (no up leg)
    rho = bottom + second + steady * (Ny - 3) + top
    search = s['t_search_row'] * Ny
    return rho + search, rho, search


def blas_threads():
    """effective BLAS thread count of this process (asserted to be 1 in the CPU legs)"""
    try:
        from threadpoolctl import threadpool_info
        n = [int(p.get('num_threads', 1)) for p in threadpool_info() if p.get('user_api') in ('blas', 'openmp')]
        return max(n) if n else 1
    except Exception:
        return None


def cpu_sample(J):
    s = _sample_once(J)
    total, rho, search = extrapolate(s)
    return {'value': total, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'blas_threads': blas_threads(),
            'seconds_rhoT_est': rho, 'seconds_search_est': search,
            'branch_marginals_per_s': s['marginals_row'] / s['t_search_row'],
            'sample': 'numpy port of the reference (oracle/), 1 BLAS thread: boundary-MPS build of a 4-row slab of the 16 lattice rows '
                      '(bond 16 -> 256 -> 512; the interior row is the steady-state cost, x13) + branch-and-bound over the first '
                      'lattice row (x16); same instance, Nx, M, Dmax'}


def _worker(rank, q, warm=False):
    # the thread pins at the top of this file were applied when the spawned child re-imported it (before numpy)
    nthr = blas_threads()
    assert nthr in (None, 1), 'BLAS thread pinning failed in the worker: %r threads' % nthr
    if warm:
        # untimed warm-up step: imports, page cache and BLAS initialisation on a small instance (L = 128)
        import warnings
        warnings.filterwarnings('ignore')
        from conftest import droplet_couplings
        from oracle import RefSolver
        RefSolver(mode='Ising', Nx=4, Ny=4, Nc=8, J=droplet_couplings(128), beta=3).search_ground_state(M=64, Dmax=8)
        q.put(None)
        return
    r = _sample_once(instance_couplings(rank), cols=REF_COLS)
    r['blas_threads'] = nthr
    q.put(r)


REF_COLS = None   # full lattice width: the share of edge sites (small bonds) is not linear in the number of columns


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    import psutil
    cores = len(os.sched_getaffinity(0))
    mem_gb = psutil.virtual_memory().available / 2 ** 30
    # one single-thread worker per host core (BLAS threads pinned to 1 at the top of this file, asserted in the worker)
    workers = int(max(1, min(cores, mem_gb // 4, 64)))
    ctx = mp.get_context('spawn')

    def step(warm=False):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, q, warm)) for r in range(workers)]
        t0 = time.perf_counter()
        for p in procs:
            p.start()
        res = [q.get() for _ in procs]
        for p in procs:
            p.join()
        return time.perf_counter() - t0, res

    for _ in range(args.warmup):
        step(warm=True)
    times, results = [], []
    for _ in range(args.steps):
        t, r = step()
        times.append(t); results.extend(r)
    # every worker processed one bounded sample; scale each worker's timings to a whole instance
    per_instance = float(np.mean([extrapolate(r)[0] for r in results]))
    value = per_instance / workers               # instances run concurrently, one per core
    marg = float(np.mean([r['marginals_row'] / r['t_search_row'] for r in results])) * workers
    cpu = {'value': value, 'unit': UNIT, 'cores': workers, 'kind': 'port',
           'blas_threads_per_worker': sorted({r.get('blas_threads') for r in results}, key=str),
           'sample': 'numpy port of the reference (oracle/): %d concurrent single-thread workers (1 BLAS thread each is the '
                     'fastest setting, SURVEY.md section 6), each timing a 4-row slab of the 16 lattice rows of the boundary-MPS build '
                     '(steady-state interior row x13) and the first lattice row of the search (x16); value = extrapolated seconds per '
                     'instance / workers; warm-up steps run a small L=128 instance per worker' % workers}
    out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': False, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f64', 'data': 'droplet instance 001 + synthetic couplings on the same chimera pattern',
           'config': {'workload': 'e01 ground-state search L=2048 (16x16x8 chimera), M=2^10, Dmax=32, beta=3, P_cutoff=1e-8, no preconditioning'},
           'single_thread_seconds_per_instance': per_instance, 'branch_marginals_per_s': marg,
           'cpu_baseline': cpu, 'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=24, help='independent instances solved concurrently per GPU')
    ap.add_argument('--no-extra', action='store_true', help='skip the one-off runs of configs 4 (M=2^12) and 5 (Gibbs)')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)
