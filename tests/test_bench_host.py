"""CPU tests of bench.py's contract: the reference arm (numpy port on the host cores) prints one JSON line with the keys
the driver reads, only rank 0 works under torchrun, and the GPU arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

BENCH = os.path.join(ROOT, 'bench.py')


def _run(args, env=None, timeout=900):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_reference_arm_line():
    out = _run(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0'])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 's/instance' and d['higher_is_better'] is False
    assert d['steps'] == 1 and d['warmup'] == 0 and d['n_gpus'] == 1 and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cpu = d['cpu_baseline']
    assert cpu['kind'] == 'port' and cpu['value'] == d['value'] and cpu['cores'] >= 1 and 'sample' in cpu
    assert cpu['blas_threads_per_worker'] == [1]                # the pinning that round 1 got wrong
    assert 'L=2048' in d['config']['workload'] and 'M=2^10' in d['config']['workload']


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert out.returncode == 0 and out.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without a GPU')
def test_gpu_arm_has_no_cpu_fallback():
    out = _run(['--steps', '1', '--warmup', '0'], timeout=300)
    assert out.returncode != 0 and 'no CPU fallback' in out.stderr
    assert not any(l.startswith('{') for l in out.stdout.splitlines())
