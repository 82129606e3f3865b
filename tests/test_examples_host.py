"""The example scripts mirror the command-line flags of the reference's examples/e01 ... e06 (CPU: argument parsing only)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EX = os.path.join(ROOT, 'examples')
SCRIPTS = sorted(f for f in os.listdir(EX) if f.startswith('e0') and f.endswith('.py'))


def test_all_six_examples_present():
    assert [s[:3] for s in SCRIPTS] == ['e01', 'e02', 'e03', 'e04', 'e05', 'e06']


@pytest.mark.parametrize('script', SCRIPTS)
def test_help_runs_without_gpu(script):
    out = subprocess.run([sys.executable, os.path.join(EX, script), '-h'], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert '-D' in out.stdout and '-M' in out.stdout


def test_reference_defaults_and_file_name():
    sys.path.insert(0, EX)
    try:
        import _common
        import e03_search_spectrum_droplet_instances as e03
    finally:
        sys.path.remove(EX)
    a = _common.parser('x', spectrum=True).parse_args([])
    assert (a.L, a.ins, a.r, a.b, a.D, a.M, a.P, a.dE, a.hd, a.max_st, a.ee, a.pre, a.s) == \
        (128, 1, 0, 3, 48, 1024, 1e-8, 1.0, 0, 2 ** 20, 1, True, False)
    assert _common.parser('x', sampling=True).parse_args([]).b == 1
    a = _common.parser('x', spectrum=True).parse_args('-L 1152 -ins 2 -no-pre -ee 3 -s'.split())
    assert (a.L, a.ins, a.pre, a.ee, a.s) == (1152, 2, False, 3, True)
    a = _common.parser('x', spectrum=True).parse_args(['-L', '1152'])
    # the name the reference's e03 gives its saved spectrum
    assert os.path.basename(e03.file_name(a, 'results')) == \
        'L=1152_ins=001_r=0_beta=3.00_D=48_M=1024_P=1.00e-08_ee=1_dE=1.000_hd=0_pre=1.npy'


def test_instance_loading_matches_fixture():
    """examples/_common.droplet_couplings reads an instance file the way the reference's e01:57-65 does; the committed
    fixture holds the same couplings (skipped where the reference's instances/ folder is absent, e.g. on the GPU box)"""
    inst = '/root/reference/instances'
    if not os.path.isdir(inst):
        pytest.skip('no instances/ folder here')
    from conftest import droplet_couplings
    sys.path.insert(0, EX)
    try:
        import _common
    finally:
        sys.path.remove(EX)
    a = _common.parser('x').parse_args(['--instances', inst, '-L', '128', '-ins', '1'])
    assert _common.droplet_couplings(a) == droplet_couplings(128, 1)
