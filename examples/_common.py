"""Shared pieces of the example scripts: the reference's command-line flags (examples/e01 ... e06 of marekrams/tnac4o) and
the instance files.  The instance directory is the reference's `instances/` folder; point to it with --instances or the
environment variable TNAC4O_INSTANCES (the files are not part of this repository)."""
import argparse
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPES = {128: (4, 4), 512: (8, 8), 1152: (12, 12), 2048: (16, 16)}


def parser(description, sampling=False, spectrum=False):
    p = argparse.ArgumentParser(description=description)
    p.add_argument('--instances', default=os.environ.get('TNAC4O_INSTANCES', os.path.join(ROOT, 'instances')),
                   help='the instances/ directory of the reference repository')
    p.add_argument('-L', type=int, choices=[128, 512, 1152, 2048], default=128, help='size of the chimera graph')
    p.add_argument('-ins', type=int, default=1, help='instance number (1-100)')
    p.add_argument('-r', type=int, default=0, help='rotate the graph by 90 degrees r times')
    p.add_argument('-b', type=float, default=1 if sampling else 3, help='inverse temperature')
    p.add_argument('-D', type=int, default=48, help='maximal bond dimension of the boundary MPS')
    p.add_argument('-M', type=int, default=2 ** 10, help='partial states kept by the branch and bound (or number of samples)')
    if not sampling:
        p.add_argument('-P', type=float, default=1e-8, help='cut-off on the range of relative probabilities')
    if spectrum:
        p.add_argument('-dE', type=float, default=1.0, help='limit on the excitation energy')
        p.add_argument('-hd', type=int, default=0, help='lower limit of the Hamming distance between states while merging')
        p.add_argument('-max_st', type=int, default=2 ** 20, help='limit on the number of reconstructed states')
        p.add_argument('-ee', type=int, default=1, choices=[1, 2, 3], help='strategy used to compress droplets')
    p.add_argument('-no-pre', dest='pre', action='store_false', help='do not use preconditioning')
    p.add_argument('-s', dest='s', action='store_true', help='save results to ./results/')
    p.set_defaults(pre=True, s=False)
    return p


def droplet_couplings(args):
    """load -> 0-based indices -> round to multiples of 1/75, as examples/e01:57-65"""
    import tnac4o_b200 as tnac4o
    fn = os.path.join(args.instances, 'Chimera_droplet_instances', 'chimera%d_spinglass_power' % args.L, '%03d.txt' % args.ins)
    return tnac4o.round_Jij(tnac4o.Jij_f2p(tnac4o.load_Jij(fn)), 1 / 75)


def setup_logging():
    logging.basicConfig(level='INFO')
