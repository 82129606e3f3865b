#!/usr/bin/env python
"""Ground-state energy and degeneracy of a J124 instance from four rotations
(examples/e06_search_gs_degeneracy_J124.py of the reference)."""
import argparse
import os
import time

from _common import ROOT, setup_logging

if __name__ == '__main__':
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument('--instances', default=os.environ.get('TNAC4O_INSTANCES', os.path.join(ROOT, 'instances')))
    p.add_argument('-C', type=int, choices=[8, 12, 16], default=8)
    p.add_argument('-ins', type=int, default=1)
    p.add_argument('-b', type=float, default=0.75)
    p.add_argument('-D', type=int, default=48)
    p.add_argument('-M', type=int, default=2 ** 12)
    p.add_argument('-P', type=float, default=1e-8)
    p.add_argument('-s', dest='s', action='store_true')
    p.add_argument('-no-pre', dest='pre', action='store_false')
    p.set_defaults(pre=True, s=False)
    args = p.parse_args()
    setup_logging()
    import tnac4o_b200 as tnac4o
    from tnac4o_b200 import drivers
    J = tnac4o.Jij_f2p(tnac4o.load_Jij(os.path.join(args.instances, 'Chimera_J124', 'C=%d_J124' % args.C, '%03d.txt' % args.ins)))
    t0 = time.time()
    E, deg, per = drivers.search_gs_degeneracy(J, args.C, args.C, Nc=8, beta=args.b, D=args.D, M=args.M, relative_P_cutoff=args.P,
                                               precondition=args.pre)
    for rot, e, d in per:
        print('Rotation %1d: energy %1d, degeneracy %1d' % (rot, e, d))
    print('Best found energy and its degeneracy for J124 instances on chimera graph C%1d, instance %1d (%.1f s)' % (args.C, args.ins,
                                                                                                             time.time() - t0))
    print('Energy = %1d' % E)
    print('Degeneracy = %1d' % deg)
    if args.s:
        drivers.write_gs_degeneracy_txt(os.path.join(drivers.results_dir(), 'J124_C=%1d_ins=%03d_beta=%0.2f_D=%1d_M=%1d_pre=%1d.txt'
                                                     % (args.C, args.ins, args.b, args.D, args.M, args.pre)), E, deg)
