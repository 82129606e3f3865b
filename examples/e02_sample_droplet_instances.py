#!/usr/bin/env python
"""Gibbs sampling of a droplet instance (examples/e02_sample_droplet_instances.py of the reference)."""
import os
import time

from _common import SHAPES, droplet_couplings, parser, setup_logging

if __name__ == '__main__':
    args = parser(__doc__, sampling=True).parse_args()
    setup_logging()
    from tnac4o_b200 import drivers
    Nx, Ny = SHAPES[args.L]
    t0 = time.time()
    ins = drivers.gibbs_sampling(droplet_couplings(args), Nx, Ny, rot=args.r, beta=args.b, D=args.D, M=args.M, precondition=args.pre)
    ins.logger.info('Total time : %.2f seconds', time.time() - t0)
    ins.show_solution(state=False)
    if args.s:
        fn = os.path.join(drivers.results_dir(), 'gibbs_L=%1d_ins=%03d_r=%1d_beta=%0.2f_D=%1d_M=%1d_pre=%1d.txt'
                          % (args.L, args.ins, args.r, args.b, args.D, args.M, args.pre))
        drivers.write_states_txt(ins, fn)
