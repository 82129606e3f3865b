#!/usr/bin/env python
"""Low-energy spectrum of a droplet instance (examples/e03_search_spectrum_droplet_instances.py of the reference)."""
import os
import time

from _common import SHAPES, droplet_couplings, parser, setup_logging


def file_name(args, results):
    return os.path.join(results, 'L=%1d_ins=%03d_r=%1d_beta=%0.2f_D=%1d_M=%1d_P=%0.2e_ee=%1d_dE=%0.3f_hd=%1d_pre=%1d.npy'
                        % (args.L, args.ins, args.r, args.b, args.D, args.M, args.P, args.ee, args.dE, args.hd, args.pre))


if __name__ == '__main__':
    args = parser(__doc__, spectrum=True).parse_args()
    setup_logging()
    from tnac4o_b200 import drivers
    Nx, Ny = SHAPES[args.L]
    t0 = time.time()
    ins = drivers.search_spectrum(droplet_couplings(args), Nx, Ny, rot=args.r, beta=args.b, D=args.D, M=args.M, relative_P_cutoff=args.P,
                                  excitations_encoding=args.ee, dE=args.dE, hd=args.hd, precondition=args.pre)
    ins.logger.info('Total time : %.2f seconds', time.time() - t0)
    if args.s:      # saved before decoding, as the reference does
        ins.save(file_name(args, drivers.results_dir()))
    t0 = time.time()
    ins.decode_low_energy_states(max_dEng=args.dE, max_states=args.max_st)
    ins.logger.info('Decoding spectrum elapse time : %.2f seconds', time.time() - t0)
    ins.show_solution(state=False)
