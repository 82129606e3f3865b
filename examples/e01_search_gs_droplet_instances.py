#!/usr/bin/env python
"""Ground-state search of a droplet instance (examples/e01_search_gs_droplet_instances.py of the reference)."""
import time

from _common import SHAPES, droplet_couplings, parser, setup_logging

if __name__ == '__main__':
    args = parser(__doc__).parse_args()
    setup_logging()
    from tnac4o_b200 import drivers
    Nx, Ny = SHAPES[args.L]
    t0 = time.time()
    ins = drivers.search_gs(droplet_couplings(args), Nx, Ny, rot=args.r, beta=args.b, D=args.D, M=args.M, relative_P_cutoff=args.P,
                            precondition=args.pre)
    ins.logger.info('Total time : %.2f seconds', time.time() - t0)
    ins.show_solution(state=False)
    print('Solution [1 -> spin up: si=+1; 0 -> spin down: si=-1]:')
    print(ins.binary_states())
