#!/usr/bin/env python
"""Load a spectrum saved by e03 -s (of this package or of the reference) and decode it into states
(examples/e04_load_spectrum_droplet_instances.py of the reference)."""
import time

from _common import droplet_couplings, parser, setup_logging
from e03_search_spectrum_droplet_instances import file_name

if __name__ == '__main__':
    args = parser(__doc__, spectrum=True).parse_args()
    setup_logging()
    from tnac4o_b200 import drivers
    t0 = time.time()
    try:
        ins, error = drivers.load_spectrum(file_name(args, drivers.results_dir()), J=droplet_couplings(args), dE=args.dE,
                                           max_states=args.max_st)
    except FileNotFoundError:
        raise SystemExit('First run e03_search_spectrum_droplet_instances.py with option -s')
    ins.logger.info('Decoding spectrum elapse time : %.2f seconds', time.time() - t0)
    ins.show_solution()
    print('Consistency of different ways to calculate energies.')
    print('For ee = 2 or 3 expected difference is ~1e-6 due to applied noise.')
    print('Difference = ', error)
