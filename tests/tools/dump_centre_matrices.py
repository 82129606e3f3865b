"""Dump the 512 x 512 centre matrices the oracle hands to its SVD while compressing a 4-row slab of L=2048 instance 001
(build container only; output under /tmp/svdsim).  usage: python tests/tools/dump_centre_matrices.py <beta>"""
import os, sys, numpy as np, warnings
warnings.filterwarnings('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['OPENBLAS_NUM_THREADS'] = '1'
import bench
import oracle.mps_ref as mr
beta = float(sys.argv[1])
bench.CFG['beta'] = beta
dump = []
orig = mr.ref_svd
def spy(T):
    if T.shape == (512, 512) and len(dump) < 40:
        dump.append(T.copy())
    return orig(T)
mr.ref_svd = spy
bench._sample_once(bench.instance_couplings(0))
os.makedirs('/tmp/svdsim', exist_ok=True)
np.savez_compressed('/tmp/svdsim/C_beta%g.npz' % beta, *dump)
print('dumped', len(dump))
