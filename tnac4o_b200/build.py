"""In-tree build of the CUDA library (sm_100a only): python tnac4o_b200/build.py

nvcc cross-compiles without a GPU; the resulting tnac4o_b200/lib/libtnac4o_b200.so travels to the GPU box with the
repository snapshot.  Objects are rebuilt only when their source (or a header) is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'lib')
LIB = os.path.join(OUT, 'libtnac4o_b200.so')
SOURCES = ['api.cu', 'gemm.cu', 'gemm_tma.cu', 'qr.cu', 'svd.cu', 'mps_ops.cu', 'mps_native.cu', 'sort.cu', 'search.cu', 'search_native.cu', 'droplet.cu', 'droplet_book.cu', 'decode.cu']
HEADERS = [os.path.join(CSRC, 'common.cuh'), os.path.join(HERE, '..', 'include', 'tnac4o_b200.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OUT, 'obj', src.replace('.cu', '.o'))
    path = os.path.join(CSRC, src)
    if _stale(obj, [path] + HEADERS):
        subprocess.run([NVCC] + FLAGS + ['-c', path, '-o', obj], check=True)
    return obj


def build(verbose=False):
    os.makedirs(os.path.join(OUT, 'obj'), exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(_compile, SOURCES))
    if _stale(LIB, objs):
        subprocess.run([NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'], check=True)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(verbose=True)
    sys.exit(0)
