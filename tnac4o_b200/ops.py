"""Tensor-level wrappers over the C ABI: every function takes CUDA float64 torch tensors (row-major) and
launches hand-written sm_100a kernels on torch's current stream.  No torch math is used on the data path.
"""
import ctypes

import torch

from . import _native as nat
from ._native import Context, check, lib, ptr

F64 = torch.float64


def _ctx(t):
    return Context.get(t.device)


def empty(shape, like=None, dtype=F64, device=None):
    return torch.empty(shape, dtype=dtype, device=device if device is not None else like.device)


def gemm(A, B, transA=False, transB=False, out=None, alpha=1.0, beta=0.0):
    """out = alpha * op(A) @ op(B) + beta * out for 2-d tensors whose rows are contiguous (tn_gemm)."""
    assert A.dim() == 2 and B.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1
    M, K = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if transB else (B.shape[0], B.shape[1])
    assert K == K2, (A.shape, B.shape, transA, transB)
    if out is None:
        out = torch.empty((M, N), dtype=F64, device=A.device)
        beta = 0.0
    assert out.shape == (M, N) and out.stride(1) == 1
    c = _ctx(A)
    check(lib.tn_gemm(c.handle, c.stream, int(transA), int(transB), M, N, K, alpha, ptr(A), max(A.stride(0), 1), 0,
                      ptr(B), max(B.stride(0), 1), 0, beta, ptr(out), max(out.stride(0), 1), 0, 1))
    return out


def transpose(A):
    assert A.dim() == 2 and A.stride(1) == 1
    m, n = A.shape
    out = torch.empty((n, m), dtype=F64, device=A.device)
    c = _ctx(A)
    check(lib.tn_transpose(c.handle, c.stream, m, n, ptr(A), max(A.stride(0), 1), ptr(out), m))
    return out


def qr_pos(A):
    """A (m x n, contiguous, DESTROYED) -> Q (m x k), R (k x n) with diag(R) >= 0, and the bit pattern of max|R|."""
    assert A.dim() == 2 and A.is_contiguous()
    m, n = A.shape
    k = min(m, n)
    Q = torch.empty((m, k), dtype=F64, device=A.device)
    R = torch.empty((k, n), dtype=F64, device=A.device)
    bits = torch.empty(1, dtype=torch.int64, device=A.device)
    c = _ctx(A)
    check(lib.tn_qr_pos(c.handle, c.stream, m, n, ptr(A), n, ptr(Q), k, ptr(R), n, ptr(bits)))
    return Q, R, bits


def maxabs_bits(x):
    bits = torch.empty(1, dtype=torch.int64, device=x.device)
    c = _ctx(x)
    check(lib.tn_maxabs(c.handle, c.stream, ptr(x), x.numel(), ptr(bits)))
    return bits


def pow2_scale_(x, bits, log2_accum=None):
    """x /= 2^floor(log2 max|x|) in place (mps.nfactor); a 1-element x becomes exactly 1."""
    assert x.is_contiguous()
    c = _ctx(x)
    check(lib.tn_pow2_scale(c.handle, c.stream, ptr(x), x.numel(), ptr(bits), ptr(log2_accum)))
    return x


last_svd_sweeps = 0


def svd(C, want_vectors=True):
    """thin SVD of a contiguous 2-d tensor by Jacobi rotations -> (U, S, Vt) or S"""
    global last_svd_sweeps
    assert C.dim() == 2 and C.stride(1) == 1
    m, n = C.shape
    k = min(m, n)
    S = torch.empty(k, dtype=F64, device=C.device)
    U = torch.empty((m, k), dtype=F64, device=C.device) if want_vectors else None
    Vt = torch.empty((k, n), dtype=F64, device=C.device) if want_vectors else None
    sweeps = ctypes.c_int(0)
    c = _ctx(C)
    check(lib.tn_svd(c.handle, c.stream, m, n, ptr(C), max(C.stride(0), 1), ptr(U), k, ptr(S), ptr(Vt), n,
                     int(want_vectors), ctypes.byref(sweeps)))
    last_svd_sweeps = sweeps.value
    return (U, S, Vt) if want_vectors else S


def truncation_rank(S, tol, Dmax):
    keep, lost = ctypes.c_int(0), ctypes.c_double(0.0)
    c = _ctx(S)
    dmax = int(min(Dmax, 2 ** 30))
    check(lib.tn_truncation_rank(c.handle, c.stream, ptr(S), S.numel(), float(tol), dmax, ctypes.byref(keep),
                                 ctypes.byref(lost)))
    return keep.value, lost.value


def mpo_apply(A, W, conj=True):
    """A (Dl, dp, Dr), W (wl, d_out, wr, d_in) -> MPO applied on the physical leg (mps.py:753-763)."""
    A = A.contiguous()
    W = W.contiguous()
    Dl, dp, Dr = A.shape
    if conj:
        wl, dp2, wr, du = W.shape
    else:
        wl, du, wr, dp2 = W.shape
    assert dp == dp2, (A.shape, W.shape, conj)
    out = torch.empty((Dl * wl, du, Dr * wr), dtype=F64, device=A.device)
    c = _ctx(A)
    check(lib.tn_mpo_apply(c.handle, c.stream, int(conj), Dl, dp, Dr, wl, wr, du, ptr(A), ptr(W), ptr(out)))
    return out


def diff_norm(a, b):
    out = torch.empty(1, dtype=F64, device=a.device)
    c = _ctx(a)
    check(lib.tn_diff_norm(c.handle, c.stream, ptr(a), ptr(b), a.numel(), ptr(out)))
    return out


def sort_capacity(n):
    return int(lib.tn_sort_capacity_for(int(n)))


def launch_count(device=None):
    """kernels launched so far on `device` by all host threads of this process"""
    return Context.total_launches(device)
