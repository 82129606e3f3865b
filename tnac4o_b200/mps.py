"""Boundary-MPS machinery on the GPU, behind the reference's own interface.

Mirrors the part of /root/reference/tnac4o/mps.py that the solver uses: class and
method names, argument meaning and the fixed truncation schedule of ``compress_mps`` are the reference's; every
tensor lives in HBM as a contiguous float64 torch tensor and every arithmetic step is a kernel of
libtnac4o_b200.so (DMMA GEMMs, cluster Householder QR, Jacobi SVD).  The only host synchronisations are the
rank decisions of ``truncateC`` and the once-per-sweep convergence test of ``variational_compress``.
"""
import numpy as np
import torch

from . import ops

F64 = torch.float64


def svd(T):
    """thin SVD with the reference's sign rule (mps.py:24-40); returns U, S, V (V = rows of right vectors)"""
    return ops.svd(T.contiguous(), want_vectors=True)


def qr(T):
    """economic QR with non-negative diagonal of R (mps.py:43-59)"""
    Q, R, _ = ops.qr_pos(T.contiguous().clone())
    return Q, R


def svd_S(T):
    """singular values only (mps.py:62-73)"""
    return ops.svd(T.contiguous(), want_vectors=False)


def nfactor(T):
    """largest |entry| floored to a power of two (mps.py:76-85); returns a host float (synchronises)"""
    bits = ops.maxabs_bits(T.contiguous())
    return 2.0 ** (int(bits.item() >> 52) - 1023)


class MPS:
    """Matrix product state with the centre-matrix bookkeeping of the reference (mps.py:96-173)."""

    def __init__(self, d=2, L=2, Dmax=2, initial='X', canonise='left', device=None):
        if initial != 'X':
            raise NotImplementedError("only the product state initial='X' is on the contraction path (SURVEY.md section 2 row 9)")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.L = L
        if isinstance(d, int):
            d = [d]
        d = (list(d) * ((L + len(d) - 1) // len(d)))[:L]
        self.d = d
        self.zero = np.finfo(float).eps
        self.dtype = 'float64'
        self.D = self._Dset(Dmax, d)
        self.A = []
        for n in range(L):
            A = torch.zeros((self.D[n], d[n], self.D[n + 1]), dtype=F64, device=self.device)
            A[0, :, 0] = 1.0 / np.sqrt(d[n])
            self.A.append(A)
        self.C = torch.ones((1, 1), dtype=F64, device=self.device)
        self.pC = L
        self._log2_normC = torch.zeros(1, dtype=F64, device=self.device)
        self.reset_R()
        self.reset_S()
        self.discarded = [0] * (L + 1)
        if canonise == 'left':
            self.canonise_left()
        elif canonise == 'right':
            self.canonise_right()
        self._log2_normC.zero_()

    # -- bookkeeping -----------------------------------------------------------------------
    @property
    def normC(self):
        return float(2.0 ** self._log2_normC.item())

    @staticmethod
    def _Dset(Dmax, d):
        """mps.py:644-653"""
        L = len(d)
        D = [1] * (L + 1)
        for n in range(L):
            D[n + 1] = min(D[n] * d[n], Dmax)
        D[-1] = 1
        for n in range(L - 1, -1, -1):
            D[n] = min(D[n + 1] * d[n], Dmax, D[n])
        return D

    def _ones(self, *shape):
        return torch.ones(shape, dtype=F64, device=self.device)

    def _one_S(self, D):
        S = torch.zeros(D, dtype=F64, device=self.device)
        S[0] = 1.0
        return S

    def reset_R(self):
        self.R = [self._ones(1, 1) for _ in range(self.L + 2)]
        self.R[-1] = None

    def reset_S(self):
        self.S = [self._one_S(1) for _ in range(self.L + 1)]

    def copy(self):
        """tensors are duplicated, environments shared (mps.py:159-173)"""
        phi = MPS.__new__(MPS)
        phi.device, phi.L, phi.d, phi.zero, phi.dtype = self.device, self.L, self.d[:], self.zero, self.dtype
        phi.A = [a.clone() for a in self.A]
        phi.C = self.C.clone()
        phi.pC = self.pC
        phi._log2_normC = self._log2_normC.clone()
        phi.D = self.D[:]
        phi.R = self.R[:]
        phi.S = [self._one_S(1) for _ in range(self.L + 1)]
        phi.discarded = [0] * (self.L + 1)
        return phi

    # -- MPO application -------------------------------------------------------------------
    def apply_mpo(self, M, Hconj=False):
        """psi = H psi, or H^dagger psi (mps.py:353-359, 753-763)"""
        for n in range(self.L):
            if M.support[n]:
                self.A[n] = ops.mpo_apply(self.A[n], M.W[n], conj=Hconj)
                self.D[n], self.d[n], self.D[n + 1] = self.A[n].shape

    def apply_diagonalO(self, diagO, n):
        """scale the physical leg of site n (mps.py:361-366)"""
        scale = torch.as_tensor(np.asarray(diagO, dtype=np.float64), device=self.device)
        self.A[n] = (self.A[n] * scale[None, :, None]).contiguous()

    # -- moving the centre -----------------------------------------------------------------
    def attach_AC(self):
        """A[pC-1] . C -> A[pC-1] (mps.py:368-373)"""
        n = self.pC - 1
        Dl, d, Dr = self.A[n].shape
        self.A[n] = ops.gemm(self.A[n].view(Dl * d, Dr), self.C).view(Dl, d, self.C.shape[1])

    def attach_CA(self):
        """C . A[pC] -> A[pC] (mps.py:375-380)"""
        n = self.pC
        Dl, d, Dr = self.A[n].shape
        self.A[n] = ops.gemm(self.C, self.A[n].view(Dl, d * Dr)).view(self.C.shape[0], d, Dr)

    def orth_left(self, n):
        """QR of (Dl d, Dr); C <- R / nfactor(R) (mps.py:532-539, 772-785)"""
        Dl, d, Dr = self.A[n].shape
        Q, R, bits = ops.qr_pos(self.A[n].view(Dl * d, Dr))       # A[n] is consumed
        ops.pow2_scale_(R, bits, self._log2_normC)
        self.A[n] = Q.view(Dl, d, R.shape[0])
        self.C = R
        self.D[n + 1] = R.shape[0]
        self.pC = n + 1

    def orth_right(self, n):
        """QR of the transposed (d Dr, Dl) matrix; C <- R^T / nfactor (mps.py:541-548, 787-800)"""
        Dl, d, Dr = self.A[n].shape
        At = ops.transpose(self.A[n].view(Dl, d * Dr))
        Q, R, bits = ops.qr_pos(At)
        ops.pow2_scale_(R, bits, self._log2_normC)
        k = R.shape[0]
        self.A[n] = ops.transpose(Q).view(k, d, Dr)
        self.C = ops.transpose(R)
        self.D[n] = k
        self.pC = n

    def truncateC(self, Dmax, tol=None):
        """SVD truncation of the centre matrix at an interior bond (mps.py:562-585, 802-811)"""
        if not (0 < self.pC < self.L):
            return 0.
        if tol is None:
            tol = self.zero
        U, S, V = ops.svd(self.C, want_vectors=True)
        keep, discarded = ops.truncation_rank(S, max(np.finfo(float).eps, tol), Dmax)
        p = self.pC
        Dl, d, Dr = self.A[p - 1].shape
        self.A[p - 1] = ops.gemm(self.A[p - 1].view(Dl * d, Dr), U[:, :keep]).view(Dl, d, keep)
        Dl, d, Dr = self.A[p].shape
        self.A[p] = ops.gemm(V[:keep], self.A[p].view(Dl, d * Dr)).view(keep, d, Dr)
        self.C = torch.diag(S[:keep]).contiguous()
        self.D[p] = keep
        self.discarded[p] = max(self.discarded[p], discarded)
        return discarded

    def canonise_left(self, compress=False, Dmax=np.inf, tol=None):
        """mps.py:202-218"""
        self.C, self.pC = self._ones(1, 1), 0
        for n in range(self.L):
            self.attach_CA()
            self.orth_left(n)
            if compress:
                self.truncateC(Dmax, tol)
        self.R[-1] = None

    def canonise_right(self, compress=False, Dmax=np.inf, tol=None):
        """mps.py:220-236"""
        self.C, self.pC = self._ones(1, 1), self.L
        for n in range(self.L - 1, -1, -1):
            self.attach_AC()
            self.orth_right(n)
            if compress:
                self.truncateC(Dmax, tol)
        self.R[-1] = None

    # -- mixed environments <self|phi> -----------------------------------------------------
    @staticmethod
    def _mps_RL(RL, A, Ac, T=None):
        """RL' = Ac^T (RL . A) (mps.py:655-658); RL is (D_self, D_phi)"""
        if T is None:
            T = ops.gemm(RL, A.view(A.shape[0], -1))                          # (Dc, d * Dr_phi)
        return ops.gemm(Ac.view(-1, Ac.shape[2]), T.view(-1, A.shape[2]), transA=True)

    @staticmethod
    def _mps_RR(RR, A, Ac):
        """RR' = (A . RR) Ac^T (mps.py:660-663); RR is (D_phi, D_self)"""
        T = ops.gemm(A.view(-1, A.shape[2]), RR)                              # (Dl_phi * d, Dc_r)
        return ops.gemm(T.view(A.shape[0], -1), Ac.view(Ac.shape[0], -1), transB=True)

    def update_RL_mix(self, phi, n, T=None):
        """mps.py:436-444"""
        new = self._mps_RL(self.R[n], phi.A[n], self.A[n], T)
        if n == self.L - 1:
            self.R[self.L + 1] = new.view(-1)[:1]
        else:
            self.R[n + 1] = new

    def update_RR_mix(self, phi, n):
        """mps.py:418-426"""
        new = self._mps_RR(self.R[n + 1], phi.A[n], self.A[n])
        if n == 0:
            self.R[self.L + 1] = new.view(-1)[:1]
        else:
            self.R[n] = new

    def setup_RL_mix(self, phi):
        """mps.py:446-452; returns <self|phi> as a 1-element device tensor"""
        for n in range(self.L):
            self.update_RL_mix(phi, n)
        return self.R[-1]

    def bond_env_mix(self, phi, n):
        """environment of the physical leg of site n in <self|phi> (mps.py:454-458, 765-769)"""
        A, Ac = phi.A[n], self.A[n]
        T1 = ops.gemm(self.R[n], A.view(A.shape[0], -1))                              # (Dc_l, d * Dr)
        T2 = ops.gemm(T1.view(-1, A.shape[2]), self.R[n + 1]).view(Ac.shape[0], A.shape[1], -1)   # (Dc_l, d, Dc_r)
        # env[p, p'] = sum_{a, b} T2[a, p, b] Ac[a, p', b]: batched over the physical leg via two permuted copies
        X = T2.permute(1, 0, 2).contiguous().view(A.shape[1], -1)
        Y = Ac.permute(1, 0, 2).contiguous().view(Ac.shape[1], -1)
        return ops.gemm(X, Y, transB=True)

    def expectation_mix(self, phi, n):
        """<self|phi> from the environments of site n (mps.py:587-591, 694-698)"""
        A, Ac = phi.A[n], self.A[n]
        T1 = ops.gemm(self.R[n], A.view(A.shape[0], -1))
        T2 = ops.gemm(T1.view(-1, A.shape[2]), self.R[n + 1])
        return ops.gemm(T2.view(1, -1), Ac.reshape(1, -1), transB=True).view(-1)

    def optimise_site(self, phi, n):
        """A[n] <- R[n] . phi.A[n] . R[n+1] (mps.py:617-621, 748-751); returns the first product for reuse"""
        A = phi.A[n]
        T1 = ops.gemm(self.R[n], A.view(A.shape[0], -1))
        out = ops.gemm(T1.view(-1, A.shape[2]), self.R[n + 1])
        self.A[n] = out.view(self.R[n].shape[0], A.shape[1], self.R[n + 1].shape[1])
        return T1

    def update_S(self):
        """Schmidt values of C; returns ||S_old - S_new||_2 as a device tensor (mps.py:550-560)"""
        S = ops.svd(self.C, want_vectors=False)
        if self.S[self.pC].numel() != S.numel():
            self.S[self.pC] = self._one_S(S.numel())
        dS = ops.diff_norm(self.S[self.pC], S)
        self.S[self.pC] = S
        return dS

    def variational_compress(self, phi, tol=None, max_sweeps=1, verbose=False):
        """one-site variational fit to phi (mps.py:238-279)"""
        if tol is None:
            tol = self.zero
        overlap = self.setup_RL_mix(phi)
        sweeps, diff = 0, 1.
        while diff > tol:
            if sweeps >= max_sweeps:
                return float(overlap.item())
            for n in range(self.L - 1, 0, -1):
                self.optimise_site(phi, n)
                self.orth_right(n)
                self.update_S()
                self.update_RR_mix(phi, n)
            dmax = torch.zeros(1, dtype=F64, device=self.device)
            for n in range(self.L):
                T1 = self.optimise_site(phi, n)
                self.orth_left(n)
                dmax = torch.maximum(dmax, self.update_S())
                self.update_RL_mix(phi, n, T1)
            diff = float(dmax.item())          # the only host read of the sweep
            overlap = self.R[-1]
            sweeps += 1
            if verbose:
                print("Sweep: %i Overlap: %.16f diff_S: %.4e" % (sweeps, float(overlap.item()), diff))
        return float(overlap.item())

    def compress_mps(self, Dmax=np.inf, tolS=None, tolV=None, max_sweeps=4, graduate_truncation=True, verbose=False):
        """the reference's fixed truncation schedule (mps.py:175-200)"""
        self.canonise_right()
        phi = self.copy()
        self.discarded = [0] * (self.L + 1)
        if graduate_truncation:
            self.canonise_left(compress=True, Dmax=Dmax * 4, tol=tolS / 10)
            self.variational_compress(phi, tol=tolV, max_sweeps=1, verbose=verbose)
            self.canonise_right(compress=True, Dmax=Dmax * 2, tol=tolS / 2)
        self.canonise_left(compress=True, Dmax=Dmax, tol=tolS)
        return self.variational_compress(phi, tol=tolV, max_sweeps=max_sweeps, verbose=verbose)


def apply_mpo_and_compress(psi, M, Hconj=True, Dmax=np.inf, tolS=1e-16, tolV=1e-10, max_sweeps=4, graduate_truncation=True):
    """``phi = psi.copy(); phi.apply_mpo(M, Hconj); overlap = phi.compress_mps(...)`` (tnac4o.py:1688-1693) as ONE call into
    the native row driver (csrc/mps_native.cu), which runs the same kernel sequence as the methods above without the
    interpreter in the loop.  Returns (phi, overlap); phi is left-canonical with ``discarded`` filled in."""
    import ctypes
    from ._native import Context, check, lib
    L = psi.L
    dev = psi.device
    c = Context.get(dev)
    PtrArr, IntArr = ctypes.c_void_p * L, ctypes.c_int * L
    A = [a.contiguous() for a in psi.A]
    W = [w.contiguous() for w in M.W]
    if not all(M.support):
        raise ValueError('apply_mpo_and_compress needs an MPO tensor on every site')
    wl = [w.shape[0] for w in W]
    wr = [w.shape[2] for w in W]
    du = [w.shape[3] if Hconj else w.shape[1] for w in W]
    handle = ctypes.c_void_p()
    dmax = float(min(Dmax, 2.0 ** 40))
    check(lib.tn_row_compress(c.handle, c.stream, L, PtrArr(*[a.data_ptr() for a in A]), IntArr(*[a.shape[0] for a in A]),
                              IntArr(*[a.shape[1] for a in A]), IntArr(*[a.shape[2] for a in A]),
                              PtrArr(*[w.data_ptr() for w in W]), IntArr(*wl), IntArr(*wr), IntArr(*du), int(bool(Hconj)),
                              dmax, float(tolS), float(tolV), int(max_sweeps), int(bool(graduate_truncation)),
                              ctypes.byref(handle)))
    try:
        D, d = (ctypes.c_int * (L + 1))(), IntArr()
        check(lib.tn_row_shapes(handle, D, d))
        phi = MPS.__new__(MPS)
        phi.device, phi.L, phi.zero, phi.dtype = dev, L, psi.zero, psi.dtype
        phi.D, phi.d = list(D), list(d)
        phi.A = [torch.empty((D[n], d[n], D[n + 1]), dtype=F64, device=dev) for n in range(L)]
        overlap, log2norm = ctypes.c_double(0.0), ctypes.c_double(0.0)
        disc = (ctypes.c_double * (L + 1))()
        check(lib.tn_row_fetch(handle, PtrArr(*[a.data_ptr() for a in phi.A]), ctypes.byref(overlap), disc,
                               ctypes.byref(log2norm)))
    finally:
        lib.tn_row_free(handle)
    phi.C = torch.ones((1, 1), dtype=F64, device=dev)
    phi.pC = L
    phi._log2_normC = psi._log2_normC + log2norm.value
    phi.reset_R()
    phi.reset_S()
    phi.discarded = [float(x) for x in disc]
    return phi, overlap.value


class MPO:
    """holder of rank-4 tensors W[n] with legs (left, out, right, in) (mps.py:818-884)"""

    def __init__(self, d=2, dout=None, L=2):
        self.L = L
        self.W = [None] * L
        self.support = [0] * L
        self.din = [d] * L if isinstance(d, int) else list(d)
        self.dout = list(self.din) if dout is None else ([dout] * L if isinstance(dout, int) else list(dout))

    def set_direct(self, W, n):
        """mps.py:859-865"""
        self.support[n] = 1
        self.W[n] = W
        self.dout[n], self.din[n] = W.shape[1], W.shape[3]
