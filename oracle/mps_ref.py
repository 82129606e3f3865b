"""Oracle restatement of the boundary-MPS machinery (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function names the reference lines it follows (paths relative to
/root/reference/tnac4o/).  Tensors are float64, C order, ``A[n]`` has legs
(left bond, physical, right bond) exactly like mps.py:96-173.
"""
import numpy as np
import scipy.linalg as sla


# ----------------------------------------------------------------------------- small wrappers
def ref_nfactor(T):
    """Largest |entry| floored to a power of two by exponent-field extraction (mps.py:76-85)."""
    top = np.float64(np.max(np.abs(T)))
    biased = np.abs(top).view(np.int64) >> 52
    return 2.0 ** (biased - 1023)


def ref_qr(T):
    """Economic QR with the diagonal of R made non-negative (mps.py:43-59)."""
    Q, R = sla.qr(T, mode='economic')
    sg = np.sign(np.diag(R).real)
    sg[sg == 0] = 1
    return Q * sg, sg[:, None] * R


def ref_svd(T):
    """Thin SVD, gesdd with gesvd fall-back, plus the reference's sign convention (mps.py:24-40)."""
    try:
        U, S, V = sla.svd(T, full_matrices=False)
    except sla.LinAlgError:
        U, S, V = sla.svd(T, full_matrices=False, lapack_driver='gesvd')
    flip = (np.abs(U.min(0)) > U.max(0)) & (np.abs(V.min(1)) > V.max(1))
    U[:, flip] *= -1
    V[flip] *= -1
    return U, S, V


def ref_svd_vals(T):
    """Singular values only (mps.py:62-73)."""
    try:
        return sla.svd(T, full_matrices=False, compute_uv=False)
    except sla.LinAlgError:
        return sla.svd(T, full_matrices=False, compute_uv=False, lapack_driver='gesvd')


def _unit_schmidt(D):
    s = np.zeros(D)
    s[0] = 1.0
    return s


# ----------------------------------------------------------------------------- the MPS object
class RefMPS:
    """Just the part of mps.MPS the solver touches.

    State: ``A`` (list of rank-3 tensors), centre matrix ``C`` at bond ``pC``, ``normC``,
    mixed environments ``R`` (length L+2, last entry = scalar overlap), Schmidt values ``S``
    and the per-bond ``discarded`` weights.
    """

    def __init__(self, L, d=1):
        # product state with all bonds 1: what mps.MPS(d=1, L=Nx, Dmax=1, initial='X') builds
        # (mps.py:108-157, 629-638) followed by the default canonise_left, which leaves each
        # tensor equal to [[[1]]] for d = 1.
        self.L = L
        self.A = [np.full((1, d, 1), 1.0 / np.sqrt(d)) for _ in range(L)]
        self.C = np.ones((1, 1))
        self.pC = L
        self.normC = 1.0
        self.R = [np.ones((1, 1)) for _ in range(L + 2)]
        self.R[-1] = None
        self.S = [_unit_schmidt(1) for _ in range(L + 1)]
        self.discarded = [0] * (L + 1)
        if d != 1:
            self.canonise_left()
            self.normC = 1.0

    # -- bookkeeping ---------------------------------------------------------
    @property
    def D(self):
        return [self.A[0].shape[0]] + [a.shape[2] for a in self.A]

    def copy(self):
        """Deep copy of tensors; environments shared by reference like mps.py:159-173."""
        other = RefMPS.__new__(RefMPS)
        other.L = self.L
        other.A = [a.copy() for a in self.A]
        other.C = self.C.copy()
        other.pC = self.pC
        other.normC = self.normC
        other.R = self.R[:]
        other.S = [_unit_schmidt(1) for _ in range(self.L + 1)]   # copy() starts from a fresh Dmax=1 object
        other.discarded = [0] * (self.L + 1)
        return other

    # -- MPO application -----------------------------------------------------
    def apply_mpo(self, W, conj=True):
        """A[n] <- W[n] applied on the physical leg (mps.py:353-359, 753-763).

        ``W[n]`` has legs (l, d_out, r, d_in).  conj=True contracts the MPS leg with ``d_out``
        and keeps ``d_in`` (boundary from the top); combined bonds are MPS-index-major.
        conj=False contracts with ``d_in`` and the combined bonds are MPO-index-major.
        """
        for n in range(self.L):
            A, Wn = self.A[n], W[n]
            if conj:
                T = np.tensordot(A, Wn, axes=(1, 1))          # a b l r u
                T = T.transpose(0, 2, 4, 1, 3)                  # a l u b r
            else:
                T = np.tensordot(Wn, A, axes=(3, 1))          # l o r a b
                T = T.transpose(0, 3, 1, 2, 4)                  # l a o r b
            s = T.shape
            self.A[n] = np.reshape(T, (s[0] * s[1], s[2], s[3] * s[4]))

    def apply_diagonal(self, diag, n):
        """Scale the physical leg of site n (mps.py:361-366)."""
        for k in range(len(diag)):
            self.A[n][:, k, :] *= diag[k]

    # -- moving the centre ---------------------------------------------------
    def absorb_right(self):
        """A[pC-1] <- A[pC-1] . C   (mps.py:368-373, 740-742)."""
        n = self.pC - 1
        self.A[n] = np.tensordot(self.A[n], self.C, axes=(2, 0))

    def absorb_left(self):
        """A[pC] <- C . A[pC]   (mps.py:375-380, 744-746)."""
        n = self.pC
        self.A[n] = np.tensordot(self.C, self.A[n], axes=(1, 0))

    def orth_left(self, n):
        """QR of (Dl*d, Dr); centre moves to bond n+1 (mps.py:532-539, 772-785)."""
        Dl, d, Dr = self.A[n].shape
        Q, C = ref_qr(self.A[n].reshape(Dl * d, Dr))
        nC = ref_nfactor(C)
        if C.shape == (1, 1):
            Q *= np.sign(C.flat[0])
            C = np.ones((1, 1))
        else:
            C = C / nC
        self.A[n] = Q.reshape(Dl, d, C.shape[0])
        self.C = C
        self.normC *= nC
        self.pC = n + 1

    def orth_right(self, n):
        """QR of the transposed (d*Dr, Dl) matrix; centre moves to bond n (mps.py:541-548, 787-800)."""
        Dl, d, Dr = self.A[n].shape
        Q, C = ref_qr(self.A[n].reshape(Dl, d * Dr).T)
        nC = ref_nfactor(C)
        if C.shape == (1, 1):
            Q *= np.sign(C.flat[0])
            C = np.ones((1, 1))
        else:
            C = C.T / nC
        self.A[n] = Q.T.reshape(C.shape[1], d, Dr)
        self.C = C
        self.normC *= nC
        self.pC = n

    def truncate_centre(self, Dmax, tol):
        """SVD-truncate C at an interior bond (mps.py:562-585, 802-811)."""
        if not (0 < self.pC < self.L):
            return 0.0
        U, S, V = ref_svd(self.C)
        tol = max(np.finfo(float).eps, tol)
        keep = min(int(np.sum(S > S[0] * tol)), Dmax)
        lost = np.sqrt(np.sum(S[keep:] ** 2)) / S[0]
        p = self.pC
        self.A[p - 1] = np.tensordot(self.A[p - 1], U[:, :keep], axes=(2, 0))
        self.A[p] = np.tensordot(V[:keep, :], self.A[p], axes=(1, 0))
        self.C = np.diag(S[:keep])
        self.discarded[p] = max(self.discarded[p], lost)
        return lost

    def canonise_left(self, compress=False, Dmax=np.inf, tol=None):
        """Left-to-right sweep (mps.py:202-218)."""
        self.C, self.pC = np.ones((1, 1)), 0
        for n in range(self.L):
            self.absorb_left()
            self.orth_left(n)
            if compress:
                self.truncate_centre(Dmax, tol)
        self.R[-1] = None

    def canonise_right(self, compress=False, Dmax=np.inf, tol=None):
        """Right-to-left sweep (mps.py:220-236)."""
        self.C, self.pC = np.ones((1, 1)), self.L
        for n in range(self.L - 1, -1, -1):
            self.absorb_right()
            self.orth_right(n)
            if compress:
                self.truncate_centre(Dmax, tol)
        self.R[-1] = None

    # -- mixed environments <self|phi> ---------------------------------------
    @staticmethod
    def env_left(RL, A, Ac):
        """mps.py:655-658."""
        return np.tensordot(Ac, np.tensordot(RL, A, axes=(1, 0)), axes=([0, 1], [0, 1]))

    @staticmethod
    def env_right(RR, A, Ac):
        """mps.py:660-663."""
        return np.tensordot(np.tensordot(A, RR, axes=(2, 0)), Ac, axes=([1, 2], [1, 2]))

    def push_left_env(self, phi, n):
        """mps.py:436-444."""
        new = self.env_left(self.R[n], phi.A[n], self.A[n])
        if n == self.L - 1:
            self.R[self.L + 1] = new.flat[0]
        else:
            self.R[n + 1] = new

    def push_right_env(self, phi, n):
        """mps.py:418-426."""
        new = self.env_right(self.R[n + 1], phi.A[n], self.A[n])
        if n == 0:
            self.R[self.L + 1] = new.flat[0]
        else:
            self.R[n] = new

    def bond_env(self, phi, n):
        """Environment of the physical leg of site n in <self|phi> (mps.py:454-458, 765-769)."""
        T = np.tensordot(np.tensordot(self.R[n], phi.A[n], axes=(1, 0)), self.R[n + 1], axes=(2, 0))
        return np.tensordot(T, self.A[n], axes=([0, 2], [0, 2]))

    def site_overlap(self, phi, n):
        """mps.py:587-591, 694-698."""
        T = np.tensordot(np.tensordot(self.R[n], phi.A[n], axes=(1, 0)), self.R[n + 1], axes=(2, 0))
        return np.tensordot(T, self.A[n], axes=((0, 1, 2), (0, 1, 2)))

    def refresh_schmidt(self):
        """mps.py:550-560: singular values of C, returns the 2-norm of the change."""
        S = ref_svd_vals(self.C)
        if self.S[self.pC].size != S.size:
            self.S[self.pC] = _unit_schmidt(S.size)
        change = np.sqrt(np.sum((self.S[self.pC] - S) ** 2))
        self.S[self.pC] = S
        return change

    def variational_compress(self, phi, tol=None, max_sweeps=1):
        """One-site variational fit of self to phi (mps.py:238-279, 617-621, 748-751)."""
        if tol is None:
            tol = np.finfo(float).eps
        for n in range(self.L):
            self.push_left_env(phi, n)
        overlap = self.R[-1]
        sweeps, diff = 0, 1.0
        while diff > tol:
            if sweeps >= max_sweeps:
                return overlap
            for n in range(self.L - 1, 0, -1):
                self._fit_site(phi, n)
                self.orth_right(n)
                self.refresh_schmidt()
                self.push_right_env(phi, n)
            diff = 0.0
            for n in range(self.L):
                self._fit_site(phi, n)
                self.orth_left(n)
                diff = np.maximum(diff, self.refresh_schmidt())
                self.push_left_env(phi, n)
            overlap = self.R[-1]
            sweeps += 1
        return overlap

    def _fit_site(self, phi, n):
        T = np.tensordot(self.R[n], phi.A[n], axes=(1, 0))
        self.A[n] = np.tensordot(T, self.R[n + 1], axes=(2, 0))

    def compress(self, Dmax, tolS, tolV, max_sweeps, graduate=True):
        """The fixed truncation schedule of mps.py:175-200."""
        self.canonise_right()
        phi = self.copy()
        self.discarded = [0] * (self.L + 1)
        if graduate:
            self.canonise_left(compress=True, Dmax=Dmax * 4, tol=tolS / 10)
            self.variational_compress(phi, tol=tolV, max_sweeps=1)
            self.canonise_right(compress=True, Dmax=Dmax * 2, tol=tolS / 2)
        self.canonise_left(compress=True, Dmax=Dmax, tol=tolS)
        return self.variational_compress(phi, tol=tolV, max_sweeps=max_sweeps)
