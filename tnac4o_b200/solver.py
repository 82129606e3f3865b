"""The solver object: the reference's ``tnac4o(mode, Nx, Ny, Nc, J, beta)`` API on top of the sm_100a kernels.

Drop-in for /root/reference/tnac4o/tnac4o.py on the contraction path: same constructor, same entry points
(``search_ground_state``, ``gibbs_sampling`` = ``sample``, ``search_low_energy_spectrum``,
``decode_low_energy_states``, ``precondition``, ``rotate_graph``, ``save`` / ``load``, ``binary_states``), same
result attributes and dtypes.  Host code prepares O(L) tables and steers the row / site loop; every
per-branch quantity lives in HBM and is produced by kernels of libtnac4o_b200.so.  There is no CPU fallback.
"""
import ctypes
import logging
import time

import numpy as np
import scipy.linalg
import scipy.sparse
import torch

from . import mps, ops
from ._native import Context, check, lib, ptr
from .model import HostTables, IsingLattice, RMFLattice, cell_bits, upload_site_tables, upper_triangular

F64 = torch.float64
_NEG_INF_BITS = 0x000FFFFFFFFFFFFF          # order-preserving encoding of -inf (common.cuh: ordered_bits)


def _decode_ordered(bits):
    bits = int(bits) & 0xFFFFFFFFFFFFFFFF
    raw = (bits & 0x7FFFFFFFFFFFFFFF) if (bits >> 63) else (~bits & 0xFFFFFFFFFFFFFFFF)
    return float(np.array([raw], dtype=np.uint64).view(np.float64)[0])


def load(file_name):
    """Load a solution saved with :meth:`tnac4o.save` (tnac4o.py:31-75); couplings are not part of the file."""
    d = np.load(file_name, allow_pickle=True).item()
    ins = tnac4o(mode=d.get('mode'), Nx=d.get('Nx'), Ny=d.get('Ny'), Nc=d.get('Nc'), beta=d.get('beta'))
    for key in ('energy', 'probability', 'degeneracy', 'states', 'discarded_probability', 'negative_probability'):
        setattr(ins, key, d.get(key))
    ins.ind0 = d.get('ind')
    ins.adj = np.zeros((0, 0))
    if d.get('excitations_encoding') is not None:
        ins.excitations_encoding = d.get('excitations_encoding')
        ins.d, ins.invd, ins.el, ins.free_d = d.get('d'), d.get('invd'), d.get('el'), d.get('free_d')
        if ins.excitations_encoding > 1:
            ins.adj = d.get('adj')
    return ins


class tnac4o:
    r"""Tensor-network solver for Ising problems on a quasi-2d lattice (see the reference docstring,
    tnac4o.py:78-143).  ``mode='Ising'``: spin index :math:`i = k N_x N_c + l N_c + m`, ``J`` a list of ``[i, j, Jij]``;
    ``mode='RMF'``: a Random Markov Field on the Ny x Nx grid, ``J = {'fun': ..., 'fac': ..., 'N': ..., 'Nx', 'Ny'}``
    (tnac4o.py:104-118)."""

    def __init__(self, mode='Ising', Nx=4, Ny=4, Nc=8, beta=1, J=None, device=None):
        if mode not in ('Ising', 'RMF'):
            raise ValueError("mode is 'Ising' or 'RMF'")
        if mode == 'RMF':
            Nc = 1
        if Nc > 8:
            raise ValueError('Single cluster is too large (cell states are stored as one byte: Nc <= 8).')
        self.mode, self.beta = mode, beta
        self.Nx_model, self.Ny_model = Nx, Ny
        self.Nx, self.Ny, self.Nc = Nx, Ny, Nc
        self.indtype = np.int8
        self.L = Nx * Ny * Nc
        self.order = np.arange(Nx * Ny)
        self.order_i = np.arange(Nx * Ny)
        self.logger = logging.getLogger('tnac4o')
        self.energy, self.probability = np.zeros(0), np.zeros(0)
        self.rotation, self.degeneracy = 0, 0
        self.states = np.zeros((0, Nx * Ny), dtype=self.indtype)
        self.discarded_probability, self.negative_probability = -np.inf, 0.
        self.device = None if device is None else torch.device(device)
        self.stats = {}
        self._sites = None
        self._host = None
        self.native_rows = True      # boundary-MPS rows through the native driver (False: Python MPS methods)
        self.native_search = True    # branch-and-bound loop through the native driver (False: the Python loop below)
        self.build_rhoT0 = False     # the reference also contracts the last row (rhoT[0] / rhoB[Ny]), which nothing reads
        if J is not None and mode == 'Ising':
            self.J = upper_triangular(J, self.L)
            self.J0 = self.J.copy()
            self._divide_couplings()
            self.ind0 = [[self.lat.ind[ny][nx] for nx in range(Nx)] for ny in range(Ny)]
            self.active = int(sum(len(a) for row in self.ind0 for a in row))
        elif J is not None:
            self.J = J.copy()                       # tnac4o.py:192-197
            self.Nrmf = np.asarray(J['N'], dtype=int).copy()
            if int(np.max(self.Nrmf)) > 256:
                raise ValueError('RMF: at most 256 states per site (states are stored as one byte)')
            self.J0, self.ind0 = [], []
            self._divide_couplings()

    # ------------------------------------------------------------------ model preparation (host)
    def _divide_couplings(self):
        """tnac4o.py:1391-1457"""
        if self.mode == 'RMF':
            self.lat = RMFLattice(self.J, self.Nx, self.Ny, self.Nrmf)
            for name in ('ind', 'N', 'sN', 'sl', 'sd', 'sr', 'su', 'll', 'lu', 'lr', 'ld'):
                setattr(self, name, getattr(self.lat, name))
            self._reset_X()
            return
        self.lat = IsingLattice(self.J, self.Nx, self.Ny, self.Nc)
        for name in ('ind', 'N', 'sN', 'sl', 'sd', 'sr', 'su', 'lr', 'ld', 'id', 'ir', 'Jin', 'Jl', 'Ju'):
            setattr(self, name, getattr(self.lat, name))
        self._reset_X()

    def _reset_X(self):
        """identity gauges (tnac4o.py:1811-1820)"""
        Ny, Nx = self.Ny, self.Nx
        self.Xu = np.ones((Ny, Nx, np.max(self.ld)))
        self.Xd = np.ones((Ny, Nx, np.max(self.ld)))
        self.Xl = np.ones((Ny, Nx, np.max(self.lr)))
        self.Xr = np.ones((Ny, Nx, np.max(self.lr)))
        self.overlaps_ud = np.empty(shape=[0, Ny - 1])
        self._sites = None
        self._host = None

    @staticmethod
    def _quarter_turn(Nx, Ny):
        """cell permutation of one quarter turn: cell a = ny * Nx + nx moves to turned[a] = (Nx - 1 - nx) * Ny + ny of the
        Ny-wide rotated lattice; returns (turned, back) with back[turned] = arange"""
        ny, nx = np.divmod(np.arange(Nx * Ny), Nx)
        turned = (Nx - 1 - nx) * Ny + ny
        back = np.empty_like(turned)
        back[turned] = np.arange(Nx * Ny)
        return turned, back

    def rotate_graph(self, rot=1):
        """quarter turns of the lattice, cumulative (tnac4o.py:290-340): every turn relabels the spins by the cell
        permutation, folds the coupling matrix back to its upper triangle and composes the cell order"""
        if self.mode == 'RMF':
            return self._rotate_rmf(rot)
        for _ in range(rot):
            turned, back = self._quarter_turn(self.Nx, self.Ny)
            # rotated J[i, j] = J[where[i], where[j]] with where = spin k of cell a -> spin k of cell turned[a]
            where = (turned[:, None] * self.Nc + np.arange(self.Nc)).ravel()
            Jc = scipy.sparse.coo_matrix(self.J)
            to = np.empty(self.L, dtype=np.int64)
            to[where] = np.arange(self.L)
            r, c = to[Jc.row], to[Jc.col]
            self.J = scipy.sparse.csr_matrix((Jc.data, (np.minimum(r, c), np.maximum(r, c))), shape=Jc.shape)
            self.Nx, self.Ny = self.Ny, self.Nx
            self.order = back[self.order]
            self.rotation += 1
        self.order_i[self.order] = np.arange(self.Nx * self.Ny)
        self.rotation = np.mod(self.rotation, 4)
        self._divide_couplings()

    def _rotate_rmf(self, rot):
        """site (ny, nx) -> (Nx - 1 - nx, ny): factor keys, sizes and the cell order (tnac4o.py:315-336)"""
        for _ in range(rot):      # (the reference does not advance `rotation` in this mode; kept: it only labels outputs)
            Nx, Ny = self.Nx, self.Ny
            turned, _ = self._quarter_turn(Nx, Ny)
            turn = lambda ny, nx: (Nx - nx - 1, ny)
            self.J['fac'] = {(turn(*key) if len(key) == 2 else turn(*key[:2]) + turn(*key[2:])): val
                             for key, val in self.J['fac'].items()}
            self.Nrmf = np.ascontiguousarray(np.asarray(self.Nrmf)[:, ::-1].T)      # sizes[Nx - 1 - nx, ny] = N[ny, nx]
            self.Nx, self.Ny = Ny, Nx
            self.order = turned[self.order]          # (this mode composes with the forward map, as the reference does)
        self.order_i[self.order] = np.arange(self.Nx * self.Ny)
        self.rotation = np.mod(self.rotation, 4)
        self._divide_couplings()

    def add_noise(self, amplitude=1e-7):
        """uniform noise in [-amplitude, amplitude] on every stored coupling / every one-site RMF function
        (tnac4o.py:917-941; consumes the global numpy RNG in the reference's order: stored entries row by row)"""
        self.logger.info('Adding noise to the coupling with ampliture %.2e', amplitude)
        draw = lambda n: (np.random.rand(n) * 2 - 1) * amplitude
        if self.mode == 'RMF':
            fun = {key: np.array(value, dtype=float) for key, value in self.J['fun'].items()}
            for f in fun.values():
                if f.ndim == 1:
                    f += draw(f.shape[0])
            self.J['fun'] = fun
        else:
            rows, cols = self.J.nonzero()
            self.J = scipy.sparse.csr_matrix(self.J + scipy.sparse.csr_matrix((draw(len(rows)), (rows, cols)), shape=self.J.shape))
        self._divide_couplings()

    # ------------------------------------------------------------------ device tables
    def _dev(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError('tnac4o_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
            self.device = torch.device('cuda', torch.cuda.current_device())
        return self.device

    def _host_tables(self):
        """small per-site tables on the host (pinned), rebuilt when beta or the gauges change"""
        if self._host is None or self._host_beta != self.beta:
            self._host = HostTables(self.lat, self.beta, (self.Xu, self.Xl, self.Xr, self.Xd))
            self._host_beta = self.beta
            self._sites = None
        return self._host

    def _upload_sites(self):
        dev = self._dev()
        host = self._host_tables()
        self._sites, self._site_buffers = upload_site_tables(host, self.Ny, self.Nx, dev)
        self._sites_beta = self.beta
        self.stats['h2d_bytes'] = host.nbytes
        return self._sites

    def _site_tables(self):
        if self._sites is None or self._sites_beta != self.beta:
            self._upload_sites()
        return self._sites

    def drop_device_tables(self):
        """forget the device copies of the site tables (the next call uploads the host tables again)"""
        self._sites = None

    # ------------------------------------------------------------------ boundary MPS
    def _row_mpo(self, ny):
        sites = self._site_tables()
        At = mps.MPO(L=self.Nx)
        for nx in range(self.Nx):
            At.set_direct(sites[ny][nx].Wmpo, nx)
        return At

    def _setup_rhoT(self, graduate_truncation=True, Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """boundary MPS of rows ny..Ny-1 for every ny, from the top (tnac4o.py:1674-1695)"""
        dev = self._dev()
        self.rhoT = [None] * (self.Ny + 1)
        self.rhoT_overlap = [1] * (self.Ny + 1)
        self.rhoT_discarded = [0] * (self.Ny + 1)
        self.rhoT[-1] = mps.MPS(d=1, L=self.Nx, Dmax=1, initial='X', device=dev)
        # rhoT[0] (all rows contracted) is never read by the searches or by the preconditioning -- the reference
        # builds it anyway (tnac4o.py:1683); here it is skipped unless `build_rhoT0` is set
        for ny in range(self.Ny - 1, -1 if self.build_rhoT0 else 0, -1):
            if self.native_rows:
                psi, self.rhoT_overlap[ny] = mps.apply_mpo_and_compress(
                    self.rhoT[ny + 1], self._row_mpo(ny), Hconj=True, Dmax=Dmax, tolS=tolS, tolV=tolV, max_sweeps=max_sweeps,
                    graduate_truncation=graduate_truncation)
            else:       # the same kernel sequence driven from Python through the reference's MPS interface
                psi = self.rhoT[ny + 1].copy()
                psi.apply_mpo(self._row_mpo(ny), Hconj=True)
                self.rhoT_overlap[ny] = psi.compress_mps(Dmax=Dmax, tolS=tolS, tolV=tolV, max_sweeps=max_sweeps,
                                                         graduate_truncation=graduate_truncation, verbose=False)
            self.rhoT_discarded[ny] = max(psi.discarded)
            self.rhoT[ny] = psi

    def _setup_rhoB(self, graduate_truncation=True, Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """the same from the bottom (tnac4o.py:1697-1718); used by the preconditioning only"""
        dev = self._dev()
        self.rhoB = [None] * (self.Ny + 1)
        self.rhoB[0] = mps.MPS(d=1, L=self.Nx, Dmax=1, initial='X', device=dev)
        # rhoB[Ny] is never read by the preconditioning (tnac4o.py:1838 loops over the interior cuts only)
        for ny in range(self.Ny if self.build_rhoT0 else self.Ny - 1):
            if self.native_rows:
                psi, _ = mps.apply_mpo_and_compress(self.rhoB[ny], self._row_mpo(ny), Hconj=False, Dmax=Dmax, tolS=tolS,
                                                    tolV=tolV, max_sweeps=max_sweeps, graduate_truncation=graduate_truncation)
            else:
                psi = self.rhoB[ny].copy()
                psi.apply_mpo(self._row_mpo(ny), Hconj=False)
                psi.compress_mps(Dmax=Dmax, tolS=tolS, tolV=tolV, max_sweeps=max_sweeps,
                                 graduate_truncation=graduate_truncation, verbose=False)
            self.rhoB[ny + 1] = psi

    # ------------------------------------------------------------------ preconditioning
    def precondition(self, mode='balancing', steps=2, beta_cond=[], Dmax_cond=[], max_scale=1024,
                     graduate_truncation=False, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """balancing heuristic on the vertical bonds (tnac4o.py:342-379)"""
        if mode != 'balancing':
            return
        if not beta_cond:
            beta_cond = [self.beta * 2. ** (nn - steps) for nn in range(steps)]
        if not Dmax_cond:
            Dmax_cond = [8] * len(beta_cond)
        main_beta = self.beta
        for b, D in zip(beta_cond, Dmax_cond):
            self.beta = b
            self.logger.info('Preconditioning with beta = %.2f', self.beta)
            keep_time = time.time()
            self._update_conditioning(direction='ud', Dmax=D, graduate_truncation=graduate_truncation, tolS=tolS,
                                      tolV=tolV, max_sweeps=max_sweeps, max_scale=max_scale)
            self.logger.info('Elapsed: %.2f seconds', time.time() - keep_time)
        self.beta = main_beta
        self._sites = None
        self._host = None

    def _update_conditioning(self, direction='ud', graduate_truncation=False, Dmax=8, tolS=1e-16, tolV=1e-10,
                             max_sweeps=4, max_scale=1024):
        """tnac4o.py:1824-1918.  Bond environments (a few 16 x 16 matrices) are balanced on the host with the same
        LAPACK routine as the reference (dgebal through scipy); all MPS work runs on the device."""
        if direction != 'ud':
            raise NotImplementedError("only direction='ud' is reachable in the reference (tnac4o.py:374-377)")
        cap = 2.0 ** np.floor(np.log2(np.sqrt(max_scale)))
        self._sites = None
        self._host = None
        self._setup_rhoT(graduate_truncation, Dmax, tolS, tolV, max_sweeps)
        self._setup_rhoB(graduate_truncation, Dmax, tolS, tolV, max_sweeps)
        Nx = self.Nx
        overlaps = np.ones((2, self.Ny - 1))

        def fro(t):
            return float(torch.linalg.vector_norm(t).item())

        def normalise_(t):
            t *= 1.0 / fro(t)

        for ny in range(1, self.Ny):
            bot, top = self.rhoB[ny], self.rhoT[ny]
            for nx in range(Nx):
                bot.update_RL_mix(top, nx)
                normalise_(bot.R[nx + 1])

            def rebalance(nx):
                env = bot.bond_env_mix(top, nx).cpu().numpy()
                _, scale = scipy.linalg.matrix_balance(env, permute=False, separate=True)
                scale = np.minimum(np.maximum(scale[0], 1 / cap), cap)
                o1 = float(bot.expectation_mix(top, nx).item()) / (fro(bot.A[nx]) * fro(top.A[nx]))
                bot.apply_diagonalO(scale, nx)
                top.apply_diagonalO(1 / scale, nx)
                o2 = float(bot.expectation_mix(top, nx).item()) / (fro(bot.A[nx]) * fro(top.A[nx]))
                if o1 < overlaps[0, ny - 1]:
                    overlaps[0, ny - 1] = o1
                    overlaps[1, ny - 1] = max(o1, o2)
                width = self.ld[ny - 1, nx]
                self.Xd[ny - 1, nx, :width] *= scale
                self.Xu[ny, nx, :width] *= 1 / scale

            for nx in range(Nx - 1, -1, -1):
                rebalance(nx)
                if nx > 0:
                    bot.orth_right(nx)
                    bot.attach_AC()
                    top.orth_right(nx)
                    top.attach_AC()
                    bot.update_RR_mix(top, nx)
                    normalise_(bot.R[nx])
            for nx in range(Nx):
                rebalance(nx)
                if nx < Nx - 1:
                    bot.orth_left(nx)
                    bot.attach_CA()
                    top.orth_left(nx)
                    top.attach_CA()
                    bot.update_RL_mix(top, nx)
                    normalise_(bot.R[nx + 1])
        self.overlaps_ud = np.vstack([self.overlaps_ud, overlaps])
        self.rhoB = []
        self._sites = None
        self._host = None

    def _check_limits(self, M, Dmax):
        """limits of the device path that the reference API does not have, checked up front"""
        if self.Nx > 63:
            raise ValueError('Nx = %d: the packed boundary key holds at most 64 positions (Nx <= 63)' % self.Nx)
        if Dmax > 256:
            raise ValueError('Dmax = %d: left environments are materialised for bond dimensions up to 256' % Dmax)
        if int(M) * int(np.max(self.N)) >= 2 ** 31:
            raise ValueError('M x (states per cell) = %d x %d exceeds the 32-bit candidate index' % (M, int(np.max(self.N))))
        if int(M) > 65535 * 64:
            raise ValueError('M = %d: the grouped GEMM of the right environments places row tiles in gridDim.z (M <= 4 193 280)' % M)
        for ny in range(self.Ny):
            self._key_offsets(ny, self.Nx - 1)          # raises when a boundary row needs more than 128 bits

    # ------------------------------------------------------------------ search machinery
    def _key_offsets(self, ny, nx):
        """bit offsets of the Nx+1 boundary indices inside the 128-bit merge key after site (ny, nx)"""
        widths = []
        for j in range(self.Nx + 1):
            if j <= nx:
                w = self.sd[ny][j]
            elif j == nx + 1:
                w = self.sr[ny][nx]
            else:
                w = self.su[ny][j - 1]
            widths.append(int(w))
        offs = np.concatenate(([0], np.cumsum(widths)[:-1])).astype(np.uint8)
        if sum(widths) > 128:
            raise ValueError('boundary index row needs %d bits; the merge key holds 128' % sum(widths))
        return offs

    class _Branches:
        """device arrays of the live partial configurations"""

        def __init__(self, cap, Nx, nsites, Dcap, dev):
            self.vind = torch.zeros((cap, Nx + 1), dtype=torch.uint8, device=dev)
            self.states = torch.zeros((cap, nsites), dtype=torch.uint8, device=dev)
            self.root = torch.zeros(cap, dtype=torch.int32, device=dev)
            self.Eng = torch.zeros(cap, dtype=F64, device=dev)
            self.prob = torch.zeros(cap, dtype=F64, device=dev)
            self.deg = torch.ones(cap, dtype=torch.int64, device=dev)
            self.RL = torch.ones((cap * Dcap,), dtype=F64, device=dev)
            self.n = 1

    def _alloc_search(self, cap, nsmax, Dcap, shards=None):
        dev = self._dev()
        nsites = self.Nx * self.Ny
        ws = {'shards': shards}
        ws['cur'] = self._Branches(cap, self.Nx, nsites, Dcap, dev)
        ws['nxt'] = self._Branches(cap, self.Nx, nsites, Dcap, dev)
        ncand = cap * nsmax
        kcap = ops.sort_capacity(ncand)
        pad = 0 if shards is None else shards.world        # gathered buffers hold world equal chunks (parallel.BranchShards)
        ws['cand'] = torch.empty((cap + pad) * nsmax, dtype=F64, device=dev)
        ws['flag'] = torch.empty(cap + pad, dtype=F64, device=dev)
        ws['surv'] = torch.empty(ncand, dtype=torch.int32, device=dev)
        for name in ('khi', 'klo', 'ktie'):
            ws[name] = torch.empty(kcap, dtype=torch.int64, device=dev)
        for name in ('parent', 'cell', 'g_rep', 'g_start', 'g_size', 'sel'):
            ws[name] = torch.empty(ncand, dtype=torch.int32, device=dev)
        for name in ('Enew', 'Pnew', 'g_prob', 'g_E'):
            ws[name] = torch.empty(ncand, dtype=F64, device=dev)
        ws['g_deg'] = torch.empty(ncand, dtype=torch.int64, device=dev)
        ws['count'] = torch.zeros(1, dtype=torch.int32, device=dev)
        ws['maxbits'] = torch.zeros(1, dtype=torch.int64, device=dev)
        ws['pdbits'] = torch.full((1,), _NEG_INF_BITS, dtype=torch.int64, device=dev)
        ws['gmin'] = torch.zeros(1, dtype=F64, device=dev)
        return ws

    def _setup_RR(self, br, ny, shards=None):
        """right environments of row ny for every row-start branch, all levels (tnac4o.py:1768-1784).
        RRat[nx] covers sites nx..Nx-1; site nx of the search uses RRat[nx + 1].  With ``shards`` the environments are
        REPLICATED: every rank contracts all row-start branches (67 MFLOP per level plus 0.5 MFLOP per branch on the DMMA
        path -- cheaper than exchanging 16 MB per level, and a branch met later in the row may descend from any of them)."""
        dev = self._dev()
        c = Context.get(dev)
        sites = self._site_tables()[ny]
        A = self.rhoT[ny + 1].A
        nb = br.n
        RRat = [None] * (self.Nx + 1)
        RRat[self.Nx] = torch.ones((nb, 1, 1), dtype=F64, device=dev)
        for nx in range(self.Nx - 1, 0, -1):
            Dl, nd, Dr = A[nx].shape
            out = torch.empty((nb, Dl, sites[nx].nl), dtype=F64, device=dev)
            up = br.vind[:, nx + 1:]
            check(lib.tn_rr_level(c.handle, c.stream, sites[nx].ref, nb, Dl, Dr, ptr(A[nx]), ptr(RRat[nx + 1]),
                                  up.data_ptr(), br.vind.stride(0), ptr(out)))
            RRat[nx] = out
        br.root[:nb] = torch.arange(nb, dtype=torch.int32, device=dev)
        return RRat

    def _site_marginals(self, ws, br, RRat, ny, nx, want_P=False):
        """T1 = RL . A on the DMMA path, then the fused marginal kernel; returns (Dl, nd, Dr) and optional P"""
        dev = self._dev()
        c = Context.get(dev)
        site = self._site_tables()[ny][nx]
        A = self.rhoT[ny + 1].A[nx]
        Dl, nd, Dr = A.shape
        B = br.n
        shards = ws.get('shards')
        if shards is not None:
            return self._site_marginals_sharded(ws, br, RRat, site, A, nx, shards)
        T1 = ops.gemm(br.RL[:B * Dl].view(B, Dl), A.view(Dl, nd * Dr))
        P = torch.empty((B, site.nS), dtype=F64, device=dev) if want_P else None
        cand = None if want_P == 'only' else ws['cand']
        check(lib.tn_marginals(c.handle, c.stream, site.ref, B, Dr, ptr(T1), ptr(RRat[nx + 1]), ptr(br.root), ptr(br.vind),
                               br.vind.stride(0), nx, ptr(br.prob) if cand is not None else None, ptr(cand), ptr(ws['flag']),
                               ptr(ws['maxbits']) if cand is not None else None, ptr(P)))
        ws['gmin'] = torch.minimum(ws['gmin'], ws['flag'][:B].min())
        self.stats['marginals'] = self.stats.get('marginals', 0) + B
        return P

    def _site_marginals_sharded(self, ws, br, RRat, site, A, nx, shards):
        """the same kernels on this rank's slice [lo, hi) of the branches, then the first half of the site's exchange step
        (SURVEY.md section 8e): the best candidate log-probability over all ranks (8-byte max-reduce).  The candidates
        themselves stay where they were computed; _select_sharded exchanges only those that survive the cut-off."""
        c = Context.get(self._dev())
        Dl, nd, Dr = A.shape
        B = br.n
        lo, hi = shards.slice(B)
        nS = site.nS
        if hi > lo:
            T1 = ops.gemm(br.RL[:B * Dl].view(B, Dl)[lo:hi], A.view(Dl, nd * Dr))
            check(lib.tn_marginals(c.handle, c.stream, site.ref, hi - lo, Dr, ptr(T1), ptr(RRat[nx + 1]), ptr(br.root[lo:]),
                                   ptr(br.vind[lo:]), br.vind.stride(0), nx, ptr(br.prob[lo:]), ptr(ws['cand'][lo * nS:]),
                                   ptr(ws['flag'][lo:]), ptr(ws['maxbits']), None))
            ws['gmin'] = torch.minimum(ws['gmin'], ws['flag'][lo:hi].min())
        else:
            ws['maxbits'].zero_()                      # the smallest ordered encoding: this rank contributes nothing
        shards.allreduce_max_ordered_(ws['maxbits'])
        self.stats['marginals'] = self.stats.get('marginals', 0) + (hi - lo)
        return None

    def _select_sharded(self, ws, B, nS, relative_P_cutoff, shards):
        """second half of the exchange step: every rank applies the relative cut-off (tnac4o.py:456-465) to ITS slice of the
        candidates, then the survivors -- (candidate id, log-probability), 12 bytes each, typically 10^3-10^4 of the
        B x 256 candidates -- are all-gathered, so that every rank holds the survivor list and their values the
        single-GPU path holds at this point.  Returns the number of survivors."""
        dev = self._dev()
        c = Context.get(dev)
        lo, hi = shards.slice(B)
        K = ctypes.c_int(0)
        if hi > lo:
            check(lib.tn_select(c.handle, c.stream, ptr(ws['cand'][lo * nS:]), (hi - lo) * nS, ptr(ws['maxbits']),
                                float(relative_P_cutoff), ptr(ws['surv']), ptr(ws['count']), ptr(ws['pdbits']), ctypes.byref(K)))
        k_loc = K.value
        ids = ws['surv'][:k_loc] + lo * nS                                   # global candidate ids of my survivors
        vals = ws['cand'][ids.long()]
        counts = shards.allgather_small(torch.tensor([k_loc], dtype=torch.int64, device=dev)).tolist()
        kmax = max(counts)
        pad_i = torch.zeros(kmax, dtype=torch.int32, device=dev)
        pad_v = torch.zeros(kmax, dtype=F64, device=dev)
        pad_i[:k_loc], pad_v[:k_loc] = ids, vals
        all_i = shards.allgather_small(pad_i).view(shards.world, kmax)
        all_v = shards.allgather_small(pad_v).view(shards.world, kmax)
        keep = torch.cat([all_i[r, :n] for r, n in enumerate(counts)])
        kept_v = torch.cat([all_v[r, :n] for r, n in enumerate(counts)])
        Ktot = int(keep.numel())
        ws['surv'][:Ktot] = keep
        ws['cand'][keep.long()] = kept_v                                     # expand reads cand[id] of every survivor
        return Ktot

    def _site_step(self, ws, RRat, ny, nx, M, relative_P_cutoff, min_dEng):
        """one site of the branch-and-bound: select -> expand -> merge -> top-M -> materialise (tnac4o.py:456-535)"""
        dev = self._dev()
        c = Context.get(dev)
        br, nxt = ws['cur'], ws['nxt']
        site = self._site_tables()[ny][nx]
        A = self.rhoT[ny + 1].A[nx]
        Dl, nd, Dr = A.shape
        B = br.n
        ncand = B * site.nS
        if ws.get('shards') is not None:
            K = self._select_sharded(ws, B, site.nS, relative_P_cutoff, ws['shards'])
        else:
            K = ctypes.c_int(0)
            check(lib.tn_select(c.handle, c.stream, ptr(ws['cand']), ncand, ptr(ws['maxbits']), float(relative_P_cutoff),
                                ptr(ws['surv']), ptr(ws['count']), ptr(ws['pdbits']), ctypes.byref(K)))
            K = K.value
        if K < 1:
            raise RuntimeError('no candidate survived the cut-off at site (%d, %d): all marginals are NaN' % (ny, nx))
        offs = self._key_offsets(ny, nx)
        check(lib.tn_expand(c.handle, c.stream, site.ref, K, nx, int(nx > 0), int(ny > 0), self.Nx + 1,
                            offs.ctypes.data_as(ctypes.c_void_p), ptr(ws['surv']), ptr(br.vind), br.vind.stride(0),
                            ptr(br.Eng), ptr(ws['cand']), ptr(ws['khi']), ptr(ws['klo']), ptr(ws['ktie']), ptr(ws['parent']),
                            ptr(ws['cell']), ptr(ws['Enew']), ptr(ws['Pnew'])))
        G = ctypes.c_int(0)
        check(lib.tn_merge(c.handle, c.stream, K, ptr(ws['khi']), ptr(ws['klo']), ptr(ws['ktie']), ptr(ws['Enew']),
                           ptr(ws['Pnew']), ptr(ws['parent']), ptr(br.deg), float(min_dEng), ptr(ws['g_rep']),
                           ptr(ws['g_deg']), ptr(ws['g_prob']), ptr(ws['g_E']), ptr(ws['g_start']), ptr(ws['g_size']),
                           ctypes.byref(G)))
        G = G.value
        groups = None
        if ws.get('want_groups'):
            # member lists in sorted order are needed by the droplet pass before the key arrays are reused
            groups = {'order': (ws['ktie'][:K] & 0xFFFFFFFF).to(torch.int32), 'K': K, 'G': G}
        check(lib.tn_topm(c.handle, c.stream, G, M, ptr(ws['g_prob']), ptr(ws['khi']), ptr(ws['klo']), ptr(ws['ktie']),
                          ptr(ws['sel']), ptr(ws['pdbits'])))
        Bn = min(G, M)
        check(lib.tn_materialise(c.handle, c.stream, site.ref, Bn, nx, ny * self.Nx + nx, self.Nx * self.Ny,
                                 br.vind.stride(0), Dl, Dr, ptr(ws['sel']), ptr(ws['g_rep']), ptr(ws['g_deg']),
                                 ptr(ws['g_prob']), ptr(ws['parent']), ptr(ws['cell']), ptr(ws['Enew']), ptr(br.vind),
                                 ptr(br.states), ptr(br.root), ptr(br.RL), ptr(A), ptr(nxt.vind), ptr(nxt.states),
                                 ptr(nxt.root), ptr(nxt.Eng), ptr(nxt.prob), ptr(nxt.deg), ptr(nxt.RL)))
        nxt.n = Bn
        ws['cur'], ws['nxt'] = nxt, br
        self.stats['candidates_kept'] = self.stats.get('candidates_kept', 0) + K
        return groups

    def _max_bond(self):
        return max(max(psi.D) for psi in self.rhoT if psi is not None)

    def _finish_search(self, ws, t_rho, t0):
        br = ws['cur']
        n = br.n
        shards = ws.get('shards')
        if shards is not None:
            shards.allreduce_min_(ws['gmin'])
            shards.allreduce_max_ordered_(ws['pdbits'])          # largest discarded log-probability over all slices
            self.stats['marginals_this_rank'] = self.stats.get('marginals', 0)
            tot = torch.tensor([float(self.stats.get('marginals', 0))], dtype=F64, device=self._dev())
            self.stats['marginals'] = int(shards.allreduce_sum_(tot).item())
            self.stats['bytes_gathered'] = shards.bytes_gathered
        torch.cuda.current_stream(self._dev()).synchronize()
        self.stats['seconds_rhoT'] = t_rho
        self.stats['seconds_search'] = time.time() - t0
        self.energy = br.Eng[:n].cpu().numpy()
        self.degeneracy = int(br.deg[0].item())
        states = br.states[:n].cpu().numpy().view(np.int8)
        self.states = states[:, self.order]
        self.probability = br.prob[:n].cpu().numpy()
        self.discarded_probability = _decode_ordered(ws['pdbits'].item())
        self.negative_probability = min(float(ws['gmin'].item()), 0)

    def _native_search(self, M, relative_P_cutoff, min_dEng, t_rho, t0):
        """the whole row / site loop of search_ground_state inside the library (csrc/search_native.cu): the same kernel
        sequence as _setup_RR / _site_marginals / _site_step below, without the interpreter between launches"""
        from ._native import TnSite
        dev = self._dev()
        c = Context.get(dev)
        Nx, Ny = self.Nx, self.Ny
        nsites = Nx * Ny
        sites = self._site_tables()
        site_arr = (TnSite * nsites)(*[sites[ny][nx].c for ny in range(Ny) for nx in range(Nx)])
        ptrs, dims = [], []
        for ny in range(Ny + 1):
            psi = self.rhoT[ny]
            ptrs += [0] * Nx if psi is None else [a.data_ptr() for a in psi.A]
            dims += [0] * (Nx + 1) if psi is None else [psi.A[0].shape[0]] + [a.shape[2] for a in psi.A]
        A_arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        D_arr = (ctypes.c_int * len(dims))(*dims)
        offs = np.concatenate([self._key_offsets(ny, nx) for ny in range(Ny) for nx in range(Nx)]).astype(np.uint8)
        states = torch.empty((M, nsites), dtype=torch.uint8, device=dev)
        Eng = torch.empty(M, dtype=F64, device=dev)
        prob = torch.empty(M, dtype=F64, device=dev)
        deg = torch.empty(M, dtype=torch.int64, device=dev)
        n, pd, neg, marg = ctypes.c_int(0), ctypes.c_double(0.0), ctypes.c_double(0.0), ctypes.c_int64(0)
        check(lib.tn_search_ground_state(c.handle, c.stream, Nx, Ny, site_arr, A_arr, D_arr,
                                         offs.ctypes.data_as(ctypes.c_void_p), int(M), float(relative_P_cutoff),
                                         float(min_dEng), ptr(states), ptr(Eng), ptr(prob), ptr(deg), ctypes.byref(n),
                                         ctypes.byref(pd), ctypes.byref(neg), ctypes.byref(marg)))
        n = n.value
        self.stats['marginals'] = int(marg.value)
        self.stats['seconds_rhoT'] = t_rho
        self.stats['seconds_search'] = time.time() - t0
        self.energy = Eng[:n].cpu().numpy()
        self.degeneracy = int(deg[0].item())
        self.states = states[:n].cpu().numpy().view(np.int8)[:, self.order]
        self.probability = prob[:n].cpu().numpy()
        self.discarded_probability = pd.value
        self.negative_probability = min(neg.value, 0)

    def _replicate_rhoT(self, shards):
        """Right environments are replicated: every rank builds the boundary MPS with the same
        deterministic kernels; rank 0's tensors are then broadcast so that the replicas are identical by construction."""
        shapes = [None if psi is None else [tuple(a.shape) for a in psi.A] for psi in self.rhoT]
        if not shards.same_everywhere(shapes):
            raise RuntimeError('boundary MPS bond dimensions differ between ranks')
        for psi in self.rhoT:
            if psi is not None:
                for a in psi.A:
                    shards.broadcast_(a, src=0)

    def search_ground_state(self, M=2 ** 10, relative_P_cutoff=1e-6, min_dEng=1e-12, graduate_truncation=True,
                            Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20, shards=None):
        """Branch-and-bound search for the most probable state (tnac4o.py:381-551).  Returns the energies.
        ``shards`` (a :class:`tnac4o_b200.parallel.BranchShards`) spreads the branch batch of this one search over the
        ranks of a process group; every rank returns the same result as the single-GPU call."""
        if shards is not None and shards.world == 1:
            shards = None
        if shards is not None:
            shards.bytes_gathered = 0
        self._check_limits(M, Dmax)
        dev = self._dev()
        c = Context.get(dev)
        t0 = time.time()
        self.stats = {}
        self.logger.info('Searching ground state with beta = %.2f', self.beta)
        self.logger.info('Preprocesing ... ')
        self._setup_rhoT(graduate_truncation=graduate_truncation, Dmax=Dmax, tolS=tolS, tolV=tolV, max_sweeps=max_sweeps)
        torch.cuda.current_stream(dev).synchronize()
        if shards is not None:
            self._replicate_rhoT(shards)
            torch.cuda.current_stream(dev).synchronize()
        t_rho = time.time() - t0
        self.logger.info('Elapsed: %.2f seconds', t_rho)
        t0 = time.time()
        self.logger.info('Searching ... ')
        if self.native_search and shards is None:
            self._native_search(M, relative_P_cutoff, min_dEng, t_rho, t0)
            self.logger.info('Elapsed search total: %.2f seconds', self.stats['seconds_rhoT'] + self.stats['seconds_search'])
            return self.energy
        ws = self._alloc_search(M, int(np.max(self.N)), self._max_bond(), shards)
        for ny in range(self.Ny):
            keep_time = time.time()
            self.logger.info('Row %d / %d', ny + 1, self.Ny)
            br = ws['cur']
            RRat = self._setup_RR(br, ny, shards)
            br.RL[:br.n] = 1.0
            for nx in range(self.Nx):
                self._site_marginals(ws, ws['cur'], RRat, ny, nx)
                self._site_step(ws, RRat, ny, nx, M, relative_P_cutoff, min_dEng)
            br = ws['cur']
            check(lib.tn_row_shift(c.handle, c.stream, br.n, br.vind.stride(0), ptr(br.vind)))
            self.logger.info('Elapsed: %.2f seconds', time.time() - keep_time)
        self._finish_search(ws, t_rho, t0)
        self.logger.info('Elapsed search total: %.2f seconds', self.stats['seconds_rhoT'] + self.stats['seconds_search'])
        return self.energy

    def gibbs_sampling(self, M=2 ** 10, graduate_truncation=True, Dmax=32, tolS=1e-15, tolV=1e-10, max_sweeps=20,
                       shard=None):
        """Sample M configurations from the Boltzmann distribution (tnac4o.py:553-650).  One np.random.rand(M)
        draw per site, in the reference's order, feeds the inverse-CDF kernel.  ``shard=(rank, world)`` keeps only
        this rank's slice of the M samples (every rank draws the same M uniforms, so the union over ranks equals the
        single-GPU result); gather with :func:`tnac4o_b200.parallel.gather_samples`."""
        from .parallel import UniformStream
        stream = UniformStream(M, *(shard or (0, 1)))
        M = stream.hi - stream.lo
        self._check_limits(M, Dmax)
        dev = self._dev()
        c = Context.get(dev)
        t0 = time.time()
        self.stats = {}
        self.logger.info('Preprocesing ... ')
        self._setup_rhoT(graduate_truncation=graduate_truncation, Dmax=Dmax, tolS=tolS, tolV=tolV, max_sweeps=max_sweeps)
        torch.cuda.current_stream(dev).synchronize()
        t_rho = time.time() - t0
        t0 = time.time()
        nsites = self.Nx * self.Ny
        Dcap = self._max_bond()
        cur = self._Branches(M, self.Nx, nsites, Dcap, dev)
        nxt = self._Branches(M, self.Nx, nsites, Dcap, dev)
        cur.n = nxt.n = M
        ws = {'flag': torch.empty(M, dtype=F64, device=dev), 'gmin': torch.ones(1, dtype=F64, device=dev), 'cand': None,
              'maxbits': None}
        parent = torch.empty(M, dtype=torch.int32, device=dev)
        cell = torch.empty(M, dtype=torch.int32, device=dev)
        Enew = torch.empty(M, dtype=F64, device=dev)
        self.logger.info('Sampling ... ')
        for ny in range(self.Ny):
            RRat = self._setup_RR(cur, ny)
            cur.RL[:M] = 1.0
            sites = self._site_tables()[ny]
            for nx in range(self.Nx):
                A = self.rhoT[ny + 1].A[nx]
                Dl, nd, Dr = A.shape
                P = self._site_marginals(ws, cur, RRat, ny, nx, want_P='only')
                uni = torch.from_numpy(np.ascontiguousarray(stream.draw())).to(dev)
                check(lib.tn_sample(c.handle, c.stream, sites[nx].ref, M, nx, int(nx > 0), int(ny > 0), ptr(P), ptr(uni),
                                    ptr(cur.vind), cur.vind.stride(0), ptr(cur.Eng), ptr(parent), ptr(cell), ptr(Enew)))
                check(lib.tn_materialise(c.handle, c.stream, sites[nx].ref, M, nx, ny * self.Nx + nx, nsites,
                                         cur.vind.stride(0), Dl, Dr, None, None, None, None, ptr(parent), ptr(cell),
                                         ptr(Enew), ptr(cur.vind), ptr(cur.states), ptr(cur.root), ptr(cur.RL), ptr(A),
                                         ptr(nxt.vind), ptr(nxt.states), ptr(nxt.root), ptr(nxt.Eng), None, None,
                                         ptr(nxt.RL)))
                cur, nxt = nxt, cur
            check(lib.tn_row_shift(c.handle, c.stream, M, cur.vind.stride(0), ptr(cur.vind)))
        torch.cuda.current_stream(dev).synchronize()
        self.stats['seconds_rhoT'], self.stats['seconds_search'] = t_rho, time.time() - t0
        self.energy = cur.Eng[:M].cpu().numpy()
        self.degeneracy = 0
        self.states = cur.states[:M].cpu().numpy().astype(int)[:, self.order]
        self.probability = np.zeros(1)
        self.discarded_probability = 0
        self.negative_probability = min(float(ws['gmin'].item()), 0)
        return self.energy

    sample = gibbs_sampling

    # ------------------------------------------------------------------ low-energy spectrum (droplets)
    def search_low_energy_spectrum(self, excitations_encoding=1, M=2 ** 10, relative_P_cutoff=1e-6, max_dEng=0.,
                                   lim_hd=0, min_dEng=1e-12, graduate_truncation=True, Dmax=32, tolS=1e-16,
                                   tolV=1e-10, max_sweeps=20, shards=None):
        """Ground state plus the hierarchy of droplets recorded while merging (tnac4o.py:652-1358).
        ``excitations_encoding=1``: independence from the snake order (727-915); ``2`` and ``3``: independence from the
        adjacency of the coupling graph, hierarchical or one layer (943-1358; host structure in droplets.py)."""
        if excitations_encoding not in (1, 2, 3):
            raise NotImplementedError('Available droplets handling strategies are excitations_encoding = 1, 2, 3.')
        self.excitations_encoding = excitations_encoding
        if shards is not None and shards.world == 1:
            shards = None
        self._check_limits(M, Dmax)
        dev = self._dev()
        c = Context.get(dev)
        t0 = time.time()
        self.stats = {}
        self.logger.info('Preprocesing ... ')
        self._setup_rhoT(graduate_truncation=graduate_truncation, Dmax=Dmax, tolS=tolS, tolV=tolV, max_sweeps=max_sweeps)
        torch.cuda.current_stream(dev).synchronize()
        if shards is not None:
            self._replicate_rhoT(shards)
            torch.cuda.current_stream(dev).synchronize()
        t_rho = time.time() - t0
        t0 = time.time()
        ws = self._alloc_search(M, int(np.max(self.N)), self._max_bond(), shards)
        ws['want_groups'] = True
        ws['gmin'] = torch.ones(1, dtype=F64, device=dev)
        self._exc_initialise()
        book = None
        if excitations_encoding == 1:
            self._book_open(M)
        if excitations_encoding > 1:
            from .droplets import AdjacencyDroplets
            book = AdjacencyDroplets(excitations_encoding, self.mode)
            if self.mode == 'RMF':
                book.set_grid(self.Nx, self.Ny)
            else:
                book.set_adjacency(self.J, [self.ind[ny][nx] for ny in range(self.Ny) for nx in range(self.Nx)])
        nsites = self.Nx * self.Ny
        self.logger.info('Searching ... ')
        for ny in range(self.Ny):
            self.logger.info('Layer %d / %d', ny + 1, self.Ny)
            RRat = self._setup_RR(ws['cur'], ny, shards)
            ws['cur'].RL[:ws['cur'].n] = 1.0
            for nx in range(self.Nx):
                self._site_marginals(ws, ws['cur'], RRat, ny, nx)
                old = ws['cur']
                groups = self._site_step(ws, RRat, ny, nx, M, relative_P_cutoff, min_dEng)
                if book is None:
                    self._record_droplets(ws, old, groups, ny * self.Nx + nx, nsites, max_dEng, lim_hd)
                else:
                    book.site_update(*self._merged_branches(ws, old, groups, ny * self.Nx + nx, nsites, max_dEng),
                                     max_dEng, lim_hd)
                    book.end_of_site()
            if book is not None:
                book.end_of_row()
            br = ws['cur']
            check(lib.tn_row_shift(c.handle, c.stream, br.n, br.vind.stride(0), ptr(br.vind)))
        self._finish_search(ws, t_rho, t0)
        if book is not None:
            book.finish(self.order_i, lim_hd)
            self.d, self.invd, self.el, self.free_d = book.d, book.invd, book.el, book.free_d
            # decoding works in the model's orientation (tnac4o.py:1131, 1356)
            if self.mode == 'RMF':
                book.set_grid(self.Nx_model, self.Ny_model)
            else:
                book.set_adjacency(self.J0, [self.ind0[ny][nx] for ny in range(self.Ny_model) for nx in range(self.Nx_model)])
            self.adj = book.adj
            return self.energy
        self._book_close()
        for key, (dpos, dstate) in self.d.items():
            dpos = self.order_i[dpos]
            srt = dpos.argsort()
            self.d[key] = (dpos[srt], dstate[srt])
        return self.energy

    def _merged_branches(self, ws, old, groups, last, nsites, max_dEng):
        """What the adjacency encodings record at one site (tnac4o.py:1063-1075, 1251-1262): for every kept branch the
        old branch of its winner, and for every branch merged into it within max_dEng (old branch, dE, dpos, dstate).
        The XOR differences come from one tn_xor_diff launch, as for encoding 1."""
        dev = self._dev()
        c = Context.get(dev)
        K, G = groups['K'], groups['G']
        order = groups['order'].cpu().numpy()
        Bn = ws['cur'].n
        host = lambda t, n: t[:n].cpu().numpy()
        g_rep, g_start, g_size = host(ws['g_rep'], G), host(ws['g_start'], G), host(ws['g_size'], G)
        g_E = host(ws['g_E'], G)
        sel = host(ws['sel'], Bn)
        Enew = host(ws['Enew'], K)
        parent, cell = host(ws['parent'], K), host(ws['cell'], K)
        pw, pm, pg = [], [], []
        for j, g in enumerate(sel):
            if g_size[g] > 1:
                members = order[g_start[g]:g_start[g] + g_size[g]]
                gap = Enew[members] - g_E[g]
                for m in members[(gap <= max_dEng) & (members != g_rep[g])]:
                    pw.append(g_rep[g]); pm.append(m); pg.append(j)
        merged = [[] for _ in sel]
        if pw:
            npairs = len(pw)
            i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device=dev)
            row_a, row_b = i32(parent[pw]), i32(parent[pm])
            cell_a, cell_b = i32(cell[pw]), i32(cell[pm])
            out_pos = torch.empty((npairs, nsites), dtype=torch.int16, device=dev)
            out_xor = torch.empty((npairs, nsites), dtype=torch.uint8, device=dev)
            out_len = torch.empty(npairs, dtype=torch.int32, device=dev)
            check(lib.tn_xor_diff(c.handle, c.stream, npairs, nsites, last, ptr(old.states), ptr(row_a), ptr(cell_a),
                                  ptr(row_b), ptr(cell_b), ptr(out_pos), ptr(out_xor), ptr(out_len)))
            out_pos, out_xor, out_len = out_pos.cpu().numpy(), out_xor.cpu().numpy().view(np.int8), out_len.cpu().numpy()
            for k in range(npairs):
                m, g = pm[k], sel[pg[k]]
                merged[pg[k]].append((int(parent[m]), Enew[m] - g_E[g], out_pos[k, :out_len[k]].astype(np.int64),
                                      out_xor[k, :out_len[k]].copy()))
        return [int(parent[g_rep[g]]) for g in sel], merged

    def _record_droplets(self, ws, old, groups, last, nsites, max_dEng, lim_hd):
        """excitation lists of the new branches (tnac4o.py:844-882) on the device (csrc/droplet_book.cu): pairs of winner and
        merged-away member, their XOR differences, the dictionary of droplet shapes, the new tree nodes with their pruned
        sub-excitations and the per-branch lists are all produced by integer kernels from the arrays the merge left in
        HBM; only two counters per site come back to the host"""
        c = Context.get(self._dev())
        K = groups['K']
        Bn = ws['cur'].n
        check(lib.tn_book_site(c.handle, self._book, int(last), int(K), int(Bn), ptr(groups['order']), ptr(ws['g_rep']),
                               ptr(ws['g_start']), ptr(ws['g_size']), ptr(ws['g_E']), ptr(ws['g_prob']), ptr(ws['sel']),
                               ptr(ws['Enew']), ptr(ws['Pnew']), ptr(ws['parent']), ptr(ws['cell']), ptr(old.states),
                               float(max_dEng), int(lim_hd) if self.mode == 'Ising' else -int(lim_hd)))

    @staticmethod
    def _materialise_pools(dE, dP, key, first, last, cptr, ccnt, cnode, cbud, roots):
        """node / children pools of csrc/droplet_book.cu -> the reference's nested tuples ((dE, key, first, last, dlogP),
        sub-excitations).  A child entry (node, budget) stands for _exc_cut_energy(node, budget) (tnac4o.py:2071-2079): its
        own children are kept while dE <= budget and viewed with min(their stored budget, budget - dE) -- nested prunings
        compose to exactly what the reference's eager recursion leaves.  Returns (list of trees, set of keys used)."""
        used = set()

        def build(n, budget):
            used.add(int(key[n]))
            kids = []
            for e in range(cptr[n], cptr[n] + ccnt[n]):
                ch = cnode[e]
                if dE[ch] <= budget:
                    kids.append(build(ch, min(cbud[e], budget - dE[ch])))
            return ((dE[n], int(key[n]), int(first[n]), int(last[n]), dP[n]), tuple(kids))

        return [build(int(n), np.inf) for n in roots], used

    def _book_open(self, M):
        c = Context.get(self._dev())
        handle = ctypes.c_void_p()
        check(lib.tn_book_create(c.handle, c.stream, self.Nx * self.Ny, int(M), ctypes.byref(handle)))
        self._book = handle

    def _book_close(self):
        """pools -> the reference's host structures: el (nested tuples of branch 0, pruned with the budgets the device kept
        lazily: _exc_cut_energy, tnac4o.py:2071-2079), d / invd / free_d (shapes the tree references)"""
        c = Context.get(self._dev())
        sizes = (ctypes.c_int64 * 6)()
        check(lib.tn_book_sizes(c.handle, self._book, sizes))
        nn, nc, nsh, nel, n0, npairs = (int(x) for x in sizes)
        f8 = lambda n: np.zeros(max(n, 1), dtype=np.float64)
        i4 = lambda n: np.zeros(max(n, 1), dtype=np.int32)
        dE, dP, key, first, last, cptr, ccnt = f8(nn), f8(nn), i4(nn), i4(nn), i4(nn), i4(nn), i4(nn)
        cnode, cbud = i4(nc), f8(nc)
        sptr, spos, sxor = i4(nsh + 1), np.zeros(max(nel, 1), dtype=np.int16), np.zeros(max(nel, 1), dtype=np.uint8)
        list0 = i4(n0)
        hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        check(lib.tn_book_export(c.handle, self._book, hp(dE), hp(dP), hp(key), hp(first), hp(last), hp(cptr), hp(ccnt),
                                 hp(cnode), hp(cbud), hp(sptr), hp(spos), hp(sxor), hp(list0)))
        lib.tn_book_free(self._book)
        self._book = None
        self.stats['droplet_pairs'] = npairs
        self.stats['droplet_nodes'] = nn
        self.el, used = self._materialise_pools(dE, dP, key, first, last, cptr, ccnt, cnode, cbud, list0[:n0])
        self.d, self.invd = {}, {}
        for k in sorted(used):
            dpos = spos[sptr[k]:sptr[k + 1]].astype(np.int64)
            dstate = sxor[sptr[k]:sptr[k + 1]].view(np.int8).copy()
            self.d[k] = (dpos, dstate)
            self.invd.setdefault(self._exc_get_sh((dpos, dstate)), []).append(k)
        self.free_d = nsh

    # ---- droplet dictionary / tree (host objects in the reference's save format; tnac4o.py:2012-2079, 2249-2285)
    def _exc_initialise(self):
        self.d, self.invd, self.el, self.free_d = {}, {}, [[]], 0

    @staticmethod
    def _exc_get_sh(exc):
        return (exc[0][0], exc[1][0], exc[0][-1], exc[1][-1])

    def _exc_hd(self, dstate):
        """tnac4o.py:2143-2150"""
        if self.mode == 'RMF':
            return sum(bin(int(st)).count('1') for st in dstate)
        return len(dstate)

    def _adjacency_book(self):
        """droplet structure of encodings 2 / 3 around the solver's (or a loaded file's) d / el / adj"""
        from .droplets import AdjacencyDroplets
        book = AdjacencyDroplets(self.excitations_encoding, self.mode)
        if self.mode == 'RMF':
            book.set_grid(self.Nx_model, self.Ny_model)
        else:
            book.set_adjacency(self.adj, [self.ind0[ny][nx] for ny in range(self.Ny_model) for nx in range(self.Nx_model)])
        book.d, book.invd, book.el, book.free_d = self.d, self.invd, self.el, self.free_d
        return book

    def _exc_unpack(self, max_dEng=0., max_states=np.inf):
        """droplet combinations below max_dEng for the adjacency encodings 2 and 3 (tnac4o.py:2337-2377); encoding 1 is
        enumerated on the device (_decode_device)"""
        return self._adjacency_book().unpack(max_dEng=max_dEng, max_states=max_states)

    def _flatten_tree(self, slot):
        """nested excitation tuples -> flat node arrays for tn_decode_enumerate: node 0 is the root the reference puts
        under every stack (tnac4o.py:2308: dE 0, first -1, last N - 1); children keep their order"""
        nsites = self.Nx_model * self.Ny_model
        dE, key, first, last, kids = [0.0], [0], [-1], [nsites - 1], [[]]
        todo = [(0, self.el)]
        while todo:
            node, children = todo.pop()
            for exc in children:
                head = exc[0]
                k = len(dE)
                dE.append(float(head[0])); key.append(slot[head[1]]); first.append(int(head[2])); last.append(int(head[3]))
                kids.append([])
                kids[node].append(k)
                todo.append((k, exc[1]))
        child_ptr = np.zeros(len(dE) + 1, dtype=np.int32)
        child_ptr[1:] = np.cumsum([len(c) for c in kids])
        child_idx = np.array([c for cs in kids for c in cs], dtype=np.int32)
        return (np.array(dE, dtype=np.float64), np.array(key, dtype=np.int32), np.array(first, dtype=np.int32),
                np.array(last, dtype=np.int32), child_ptr, child_idx)

    def _droplet_csr_host(self):
        """dictionary of droplet shapes as CSR arrays (numpy) + key -> slot map"""
        keys = sorted(self.d)
        slot = {k: i for i, k in enumerate(keys)}
        drop_ptr = np.zeros(len(keys) + 1, dtype=np.int32)
        for i, k in enumerate(keys):
            drop_ptr[i + 1] = drop_ptr[i] + len(self.d[k][0])
        drop_pos = np.concatenate([self.d[k][0] for k in keys]).astype(np.int16) if keys else np.zeros(1, np.int16)
        drop_xor = np.concatenate([self.d[k][1] for k in keys]).astype(np.int8).view(np.uint8) if keys else np.zeros(1, np.uint8)
        return slot, drop_ptr, drop_pos, drop_xor

    def _droplet_csr(self):
        dev = self._dev()
        slot, drop_ptr, drop_pos, drop_xor = self._droplet_csr_host()
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return slot, up(drop_ptr), up(drop_pos), up(drop_xor)

    def _decode_device(self, max_dEng, max_states):
        """encoding 1: level-synchronous expansion of the flattened tree on the device (csrc/decode.cu) -> sorted
        excitation energies and states"""
        dev = self._dev()
        c = Context.get(dev)
        nsites = self.Nx_model * self.Ny_model
        slot, drop_ptr, drop_pos, drop_xor = self._droplet_csr()
        dE, key, first, last, child_ptr, child_idx = self._flatten_tree(slot)
        hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        handle, count = ctypes.c_void_p(), ctypes.c_int64(0)
        cap = int(min(max_states, 2 ** 62))
        check(lib.tn_decode_enumerate(c.handle, c.stream, nsites, len(dE), hp(dE), hp(key), hp(first), hp(last), hp(child_ptr),
                                      hp(child_idx), float(max_dEng), cap, ctypes.byref(handle), ctypes.byref(count)))
        try:
            n = int(count.value)
            ground = torch.from_numpy(np.ascontiguousarray(self.states[0]).view(np.uint8).copy()).to(dev)
            Eng = torch.empty(n, dtype=F64, device=dev)
            states = torch.empty((n, nsites), dtype=torch.uint8, device=dev)
            check(lib.tn_decode_fetch(c.handle, handle, n, ptr(ground), ptr(drop_ptr), ptr(drop_pos), ptr(drop_xor), ptr(Eng),
                                      ptr(states)))
            torch.cuda.current_stream(dev).synchronize()
        finally:
            lib.tn_decode_free(handle)
        return Eng, states

    def decode_low_energy_states(self, max_dEng=0., max_states=1024):
        """Expand the droplet tree into states (tnac4o.py:1360-1389).  Encoding 1: enumeration, top-max_states cut, energy
        sort and the XOR of droplet shapes onto the ground state all run on the device (tn_decode_enumerate / _fetch);
        encodings 2 and 3: host enumeration (droplets.py) + one XOR kernel over all states (tn_apply_droplets)."""
        dev = self._dev()
        c = Context.get(dev)
        t0 = time.time()
        if getattr(self, 'excitations_encoding', 1) == 1:
            Eng, states = self._decode_device(max_dEng, max_states)
            Eng = Eng.cpu().numpy()
            self.energy = Eng + self.energy[0]
            self.states = states.cpu().numpy().view(np.int8)
            self.stats['seconds_decode'] = time.time() - t0
            return Eng[0]
        Eng, flip = self._exc_unpack(max_dEng=max_dEng, max_states=max_states)
        order = Eng.argsort()
        Eng = Eng[order]
        count = min(max_states, len(Eng))
        nsites = self.Nx * self.Ny
        slot, drop_ptr, drop_pos, drop_xor = self._droplet_csr()
        flip_ptr = np.zeros(count + 1, dtype=np.int32)
        flat = []
        for i in range(count):
            f = flip[order[i]]
            flat.extend(slot[k] for k in f)
            flip_ptr[i + 1] = len(flat)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        ground = up(np.ascontiguousarray(self.states[0]).view(np.uint8))
        out = torch.empty((count, nsites), dtype=torch.uint8, device=dev)
        t = [up(flip_ptr), up(np.asarray(flat if flat else [0], dtype=np.int32))]
        check(lib.tn_apply_droplets(c.handle, c.stream, count, nsites, ptr(ground), ptr(t[0]), ptr(t[1]), ptr(drop_ptr),
                                    ptr(drop_pos), ptr(drop_xor), ptr(out)))
        self.energy = Eng + self.energy[0]
        self.states = out.cpu().numpy().view(np.int8)
        self.stats['seconds_decode'] = time.time() - t0
        return Eng[0]

    # ------------------------------------------------------------------ results / files
    def binary_states(self, number=-1):
        """cell states -> spins: 1 up, 0 down, 2 inactive (tnac4o.py:261-286)"""
        ns = self.states.shape[0]
        ns = ns + number + 1 if number < 0 else min(number, ns)
        if self.mode == 'RMF':
            return self.states[:ns]
        out = np.zeros((ns, self.L), dtype=np.int8) + 2
        k = -1
        for ny in range(self.Ny_model):
            for nx in range(self.Nx_model):
                k += 1
                spins = self.ind0[ny][nx]
                out[:, spins] = (1 - cell_bits(len(spins)))[self.states[:ns, k]]
        return out

    def save(self, file_name):
        """np.save of the result dictionary in the reference's layout (tnac4o.py:200-233)"""
        d = {'mode': self.mode, 'rotation': self.rotation, 'energy': self.energy, 'probability': self.probability,
             'degeneracy': self.degeneracy, 'states': self.states, 'discarded_probability': self.discarded_probability,
             'negative_probability': self.negative_probability, 'Nx': self.Nx_model, 'Ny': self.Ny_model, 'Nc': self.Nc,
             'beta': self.beta, 'ind': self.ind0}
        if hasattr(self, 'excitations_encoding'):
            d.update({'excitations_encoding': self.excitations_encoding, 'd': self.d, 'invd': self.invd, 'el': self.el,
                      'free_d': self.free_d})
            if self.excitations_encoding > 1:
                d['adj'] = scipy.sparse.csr_matrix(self.adj)
        np.save(file_name, d)

    def show_properties(self):
        print("L:     ", self.L)
        print("Ny:    ", self.Ny)
        print("Nx:    ", self.Nx)
        print("Beta:  ", self.beta)

    def show_solution(self, state=False):
        if len(self.energy) > 0:
            print("Energy            : %4.6f" % self.energy[0])
            print("Degeneracy        : %2d" % self.degeneracy)
            print("log2(Probability) : %0.2e" % self.probability[0])
            print("Discarder log2(P) : %0.2e" % self.discarded_probability)
            print("Min P (err)       : %0.2e" % self.negative_probability)
            print("# of states       : %1d" % len(self.energy))
            print("Rotation/direction: %1d" % self.rotation)
            if state:
                print(self.states[0])
        else:
            print('No solution to show.')

    def exc_print(self):
        self._exc_print(self.el, 1)

    def _exc_print(self, el, layer):
        for exc in el:
            kk = self.d[exc[0][1]]
            print((3 * layer - 3) * ' ' + "|- %0.4f " % (exc[0][0]) + ' : ' + ' '.join(map(str, kk[0])) + ' | ' +
                  ' '.join(map(str, kk[1])))
            self._exc_print(exc[1], layer + 1)
