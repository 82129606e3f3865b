"""Host-side model preparation: coupling blocks, leg sizes and the per-site constant tables that are uploaded
to the GPU once per (model, gauge).  O(L) numpy work, kept on the host on purpose:
the energy tables must be BIT-identical to the reference's, so they are produced by the same numpy
expressions (tnac4o.py:1513-1529); the Boltzmann weights follow tnac4o.py:1566-1607 (energies added first, one
exp, gauges multiplied in the order Xu, Xl, Xr, Xd) but are never expanded to the dense 5-leg tensor.
"""
import ctypes

import numpy as np
import scipy.sparse
import torch

from ._native import TnSite


def cell_bits(n):
    """bit_a(s) for all 2^n states (tnac4o.py:1461-1467: conf = 1 - bit, first spin fastest)"""
    s = np.arange(2 ** n)[:, None]
    return ((s >> np.arange(n)[None, :]) & 1).astype(np.int8)


def cell_spins(n):
    """sigma_a(s) = 1 - 2 bit_a(s), int8 like ``2 * _cluster_configurations(n) - 1``"""
    return (1 - 2 * cell_bits(n)).astype(np.int8)


def pext_table(n, positions):
    """bond index selected by a cell state: bits at ``positions`` packed little-endian (tnac4o.py:1469-1487)"""
    bits = cell_bits(n).astype(np.int64)
    out = np.zeros(2 ** n, dtype=np.int64)
    for j, a in enumerate(positions):
        out += bits[:, a] << j
    return out


class IsingLattice:
    """Couplings of an Ny x Nx lattice of Nc-spin cells split into per-cell blocks (tnac4o.py:1391-1457)."""

    def __init__(self, J, Nx, Ny, Nc):
        self.Nx, self.Ny, self.Nc = Nx, Ny, Nc
        self.J = J          # scipy sparse, upper triangular
        self.divide()

    def active_spins(self, ny, nx):
        """spins of a cell with a non-zero coupling (threshold 1e-12, tnac4o.py:1408-1411)"""
        ind = self.Nc * (self.Nx * ny + nx) + np.arange(self.Nc)
        weight = np.sum(np.abs(self.J[ind, :].toarray()), axis=1) + np.sum(np.abs(self.J[:, ind].toarray()), axis=0)
        return ind[np.nonzero(weight > 1e-12)]

    def divide(self):
        Ny, Nx = self.Ny, self.Nx
        self.ind = [[self.active_spins(ny, nx) for nx in range(Nx)] for ny in range(Ny)]
        self.sN = np.array([[len(self.ind[ny][nx]) for nx in range(Nx)] for ny in range(Ny)], dtype=int)
        self.N = 2 ** self.sN
        none = np.zeros(0, dtype=int)
        self.Jin = [[None] * Nx for _ in range(Ny)]
        self.Jl = [[np.zeros((self.sN[ny][nx], 0)) for nx in range(Nx)] for ny in range(Ny)]
        self.Ju = [[np.zeros((self.sN[ny][nx], 0)) for nx in range(Nx)] for ny in range(Ny)]
        self.id = [[none] * Nx for _ in range(Ny)]
        self.ir = [[none] * Nx for _ in range(Ny)]
        self.sl, self.sd, self.sr, self.su = (np.zeros((Ny, Nx), dtype=int) for _ in range(4))
        for ny in range(Ny):
            for nx in range(Nx):
                here = self.ind[ny][nx]
                self.Jin[ny][nx] = self.J[here, :][:, here].toarray()
                if nx > 0:
                    block = self.J[self.ind[ny][nx - 1]][:, here].toarray()
                    rows = np.nonzero(np.sum(np.abs(block), axis=1))[0]
                    self.Jl[ny][nx] = block[rows].T
                    self.ir[ny][nx - 1] = rows
                    self.sr[ny][nx - 1] = self.sl[ny][nx] = len(rows)
                if ny > 0:
                    block = self.J[self.ind[ny - 1][nx]][:, here].toarray()
                    rows = np.nonzero(np.sum(np.abs(block), axis=1))[0]
                    self.Ju[ny][nx] = block[rows].T
                    self.id[ny - 1][nx] = rows
                    self.sd[ny - 1][nx] = self.su[ny][nx] = len(rows)
        self.ll, self.lu = 2 ** self.sl, 2 ** self.su
        self.lr, self.ld = 2 ** self.sr, 2 ** self.sd

    # ---- per-site tables ---------------------------------------------------------------------
    def energy_tables(self, ny, nx):
        """Es[s], Esl[s, l], Esu[s, u] -- the reference's expressions verbatim in arithmetic (tnac4o.py:1512-1529)"""
        st = cell_spins(self.sN[ny][nx])
        Jin = self.Jin[ny][nx]
        Es = 1. * np.sum(np.dot(st, np.triu(Jin, 1)) * st, 1) + np.dot(st, Jin.diagonal())
        Esl = np.dot(np.dot(st, self.Jl[ny][nx]), cell_spins(self.sl[ny][nx]).T)
        Esu = np.dot(np.dot(st, self.Ju[ny][nx]), cell_spins(self.su[ny][nx]).T)
        return Es, Esl, Esu

    def boltzmann(self, ny, nx, beta, Xu, Xl, Xr, Xd):
        """Wc[s, l, u] with all four gauges folded in, and the bond maps d(s), r(s) (tnac4o.py:1566-1607)"""
        n = self.sN[ny][nx]
        L1, L4 = self.sl[ny][nx], self.su[ny][nx]
        st = cell_spins(n)
        Jin = self.Jin[ny][nx]
        Es = np.sum(np.dot(st, np.triu(Jin, 1)) * st, 1) + np.dot(st, Jin.diagonal())
        Es = beta * (np.min(Es) - Es)
        E1 = np.dot(np.dot(st, self.Jl[ny][nx]), cell_spins(L1).T)
        E1 = beta * (np.min(E1) - E1)
        E4 = np.dot(np.dot(st, self.Ju[ny][nx]), cell_spins(L4).T)
        E4 = beta * (np.min(E4) - E4)
        Wc = np.exp((Es[:, None, None] + E1[:, :, None]) + E4[:, None, :])
        Wc = Wc * Xu[None, None, :2 ** L4]
        Wc = Wc * Xl[None, :2 ** L1, None]
        dmap = pext_table(n, self.id[ny][nx])
        rmap = pext_table(n, self.ir[ny][nx])
        Wc = Wc * Xr[rmap][:, None, None]
        Wc = Wc * Xd[dmap][:, None, None]
        return Wc, dmap, rmap

    def exponents(self, ny, nx, beta):
        """beta-scaled, shifted energies E0[s], E1[s, l], E4[s, u] whose sum is exponentiated (tnac4o.py:1571-1583) and
        the bond maps; the small host-side input of the device table builder"""
        n = self.sN[ny][nx]
        st = cell_spins(n)
        Jin = self.Jin[ny][nx]
        Es = np.sum(np.dot(st, np.triu(Jin, 1)) * st, 1) + np.dot(st, Jin.diagonal())
        E0 = beta * (np.min(Es) - Es)
        E1 = np.dot(np.dot(st, self.Jl[ny][nx]), cell_spins(self.sl[ny][nx]).T)
        E1 = beta * (np.min(E1) - E1)
        E4 = np.dot(np.dot(st, self.Ju[ny][nx]), cell_spins(self.su[ny][nx]).T)
        E4 = beta * (np.min(E4) - E4)
        return E0, E1, E4, pext_table(n, self.id[ny][nx]), pext_table(n, self.ir[ny][nx])

    @staticmethod
    def traced(Wc, dmap, rmap, nd, nr):
        """sum over the cell state: legs (l, d, r, u); terms added in ascending s like np.sum(axis=0) (tnac4o.py:1686)"""
        W = np.zeros((Wc.shape[1], nd, nr, Wc.shape[2]))
        for s in range(Wc.shape[0]):
            W[:, dmap[s], rmap[s], :] += Wc[s]
        return W


class RMFLattice:
    """Random Markov Field on an Ny x Nx grid (mode='RMF', tnac4o.py:160-163, 1446-1456): one variable per site with N[ny][nx]
    local states, factors ``J['fun'][J['fac'][key]]`` on sites (key = (ny, nx): vector) and on nearest-neighbour bonds
    (key = (ny1, nx1, ny2, nx2): matrix indexed [state 1, state 2]).  Same interface as :class:`IsingLattice`; a bond index
    IS the neighbouring state, so the bond maps are the identity (or 0 on legs without a factor)."""

    def __init__(self, J, Nx, Ny, N):
        self.Nx, self.Ny, self.Nc = Nx, Ny, 1
        self.J = J
        self.N = np.asarray(N, dtype=int).copy()
        self.divide()

    def divide(self):
        Ny, Nx, fac = self.Ny, self.Nx, self.J['fac']
        self.ll, self.lu, self.lr, self.ld = (np.ones((Ny, Nx), dtype=int) for _ in range(4))
        for ny in range(Ny):
            for nx in range(Nx):
                if ((ny, nx - 1, ny, nx) in fac) or ((ny, nx, ny, nx - 1) in fac):
                    self.ll[ny, nx] = self.N[ny][nx - 1]
                if ((ny, nx, ny, nx + 1) in fac) or ((ny, nx + 1, ny, nx) in fac):
                    self.lr[ny, nx] = self.N[ny][nx + 1]
                if ((ny - 1, nx, ny, nx) in fac) or ((ny, nx, ny - 1, nx) in fac):
                    self.lu[ny, nx] = self.N[ny - 1][nx]
                if ((ny, nx, ny + 1, nx) in fac) or ((ny + 1, nx, ny, nx) in fac):
                    self.ld[ny, nx] = self.N[ny + 1][nx]
                for leg in (self.lr[ny, nx], self.ld[ny, nx]):
                    # the reference's tensor only exists for legs of size 1 or N (tnac4o.py:1648-1665)
                    if leg not in (1, self.N[ny][nx]):
                        raise ValueError('RMF: neighbouring sites coupled by a factor must have the same number of states')
        width = lambda a: np.ceil(np.log2(np.maximum(a, 1))).astype(int)       # bits of a bond index in the merge key
        self.sl, self.su, self.sr, self.sd = width(self.ll), width(self.lu), width(self.lr), width(self.ld)
        self.sN = width(self.N)
        self.ind = [[np.array([ny * Nx + nx]) for nx in range(Nx)] for ny in range(Ny)]

    def _bond(self, a, b, shape):
        """factor between site a = (ny, nx) (row index) and its left / upper neighbour b (column index)"""
        fac, fun = self.J['fac'], self.J['fun']
        if b + a in fac:
            return np.asarray(fun[fac[b + a]], dtype=float).T
        if a + b in fac:
            return np.asarray(fun[fac[a + b]], dtype=float)
        return np.zeros(shape)

    def energy_tables(self, ny, nx):
        """Es[s], Esl[s, left state], Esu[s, upper state] (tnac4o.py:1532-1557)"""
        N = int(self.N[ny][nx])
        fac, fun = self.J['fac'], self.J['fun']
        Es = np.asarray(fun[fac[(ny, nx)]], dtype=float).reshape(N) if (ny, nx) in fac else np.zeros(N)
        Esl = self._bond((ny, nx), (ny, nx - 1), (N, int(self.ll[ny, nx]))) if nx > 0 else np.zeros((N, 1))
        Esu = self._bond((ny, nx), (ny - 1, nx), (N, int(self.lu[ny, nx]))) if ny > 0 else np.zeros((N, 1))
        return Es, Esl, Esu

    def exponents(self, ny, nx, beta):
        """beta (min - E) for the site, left-bond and up-bond factors, and the bond maps d(s), r(s) = s mod leg size
        (tnac4o.py:1609-1638, 1477-1489)"""
        N = int(self.N[ny][nx])
        Es, E1, E4 = self.energy_tables(ny, nx)
        if E1.shape[1] != self.ll[ny, nx]:
            E1 = np.zeros((N, int(self.ll[ny, nx])))
        if E4.shape[1] != self.lu[ny, nx]:
            E4 = np.zeros((N, int(self.lu[ny, nx])))
        E0 = beta * (np.min(Es) - Es)
        E1 = beta * (np.min(E1) - E1)
        E4 = beta * (np.min(E4) - E4)
        s = np.arange(N)
        return E0, E1, E4, np.mod(s, self.ld[ny, nx]), np.mod(s, self.lr[ny, nx])


class HostTables:
    """Host half of the per-site constants: every small table of every site packed into ONE pinned float64 buffer and
    one uint8 buffer, so that the upload is two copies per instance (~34 MB at L = 2048)."""

    FIELDS = ('E0', 'E1', 'E4', 'Xu', 'Xl', 'Xr', 'Xd', 'Es', 'Esl', 'Esu')

    def __init__(self, lattice, beta, X):
        Xu, Xl, Xr, Xd = X
        self.sites, chunks, maps = [], [], []
        off, moff = 0, 0
        for ny in range(lattice.Ny):
            for nx in range(lattice.Nx):
                E0, E1, E4, dmap, rmap = lattice.exponents(ny, nx, beta)
                Es, Esl, Esu = lattice.energy_tables(ny, nx)
                nS, nl, nu = E0.shape[0], E1.shape[1], E4.shape[1]
                nd, nr = int(lattice.ld[ny][nx]), int(lattice.lr[ny][nx])
                parts = (E0, E1.reshape(nS, nl), E4.reshape(nS, nu), Xu[ny][nx][:nu], Xl[ny][nx][:nl], Xr[ny][nx][:nr],
                         Xd[ny][nx][:nd], Es, Esl.reshape(nS, nl), Esu.reshape(nS, nu))
                meta = {'dims': (nS, nl, nd, nr, nu), 'off': {}, 'moff': moff}
                for name, a in zip(self.FIELDS, parts):
                    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
                    meta['off'][name] = (off, a.size)
                    chunks.append(a)
                    off += a.size
                maps.append(np.ascontiguousarray(dmap, dtype=np.uint8))
                maps.append(np.ascontiguousarray(rmap, dtype=np.uint8))
                moff += 2 * nS
                self.sites.append(meta)
        pin = torch.cuda.is_available()
        self.f64 = torch.from_numpy(np.concatenate(chunks))
        self.u8 = torch.from_numpy(np.concatenate(maps))
        if pin:
            self.f64, self.u8 = self.f64.pin_memory(), self.u8.pin_memory()
        self.nbytes = self.f64.numel() * 8 + self.u8.numel()


class SiteTables:
    """Device half: views into the uploaded buffers plus the Boltzmann-weight tables built on the device
    (tn_build_site_tables), and the `tn_site` descriptor handed to the kernels."""

    def __init__(self, meta, dbuf, dmaps, device):
        from ._native import Context, check, lib
        self.nS, self.nl, self.nd, self.nr, self.nu = meta['dims']
        view = lambda name: dbuf[meta['off'][name][0]:meta['off'][name][0] + meta['off'][name][1]]
        self.dmap = dmaps[meta['moff']:meta['moff'] + self.nS]
        self.rmap = dmaps[meta['moff'] + self.nS:meta['moff'] + 2 * self.nS]
        self.Es, self.Esl, self.Esu = view('Es'), view('Esl'), view('Esu')
        f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device=device)
        self.Wlu = f64(self.nl, self.nu, self.nS)                                    # [l][u][s]
        self.WtrU = f64(self.nu, self.nl, self.nd, self.nr)                          # [u][l][d][r]
        self.Wmpo = f64(self.nl, self.nd, self.nr, self.nu)                          # (l, d, r, u) for the MPO
        c = Context.get(device)
        check(lib.tn_build_site_tables(c.handle, c.stream, self.nS, self.nl, self.nd, self.nr, self.nu,
                                       *[view(n).data_ptr() for n in ('E0', 'E1', 'E4', 'Xu', 'Xl', 'Xr', 'Xd')],
                                       self.dmap.data_ptr(), self.rmap.data_ptr(),
                                       self.Wlu.data_ptr(), self.WtrU.data_ptr(), self.Wmpo.data_ptr()))
        self.c = TnSite(self.nS, self.nl, self.nd, self.nr, self.nu, self.Wlu.data_ptr(), self.WtrU.data_ptr(),
                        self.dmap.data_ptr(), self.rmap.data_ptr(), self.Es.data_ptr(), self.Esl.data_ptr(),
                        self.Esu.data_ptr())
        self.ref = ctypes.byref(self.c)


def upload_site_tables(host, Ny, Nx, device):
    """two host-to-device copies (pinned, asynchronous on the current stream), then the device table builder per site"""
    dbuf = host.f64.to(device, non_blocking=True)
    dmaps = host.u8.to(device, non_blocking=True)
    flat = [SiteTables(meta, dbuf, dmaps, device) for meta in host.sites]
    sites = [[flat[ny * Nx + nx] for nx in range(Nx)] for ny in range(Ny)]
    return sites, (dbuf, dmaps)


def upper_triangular(J, L):
    """coupling list [i, j, Jij] -> upper-triangular float sparse matrix (tnac4o.py:176-181)"""
    ii, jj, vv = zip(*J)
    full = scipy.sparse.coo_matrix((vv, (ii, jj)), shape=(L, L))
    return (scipy.sparse.triu(full) + scipy.sparse.tril(full, -1).T).astype(dtype=float, copy=False)
