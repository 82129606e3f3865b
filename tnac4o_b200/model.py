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


class SiteTables:
    """Device copies of one site's constants plus the `tn_site` descriptor handed to the kernels.  Only the small
    exponent / energy tables cross PCIe (~130 KiB per chimera site); the 1.5 MiB of Boltzmann-weight tables are
    built on the device (tn_build_site_tables)."""

    def __init__(self, lattice, ny, nx, beta, X, device):
        from ._native import Context, check, lib
        Xu, Xl, Xr, Xd = X
        E0, E1, E4, dmap, rmap = lattice.exponents(ny, nx, beta)
        self.nS = E0.shape[0]
        self.nl, self.nu = E1.shape[1], E4.shape[1]
        self.nd, self.nr = int(2 ** lattice.sd[ny][nx]), int(2 ** lattice.sr[ny][nx])
        Es, Esl, Esu = lattice.energy_tables(ny, nx)
        self.host_dmap, self.host_rmap = dmap, rmap
        dev = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
        f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device=device)
        self.dmap = dev(dmap, np.uint8)
        self.rmap = dev(rmap, np.uint8)
        self.Es, self.Esl, self.Esu = dev(Es), dev(Esl.reshape(self.nS, self.nl)), dev(Esu.reshape(self.nS, self.nu))
        self.Wlu = f64(self.nl, self.nu, self.nS)                                    # [l][u][s]
        self.WtrU = f64(self.nu, self.nl, self.nd, self.nr)                          # [u][l][d][r]
        self.Wmpo = f64(self.nl, self.nd, self.nr, self.nu)                          # (l, d, r, u) for the MPO
        small = [dev(E0), dev(E1.reshape(self.nS, self.nl)), dev(E4.reshape(self.nS, self.nu)), dev(Xu[ny][nx][:self.nu]),
                 dev(Xl[ny][nx][:self.nl]), dev(Xr[ny][nx][:self.nr]), dev(Xd[ny][nx][:self.nd])]
        c = Context.get(device)
        check(lib.tn_build_site_tables(c.handle, c.stream, self.nS, self.nl, self.nd, self.nr, self.nu,
                                       *[t.data_ptr() for t in small], self.dmap.data_ptr(), self.rmap.data_ptr(),
                                       self.Wlu.data_ptr(), self.WtrU.data_ptr(), self.Wmpo.data_ptr()))
        self._keep = small           # alive until the builder kernels have run (same stream as every later use)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in small + [self.dmap, self.rmap, self.Es, self.Esl, self.Esu])
        self.c = TnSite(self.nS, self.nl, self.nd, self.nr, self.nu, self.Wlu.data_ptr(), self.WtrU.data_ptr(),
                        self.dmap.data_ptr(), self.rmap.data_ptr(), self.Es.data_ptr(), self.Esl.data_ptr(),
                        self.Esu.data_ptr())
        self.ref = ctypes.byref(self.c)


def upper_triangular(J, L):
    """coupling list [i, j, Jij] -> upper-triangular float sparse matrix (tnac4o.py:176-181)"""
    ii, jj, vv = zip(*J)
    full = scipy.sparse.coo_matrix((vv, (ii, jj)), shape=(L, L))
    return (scipy.sparse.triu(full) + scipy.sparse.tril(full, -1).T).astype(dtype=float, copy=False)
