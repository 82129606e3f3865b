"""Oracle restatement of the solver hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/tnac4o/tnac4o.py (Ising mode).  Differences from the reference are
confined to representation, never to arithmetic order on quantities that are compared
bit-for-bit:

* the dense rank-5 PEPS tensor of ``_peps_tensor`` (tnac4o.py:1562-1607, 128 MiB per chimera
  site, 65 536 non-zeros) is kept as the compact table ``Wc[s, l, u]`` together with the
  bond-index maps ``d(s)``, ``r(s)``; the traced MPO tensor and the per-branch slices are
  assembled from it in the same summation order (adding exact zeros is exact);
* dictionaries keyed by index tuples are kept, as in the reference, for the left / right
  environments, so that the set of contractions performed is the same.
"""
import itertools
import logging

import numpy as np
import scipy.linalg
import scipy.sparse

from .mps_ref import RefMPS, ref_nfactor


def cell_bits(n):
    """(2^n, n) int8 table, bit_a(s) of state s; the reference's conf = 1 - bits (tnac4o.py:1461-1467)."""
    s = np.arange(2 ** n)[:, None]
    return ((s >> np.arange(n)[None, :]) & 1).astype(np.int8)


def cell_spins(n):
    """sigma_a(s) = 1 - 2 bit_a(s) as int8, equal to ``2 * _cluster_configurations(n) - 1``."""
    return (1 - 2 * cell_bits(n)).astype(np.int8)


def pext_table(n, positions):
    """index of the sub-configuration on ``positions`` for every cell state (tnac4o.py:1469-1487)."""
    bits = cell_bits(n).astype(np.int64)
    out = np.zeros(2 ** n, dtype=np.int64)
    for j, a in enumerate(positions):
        out += bits[:, a] << j
    return out


class RefSolver:
    """numpy oracle with the reference's object API (tnac4o.py:78-198)."""

    def __init__(self, mode='Ising', Nx=4, Ny=4, Nc=8, beta=1, J=None):
        if mode != 'Ising':
            raise NotImplementedError('oracle covers mode="Ising" only (SURVEY.md section 2 row 24)')
        if Nc > 8:
            raise ValueError('oracle restates the int8 path (Nc <= 8) only')
        self.mode, self.beta = mode, beta
        self.Nx_model, self.Ny_model = Nx, Ny
        self.Nx, self.Ny, self.Nc = Nx, Ny, Nc
        self.indtype = np.int8
        self.L = Nx * Ny * Nc
        self.order = np.arange(Nx * Ny)
        self.order_i = np.arange(Nx * Ny)
        self.logger = logging.getLogger('tnac4o.oracle')
        self.energy, self.probability = np.zeros(0), np.zeros(0)
        self.rotation, self.degeneracy = 0, 0
        self.states = np.zeros((0, Nx * Ny), dtype=self.indtype)
        self.trace = None          # optional callable(tag, **arrays) used by the parity tests
        if J is not None:
            ii, jj, vv = zip(*J)
            full = scipy.sparse.coo_matrix((vv, (ii, jj)), shape=(self.L, self.L))
            self.J = (scipy.sparse.triu(full) + scipy.sparse.tril(full, -1).T).astype(float)
            self.J0 = self.J.copy()
            self.ind0 = [[self._active(self.J, ny, nx, Nx) for nx in range(Nx)] for ny in range(Ny)]
            self._divide_couplings()

    # ------------------------------------------------------------------ model preparation
    def _active(self, J, ny, nx, Nx):
        """spins of the cell with any non-zero coupling (tnac4o.py:185-191, 1406-1413)."""
        ind = self.Nc * (Nx * ny + nx) + np.arange(self.Nc)
        weight = np.sum(np.abs(J[ind, :].toarray()), axis=1) + np.sum(np.abs(J[:, ind].toarray()), axis=0)
        return ind[np.nonzero(weight > 1e-12)]

    def _divide_couplings(self):
        """per-cell coupling blocks and leg sizes (tnac4o.py:1391-1457)."""
        Ny, Nx = self.Ny, self.Nx
        self.ind = [[self._active(self.J, ny, nx, Nx) for nx in range(Nx)] for ny in range(Ny)]
        self.sN = np.array([[len(self.ind[ny][nx]) for nx in range(Nx)] for ny in range(Ny)])
        self.N = 2 ** self.sN
        empty = np.zeros(0, dtype=int)
        self.Jin = [[None] * Nx for _ in range(Ny)]
        self.Jl = [[np.zeros((self.sN[ny][nx], 0)) for nx in range(Nx)] for ny in range(Ny)]
        self.Ju = [[np.zeros((self.sN[ny][nx], 0)) for nx in range(Nx)] for ny in range(Ny)]
        self.id = [[empty] * Nx for _ in range(Ny)]
        self.ir = [[empty] * Nx for _ in range(Ny)]
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
        self.lr, self.ld = 2 ** self.sr, 2 ** self.sd
        self._reset_X()

    def _reset_X(self):
        """identity gauges on all bonds (tnac4o.py:1811-1820)."""
        Ny, Nx = self.Ny, self.Nx
        self.Xu = np.ones((Ny, Nx, np.max(self.ld)))
        self.Xd = np.ones((Ny, Nx, np.max(self.ld)))
        self.Xl = np.ones((Ny, Nx, np.max(self.lr)))
        self.Xr = np.ones((Ny, Nx, np.max(self.lr)))
        self.overlaps_ud = np.empty(shape=[0, Ny - 1])

    def rotate_graph(self, rot=1):
        """quarter turns of the lattice (tnac4o.py:290-340)."""
        for _ in range(rot):
            self.rotation += 1
            Nx, Ny, Nc = self.Nx, self.Ny, self.Nc
            spin_map = np.arange(self.L)
            order = np.arange(Nx * Ny)
            order_i = np.arange(Nx * Ny)
            for nx in range(Nx):
                for ny in range(Ny):
                    src = (ny * Nx + nx) * Nc + np.arange(Nc)
                    dst = ((Nx - nx - 1) * Ny + ny) * Nc + np.arange(Nc)
                    spin_map[src] = dst
                    a, b = ny * Nx + nx, (Nx - nx - 1) * Ny + ny
                    order[a], order_i[b] = b, a
            self.Nx, self.Ny = Ny, Nx
            self.J = self.J[spin_map, :][:, spin_map]
            self.J = scipy.sparse.triu(self.J) + scipy.sparse.tril(self.J, -1).T
            self.order = order_i[self.order]
        self.order_i[self.order] = np.arange(self.Nx * self.Ny)
        self.rotation = np.mod(self.rotation, 4)
        self._divide_couplings()

    # ------------------------------------------------------------------ per-site tables
    def bond_down(self, s, ny, nx):
        """tnac4o.py:1469-1476."""
        return pext_table(self.sN[ny][nx], self.id[ny][nx])[s]

    def bond_right(self, s, ny, nx):
        """tnac4o.py:1480-1487."""
        return pext_table(self.sN[ny][nx], self.ir[ny][nx])[s]

    def energy_tables(self, ny, nx):
        """Es[s], Esl[s, l], Esu[s, u] with the reference's expressions (tnac4o.py:1512-1529)."""
        st = cell_spins(self.sN[ny][nx])
        Jin = self.Jin[ny][nx]
        Es = 1. * np.sum(np.dot(st, np.triu(Jin, 1)) * st, 1) + np.dot(st, Jin.diagonal())
        Esl = np.dot(np.dot(st, self.Jl[ny][nx]), cell_spins(self.sl[ny][nx]).T)
        Esu = np.dot(np.dot(st, self.Ju[ny][nx]), cell_spins(self.su[ny][nx]).T)
        return Es, Esl, Esu

    def site_weights(self, ny, nx):
        """compact PEPS data of one site: Wc[s, l, u] (gauges on all four legs folded in), d(s), r(s).

        Restates tnac4o.py:1566-1607: the three shifted energies are added first, one exp is
        taken, then the gauges are multiplied in the order Xu, Xl, Xr, Xd.
        """
        n = self.sN[ny][nx]
        L1, L4 = self.sl[ny][nx], self.su[ny][nx]
        st = cell_spins(n)
        Jin = self.Jin[ny][nx]
        Es = np.sum(np.dot(st, np.triu(Jin, 1)) * st, 1) + np.dot(st, Jin.diagonal())
        Es = self.beta * (np.min(Es) - Es)
        E1 = np.dot(np.dot(st, self.Jl[ny][nx]), cell_spins(L1).T)
        E1 = self.beta * (np.min(E1) - E1)
        E4 = np.dot(np.dot(st, self.Ju[ny][nx]), cell_spins(L4).T)
        E4 = self.beta * (np.min(E4) - E4)
        Wc = (Es[:, None, None] + E1[:, :, None]) + E4[:, None, :]
        Wc = np.exp(Wc)
        Wc = Wc * self.Xu[ny][nx][None, None, :2 ** L4]
        Wc = Wc * self.Xl[ny][nx][None, :2 ** L1, None]
        dmap = pext_table(n, self.id[ny][nx])
        rmap = pext_table(n, self.ir[ny][nx])
        Wc = Wc * self.Xr[ny][nx][rmap][:, None, None]
        Wc = Wc * self.Xd[ny][nx][dmap][:, None, None]
        return Wc, dmap, rmap

    def traced_mpo(self, ny, nx):
        """sum over the cell state of the PEPS tensor: legs (l, d, r, u) (tnac4o.py:1686)."""
        Wc, dmap, rmap = self.site_weights(ny, nx)
        W = np.zeros((Wc.shape[1], 2 ** self.sd[ny][nx], 2 ** self.sr[ny][nx], Wc.shape[2]))
        for s in range(Wc.shape[0]):            # ascending s = the order np.sum(axis=0) adds rows
            W[:, dmap[s], rmap[s], :] += Wc[s]
        return W

    # ------------------------------------------------------------------ boundary MPS
    def _setup_rhoT(self, graduate_truncation=True, Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """tnac4o.py:1674-1695."""
        self.rhoT = [None] * (self.Ny + 1)
        self.rhoT_overlap = [1] * (self.Ny + 1)
        self.rhoT_discarded = [0] * (self.Ny + 1)
        self.rhoT[-1] = RefMPS(self.Nx, d=1)
        for ny in range(self.Ny - 1, -1, -1):
            W = [self.traced_mpo(ny, nx) for nx in range(self.Nx)]
            psi = self.rhoT[ny + 1].copy()
            psi.apply_mpo(W, conj=True)
            self.rhoT_overlap[ny] = psi.compress(Dmax, tolS, tolV, max_sweeps, graduate_truncation)
            self.rhoT_discarded[ny] = max(psi.discarded)
            self.rhoT[ny] = psi

    def _setup_rhoB(self, graduate_truncation=True, Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """tnac4o.py:1697-1718."""
        self.rhoB = [None] * (self.Ny + 1)
        self.rhoB[0] = RefMPS(self.Nx, d=1)
        for ny in range(self.Ny):
            W = [self.traced_mpo(ny, nx) for nx in range(self.Nx)]
            psi = self.rhoB[ny].copy()
            psi.apply_mpo(W, conj=False)
            psi.compress(Dmax, tolS, tolV, max_sweeps, graduate_truncation)
            self.rhoB[ny + 1] = psi

    def _setup_RR(self, vind, ny):
        """right environments per unique index suffix (tnac4o.py:1768-1784)."""
        levels = [{(): np.ones((1, 1))}]
        for nx in range(self.Nx - 1, 0, -1):
            W = self.traced_mpo(ny, nx)
            new = {}
            for row in vind:
                key = tuple(row[nx + 1:])
                if key not in new:
                    T = np.tensordot(self.rhoT[ny + 1].A[nx], levels[-1][key[1:]], axes=(2, 0))
                    R = np.tensordot(T, W[:, :, :, key[0]], axes=([1, 2], [1, 2]))
                    R *= (1 / ref_nfactor(R))
                    new[key] = R
            levels.append(new)
        return levels

    @staticmethod
    def marginal(wrow, dmap, rmap, RL, AT, RR):
        """conditional probabilities of one cell for one branch (tnac4o.py:1786-1807).

        ``wrow[s]`` = Wc[s, l, u] for the branch's (l, u).  Returns (P, flag).
        """
        T1 = np.tensordot(RL, AT, axes=(0, 0))
        T2 = np.tensordot(T1, RR, axes=(1, 0))
        Pn = wrow * T2[dmap, rmap]
        low = Pn.min()
        if low < 0.:
            small = (Pn < np.abs(low))
            Pn[small] = np.abs(low)
            low *= np.sum(small)
        total = np.sum(Pn)
        if total > 0.:
            Pn *= 1. / total
            low *= 1. / total
        else:
            Pn += 1. / len(Pn)
            low = -1
        return Pn, low

    def _site_marginals(self, ny, nx, vind, RLl, RRl):
        Wc, dmap, rmap = self.site_weights(ny, nx)
        B = vind.shape[0]
        P = np.zeros((B, self.N[ny][nx]))
        flag = np.zeros(B)
        AT = self.rhoT[ny + 1].A[nx]
        for k in range(B):
            t = tuple(vind[k])
            P[k], flag[k] = self.marginal(Wc[:, t[nx], t[nx + 1]], dmap, rmap,
                                          RLl[t[:nx]], AT, RRl[self.Nx - nx - 1][t[nx + 2:]])
        return P, flag

    def _site_energy(self, states, vind_left, vind_up, ny, nx):
        """energy increment (tnac4o.py:1506-1531); left/up bond indices are read off the state rows."""
        Es, Esl, Esu = self.energy_tables(ny, nx)
        pos = ny * self.Nx + nx
        dE = Es[states[:, pos]]
        if nx > 0:
            dE += Esl[states[:, pos], self.bond_right(states[:, pos - 1], ny, nx - 1)]
        if ny > 0:
            dE += Esu[states[:, pos], self.bond_down(states[:, pos - self.Nx], ny - 1, nx)]
        return dE

    def _advance_left_env(self, RLl, vind, ny, nx):
        """tnac4o.py:528-535."""
        new = {}
        AT = self.rhoT[ny + 1].A[nx]
        for row in vind:
            key = tuple(row[:nx + 1])
            if key not in new:
                v = np.dot(RLl[key[:-1]], AT[:, key[-1], :])
                v *= (1 / ref_nfactor(v))
                new[key] = v
        return new

    # ------------------------------------------------------------------ branch and bound
    def _expand_and_cut(self, P, prob, relative_P_cutoff, pd_max):
        """log2, accumulate, relative cut-off (tnac4o.py:450-465)."""
        cand = (np.log2(P) + prob[:, None]).reshape(-1)
        order = np.arange(cand.size)
        if relative_P_cutoff > 0:
            cutoff = np.max(cand) + np.log2(relative_P_cutoff)
            keep = max(int((cand > cutoff).sum()), 1)
            if keep < cand.size:
                order = cand.argpartition(-keep - 1)
                pd_max = max(pd_max, cand[order[-keep - 1]])
                order = order[-keep:]
                cand = cand[order]
        return cand, order, pd_max

    def _merge_groups(self, vind):
        """groups of identical boundary rows: (unique rows, order, sizes) (tnac4o.py:481-485)."""
        uniq, inv = np.unique(vind, return_inverse=True, axis=0)
        inv = np.asarray(inv).reshape(-1)
        order = inv.argsort(kind='stable')
        sizes = np.bincount(inv, minlength=uniq.shape[0])
        return uniq, order, sizes

    def search_ground_state(self, M=2 ** 10, relative_P_cutoff=1e-6, min_dEng=1e-12,
                            graduate_truncation=True, Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """tnac4o.py:381-551."""
        self._setup_rhoT(graduate_truncation, Dmax, tolS, tolV, max_sweeps)
        Nx = self.Nx
        vind = np.zeros((1, Nx + 1), dtype=self.indtype)
        states = np.zeros((1, Nx * self.Ny), dtype=self.indtype)
        Eng, prob, deg = np.zeros(1), np.zeros(1), np.ones(1, dtype=int)
        pd_max, globalmin = -np.inf, 0.
        self.marginals_evaluated = 0
        for ny in range(self.Ny):
            RRl = self._setup_RR(vind, ny)
            RLl = {(): np.ones(1)}
            for nx in range(Nx):
                nst = self.N[ny][nx]
                P, flag = self._site_marginals(ny, nx, vind, RLl, RRl)
                self.marginals_evaluated += P.shape[0]
                if self.trace:
                    self.trace('marginals', ny=ny, nx=nx, vind=vind.copy(), P=P.copy(), prob=prob.copy(), Eng=Eng.copy())
                prob, order, pd_max = self._expand_and_cut(P, prob, relative_P_cutoff, pd_max)
                parent, cell = order // nst, np.mod(order, nst)
                states = states[parent]
                states[:, ny * Nx + nx] = cell
                vind = vind[parent]
                deg = deg[parent]
                vind[:, nx] = self.bond_down(cell, ny, nx)
                vind[:, nx + 1] = self.bond_right(cell, ny, nx)
                Eng = Eng[parent]
                Eng += self._site_energy(states, None, None, ny, nx)

                uniq, gorder, sizes = self._merge_groups(vind)
                G = len(sizes)
                rep = np.zeros(G, dtype=int)
                degn = np.zeros(G, dtype=int)
                probn = np.zeros(G)
                lo = 0
                for g, sz in enumerate(sizes):
                    members = gorder[lo:lo + sz]
                    lo += sz
                    Eg = Eng[members]
                    best = np.argmin(Eg)
                    rep[g] = members[best]
                    tied = members[(Eg - Eg[best] <= min_dEng)]
                    if len(tied) > 1:
                        degn[g] = sum(deg[tied])
                        probn[g] = np.mean(prob[tied])
                    else:
                        degn[g] = deg[tied[0]]
                        probn[g] = prob[tied[0]]
                vind, prob, deg = uniq, probn, degn
                states, Eng = states[rep], Eng[rep]

                if prob.size > M:
                    order = prob.argpartition(-M - 1)
                    pd_max = max(pd_max, prob[order[-M - 1]])
                    order = order[-M:]
                    vind, states, prob, Eng, deg = vind[order], states[order], prob[order], Eng[order], deg[order]
                if self.trace:
                    self.trace('branches', ny=ny, nx=nx, vind=vind.copy(), prob=prob.copy(), Eng=Eng.copy(),
                               deg=deg.copy(), states=states.copy())
                RLl = self._advance_left_env(RLl, vind, ny, nx)
                globalmin = min(globalmin, np.min(flag))
            vind[:, 1:] = vind[:, :-1]
            vind[:, 0] = 0
        self.energy = Eng
        self.degeneracy = deg[0]
        self.states = states[:, self.order]
        self.probability = prob
        self.discarded_probability = pd_max
        self.negative_probability = min(globalmin, 0)
        return Eng

    def gibbs_sampling(self, M=2 ** 10, graduate_truncation=True, Dmax=32, tolS=1e-15, tolV=1e-10,
                       max_sweeps=20, uniforms=None):
        """tnac4o.py:553-650.  ``uniforms`` (Ny*Nx, M) replaces the np.random.rand(M) draws when given."""
        self._setup_rhoT(graduate_truncation, Dmax, tolS, tolV, max_sweeps)
        Nx = self.Nx
        vind = np.zeros((M, Nx + 1), dtype=int)
        states = np.zeros((M, Nx * self.Ny), dtype=int)
        Eng = np.zeros(M)
        globalmin = 1.
        self.marginals_evaluated = 0
        for ny in range(self.Ny):
            RRl = self._setup_RR(vind, ny)
            RLl = {(): np.ones(1)}
            for nx in range(Nx):
                uniq, inv = np.unique(vind, axis=0, return_inverse=True)
                inv = np.asarray(inv).reshape(-1)
                Pu, fu = self._site_marginals(ny, nx, uniq, RLl, RRl)
                self.marginals_evaluated += uniq.shape[0]
                P, flag = Pu[inv], fu[inv]
                cdf = P.cumsum(axis=1)
                rr = np.random.rand(M) if uniforms is None else uniforms[ny * Nx + nx]
                cell = np.array([np.searchsorted(cdf[k], rr[k]) for k in range(M)], dtype=int)
                states[:, ny * Nx + nx] = cell
                vind[:, nx] = self.bond_down(cell, ny, nx)
                vind[:, nx + 1] = self.bond_right(cell, ny, nx)
                Eng += self._site_energy(states, None, None, ny, nx)
                RLl = self._advance_left_env(RLl, vind, ny, nx)
                globalmin = min(globalmin, np.min(flag))
            vind[:, 1:] = vind[:, :-1]
            vind[:, 0] = 0
        self.energy = Eng
        self.degeneracy = 0
        self.states = states[:, self.order]
        self.probability = np.zeros(1)
        self.discarded_probability = 0
        self.negative_probability = min(globalmin, 0)
        return Eng

    # ------------------------------------------------------------------ droplets (encoding 1)
    def _droplet_key(self, dpos, dstate):
        """dictionary of droplet shapes with the (first, last) semi-hash (tnac4o.py:2051-2069, 2270-2275)."""
        tag = (dpos[0], dstate[0], dpos[-1], dstate[-1])
        for k in self.invd.get(tag, []):
            if np.array_equal(dpos, self.d[k][0]) and np.array_equal(dstate, self.d[k][1]):
                return k
        k = self.free_d
        self.invd.setdefault(tag, []).append(k)
        self.d[k] = (dpos, dstate)
        self.free_d += 1
        return k

    def _prune(self, exc, budget):
        """drop sub-excitations above the energy budget, recursively (tnac4o.py:2071-2079)."""
        return (exc[0], tuple(self._prune(se, budget - se[0][0]) for se in exc[1] if se[0][0] <= budget))

    def _keys_in(self, excs):
        out = set()
        for e in excs:
            out.add(e[0][1])
            out |= self._keys_in(e[1])
        return out

    def _collect_garbage(self):
        """tnac4o.py:2249-2268."""
        live = set()
        for bel in self.el:
            live |= self._keys_in(bel)
        self.d = {k: self.d[k] for k in live}
        self.invd = {}
        for k in live:
            dpos, dstate = self.d[k]
            self.invd.setdefault((dpos[0], dstate[0], dpos[-1], dstate[-1]), []).append(k)

    # ------------------------------------------------------------------ droplets (encodings 2 and 3: adjacency graph)
    def _adj_setup(self, J, Nx, Ny, ind):
        """adjacency of the coupling graph and, per cell, the spins behind every XOR pattern (tnac4o.py:2020-2036)."""
        off = scipy.sparse.triu(J, 1) != 0
        self.adj = (off + off.T).toarray()
        self.xor2ind = []
        for ny in range(Ny):
            for nx in range(Nx):
                spins = np.asarray(ind[ny][nx])
                flipped = cell_bits(len(spins)).astype(bool)          # row x: which spins of the cell pattern x flips
                self.xor2ind.append([spins[flipped[x]] for x in range(2 ** len(spins))])

    def _shape_of(self, e):
        return self.d[e] if isinstance(e, (int, np.integer)) else e

    def _spins_of(self, shape):
        """tnac4o.py:2081-2085 (negative int8 patterns index the table from its end, i.e. as unsigned bytes)."""
        return np.hstack([self.xor2ind[pos][pat] for pos, pat in zip(*shape)])

    def _is_connected(self, shape):
        """tnac4o.py:2087-2104."""
        spins = self._spins_of(shape)
        reached, left = spins[:1], spins[1:]
        while reached.size > 0 and left.size > 0:
            hit = np.any(self.adj[reached, :][:, left], axis=0)
            reached, left = left[hit], left[~hit]
        return left.size == 0

    def _touching(self, e1, e2):
        """tnac4o.py:2122-2134."""
        a, b = self._spins_of(self._shape_of(e1)), self._spins_of(self._shape_of(e2))
        return np.any(self.adj[a, :][:, b])

    def _hd_between(self, e1, e2):
        """tnac4o.py:2157-2187 (bit counts through bin() of the signed pattern, as the reference does)."""
        (p1, s1), (p2, s2) = self._shape_of(e1), self._shape_of(e2)
        i = j = hd = 0
        while i < len(p1) and j < len(p2):
            if p1[i] == p2[j]:
                hd += bin(np.bitwise_xor(s1[i], s2[j])).count('1')
                i, j = i + 1, j + 1
            elif p1[i] < p2[j]:
                hd += bin(s1[i]).count('1')
                i += 1
            else:
                hd += bin(s2[j]).count('1')
                j += 1
        for k in range(i, len(p1)):
            hd += bin(s1[k]).count('1')
        for k in range(j, len(p2)):
            hd += bin(s2[k]).count('1')
        return hd

    def _merge_shapes(self, e1, e2):
        """XOR of two droplets on the union of their cells, cancelled cells removed (tnac4o.py:2206-2247)."""
        (p1, s1), (p2, s2) = self._shape_of(e1), self._shape_of(e2)
        pos = np.union1d(p1, p2)
        pat = np.zeros(len(pos), dtype=np.int64)
        pat[np.searchsorted(pos, p1)] = s1
        pat[np.searchsorted(pos, p2)] = np.bitwise_xor(pat[np.searchsorted(pos, p2)], np.asarray(s2, dtype=np.int64))
        keep = pat != 0
        return pos[keep].astype(np.int64), pat[keep]

    def _enumerate_adj(self, excs, max_dEng=0., max_states=np.inf, one_layer=False):
        """tnac4o.py:2337-2377: breadth-wise expansion, one popped excitation per state and pass."""
        Eng, todo, flip = [0.0], [list(excs)], [[]]
        progressed = True
        while progressed:
            progressed = False
            k = 0
            while k < len(Eng):
                if todo[k]:
                    exc = todo[k].pop()
                    if Eng[k] + exc[0][0] <= max_dEng:
                        Eng.append(Eng[k] + exc[0][0])
                        flip.append(flip[k] + [exc[0][1]])
                        rest = [x for x in todo[k] if not self._touching(x[0][1], exc[0][1])]
                        todo.append(rest)
                        if not one_layer:
                            rest.extend(list(exc[1]))
                        progressed = True
                k += 1
            if len(Eng) > max_states:
                sel = np.array(Eng).argpartition(max_states)[:max_states]
                Eng = [Eng[i] for i in sel]
                flip = [flip[i] for i in sel]
                todo = [todo[i] for i in sel]
        return np.array(Eng), flip

    def search_low_energy_spectrum(self, excitations_encoding=1, M=2 ** 10, relative_P_cutoff=1e-6,
                                   max_dEng=0., lim_hd=0, min_dEng=1e-12, graduate_truncation=True,
                                   Dmax=32, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """tnac4o.py:652-725 and the three variants 727-915 (encoding 1), 943-1133 (2), 1135-1358 (3); the search loop is
        common, the encodings differ in what is recorded when branches merge."""
        if excitations_encoding not in (1, 2, 3):
            raise NotImplementedError('Available droplets handling strategies are excitations_encoding = 1, 2, 3.')
        enc = self.excitations_encoding = excitations_encoding
        self._setup_rhoT(graduate_truncation, Dmax, tolS, tolV, max_sweeps)
        if enc > 1:
            self._adj_setup(self.J, self.Nx, self.Ny, self.ind)
        Nx = self.Nx
        vind = np.zeros((1, Nx + 1), dtype=self.indtype)
        states = np.zeros((1, Nx * self.Ny), dtype=self.indtype)
        Eng, prob, deg = np.zeros(1), np.zeros(1), np.ones(1, dtype=int)
        pd_max, globalmin = -np.inf, 1.
        self.d, self.invd, self.el, self.free_d = {}, {}, [[]], 0
        for ny in range(self.Ny):
            RRl = self._setup_RR(vind, ny)
            RLl = {(): np.ones(1)}
            for nx in range(Nx):
                nst = self.N[ny][nx]
                P, flag = self._site_marginals(ny, nx, vind, RLl, RRl)
                prob, order, pd_max = self._expand_and_cut(P, prob, relative_P_cutoff, pd_max)
                parent, cell = order // nst, np.mod(order, nst)
                states = states[parent]
                states[:, ny * Nx + nx] = cell
                vind = vind[parent]
                deg = deg[parent]
                vind[:, nx] = self.bond_down(cell, ny, nx)
                vind[:, nx + 1] = self.bond_right(cell, ny, nx)
                Eng = Eng[parent]
                Eng += self._site_energy(states, None, None, ny, nx)

                uniq, gorder, sizes = self._merge_groups(vind)
                G = len(sizes)
                rep = np.zeros(G, dtype=int)
                degn = np.zeros(G, dtype=int)
                probn, Engn = np.zeros(G), np.zeros(G)
                starts = np.concatenate(([0], np.cumsum(sizes)))
                for g in range(G):
                    members = gorder[starts[g]:starts[g + 1]]
                    Eg = Eng[members]
                    best = np.argmin(Eg)
                    rep[g], Engn[g] = members[best], Eg[best]
                    tied = members[(Eg - Engn[g] <= min_dEng)]
                    if len(tied) > 1:
                        degn[g] = sum(deg[tied])
                        probn[g] = np.mean(prob[tied])
                    else:
                        degn[g] = deg[tied[0]]
                        probn[g] = prob[tied[0]]
                if G > M:
                    keepg = probn.argpartition(-M - 1)
                    pd_max = max(pd_max, probn[keepg[-M - 1]])
                    keepg = keepg[-M:]
                else:
                    keepg = np.arange(G)

                new_el = []
                last = Nx * ny + nx
                seen_w, seen_m = [], []           # what the device path hands to its droplet book (trace for the tests)
                for g in keepg:
                    members = gorder[starts[g]:starts[g + 1]]
                    winner = rep[g]
                    bel = self.el[parent[winner]][:]
                    fresh, recs = [], []
                    for m in members:
                        gap = Eng[m] - Engn[g]
                        if gap <= max_dEng and m != winner:
                            diff = np.bitwise_xor(states[winner], states[m])
                            dpos = diff.nonzero()[0]
                            dstate = diff[dpos]
                            recs.append((parent[m], gap, dpos, dstate))
                            if enc == 1:
                                if lim_hd <= 1 or len(dstate) >= lim_hd:
                                    key = self._droplet_key(dpos, dstate)
                                    subs = [self._prune(se, max_dEng - (se[0][0] + gap))
                                            for se in self.el[parent[m]]
                                            if se[0][3] >= dpos[0] and se[0][0] + gap <= max_dEng]
                                    bel.append(((gap, key, dpos[0], last, prob[m] - probn[g]), tuple(subs)))
                            elif enc == 2:
                                # only connected differences become droplets; dependent sub-droplets follow (1071-1082)
                                if (lim_hd <= 1 or len(dstate) >= lim_hd) and self._is_connected((dpos, dstate)):
                                    key = self._droplet_key(dpos, dstate)
                                    subs = [self._prune(se, max_dEng - (se[0][0] + gap)) for se in self.el[parent[m]]
                                            if se[0][0] + gap <= max_dEng and self._touching(key, se[0][1])]
                                    bel.append(((gap, key), tuple(subs)))
                            else:
                                # one layer: the difference combined with every independent set of the droplets of the
                                # merged branch that touch it; connected combinations are stored flat (1259-1276)
                                near = [se for se in self.el[parent[m]]
                                        if se[0][0] + gap <= max_dEng and self._touching((dpos, dstate), se[0][1])]
                                sE, sflip = self._enumerate_adj(near, max_dEng - gap, one_layer=True)
                                for dE, keys in zip(sE, sflip):
                                    shape = (dpos, dstate)
                                    for k in keys:
                                        shape = self._merge_shapes(shape, k)
                                    if (lim_hd <= 1 or len(shape[1]) >= lim_hd) and self._is_connected(shape):
                                        fresh.append(((dE + gap, self._droplet_key(*shape)), ()))
                    if enc == 3:
                        bel.extend(sorted(fresh, key=lambda x: x[0][0]))
                    new_el.append(bel)
                    seen_w.append(parent[winner])
                    seen_m.append(recs)
                if self.trace:
                    self.trace('droplets', ny=ny, nx=nx, winner_parent=seen_w, merged=seen_m)
                vind, states = uniq[keepg], states[rep[keepg]]
                prob, Eng, deg = probn[keepg], Engn[keepg], degn[keepg]
                self.el = new_el
                RLl = self._advance_left_env(RLl, vind, ny, nx)
                if enc != 3:
                    self._collect_garbage()
                globalmin = min(globalmin, np.min(flag))
            if enc == 3:
                self._collect_garbage()
                if self.trace:
                    self.trace('row_end', ny=ny)
            vind[:, 1:] = vind[:, :-1]
            vind[:, 0] = 0
        if enc == 3:
            # greedy removal of near-duplicate droplets in energy order (tnac4o.py:1311-1326)
            bel = sorted(self.el[0], key=lambda x: x[0][0])
            if lim_hd > 1:
                kept = []
                for x in bel:
                    if all(self._hd_between(x[0][1], y[0][1]) >= lim_hd for y in kept):
                        kept.append(x)
                bel = kept
            self.el[0] = bel
            self._collect_garbage()
        self.energy = Eng
        self.degeneracy = deg[0]
        self.states = states[:, self.order]
        self.probability = prob
        self.discarded_probability = pd_max
        self.negative_probability = min(globalmin, 0)
        self.el = self.el[0]
        for key, (dpos, dstate) in self.d.items():
            dpos = self.order_i[dpos]
            srt = dpos.argsort()
            self.d[key] = (dpos[srt], dstate[srt])
        if enc > 1:
            self._adj_setup(self.J0, self.Nx_model, self.Ny_model, self.ind0)     # decode works in the model's orientation
        return Eng

    def _unpack(self, max_dEng, max_states):
        """enumerate droplet combinations, snake-order independence (tnac4o.py:2295-2335); encodings 2 and 3 enumerate
        through the adjacency graph (2287-2293)."""
        if self.excitations_encoding > 1:
            return self._enumerate_adj(self.el, max_dEng, max_states, one_layer=(self.excitations_encoding == 3))
        Eng, flip = [0.0], [[]]
        nsites = self.Nx_model * self.Ny_model
        stacks = [[((0, 0, -1, nsites - 1, 1), tuple(self.el))]]
        for nn in range(nsites - 1, -1, -1):
            k = 0
            while k < len(Eng):
                for ee in stacks[k][-1][1]:
                    if ee[0][3] == nn and Eng[k] + ee[0][0] <= max_dEng:
                        Eng.append(Eng[k] + ee[0][0])
                        flip.append(flip[k] + [ee[0][1]])
                        stacks.append(stacks[k] + [ee])
                    elif ee[0][3] > nn:
                        break
                k += 1
            if len(Eng) > max_states:
                sel = np.array(Eng).argpartition(max_states)[:max_states]
                Eng = [Eng[i] for i in sel]
                flip = [flip[i] for i in sel]
                stacks = [stacks[i] for i in sel]
            for k in range(len(Eng)):
                while stacks[k][-1][0][2] >= nn:
                    stacks[k].pop()
        return np.array(Eng), flip

    def decode_low_energy_states(self, max_dEng=0., max_states=1024):
        """tnac4o.py:1360-1389."""
        Eng, flip = self._unpack(max_dEng, max_states)
        ground = self.states[0]
        order = Eng.argsort()
        Eng = Eng[order]
        count = min(max_states, len(Eng))
        out = np.zeros((count, self.Nx * self.Ny), dtype=self.indtype)
        for i in range(count):
            row = ground.copy()
            for key in flip[order[i]]:
                dpos, dstate = self.d[key]
                row[dpos] = np.bitwise_xor(row[dpos], dstate)
            out[i] = row
        self.energy = Eng + self.energy[0]
        self.states = out
        return Eng[0]

    def binary_states(self, number=-1):
        """cell-state indices -> per-spin values 1 (up) / 0 (down) / 2 (inactive) (tnac4o.py:261-286)."""
        ns = self.states.shape[0]
        ns = ns + number + 1 if number < 0 else min(number, ns)
        out = np.zeros((ns, self.L), dtype=np.int8) + 2
        k = -1
        for ny in range(self.Ny_model):
            for nx in range(self.Nx_model):
                k += 1
                spins = self.ind0[ny][nx]
                out[:, spins] = (1 - cell_bits(len(spins)))[self.states[:ns, k]]
        return out

    # ------------------------------------------------------------------ preconditioning (balancing)
    def precondition(self, mode='balancing', steps=2, beta_cond=[], Dmax_cond=[], max_scale=1024,
                     graduate_truncation=False, tolS=1e-16, tolV=1e-10, max_sweeps=20):
        """tnac4o.py:342-379."""
        if mode != 'balancing':
            return
        if not beta_cond:
            beta_cond = [self.beta * 2. ** (k - steps) for k in range(steps)]
        if not Dmax_cond:
            Dmax_cond = [8] * len(beta_cond)
        target = self.beta
        for b, D in zip(beta_cond, Dmax_cond):
            self.beta = b
            self._balance_vertical(D, graduate_truncation, tolS, tolV, max_sweeps, max_scale)
        self.beta = target

    def _balance_vertical(self, Dmax, graduate, tolS, tolV, max_sweeps, max_scale):
        """tnac4o.py:1824-1918 (direction 'ud')."""
        cap = ref_nfactor(np.sqrt(max_scale))
        self._setup_rhoT(graduate, Dmax, tolS, tolV, max_sweeps)
        self._setup_rhoB(graduate, Dmax, tolS, tolV, max_sweeps)
        Nx = self.Nx
        overlaps = np.ones((2, self.Ny - 1))
        for ny in range(1, self.Ny):
            bot, top = self.rhoB[ny], self.rhoT[ny]
            for nx in range(Nx):
                bot.push_left_env(top, nx)
                bot.R[nx + 1] *= (1 / np.linalg.norm(bot.R[nx + 1]))

            def rebalance(nx):
                env = bot.bond_env(top, nx)
                _, scale = scipy.linalg.matrix_balance(env, permute=False, separate=True)
                scale = np.minimum(np.maximum(scale[0], 1 / cap), cap)
                o1 = bot.site_overlap(top, nx)
                o1 *= 1 / (np.linalg.norm(bot.A[nx]) * np.linalg.norm(top.A[nx]))
                bot.apply_diagonal(scale, nx)
                top.apply_diagonal(1 / scale, nx)
                nb, nt = np.linalg.norm(bot.A[nx]), np.linalg.norm(top.A[nx])
                o2 = bot.site_overlap(top, nx)
                o2 *= 1 / (nb * nt)
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
                    bot.absorb_right()
                    top.orth_right(nx)
                    top.absorb_right()
                    bot.push_right_env(top, nx)
                    bot.R[nx] *= (1 / np.linalg.norm(bot.R[nx]))
            for nx in range(Nx):
                rebalance(nx)
                if nx < Nx - 1:
                    bot.orth_left(nx)
                    bot.absorb_left()
                    top.orth_left(nx)
                    top.absorb_left()
                    bot.push_left_env(top, nx)
                    bot.R[nx + 1] *= (1 / np.linalg.norm(bot.R[nx + 1]))
        self.overlaps_ud = np.vstack([self.overlaps_ud, overlaps])
        self.rhoB = []
