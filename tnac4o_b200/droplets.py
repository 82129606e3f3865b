"""Adjacency-based droplet bookkeeping: ``excitations_encoding = 2`` and ``3`` of the reference
(tnac4o.py:943-1358 for the two search variants, 2020-2040 / 2087-2247 / 2337-2377 for the helpers).

Host-side structure only (dictionary of droplet shapes + excitation lists in the reference's own save format); the
search itself -- marginals, selection, merge groups, XOR differences of merged branches -- runs on the device and hands
this class one record per merged-away branch.  Encoding 1 (snake-order independence) lives in solver.py.

Format (what ``save`` writes): ``d[key] = (dpos, dstate)`` with ``dpos`` the ascending cell positions a droplet
touches and ``dstate`` the XOR pattern of the cell state there; an excitation is ``((dE, key), sub_excitations)``.
Two droplets are independent when no spin of one is coupled to a spin of the other (adjacency of the coupling graph).
"""
import numpy as np
import scipy.sparse


class AdjacencyDroplets:
    def __init__(self, encoding, mode='Ising'):
        if encoding not in (2, 3):
            raise ValueError('AdjacencyDroplets handles excitations_encoding 2 and 3')
        self.encoding = encoding
        self.mode = mode
        self.grid_Nx = None
        self.d, self.invd, self.el, self.free_d = {}, {}, [[]], 0
        self.adj = None
        self.cell_spins = None
        self._spin_cache = {}

    # ------------------------------------------------------------------ geometry
    def set_adjacency(self, J, cells):
        """J: upper-triangular sparse couplings of the lattice as currently oriented; cells: per cell (in the order the
        states are stored) the ascending global indices of its active spins (tnac4o.py:2020-2036)."""
        up = scipy.sparse.triu(J, 1) != 0
        self.adj = np.asarray((up + up.T).toarray(), dtype=bool)
        self.cell_spins = [np.asarray(c, dtype=np.int64) for c in cells]
        self._spin_cache = {}

    def set_grid(self, Nx, Ny):
        """mode='RMF': droplets live on the sites of an Ny x Nx nearest-neighbour grid (tnac4o.py:2038-2041)"""
        self.grid_Nx, self.grid_Ny = Nx, Ny
        self.adj = np.zeros((0, 0))

    def _grid_distance(self, a, b):
        """Manhattan distances between two sets of sites of the grid (tnac4o.py:2103-2109, 2136-2142)"""
        a, b = np.asarray(a), np.asarray(b)
        ax, ay, bx, by = a % self.grid_Nx, a // self.grid_Nx, b % self.grid_Nx, b // self.grid_Nx
        return np.abs(ax[:, None] - bx[None, :]) + np.abs(ay[:, None] - by[None, :])

    def spins(self, dpos, dstate):
        """global indices of the spins a droplet flips: the set bits of every XOR pattern, cell by cell"""
        if self.mode == 'RMF':
            return np.asarray(dpos, dtype=np.int64)
        parts = []
        for cell, pattern in zip(dpos, dstate):
            key = (int(cell), int(pattern) & 0xFF)
            hit = self._spin_cache.get(key)
            if hit is None:
                ids = self.cell_spins[key[0]]
                hit = ids[[(key[1] >> a) & 1 == 1 for a in range(len(ids))]]
                self._spin_cache[key] = hit
            parts.append(hit)
        return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)

    def _shape(self, e):
        return self.d[e] if isinstance(e, (int, np.integer)) else e

    def connected(self, dpos, dstate):
        """is the droplet one connected piece of the coupling graph? (tnac4o.py:2093-2104)"""
        todo = self.spins(dpos, dstate)
        front, rest = todo[:1], todo[1:]
        while front.size and rest.size:
            if self.mode == 'RMF':
                touched = (self._grid_distance(front, rest) == 1).any(axis=0)
            else:
                touched = self.adj[np.ix_(front, rest)].any(axis=0)
            front, rest = rest[touched], rest[~touched]
        return rest.size == 0

    def overlap(self, e1, e2):
        """does any spin of one droplet couple to a spin of the other? (tnac4o.py:2129-2134)"""
        a, b = self._shape(e1), self._shape(e2)
        if self.mode == 'RMF':
            return bool((self._grid_distance(a[0], b[0]) <= 1).any())
        return bool(self.adj[np.ix_(self.spins(*a), self.spins(*b))].any())

    def hamming(self, dstate):
        """size of a droplet as the reference counts it (tnac4o.py:2143-2150): touched cells in Ising mode, set bits of the
        XOR patterns in RMF mode"""
        if self.mode == 'RMF':
            return sum(self._ones(v) for v in dstate)
        return len(dstate)

    @staticmethod
    def _ones(v):
        """the reference counts bits with ``bin(x).count('1')`` on the *signed* pattern (tnac4o.py:2172-2186), i.e. the
        ones of |x|; kept as is so that the greedy removal of near-duplicates selects the same droplets"""
        return bin(int(v)).count('1')

    def hamming_between(self, e1, e2):
        """number of spins in which two droplets differ, as the reference counts it (tnac4o.py:2170-2187)"""
        (p1, s1), (p2, s2) = self._shape(e1), self._shape(e2)
        n1, n2, hd = 0, 0, 0
        if self.mode == 'RMF':          # sites on which the two patterns differ (tnac4o.py:2179-2195)
            both = {int(p): int(v) for p, v in zip(p1, s1)}
            other = {int(p): int(v) for p, v in zip(p2, s2)}
            return sum(1 for p in set(both) | set(other) if both.get(p) != other.get(p))
        while n1 < len(p1) and n2 < len(p2):
            if p1[n1] == p2[n2]:
                hd += self._ones(int(s1[n1]) ^ int(s2[n2]))
                n1, n2 = n1 + 1, n2 + 1
            elif p1[n1] < p2[n2]:
                hd += self._ones(s1[n1])
                n1 += 1
            else:
                hd += self._ones(s2[n2])
                n2 += 1
        hd += sum(self._ones(v) for v in s1[n1:]) + sum(self._ones(v) for v in s2[n2:])
        return hd

    @staticmethod
    def combine(a, b):
        """symmetric difference of two droplets: XOR of the patterns, cells that cancel are dropped (tnac4o.py:2206-2247)"""
        pat = {}
        for p, s in zip(*a):
            pat[int(p)] = int(s)
        for p, s in zip(*b):
            v = pat.get(int(p), 0) ^ int(s)
            if v == 0 and int(p) in pat:
                del pat[int(p)]
            elif v != 0:
                pat[int(p)] = v
        pos = np.array(sorted(pat), dtype=np.int64)
        return pos, np.array([pat[int(p)] for p in pos], dtype=np.int64)

    # ------------------------------------------------------------------ dictionary of shapes
    @staticmethod
    def _tag(dpos, dstate):
        return (dpos[0], dstate[0], dpos[-1], dstate[-1])

    def key_of(self, dpos, dstate):
        """key of a shape, adding it when new (tnac4o.py:2051-2069)"""
        tag = self._tag(dpos, dstate)
        for k in self.invd.get(tag, []):
            if np.array_equal(dpos, self.d[k][0]) and np.array_equal(dstate, self.d[k][1]):
                return k
        k = self.free_d
        self.invd.setdefault(tag, []).append(k)
        self.d[k] = (dpos, dstate)
        self.free_d += 1
        return k

    def _keys(self, excs, out):
        for e in excs:
            out.add(e[0][1])
            self._keys(e[1], out)
        return out

    def drop_unused_shapes(self):
        """tnac4o.py:2249-2268"""
        live = set()
        for bel in self.el:
            self._keys(bel, live)
        self.d = {k: self.d[k] for k in live}
        self.invd = {}
        for k in live:
            self.invd.setdefault(self._tag(*self.d[k]), []).append(k)

    def trim(self, exc, budget):
        """drop sub-excitations above the energy budget, recursively (tnac4o.py:2071-2079)"""
        return (exc[0], tuple(self.trim(se, budget - se[0][0]) for se in exc[1] if se[0][0] <= budget))

    # ------------------------------------------------------------------ enumeration
    def enumerate(self, excs, max_dEng=0., max_states=np.inf, one_layer=False):
        """all combinations of mutually independent droplets below max_dEng (tnac4o.py:2337-2377).  The sweep structure
        (one pop per state and pass; a pass that accepts nothing ends the enumeration) is the reference's."""
        Eng, flips, pending = [0.0], [[]], [list(excs)]
        again = True
        while again:
            again = False
            k = 0
            while k < len(Eng):
                if pending[k]:
                    exc = pending[k].pop()
                    if Eng[k] + exc[0][0] <= max_dEng:
                        Eng.append(Eng[k] + exc[0][0])
                        flips.append(flips[k] + [exc[0][1]])
                        free = [x for x in pending[k] if not self.overlap(x[0][1], exc[0][1])]
                        if not one_layer:
                            free.extend(exc[1])
                        pending.append(free)
                        again = True
                k += 1
            if len(Eng) > max_states:
                keep = np.array(Eng).argpartition(max_states)[:max_states]
                Eng = [Eng[i] for i in keep]
                flips = [flips[i] for i in keep]
                pending = [pending[i] for i in keep]
        return np.array(Eng), flips

    def unpack(self, max_dEng=0., max_states=np.inf):
        return self.enumerate(self.el, max_dEng, max_states, one_layer=(self.encoding == 3))

    # ------------------------------------------------------------------ one site of the search
    def site_update(self, winner_parent, merged, max_dEng, lim_hd):
        """New excitation lists after one site.  Branch j of the new list descends from old branch winner_parent[j];
        merged[j] lists the branches merged into it as (old branch, dE above the winner, dpos, dstate) in group order
        (tnac4o.py:1063-1090 for encoding 2, 1251-1282 for encoding 3)."""
        new_el = []
        for j, wp in enumerate(winner_parent):
            bel = list(self.el[wp])
            if self.encoding == 2:
                for old, gap, dpos, dstate in merged[j]:
                    if (lim_hd <= 1 or self.hamming(dstate) >= lim_hd) and self.connected(dpos, dstate):
                        key = self.key_of(dpos, dstate)
                        subs = [self.trim(se, max_dEng - (se[0][0] + gap)) for se in self.el[old]
                                if se[0][0] + gap <= max_dEng and self.overlap(key, se[0][1])]
                        bel.append(((gap, key), tuple(subs)))
            else:
                fresh = []
                for old, gap, dpos, dstate in merged[j]:
                    near = [se for se in self.el[old]
                            if se[0][0] + gap <= max_dEng and self.overlap((dpos, dstate), se[0][1])]
                    sub_E, sub_flip = self.enumerate(near, max_dEng - gap, one_layer=True)
                    for dE, keys in zip(sub_E, sub_flip):
                        shape = (dpos, dstate)
                        for k in keys:
                            shape = self.combine(shape, self.d[k])
                        if (lim_hd <= 1 or self.hamming(shape[1]) >= lim_hd) and self.connected(*shape):
                            fresh.append(((dE + gap, self.key_of(*shape)), ()))
                fresh.sort(key=lambda x: x[0][0])
                bel.extend(fresh)
            new_el.append(bel)
        self.el = new_el

    def end_of_site(self):
        if self.encoding == 2:
            self.drop_unused_shapes()

    def end_of_row(self):
        if self.encoding == 3:
            self.drop_unused_shapes()

    def finish(self, order_i, lim_hd):
        """close the search: encoding 3 removes near-duplicates greedily in energy order (tnac4o.py:1311-1326); the shapes
        are mapped back to the model's cell order (1337-1346)"""
        if self.encoding == 3:
            bel = sorted(self.el[0], key=lambda x: x[0][0])
            if lim_hd > 1:
                kept = []
                for x in bel:
                    if all(self.hamming_between(x[0][1], y[0][1]) >= lim_hd for y in kept):
                        kept.append(x)
                bel = kept
            self.el[0] = bel
            self.drop_unused_shapes()
        self.el = self.el[0]
        for key, (dpos, dstate) in self.d.items():
            dpos = np.asarray(order_i)[dpos]
            srt = dpos.argsort()
            self.d[key] = (dpos[srt], dstate[srt])
