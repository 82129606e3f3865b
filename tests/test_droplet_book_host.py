"""CPU test of the data-structure claim behind csrc/droplet_book.cu: keeping, per child entry, the energy budget it was
pruned with and composing nested prunings lazily as min(stored budget, outer budget - dE) gives exactly the tree the
reference builds by pruning eagerly with _exc_cut_energy (tnac4o.py:2071-2079) every time a branch is merged away
(tnac4o.py:863-875).  Random merge histories are replayed both ways: the reference's way with nested tuples, and the
device's way with flat node / children pools that solver._materialise_pools expands at the end."""
import numpy as np

from tnac4o_b200.solver import tnac4o


def cut(exc, maxdE):
    """the reference's recursion (tnac4o.py:2071-2079)"""
    return (exc[0], tuple(cut(se, maxdE - se[0][0]) for se in exc[1] if se[0][0] <= maxdE))


def replay(seed, nsites=12, nbranch=6, max_dE=2.0):
    rng = np.random.default_rng(seed)
    # eager: el[b] = list of nested tuples; lazy: pools + lists of node ids
    el = [[] for _ in range(nbranch)]
    dE, dP, key, first, last, cptr, ccnt, cnode, cbud = [], [], [], [], [], [], [], [], []
    lists = [[] for _ in range(nbranch)]
    for site in range(nsites):
        new_el, new_lists = [], []
        for j in range(nbranch):
            winner = int(rng.integers(nbranch))
            bel, blist = el[winner][:], lists[winner][:]
            for _ in range(int(rng.integers(0, 3))):              # branches merged into the winner at this site
                loser = int(rng.integers(nbranch))
                gap = float(rng.choice([0.0, 0.25, 0.5, 0.75, 1.0, 1.5, 2.5]))
                if gap > max_dE:
                    continue
                dfirst = int(rng.integers(0, site + 1))
                k = int(rng.integers(1000))
                dlogp = float(rng.standard_normal())
                # reference: tnac4o.py:866-875
                subs = [cut(sne, max_dE - (sne[0][0] + gap)) for sne in el[loser]
                        if sne[0][3] >= dfirst and sne[0][0] + gap <= max_dE]
                bel.append(((gap, k, dfirst, site, dlogp), tuple(subs)))
                # device: book_child_count / book_node_fill
                node = len(dE)
                dE.append(gap); dP.append(dlogp); key.append(k); first.append(dfirst); last.append(site)
                cptr.append(len(cnode))
                n = 0
                for nd in lists[loser]:
                    if last[nd] >= dfirst and dE[nd] + gap <= max_dE:
                        cnode.append(nd); cbud.append(max_dE - (dE[nd] + gap)); n += 1
                ccnt.append(n)
                blist.append(node)
            new_el.append(bel); new_lists.append(blist)
        el, lists = new_el, new_lists
    arr = lambda a, t: np.asarray(a if a else [0], dtype=t)
    pools = (arr(dE, np.float64), arr(dP, np.float64), arr(key, np.int32), arr(first, np.int32), arr(last, np.int32),
             arr(cptr, np.int32), arr(ccnt, np.int32), arr(cnode, np.int32), arr(cbud, np.float64))
    return el, lists, pools


def test_lazy_budgets_equal_eager_pruning():
    grew = 0
    for seed in range(40):
        el, lists, pools = replay(seed)
        for b in range(len(el)):
            got, used = tnac4o._materialise_pools(*pools, np.asarray(lists[b], dtype=np.int32))
            assert got == el[b], (seed, b)
            grew += sum(len(e[1]) for e in el[b])

            def keys(excs):
                out = set()
                for e in excs:
                    out.add(e[0][1]); out |= keys(e[1])
                return out
            assert used == keys(el[b])
    assert grew > 100                                          # the histories do build nested trees
