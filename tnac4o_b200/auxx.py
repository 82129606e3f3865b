"""Coupling-file helpers and the independent energy check, with the reference's names
(/root/reference/tnac4o/auxx.py).  Loaders are host glue; ``energy_Jij`` runs as an integer/CSR kernel
on the GPU instead of the reference's dense L x L products (auxx.py:82-107).
"""
import numpy as np
import scipy.sparse
import torch

from ._native import Context, check, lib, ptr


def load_Jij(file_name):
    """text file with lines ``i j Jij`` -> list of [i, j, Jij] (auxx.py:24-36)"""
    J = np.loadtxt(file_name)
    return [[int(row[0]), int(row[1]), float(row[2])] for row in J]


def round_Jij(J, dJ):
    """couplings rounded to multiples of dJ (auxx.py:39-50)"""
    dJ = float(dJ)
    return [[x[0], x[1], round(x[2] / dJ) * dJ] for x in J]


def minus_Jij(J):
    """auxx.py:53-63"""
    return [[x[0], x[1], -x[2]] for x in J]


def Jij_f2p(J):
    """1-based -> 0-based spin indices (auxx.py:66-79)"""
    return [[x[0] - 1, x[1] - 1, x[2]] for x in J]


def energy_Jij(J, states, device=None):
    """E = sum_{i<j} J_ij s_i s_j + sum_i J_ii s_i for states encoded 1 (up) / 0 (down) (auxx.py:82-107)."""
    states = np.ascontiguousarray(np.asarray(states), dtype=np.int8)
    L = states.shape[1]
    ii, jj, vv = zip(*J)
    full = scipy.sparse.coo_matrix((vv, (ii, jj)), shape=(L, L))
    JJ = scipy.sparse.coo_matrix(scipy.sparse.triu(full) + scipy.sparse.tril(full, -1).T)
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError('tnac4o_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        device = torch.device('cuda', torch.cuda.current_device())
    c = Context.get(device)
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
    ci, cj, cv = up(JJ.row, np.int32), up(JJ.col, np.int32), up(JJ.data, np.float64)
    bits = up(states, np.int8)
    E = torch.empty(states.shape[0], dtype=torch.float64, device=device)
    check(lib.tn_energy_ising(c.handle, c.stream, states.shape[0], L, ptr(bits), JJ.nnz, ptr(ci), ptr(cj), ptr(cv), ptr(E)))
    return E.cpu().numpy()


def energy_RMF(J, states):
    """cost function of a random-Markov-field model for rows of cell states (auxx.py:110-134): ``J['fac']`` maps a site
    (ny, nx) or a bond (ny1, nx1, ny2, nx2) to the index of its table in ``J['fun']``.  Host glue like the loaders (one
    table look-up per factor); the solver itself implements ``mode='Ising'`` only."""
    states = np.asarray(states)
    E = np.zeros(len(states))
    Nx = J['Nx']
    for where, f in J['fac'].items():
        table = np.asarray(J['fun'][f])
        if len(where) == 2:
            E += table[states[:, where[0] * Nx + where[1]]]
        elif len(where) == 4:
            E += table[states[:, where[0] * Nx + where[1]], states[:, where[2] * Nx + where[3]]]
    return E
