"""Oracle restatement of the coupling helpers and the independent energy check
(TEST INFRASTRUCTURE, see oracle/__init__.py).  Follows /root/reference/tnac4o/auxx.py.
"""
import numpy as np
import scipy.sparse


def shift_to_zero_based(J):
    """auxx.py:66-79."""
    return [[i - 1, j - 1, v] for i, j, v in J]


def round_couplings(J, dJ):
    """auxx.py:39-50: v -> round(v / dJ) * dJ with Python's round()."""
    dJ = float(dJ)
    return [[i, j, round(v / dJ) * dJ] for i, j, v in J]


def negate_couplings(J):
    """auxx.py:53-63."""
    return [[i, j, -v] for i, j, v in J]


def _upper(J, L):
    ii, jj, vv = zip(*J)
    full = scipy.sparse.coo_matrix((vv, (ii, jj)), shape=(L, L))
    return scipy.sparse.triu(full) + scipy.sparse.tril(full, -1).T


def energy_ising_dense(J, states):
    """E = s^T triu(J,1) s + diag(J).s for 0/1 encoded states, dense products in chunks of 1024 (auxx.py:82-107)."""
    L = len(states[0])
    JJ = _upper(J, L)
    st = 2 * np.array(states) - 1
    out = np.zeros(st.shape[0], dtype=float)
    offdiag = scipy.sparse.triu(JJ, 1).toarray()
    for lo in range(0, st.shape[0], 1024):
        blk = st[lo:lo + 1024]
        out[lo:lo + 1024] = np.sum(np.dot(blk, offdiag) * blk, 1) + np.dot(blk, JJ.diagonal())
    return out


def energy_ising_sparse(J, states):
    """Same quantity evaluated coupling by coupling in coordinate order (what the integer/CSR kernel does)."""
    L = len(states[0])
    JJ = scipy.sparse.coo_matrix(_upper(J, L))
    st = (2 * np.array(states) - 1).astype(np.float64)
    out = np.zeros(st.shape[0])
    for i, j, v in zip(JJ.row, JJ.col, JJ.data):
        out += v * st[:, i] * (st[:, j] if i != j else 1.0)
    return out
