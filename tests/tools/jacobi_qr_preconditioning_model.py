"""Sweep counts of the Jacobi model after QR preconditioning of the live side (W^T = Q2 R2, then Jacobi on the nc x nc
triangular factor; optionally a second QR as in LAPACK dgejsv, or column pivoting).

Measured on the 8 centre matrices of /tmp/svdsim/C_beta3.npz: 11-14 sweeps -> 7-9 sweeps, vectors of length nc instead
of 512; a second QR or pivoting gains at most one more sweep.  """
import sys, numpy as np, scipy.linalg as sla
from jacobi_sweep_model import jacobi_sweeps, DEAD
z = np.load(sys.argv[1])
def live_side(C):
    fro2 = (C**2).sum(); rn, cn = (C**2).sum(1), (C**2).sum(0); floor = DEAD*DEAD*fro2
    use_rows = (rn > floor).sum() <= (cn > floor).sum()
    V = C if use_rows else C.T
    norms = rn if use_rows else cn
    return V[norms > floor]
def sweeps_on(Mat):
    # force orientation: orthogonalise the rows of Mat (all live)
    import sim
    m, n = Mat.shape
    W = Mat.copy()
    # reuse jacobi_sweeps by building a matrix whose rows are chosen: pad so rows side has fewer live
    sw, nc, rots, S = jacobi_sweeps(np.vstack([W, np.zeros((0, n))]) if True else W)
    return sw, nc, S
for k in z.files[:8]:
    C = z[k]; ref = np.linalg.svd(C, compute_uv=False)
    W = live_side(C)                      # nc x 512
    nc = W.shape[0]
    Q2, R2 = np.linalg.qr(W.T)            # 512 x nc, nc x nc ; W = R2^T Q2^T
    out = []
    for name, Mat in (('rows(R2)', R2), ('rows(L=R2^T)', R2.T.copy())):
        # jacobi_sweeps picks the side with fewer live vectors itself; to force rows, append nothing but report
        m, n = Mat.shape
        # force: make it "wide" by appending zero columns so rows are the short side
        Mw = np.hstack([Mat, np.zeros((m, 1))])
        sw, ncc, rots, S = jacobi_sweeps(Mw)
        err = np.abs(S - ref[:len(S)]).max()/ref[0]
        out.append('%s: nc=%d sweeps=%d err=%.1e' % (name, ncc, sw, err))
    # second QR (dgejsv style): R2^T = Q3 R3 ; jacobi on rows of R3^T? (columns of R3)
    Q3, R3 = np.linalg.qr(R2.T)
    for name, Mat in (('rows(R3)', R3), ('rows(R3^T)', R3.T.copy())):
        Mw = np.hstack([Mat, np.zeros((Mat.shape[0], 1))])
        sw, ncc, rots, S = jacobi_sweeps(Mw)
        err = np.abs(S - ref[:len(S)]).max()/ref[0]
        out.append('%s: sweeps=%d err=%.1e' % (name, sw, err))
    # column-pivoted QR of W^T
    Qp, Rp, piv = sla.qr(W.T, mode='economic', pivoting=True)
    for name, Mat in (('rows(Rp)', Rp), ('rows(Rp^T)', Rp.T.copy())):
        Mw = np.hstack([Mat, np.zeros((Mat.shape[0], 1))])
        sw, ncc, rots, S = jacobi_sweeps(Mw)
        err = np.abs(S - ref[:len(S)]).max()/ref[0]
        out.append('%s: sweeps=%d err=%.1e' % (name, sw, err))
    print(k, ' | '.join(out), flush=True)
