"""numpy model of the cluster Jacobi kernel (csrc/svd.cu: same orientation rule, dead-vector floor, round-robin tournament and
rotation threshold) used to count sweeps on the real 512 x 512 centre matrices of an L=2048 boundary-MPS build.

    python tests/tools/dump_centre_matrices.py 3        # oracle run, writes /tmp/svdsim/C_beta3.npz (build container only)
    python tests/tools/jacobi_sweep_model.py /tmp/svdsim/C_beta3.npz 8

Measured (droplet instance 001, Dmax=32): beta=3: 100-234 live vectors, 11-14 sweeps, initial ordering by norm changes
nothing; beta=1: 512 live vectors, 10-12 sweeps.  Test / design infrastructure only -- the product never imports it."""
import sys, numpy as np
DEAD = 2.2e-18
def jacobi_sweeps(C, order='natural', max_sweeps=60):
    m, n = C.shape
    fro2 = (C**2).sum()
    rn, cn = (C**2).sum(1), (C**2).sum(0)
    floor = DEAD*DEAD*fro2
    use_rows = (rn > floor).sum() <= (cn > floor).sum()
    V = C if use_rows else C.T
    norms = rn if use_rows else cn
    live = np.where(norms > floor)[0]
    if order == 'desc': live = live[np.argsort(-norms[live], kind='stable')]
    elif order == 'asc': live = live[np.argsort(norms[live], kind='stable')]
    W = V[live].copy()
    nc, a = W.shape
    tol = np.sqrt(a)*2.220446049250313e-16
    nce = nc + (nc & 1); r1 = nce-1; half = nce//2
    floor2 = floor
    rots = []
    for sweep in range(max_sweeps):
        nrot = 0
        for r in range(r1):
            i = np.arange(half)
            p = (r+i) % r1
            q = np.where(i == 0, r1, (r+r1-i) % r1)
            ok = (p < nc) & (q < nc)
            p, q = p[ok], q[ok]
            xp, xq = W[p], W[q]
            app = (xp*xp).sum(1); aqq = (xq*xq).sum(1); apq = (xp*xq).sum(1)
            do = (np.abs(apq) > tol*np.sqrt(app)*np.sqrt(aqq)) & (app > floor2) & (aqq > floor2)
            if do.any():
                zeta = (aqq[do]-app[do])/(2*apq[do])
                t = np.copysign(1.0, zeta)/(np.abs(zeta)+np.sqrt(1+zeta*zeta))
                cs = 1/np.sqrt(1+t*t); sn = cs*t
                u, v = xp[do], xq[do]
                W[p[do]] = cs[:,None]*u - sn[:,None]*v
                W[q[do]] = sn[:,None]*u + cs[:,None]*v
                nrot += int(do.sum())
        rots.append(nrot)
        if nrot == 0: break
    S = np.sort(np.sqrt((W**2).sum(1)))[::-1]
    return len(rots), nc, rots, S
if __name__ == '__main__':
    z = np.load(sys.argv[1])
    tot = {}
    for k in z.files[:int(sys.argv[2]) if len(sys.argv) > 2 else 10]:
        C = z[k]
        ref = np.linalg.svd(C, compute_uv=False)
        line = []
        for order in ('natural', 'desc', 'asc'):
            sw, nc, rots, S = jacobi_sweeps(C, order)
            err = np.abs(S[:len(S)] - ref[:len(S)]).max()/ref[0]
            line.append('%s: nc=%d sweeps=%d err=%.1e' % (order, nc, sw, err))
            tot[order] = tot.get(order, 0) + sw*nc
        print(k, ' | '.join(line), flush=True)
    print(tot)
