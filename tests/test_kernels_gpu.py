"""GPU parity tests of the individual sm_100a kernels against numpy / the oracle (called through the C ABI)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def up(a, dtype=None):
    return torch.from_numpy(np.ascontiguousarray(a if dtype is None else a.astype(dtype))).to(dev())


@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (7, 5, 3), (64, 64, 16), (130, 70, 33), (512, 512, 512), (1024, 257, 96),
                                   (32, 480, 8192), (16, 16, 4096), (2048, 128, 512), (57, 59, 39),
                                   # >= 148 tiles of 128 x 128: the TMA-staged kernel (gemm_tma.cu) for the untransposed case,
                                   # with ragged M / N / K edges that the hardware zero-fills
                                   (8192, 512, 512), (4000, 650, 100), (2500, 1300, 78)])
@pytest.mark.parametrize('tA,tB', [(False, False), (True, False), (False, True), (True, True)])
def test_gemm(M, N, K, tA, tB):
    from tnac4o_b200 import ops
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if tA else (M, K))
    B = rng.standard_normal((N, K) if tB else (K, N))
    C = ops.gemm(up(A), up(B), transA=tA, transB=tB).cpu().numpy()
    ref = (A.T if tA else A) @ (B.T if tB else B)
    assert np.max(np.abs(C - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))) * np.sqrt(K)


def test_gemm_alpha_beta_and_strided_views():
    from tnac4o_b200 import ops
    rng = np.random.default_rng(5)
    A, B, C0 = rng.standard_normal((100, 40)), rng.standard_normal((40, 90)), rng.standard_normal((100, 64))
    out = up(C0)
    ops.gemm(up(A), up(B)[:, :64], out=out, alpha=-1.0, beta=1.0)
    assert np.max(np.abs(out.cpu().numpy() - (C0 - A @ B[:, :64]))) < 1e-12


@pytest.mark.parametrize('m,n', [(1, 1), (5, 3), (33, 70), (300, 7)])
def test_transpose(m, n):
    from tnac4o_b200 import ops
    A = np.random.default_rng(0).standard_normal((m, n))
    assert np.array_equal(ops.transpose(up(A)).cpu().numpy(), A.T)


def _check_qr(A, tol=1e-13):
    from tnac4o_b200 import ops
    from oracle import ref_qr
    m, n = A.shape
    k = min(m, n)
    Q, R, bits = ops.qr_pos(up(A).clone())
    Q, R = Q.cpu().numpy(), R.cpu().numpy()
    scale = max(1.0, np.max(np.abs(A)))
    assert Q.shape == (m, k) and R.shape == (k, n)
    assert np.all(np.diag(R) >= 0)
    assert np.max(np.abs(np.tril(R, -1))) == 0
    assert np.max(np.abs(Q.T @ Q - np.eye(k))) < tol * 10 * np.sqrt(m)
    assert np.max(np.abs(Q @ R - A)) < tol * 10 * scale * np.sqrt(m)
    assert np.float64(np.max(np.abs(R))).view(np.int64) == int(bits.item())
    return Q, R


@pytest.mark.parametrize('m,n', [(1, 1), (2, 1), (16, 16), (16, 512), (256, 512), (40, 12), (512, 32), (1024, 64),
                                 (2048, 128), (2048, 512), (8192, 512), (100, 57), (17, 5), (129, 130)])
def test_qr_random(m, n):
    from oracle import ref_qr
    A = np.random.default_rng(m + n).standard_normal((m, n))
    Q, R = _check_qr(A)
    Qr, Rr = ref_qr(A)
    assert np.max(np.abs(R - Rr)) < 1e-10 * np.sqrt(m)       # full column rank: QR with diag(R) > 0 is unique


def test_qr_graded_and_rank_deficient():
    rng = np.random.default_rng(3)
    U, _ = np.linalg.qr(rng.standard_normal((600, 64)))
    V, _ = np.linalg.qr(rng.standard_normal((64, 64)))
    A = (U * np.logspace(0, -17, 64)) @ V.T                  # singular values over 17 decades, like the boundary MPS
    _check_qr(A)
    B = np.zeros((300, 40))
    B[:, :10] = rng.standard_normal((300, 10))
    B[:, 20:30] = B[:, :10] @ rng.standard_normal((10, 10))   # exactly dependent and exactly zero columns
    _check_qr(B)
    _check_qr(np.zeros((50, 8)))


def _check_svd(C, rel_sv_tol=1e-10):
    from tnac4o_b200 import ops
    m, n = C.shape
    k = min(m, n)
    U, S, Vt = ops.svd(up(C), want_vectors=True)
    U, S, Vt = U.cpu().numpy(), S.cpu().numpy(), Vt.cpu().numpy()
    S_only = ops.svd(up(C), want_vectors=False).cpu().numpy()
    Sr = np.linalg.svd(C, compute_uv=False)
    nrm = max(Sr[0], 1e-300)
    assert np.all(np.diff(S) <= 0)
    assert np.max(np.abs(S - Sr)) < 1e-13 * nrm * np.sqrt(max(m, n))
    assert np.max(np.abs(S_only - Sr)) < 1e-13 * nrm * np.sqrt(max(m, n))
    assert np.max(np.abs((U * S) @ Vt - C)) < 1e-13 * nrm * max(m, n)
    live = S > 1e-16 * max(S[0], 1e-300)       # triplets below 1e-2 eps ||C|| are deflated by design (svd.cu)
    assert np.max(np.abs((U.T @ U - np.eye(k))[np.ix_(live, live)])) < 1e-12 * np.sqrt(max(m, n))
    assert np.max(np.abs((Vt @ Vt.T - np.eye(k))[np.ix_(live, live)])) < 1e-12 * np.sqrt(max(m, n))
    return S


@pytest.mark.parametrize('m,n', [(1, 1), (2, 2), (3, 5), (16, 16), (32, 32), (64, 64), (16, 256), (16, 512), (57, 57),
                                 (128, 128), (256, 256), (512, 512), (39, 64)])
def test_svd_random(m, n):
    _check_svd(np.random.default_rng(m * 3 + n).standard_normal((m, n)))


@pytest.mark.parametrize('k', [32, 128, 512])
def test_svd_graded_spectrum_relative_accuracy(k):
    """singular values spanning 16 decades (what truncateC sees, mps.py:804-806): small ones must stay accurate"""
    rng = np.random.default_rng(k)
    U, _ = np.linalg.qr(rng.standard_normal((k, k)))
    V, _ = np.linalg.qr(rng.standard_normal((k, k)))
    s = np.logspace(0, -16, k)
    C = np.triu((U * s) @ V.T @ np.diag(np.logspace(0, -3, k)))      # graded upper-triangular, like an R factor
    S = _check_svd(C)
    exact = np.linalg.svd(C.astype(np.longdouble).astype(np.float64), compute_uv=False)
    big = exact > 1e-9 * exact[0]
    assert np.max(np.abs(S[big] / exact[big] - 1)) < 1e-6


def test_truncation_rank_and_nfactor():
    from tnac4o_b200 import ops
    from oracle import ref_nfactor
    S = np.array([1.0, 0.5, 1e-3, 1e-9, 3e-16, 1e-17, 0.0])
    for tol, Dmax in ((1e-16, 32), (1e-8, 32), (1e-16, 2)):
        keep, lost = ops.truncation_rank(up(S), max(np.finfo(float).eps, tol), Dmax)
        kref = min(int(np.sum(S > S[0] * max(np.finfo(float).eps, tol))), Dmax)
        assert keep == kref and abs(lost - np.sqrt(np.sum(S[kref:] ** 2)) / S[0]) < 1e-15
    rng = np.random.default_rng(1)
    for scale in (1.0, 3e-7, 1e9, 0.99999, 2.0):
        x = rng.standard_normal((13, 7)) * scale
        t = up(x)
        ops.pow2_scale_(t, ops.maxabs_bits(t))
        assert np.array_equal(t.cpu().numpy(), x / ref_nfactor(x))
    one = up(np.array([[3.7]]))
    ops.pow2_scale_(one, ops.maxabs_bits(one))
    assert one.item() == 1.0


@pytest.mark.parametrize('conj', [True, False])
@pytest.mark.parametrize('shape', [(1, 1, 1, 1, 16, 16), (3, 16, 5, 16, 16, 16), (8, 16, 8, 1, 16, 16), (4, 4, 6, 2, 3, 5)])
def test_mpo_apply(conj, shape):
    from tnac4o_b200 import ops
    from oracle import RefMPS
    Dl, dp, Dr, wl, wr, du = shape
    rng = np.random.default_rng(sum(shape))
    A = rng.standard_normal((Dl, dp, Dr))
    W = rng.standard_normal((wl, dp, wr, du) if conj else (wl, du, wr, dp))
    psi = RefMPS(1, d=1)
    psi.A = [A.copy()]
    psi.apply_mpo([W], conj=conj)
    out = ops.mpo_apply(up(A), up(W), conj=conj).cpu().numpy()
    assert out.shape == psi.A[0].shape
    assert np.max(np.abs(out - psi.A[0])) < 1e-13 * dp


@pytest.mark.parametrize('n', [1, 2, 100, 2048, 5000, 70000])
def test_sort_keys(n):
    from tnac4o_b200 import ops
    from tnac4o_b200._native import Context, check, lib, ptr
    rng = np.random.default_rng(n)
    cap = ops.sort_capacity(n)
    hi = rng.integers(0, 4, size=n, dtype=np.int64)
    lo = rng.integers(0, 2 ** 62, size=n, dtype=np.int64) * rng.integers(0, 2, size=n)
    tie = rng.permutation(n).astype(np.int64)
    bufs = []
    for a in (hi, lo, tie):
        t = torch.zeros(cap, dtype=torch.int64, device=dev())
        t[:n] = up(a)
        bufs.append(t)
    c = Context.get(dev())
    check(lib.tn_sort_keys(c.handle, c.stream, ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]), n))
    got = np.stack([b[:n].cpu().numpy() for b in bufs], axis=1)
    order = np.lexsort((tie, lo, hi))
    assert np.array_equal(got, np.stack([hi, lo, tie], axis=1)[order])


def test_energy_ising_kernel(J128):
    import tnac4o_b200
    from oracle.auxx_ref import energy_ising_dense
    states = np.random.default_rng(0).integers(0, 2, size=(300, 128)).astype(np.int8)
    E = tnac4o_b200.energy_Jij(J128, states)
    assert np.max(np.abs(E - energy_ising_dense(J128, states))) < 1e-10


def test_device_built_site_tables_match_host_expressions(J128):
    """tn_build_site_tables against the host restatement of _peps_tensor (tnac4o_b200/model.py: boltzmann, traced)"""
    from tnac4o_b200.model import HostTables, IsingLattice, upload_site_tables, upper_triangular
    lat = IsingLattice(upper_triangular(J128, 128), 4, 4, 8)
    rng = np.random.default_rng(0)
    X = tuple(np.exp(rng.uniform(-1, 1, size=(4, 4, 16))) for _ in range(4))          # non-trivial gauges
    sites, _keep = upload_site_tables(HostTables(lat, 3.0, X), 4, 4, dev())
    for ny, nx in [(0, 0), (1, 2), (3, 3), (2, 0)]:
        t = sites[ny][nx]
        Xu, Xl, Xr, Xd = X
        Wc, dmap, rmap = lat.boltzmann(ny, nx, 3.0, Xu[ny][nx], Xl[ny][nx], Xr[ny][nx], Xd[ny][nx])
        Wtr = lat.traced(Wc, dmap, rmap, t.nd, t.nr)
        scale = np.max(Wc)
        assert np.max(np.abs(t.Wlu.cpu().numpy() - Wc.transpose(1, 2, 0))) <= 1e-15 * scale
        assert np.max(np.abs(t.Wmpo.cpu().numpy() - Wtr)) <= 4e-15 * np.max(Wtr)
        assert np.max(np.abs(t.WtrU.cpu().numpy() - Wtr.transpose(3, 0, 1, 2))) <= 4e-15 * np.max(Wtr)


def test_qr_graph_replay_on_side_stream():
    """on a non-default stream tn_qr_pos replays a captured CUDA graph per shape: same results as the plain launches"""
    from tnac4o_b200 import ops
    rng = np.random.default_rng(11)
    A = rng.standard_normal((600, 96))
    Q0, R0, b0 = ops.qr_pos(up(A).clone())                  # default stream: plain launch sequence
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):                                   # capture, then two replays
            Q1, R1, b1 = ops.qr_pos(up(A).clone())
        s.synchronize()
    assert torch.equal(Q0, Q1) and torch.equal(R0, R1) and int(b0.item()) == int(b1.item())
    _check_qr(A)


@pytest.mark.parametrize('nb', [1, 63, 1000, 5000])
@pytest.mark.parametrize('dims', [(32, 32, 16, 16, 16, 16), (32, 1, 16, 16, 1, 16), (8, 24, 16, 16, 16, 1), (5, 7, 3, 4, 2, 6)])
def test_rr_level_grouped_gemm_against_einsum(nb, dims):
    """tn_rr_level (tnac4o.py:1776-1782): RR'[b] = sum A[a,p,b'] RR[b][b',r] Wtr[l,p,r,u_b] / nfactor, every branch with
    its own up index -- grouped by u, one DMMA GEMM per group tile, power-of-two scaling per branch"""
    import ctypes
    from tnac4o_b200._native import Context, TnSite, check, lib
    Dl, Dr, nl, nd, nr, nu = dims
    rng = np.random.default_rng(nb + Dl)
    A = rng.standard_normal((Dl, nd, Dr))
    W = rng.uniform(0.1, 1.0, size=(nu, nl, nd, nr))                      # [u][l][p][r] as SiteTables.WtrU
    RR = rng.standard_normal((nb, Dr, nr)) * np.exp(rng.uniform(-30, 30, size=(nb, 1, 1)))
    vind = rng.integers(0, nu, size=(nb, 5)).astype(np.uint8)
    dA, dW, dRR, dv = up(A), up(W), up(RR), up(vind)
    out = torch.empty((nb, Dl, nl), dtype=torch.float64, device=dev())
    site = TnSite(1, nl, nd, nr, nu, None, dW.data_ptr(), None, None, None, None, None)
    c = Context.get(dev())
    check(lib.tn_rr_level(c.handle, c.stream, ctypes.byref(site), nb, Dl, Dr, dA.data_ptr(), dRR.data_ptr(),
                          dv[:, 2:].data_ptr(), dv.stride(0), out.data_ptr()))
    ref = np.einsum('apc,bcr,blpr->bal', A, RR, W[vind[:, 2]], optimize=True)
    mx = np.abs(ref).reshape(nb, -1).max(axis=1)
    ref /= (2.0 ** np.floor(np.log2(mx)))[:, None, None]
    got = out.cpu().numpy()
    assert np.max(np.abs(got - ref)) < 1e-12 * np.sqrt(Dr * nr * nd)
    assert np.all(np.abs(got).reshape(nb, -1).max(axis=1) < 2.0) and np.all(np.abs(got).reshape(nb, -1).max(axis=1) >= 1.0)


@pytest.mark.parametrize('k,decades', [(512, 6), (512, 10), (384, 4)])
def test_svd_many_live_vectors_wide_cluster(k, decades):
    """slowly decaying spectra (beta = 1 boundary MPS): 260-430 live vectors run in one 16-CTA cluster, more take the
    multi-launch path; both must meet the same accuracy"""
    rng = np.random.default_rng(k + decades)
    U, _ = np.linalg.qr(rng.standard_normal((k, k)))
    V, _ = np.linalg.qr(rng.standard_normal((k, k)))
    s = np.logspace(0, -decades * 4, k)                 # about k * 4 / decades ... live vectors above eps
    C = np.triu((U * s) @ V.T)
    S = _check_svd(C)
    exact = np.linalg.svd(C, compute_uv=False)
    big = exact > 1e-9 * exact[0]
    assert np.max(np.abs(S[big] / exact[big] - 1)) < 1e-6
