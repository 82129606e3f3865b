// Native (C++) driver of one boundary-MPS row update:  psi <- compress( MPO . psi )  with the reference's fixed
// schedule (mps.py:175-200).  It issues exactly the kernel sequence of tnac4o_b200/mps.py (the Python mirror of the
// reference's MPS class, kept for tests and for the preconditioning sweeps), but without a Python interpreter in the
// loop: ~1500 primitive calls per row run back to back on one stream, host read-backs only for the rank decisions
// of truncateC and once per variational sweep.  Temporaries come from the stream-ordered allocator.
#include <math.h>

#include <vector>

#include "common.cuh"

int tn_gemm_impl(tn_ctx* ctx, cudaStream_t st, int tA, int tB, int M, int N, int K, double alpha, const double* A,
                 int lda, int64_t sA, const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC,
                 int batch);

namespace {

// the context whose pool serves the allocations of the calling host thread (one context per thread)
thread_local tn_ctx* g_ctx = nullptr;

// device buffer owned through the stream-ordered allocator
struct Buf {
    double* p = nullptr;
    int64_t n = 0;
    cudaStream_t st = nullptr;
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    Buf(Buf&& o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
    Buf& operator=(Buf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~Buf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, st);
        p = nullptr; n = 0;
    }
    int alloc(int64_t count, cudaStream_t s) {
        release();
        st = s; n = count;
        cudaError_t e = tn_malloc_async(g_ctx, (void**)&p, (size_t)(count > 0 ? count : 1) * sizeof(double), s);
        if (e != cudaSuccess) { p = nullptr; return tn_cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__); }
        return TN_OK;
    }
};

struct Mat { Buf b; int r = 0, c = 0; };              // row-major r x c
struct Ten { Buf b; int Dl = 0, d = 0, Dr = 0; };     // (Dl, d, Dr)

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

__global__ void fill_kernel(double* x, int64_t n, double v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void diag_kernel(const double* __restrict__ S, int k, double* __restrict__ C) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k * k) C[i] = (i / k == i % k) ? S[i / k] : 0.0;
}
__global__ void unit_vec_kernel(double* x, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (i == 0) ? 1.0 : 0.0;
}
__global__ void max_into_kernel(double* acc, const double* v) { if (threadIdx.x == 0 && *v > *acc) *acc = *v; }

struct Mps {
    tn_ctx* ctx;
    cudaStream_t st;
    int L;
    std::vector<Ten> A;
    Mat C;
    int pC;
    Buf log2norm;                  // 1 double on the device
    std::vector<Mat> R;            // L + 2 mixed environments; R[L+1] holds the scalar overlap
    std::vector<Mat> S;            // Schmidt values per bond (1 x k)
    std::vector<double> discarded;

    int init(tn_ctx* c, cudaStream_t s, int len) {
        ctx = c; st = s; L = len;
        A.resize(L); R.resize(L + 2); S.resize(L + 1); discarded.assign(L + 1, 0.0);
        TRY(log2norm.alloc(1, st));
        TN_CUDA(cudaMemsetAsync(log2norm.p, 0, sizeof(double), st));
        for (int n = 0; n < L + 2; ++n) TRY(ones(R[n], 1, 1));
        for (int n = 0; n <= L; ++n) TRY(unit_S(S[n], 1));
        TRY(ones(C, 1, 1));
        pC = L;
        return TN_OK;
    }
    int ones(Mat& m, int r, int c) {
        TRY(m.b.alloc((int64_t)r * c, st));
        m.r = r; m.c = c;
        fill_kernel<<<1, 64, 0, st>>>(m.b.p, (int64_t)r * c, 1.0);
        TN_LAUNCHED(ctx);
        return TN_OK;
    }
    int unit_S(Mat& m, int k) {
        TRY(m.b.alloc(k, st));
        m.r = 1; m.c = k;
        unit_vec_kernel<<<ceil_div(k, 128), 128, 0, st>>>(m.b.p, k);
        TN_LAUNCHED(ctx);
        return TN_OK;
    }
    int gemm(int tA, int tB, int M, int N, int K, const double* a, int lda, const double* b, int ldb, double* c, int ldc) {
        tn_prof_scope prof(ctx, st, TN_P_GEMM, 2.0 * M * N * K, 8.0 * ((double)M * K + (double)K * N + (double)M * N));
        return tn_gemm_impl(ctx, st, tA, tB, M, N, K, 1.0, a, lda, 0, b, ldb, 0, 0.0, c, ldc, 0, 1);
    }
    // algorithmic flops of an economic QR with explicit Q (dgeqrf + dorgqr) and of an SVD
    static double qr_flops(double m, double n) {
        const double k = m < n ? m : n;
        return 4.0 * m * n * k - 2.0 * (m + n) * k * k + 4.0 / 3.0 * k * k * k;
    }
    static double svd_flops(double m, double n, bool vectors) {
        const double k = m < n ? m : n;
        return vectors ? 22.0 * k * k * k : 8.0 / 3.0 * k * k * k;
    }

    // ---- moving the centre (mps.py:368-380, 532-548, 772-800)
    int attach_AC() {
        Ten& a = A[pC - 1];
        Buf out;
        TRY(out.alloc((int64_t)a.Dl * a.d * C.c, st));
        TRY(gemm(0, 0, a.Dl * a.d, C.c, a.Dr, a.b.p, a.Dr, C.b.p, C.c, out.p, C.c));
        a.b = std::move(out); a.Dr = C.c;
        return TN_OK;
    }
    int attach_CA() {
        Ten& a = A[pC];
        Buf out;
        TRY(out.alloc((int64_t)C.r * a.d * a.Dr, st));
        TRY(gemm(0, 0, C.r, a.d * a.Dr, a.Dl, C.b.p, C.c, a.b.p, a.d * a.Dr, out.p, a.d * a.Dr));
        a.b = std::move(out); a.Dl = C.r;
        return TN_OK;
    }
    int orth_left(int n) {
        Ten& a = A[n];
        const int m = a.Dl * a.d, nn = a.Dr, k = m < nn ? m : nn;
        Buf Q, Rm, bits;
        TRY(Q.alloc((int64_t)m * k, st)); TRY(Rm.alloc((int64_t)k * nn, st)); TRY(bits.alloc(1, st));
        {
            tn_prof_scope prof(ctx, st, TN_P_QR, qr_flops(m, nn), 8.0 * (2.0 * m * nn + (double)k * nn));
            TRY(tn_qr_pos(ctx, st, m, nn, a.b.p, nn, Q.p, k, Rm.p, nn, (unsigned long long*)bits.p));
        }
        tn_prof_scope prof(ctx, st, TN_P_MPS_OTHER, 0.0, 16.0 * k * nn);
        TRY(tn_pow2_scale(ctx, st, Rm.p, (int64_t)k * nn, (const unsigned long long*)bits.p, log2norm.p));
        a.b = std::move(Q); a.Dr = k;
        C.b = std::move(Rm); C.r = k; C.c = nn;
        pC = n + 1;
        return TN_OK;
    }
    int orth_right(int n) {
        Ten& a = A[n];
        const int rows = a.d * a.Dr, cols = a.Dl, k = rows < cols ? rows : cols;      // QR of the (d Dr) x Dl transpose
        Buf At, Q, Rm, bits, Qt, Ct;
        TRY(At.alloc((int64_t)rows * cols, st));
        {
            tn_prof_scope prof(ctx, st, TN_P_MPS_OTHER, 0.0, 16.0 * rows * cols);
            TRY(tn_transpose(ctx, st, cols, rows, a.b.p, rows, At.p, cols));
        }
        TRY(Q.alloc((int64_t)rows * k, st)); TRY(Rm.alloc((int64_t)k * cols, st)); TRY(bits.alloc(1, st));
        {
            tn_prof_scope prof(ctx, st, TN_P_QR, qr_flops(rows, cols), 8.0 * (2.0 * rows * cols + (double)k * cols));
            TRY(tn_qr_pos(ctx, st, rows, cols, At.p, cols, Q.p, k, Rm.p, cols, (unsigned long long*)bits.p));
        }
        tn_prof_scope prof(ctx, st, TN_P_MPS_OTHER, 0.0, 16.0 * ((double)k * cols * 2 + (double)rows * k));
        TRY(tn_pow2_scale(ctx, st, Rm.p, (int64_t)k * cols, (const unsigned long long*)bits.p, log2norm.p));
        TRY(Qt.alloc((int64_t)k * rows, st));
        TRY(tn_transpose(ctx, st, rows, k, Q.p, k, Qt.p, rows));
        TRY(Ct.alloc((int64_t)cols * k, st));
        TRY(tn_transpose(ctx, st, k, cols, Rm.p, cols, Ct.p, k));
        a.b = std::move(Qt); a.Dl = k;
        C.b = std::move(Ct); C.r = cols; C.c = k;
        pC = n;
        return TN_OK;
    }
    // ---- SVD truncation of the centre matrix (mps.py:562-585, 802-811)
    int truncateC(double Dmax, double tol) {
        if (!(pC > 0 && pC < L)) return TN_OK;
        const int m = C.r, n = C.c, k = m < n ? m : n;
        Buf U, Sv, Vt;
        TRY(U.alloc((int64_t)m * k, st)); TRY(Sv.alloc(k, st)); TRY(Vt.alloc((int64_t)k * n, st));
        int sweeps = 0;
        {
            tn_prof_scope prof(ctx, st, TN_P_SVD, svd_flops(m, n, true), 8.0 * (3.0 * m * n));
            TRY(tn_svd(ctx, st, m, n, C.b.p, n, U.p, k, Sv.p, Vt.p, n, 1, &sweeps));
        }
        const double eps = 2.220446049250313e-16;
        int keep = 0;
        double lost = 0.0;
        int dmax = Dmax > 1e9 ? (1 << 30) : (int)Dmax;
        TRY(tn_truncation_rank(ctx, st, Sv.p, k, tol > eps ? tol : eps, dmax, &keep, &lost));
        if (keep < 1) { tn_set_error("truncateC: no singular value above the tolerance"); return TN_ERR_ARG; }
        Ten& al = A[pC - 1];
        Buf nl;
        TRY(nl.alloc((int64_t)al.Dl * al.d * keep, st));
        TRY(gemm(0, 0, al.Dl * al.d, keep, al.Dr, al.b.p, al.Dr, U.p, k, nl.p, keep));
        al.b = std::move(nl); al.Dr = keep;
        Ten& ar = A[pC];
        Buf nr;
        TRY(nr.alloc((int64_t)keep * ar.d * ar.Dr, st));
        TRY(gemm(0, 0, keep, ar.d * ar.Dr, ar.Dl, Vt.p, n, ar.b.p, ar.d * ar.Dr, nr.p, ar.d * ar.Dr));
        ar.b = std::move(nr); ar.Dl = keep;
        TRY(C.b.alloc((int64_t)keep * keep, st));
        C.r = C.c = keep;
        diag_kernel<<<ceil_div((int64_t)keep * keep, 256), 256, 0, st>>>(Sv.p, keep, C.b.p);
        TN_LAUNCHED(ctx);
        if (lost > discarded[pC]) discarded[pC] = lost;
        return TN_OK;
    }
    int canonise_left(bool compress, double Dmax, double tol) {
        TRY(ones(C, 1, 1)); pC = 0;
        for (int n = 0; n < L; ++n) {
            TRY(attach_CA()); TRY(orth_left(n));
            if (compress) TRY(truncateC(Dmax, tol));
        }
        return TN_OK;
    }
    int canonise_right(bool compress, double Dmax, double tol) {
        TRY(ones(C, 1, 1)); pC = L;
        for (int n = L - 1; n >= 0; --n) {
            TRY(attach_AC()); TRY(orth_right(n));
            if (compress) TRY(truncateC(Dmax, tol));
        }
        return TN_OK;
    }
    // ---- mixed environments <self|phi> (mps.py:418-452, 655-663)
    // R[n] (left) is (D_self, D_phi); R[n+1] (right) is (D_phi, D_self)
    int update_RL(const Mps& phi, int n, const Buf* T1) {
        const Ten& a = phi.A[n];
        const Ten& ac = A[n];
        Buf T;
        const double* t = nullptr;
        if (T1) t = T1->p;
        else {
            TRY(T.alloc((int64_t)R[n].r * a.d * a.Dr, st));
            TRY(gemm(0, 0, R[n].r, a.d * a.Dr, a.Dl, R[n].b.p, R[n].c, a.b.p, a.d * a.Dr, T.p, a.d * a.Dr));
            t = T.p;
        }
        Mat out;
        TRY(out.b.alloc((int64_t)ac.Dr * a.Dr, st));
        out.r = ac.Dr; out.c = a.Dr;
        TRY(gemm(1, 0, ac.Dr, a.Dr, ac.Dl * ac.d, ac.b.p, ac.Dr, t, a.Dr, out.b.p, a.Dr));
        if (n == L - 1) R[L + 1] = std::move(out); else R[n + 1] = std::move(out);
        return TN_OK;
    }
    int update_RR(const Mps& phi, int n) {
        const Ten& a = phi.A[n];
        const Ten& ac = A[n];
        const Mat& rr = R[n + 1];
        Buf T;
        TRY(T.alloc((int64_t)a.Dl * a.d * rr.c, st));
        TRY(gemm(0, 0, a.Dl * a.d, rr.c, a.Dr, a.b.p, a.Dr, rr.b.p, rr.c, T.p, rr.c));
        Mat out;
        TRY(out.b.alloc((int64_t)a.Dl * ac.Dl, st));
        out.r = a.Dl; out.c = ac.Dl;
        TRY(gemm(0, 1, a.Dl, ac.Dl, a.d * rr.c, T.p, a.d * rr.c, ac.b.p, ac.d * ac.Dr, out.b.p, ac.Dl));
        if (n == 0) R[L + 1] = std::move(out); else R[n] = std::move(out);
        return TN_OK;
    }
    // A[n] <- R[n] . phi.A[n] . R[n+1]; T1 = R[n] . phi.A[n] is returned for reuse (mps.py:617-621, 748-751)
    int optimise_site(const Mps& phi, int n, Buf& T1) {
        const Ten& a = phi.A[n];
        const Mat& rl = R[n];
        const Mat& rr = R[n + 1];
        TRY(T1.alloc((int64_t)rl.r * a.d * a.Dr, st));
        TRY(gemm(0, 0, rl.r, a.d * a.Dr, a.Dl, rl.b.p, rl.c, a.b.p, a.d * a.Dr, T1.p, a.d * a.Dr));
        Buf out;
        TRY(out.alloc((int64_t)rl.r * a.d * rr.c, st));
        TRY(gemm(0, 0, rl.r * a.d, rr.c, a.Dr, T1.p, a.Dr, rr.b.p, rr.c, out.p, rr.c));
        A[n].b = std::move(out); A[n].Dl = rl.r; A[n].d = a.d; A[n].Dr = rr.c;
        return TN_OK;
    }
    // Schmidt values of C; *dS (device) = ||S_old - S_new||_2 (mps.py:550-560)
    int update_S(double* dS) {
        const int k = C.r < C.c ? C.r : C.c;
        Mat Sn;
        TRY(Sn.b.alloc(k, st));
        Sn.r = 1; Sn.c = k;
        int sweeps = 0;
        {
            tn_prof_scope prof(ctx, st, TN_P_SVD, svd_flops(C.r, C.c, false), 8.0 * C.r * C.c);
            TRY(tn_svd(ctx, st, C.r, C.c, C.b.p, C.c, nullptr, 1, Sn.b.p, nullptr, 1, 0, &sweeps));
        }
        if (S[pC].c != k) TRY(unit_S(S[pC], k));
        TRY(tn_diff_norm(ctx, st, S[pC].b.p, Sn.b.p, k, dS));
        S[pC] = std::move(Sn);
        return TN_OK;
    }
    int read_scalar(const double* dptr, double* out) {
        double* h = (double*)((char*)ctx->pinned + 384);
        TN_CUDA(cudaMemcpyAsync(h, dptr, sizeof(double), cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaStreamSynchronize(st));
        *out = *h;
        return TN_OK;
    }
    int variational_compress(const Mps& phi, double tol, int max_sweeps, double* overlap) {
        for (int n = 0; n < L; ++n) TRY(update_RL(phi, n, nullptr));
        Buf dS, dmax;
        TRY(dS.alloc(1, st)); TRY(dmax.alloc(1, st));
        int sweeps = 0;
        double diff = 1.0;
        while (diff > tol) {
            if (sweeps >= max_sweeps) break;
            for (int n = L - 1; n > 0; --n) {
                Buf T1;
                TRY(optimise_site(phi, n, T1));
                TRY(orth_right(n));
                TRY(update_S(dS.p));
                TRY(update_RR(phi, n));
            }
            TN_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(double), st));
            for (int n = 0; n < L; ++n) {
                Buf T1;
                TRY(optimise_site(phi, n, T1));
                TRY(orth_left(n));
                TRY(update_S(dS.p));
                max_into_kernel<<<1, 32, 0, st>>>(dmax.p, dS.p);
                TN_LAUNCHED(ctx);
                TRY(update_RL(phi, n, &T1));
            }
            TRY(read_scalar(dmax.p, &diff));           // the only host read of the sweep
            ++sweeps;
        }
        return read_scalar(R[L + 1].b.p, overlap);
    }
    // deep copy of the tensors only, like MPS.copy() (mps.py:159-173)
    int copy_tensors_from(const Mps& o) {
        for (int n = 0; n < L; ++n) {
            const Ten& s = o.A[n];
            TRY(A[n].b.alloc((int64_t)s.Dl * s.d * s.Dr, st));
            A[n].Dl = s.Dl; A[n].d = s.d; A[n].Dr = s.Dr;
            TN_CUDA(cudaMemcpyAsync(A[n].b.p, s.b.p, (size_t)s.Dl * s.d * s.Dr * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        return TN_OK;
    }
    // the reference's fixed schedule (mps.py:175-200)
    int compress(double Dmax, double tolS, double tolV, int max_sweeps, int graduate, double* overlap) {
        TRY(canonise_right(false, 0, 0));
        Mps phi;
        TRY(phi.init(ctx, st, L));
        TRY(phi.copy_tensors_from(*this));
        discarded.assign(L + 1, 0.0);
        if (graduate) {
            TRY(canonise_left(true, Dmax * 4, tolS / 10));
            double ov;
            TRY(variational_compress(phi, tolV, 1, &ov));
            TRY(canonise_right(true, Dmax * 2, tolS / 2));
        }
        TRY(canonise_left(true, Dmax, tolS));
        return variational_compress(phi, tolV, max_sweeps, overlap);
    }
};

}  // namespace

struct tn_row {
    Mps psi;
    double overlap = 0.0;
};

extern "C" {

/* psi <- compress(MPO . psi): A_in[n] (Dl[n], dphys[n], Dr[n]) device tensors of the previous row's MPS, W[n] device MPO
 * tensors with legs (wl[n], d_out, wr[n], d_in) (Hconj form: d_out = dphys[n], result leg d_in = du[n]).  The result
 * stays on the device inside *out until tn_row_free. */
int tn_row_compress(tn_ctx* ctx, void* stream, int L, const double* const* A_in, const int* Dl, const int* dphys,
                    const int* Dr, const double* const* W, const int* wl, const int* wr, const int* du, int conj,
                    double Dmax, double tolS, double tolV, int max_sweeps, int graduate, tn_row** out) {
    TN_REQUIRE(ctx && out && L >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    g_ctx = ctx;
    tn_row* row = new tn_row();
    int rc = row->psi.init(ctx, st, L);
    for (int n = 0; n < L && !rc; ++n) {
        Ten& a = row->psi.A[n];
        a.Dl = Dl[n] * wl[n]; a.d = du[n]; a.Dr = Dr[n] * wr[n];
        rc = a.b.alloc((int64_t)a.Dl * a.d * a.Dr, st);
        if (!rc) {
            tn_prof_scope prof(ctx, st, TN_P_MPS_OTHER, 2.0 * a.Dl * a.d * a.Dr * dphys[n], 8.0 * a.Dl * a.d * a.Dr);
            rc = tn_mpo_apply(ctx, st, conj, Dl[n], dphys[n], Dr[n], wl[n], wr[n], du[n], A_in[n], W[n], a.b.p);
        }
    }
    if (!rc) rc = row->psi.compress(Dmax, tolS, tolV, max_sweeps, graduate, &row->overlap);
    if (rc) { cudaStreamSynchronize(st); delete row; return rc; }
    *out = row;
    return TN_OK;
}

/* shapes of the compressed row: D (L + 1 bond dimensions), d (L physical dimensions) */
int tn_row_shapes(const tn_row* row, int* D, int* d) {
    TN_REQUIRE(row && D && d, "bad arguments");
    const Mps& p = row->psi;
    D[0] = p.A[0].Dl;
    for (int n = 0; n < p.L; ++n) { D[n + 1] = p.A[n].Dr; d[n] = p.A[n].d; }
    return TN_OK;
}

/* copies the tensors into caller-owned buffers of the sizes reported by tn_row_shapes; host scalars: overlap with the
 * uncompressed state, per-bond discarded weights (L + 1), log2 of the accumulated norm (synchronises) */
int tn_row_fetch(tn_row* row, double* const* A_out, double* h_overlap, double* h_discarded, double* h_log2norm) {
    TN_REQUIRE(row && A_out, "bad arguments");
    Mps& p = row->psi;
    for (int n = 0; n < p.L; ++n) {
        const Ten& a = p.A[n];
        TN_CUDA(cudaMemcpyAsync(A_out[n], a.b.p, (size_t)a.Dl * a.d * a.Dr * sizeof(double), cudaMemcpyDeviceToDevice, p.st));
    }
    if (h_overlap) *h_overlap = row->overlap;
    if (h_discarded) for (int n = 0; n <= p.L; ++n) h_discarded[n] = p.discarded[n];
    if (h_log2norm) { int rc = p.read_scalar(p.log2norm.p, h_log2norm); if (rc) return rc; }
    return TN_OK;
}

int tn_row_free(tn_row* row) {
    if (row) {
        cudaStream_t st = row->psi.st;
        delete row;
        cudaStreamSynchronize(st);
    }
    return TN_OK;
}

}  // extern "C"
