// Economic Householder QR with non-negative diag(R) -- replaces mps.qr (mps.py:43-59).
//
// Blocked compact-WY algorithm.  The panel (m x JB) is factored by ONE thread-block cluster of 8 CTAs: the
// panel rows are split over the CTAs and kept in shared memory for the whole factorisation; the per-column
// dot products are reduced across the cluster through distributed shared memory (one cluster.sync per
// column).  Trailing updates and the accumulation of Q are DMMA GEMMs (gemm.cu).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

int tn_gemm_impl(tn_ctx* ctx, cudaStream_t st, int tA, int tB, int M, int N, int K, double alpha, const double* A,
                 int lda, int64_t sA, const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC,
                 int batch);

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int PT = 1024;       // threads per CTA

template <int JB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(PT, 1)
qr_panel_kernel(double* __restrict__ A, int lda, int m, int j0, int jb, double* __restrict__ Vall, int ldv,
                double* __restrict__ Tout) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) double P[];          // [rows_per][JB]
    __shared__ double cw[2][CL][JB];                      // per-CTA partial dots, double-buffered
    __shared__ double prow[2][JB];                        // pivot row broadcast
    __shared__ double red[PT / 32][JB];
    __shared__ double tau_s[JB];
    __shared__ double Gpart[CL][JB * JB];                 // only rank 0's copy is used
    __shared__ double Ts[JB * JB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid % JB, rg = tid / JB;
    constexpr int RG = PT / JB;
    const int m_rem = m - j0;
    const int rows_per = (m_rem + CL - 1) / CL;
    const int r_begin = rank * rows_per;                          // relative to j0
    const int nloc = max(0, min(rows_per, m_rem - r_begin));

    // load my rows of the panel
    for (int idx = tid; idx < nloc * JB; idx += PT) {
        int i = idx / JB, cc = idx % JB;
        P[idx] = (cc < jb) ? A[(int64_t)(j0 + r_begin + i) * lda + j0 + cc] : 0.0;
    }
    __syncthreads();

    for (int j = 0; j < jb; ++j) {
        const int buf = j & 1;
        const int owner = j / rows_per;          // CTA that holds the pivot row (relative row j)
        // ---- phase A: partial dots over rows strictly below the pivot
        double acc = 0.0;
        for (int i = rg; i < nloc; i += RG) {
            if (r_begin + i > j) acc += P[i * JB + j] * P[i * JB + c];
        }
        if (JB == 16) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (JB == 8) { acc += __shfl_xor_sync(0xffffffffu, acc, 16); acc += __shfl_xor_sync(0xffffffffu, acc, 8); }
        if (lane < JB) red[warp][lane] = acc;
        __syncthreads();
        if (tid < JB) {
            double s = 0.0;
            for (int w = 0; w < PT / 32; ++w) s += red[w][tid];
            for (int r = 0; r < CL; ++r) {
                double* remote = cluster.map_shared_rank(&cw[buf][rank][tid], r);
                *remote = s;
            }
            if (rank == owner) {
                double pv = P[(j - r_begin) * JB + tid];
                for (int r = 0; r < CL; ++r) {
                    double* remote = cluster.map_shared_rank(&prow[buf][tid], r);
                    *remote = pv;
                }
            }
        }
        cluster.sync();
        // ---- phase B: reflector parameters (every thread, identical arithmetic)
        double wj = 0.0, wc = 0.0;
#pragma unroll
        for (int r = 0; r < CL; ++r) { wj += cw[buf][r][j]; wc += cw[buf][r][c]; }
        const double alpha = prow[buf][j];
        double beta, tau, scale;
        if (wj == 0.0) { beta = alpha; tau = 0.0; scale = 0.0; }
        else {
            beta = -copysign(sqrt(alpha * alpha + wj), alpha);
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        const double sc = tau * (prow[buf][c] + scale * wc);     // tau * v^T a_c
        // ---- phase C: apply to the remaining panel columns
        for (int i = rg; i < nloc; i += RG) {
            int rel = r_begin + i;
            if (rel > j) {
                if (c > j) P[i * JB + c] -= (P[i * JB + j] * scale) * sc;
            } else if (rel == j) {
                if (c > j) P[i * JB + c] -= sc;
            }
        }
        __syncthreads();
        if (c == j) {
            for (int i = rg; i < nloc; i += RG) {
                int rel = r_begin + i;
                if (rel > j) P[i * JB + j] *= scale;
                else if (rel == j) P[i * JB + j] = beta;
            }
            if (rg == 0) tau_s[j] = tau;
        }
        __syncthreads();
    }

    // ---- write R rows back to A, convert the shared panel to explicit V, store V
    for (int idx = tid; idx < nloc * JB; idx += PT) {
        int i = idx / JB, cc = idx % JB;
        int rel = r_begin + i;
        double x = P[idx];
        if (cc < jb) {
            if (rel <= cc) A[(int64_t)(j0 + rel) * lda + j0 + cc] = x;      // upper triangle incl. diagonal = R
            double v = (rel > cc) ? x : (rel == cc ? 1.0 : 0.0);
            P[idx] = v;
            Vall[(int64_t)(j0 + rel) * ldv + j0 + cc] = v;
        } else {
            P[idx] = 0.0;
        }
    }
    __syncthreads();
    // ---- Gram of V for the T factor: G[a][b] = sum_rows V[.,a] V[.,b]
    {
        constexpr int NP = JB * JB;
        constexpr int SL = PT / NP;                      // row slices per (a,b) pair
        int pair = tid % NP, sl = tid / NP;
        int a = pair / JB, b = pair % JB;
        double g = 0.0;
        if (sl < SL)
            for (int i = sl; i < nloc; i += SL) g += P[i * JB + a] * P[i * JB + b];
        // reduce the SL slices through shared memory (reuse red as scratch is too small -> use Ts stages)
        __shared__ double gred[PT];
        gred[tid] = g;
        __syncthreads();
        if (tid < NP) {
            double s = 0.0;
            for (int q = 0; q < SL; ++q) s += gred[q * NP + tid];
            double* remote = cluster.map_shared_rank(&Gpart[rank][tid], 0);
            *remote = s;
        }
    }
    cluster.sync();
    if (rank == 0) {
        constexpr int NP = JB * JB;
        if (tid < NP) {
            double s = 0.0;
            for (int r = 0; r < CL; ++r) s += Gpart[r][tid];
            Gpart[0][tid] = s;            // own slot r = 0 is summed first, safe to overwrite after the loop
            Ts[tid] = 0.0;
        }
        __syncthreads();
        if (tid == 0) {
            // forward columnwise dlarft: T[j][j] = tau_j, T[0:j, j] = -tau_j * T[0:j,0:j] * G[0:j, j]
            for (int j = 0; j < jb; ++j) {
                for (int i = 0; i < j; ++i) {
                    double s = 0.0;
                    for (int k = i; k < j; ++k) s += Ts[i * JB + k] * Gpart[0][k * JB + j];
                    Ts[i * JB + j] = -tau_s[j] * s;
                }
                Ts[j * JB + j] = tau_s[j];
            }
        }
        __syncthreads();
        if (tid < NP) Tout[tid] = Ts[tid];
    }
}

__global__ void set_identity_kernel(double* Q, int ldq, int m, int k) {
    int64_t total = (int64_t)m * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / k), cc = (int)(i % k);
        Q[(int64_t)r * ldq + cc] = (r == cc) ? 1.0 : 0.0;
    }
}

// R_out = sgn * triu(A[0:k, :]),  Q[:, i] *= sgn_i,  sgn_i = -1 if R_ii < 0 else +1; optionally max|R|.
__global__ void qr_finish_kernel(const double* __restrict__ A, int lda, int m, int n, int k, double* __restrict__ Q,
                                 int ldq, double* __restrict__ R, int ldr, unsigned long long* maxabs_bits) {
    int64_t nR = (int64_t)k * n, nQ = (int64_t)m * k;
    double local_max = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nR + nQ; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < nR) {
            int r = (int)(i / n), cc = (int)(i % n);
            double sg = (A[(int64_t)r * lda + r] < 0.0) ? -1.0 : 1.0;
            double v = (cc >= r) ? sg * A[(int64_t)r * lda + cc] : 0.0;
            R[(int64_t)r * ldr + cc] = v;
            local_max = fmax(local_max, fabs(v));
        } else {
            int64_t q = i - nR;
            int r = (int)(q / k), cc = (int)(q % k);
            if (A[(int64_t)cc * lda + cc] < 0.0) Q[(int64_t)r * ldq + cc] = -Q[(int64_t)r * ldq + cc];
        }
    }
    if (maxabs_bits) {
        local_max = warp_max(local_max);
        if ((threadIdx.x & 31) == 0 && local_max > 0.0)
            atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(local_max));
    }
}

template <int JB>
int qr_impl(tn_ctx* ctx, cudaStream_t st, int m, int n, double* A, int lda, double* Q, int ldq, double* R, int ldr,
            unsigned long long* maxabs_bits) {
    const int k = min(m, n);
    const int npan = ceil_div(k, JB);
    const int wcols = max(n, k);
    size_t need = ((size_t)m * k + (size_t)npan * JB * JB + 2 * (size_t)JB * wcols) * sizeof(double);
    double* ws = (double*)tn_scratch(ctx, TN_SLOT_QR, need);
    if (!ws) return TN_ERR_NOMEM;
    double* Vall = ws;
    double* Tall = Vall + (size_t)m * k;
    double* W = Tall + (size_t)npan * JB * JB;
    double* W2 = W + (size_t)JB * wcols;

    auto kern = qr_panel_kernel<JB>;
    const int rows_per_max = ceil_div(m, CL);
    size_t smem = (size_t)rows_per_max * JB * sizeof(double);
    TN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    for (int p = 0; p < npan; ++p) {
        int j0 = p * JB, jb = min(JB, k - j0);
        kern<<<CL, PT, smem, st>>>(A, lda, m, j0, jb, Vall, k, Tall + (size_t)p * JB * JB);
        TN_LAUNCHED(ctx);
        int n2 = n - (j0 + jb);
        if (n2 > 0) {
            int mr = m - j0;
            const double* V = Vall + (size_t)j0 * k + j0;
            double* A2 = A + (size_t)j0 * lda + j0 + jb;
            int rc;
            if ((rc = tn_gemm_impl(ctx, st, 1, 0, jb, n2, mr, 1.0, V, k, 0, A2, lda, 0, 0.0, W, n2, 0, 1))) return rc;
            if ((rc = tn_gemm_impl(ctx, st, 1, 0, jb, n2, jb, 1.0, Tall + (size_t)p * JB * JB, JB, 0, W, n2, 0, 0.0, W2, n2, 0, 1))) return rc;
            if ((rc = tn_gemm_impl(ctx, st, 0, 0, mr, n2, jb, -1.0, V, k, 0, W2, n2, 0, 1.0, A2, lda, 0, 1))) return rc;
        }
    }
    // accumulate Q = H_1 ... H_p [I; 0]
    {
        int64_t total = (int64_t)m * k;
        int blocks = (int)((total + 255) / 256 < 8 * ctx->sm_count ? (total + 255) / 256 : 8 * ctx->sm_count);
        set_identity_kernel<<<blocks, 256, 0, st>>>(Q, ldq, m, k);
        TN_LAUNCHED(ctx);
    }
    for (int p = npan - 1; p >= 0; --p) {
        int j0 = p * JB, jb = min(JB, k - j0);
        int mr = m - j0, kc = k - j0;
        const double* V = Vall + (size_t)j0 * k + j0;
        double* Q2 = Q + (size_t)j0 * ldq + j0;
        int rc;
        if ((rc = tn_gemm_impl(ctx, st, 1, 0, jb, kc, mr, 1.0, V, k, 0, Q2, ldq, 0, 0.0, W, kc, 0, 1))) return rc;
        if ((rc = tn_gemm_impl(ctx, st, 0, 0, jb, kc, jb, 1.0, Tall + (size_t)p * JB * JB, JB, 0, W, kc, 0, 0.0, W2, kc, 0, 1))) return rc;
        if ((rc = tn_gemm_impl(ctx, st, 0, 0, mr, kc, jb, -1.0, V, k, 0, W2, kc, 0, 1.0, Q2, ldq, 0, 1))) return rc;
    }
    if (maxabs_bits) TN_CUDA(cudaMemsetAsync(maxabs_bits, 0, sizeof(unsigned long long), st));
    {
        int64_t total = (int64_t)k * n + (int64_t)m * k;
        int blocks = (int)((total + 255) / 256 < 8 * ctx->sm_count ? (total + 255) / 256 : 8 * ctx->sm_count);
        qr_finish_kernel<<<blocks, 256, 0, st>>>(A, lda, m, n, k, Q, ldq, R, ldr, maxabs_bits);
        TN_LAUNCHED(ctx);
    }
    return TN_OK;
}

}  // namespace

extern "C" int tn_qr_pos(tn_ctx* ctx, void* stream, int m, int n, double* A, int lda, double* Q, int ldq, double* R,
                         int ldr, unsigned long long* maxabs_bits) {
    TN_REQUIRE(ctx != nullptr, "null context");
    TN_REQUIRE(m >= 1 && n >= 1, "empty matrix");
    TN_REQUIRE(lda >= n && ldq >= (m < n ? m : n) && ldr >= n, "bad leading dimension");
    cudaStream_t st = as_stream(stream);
    const int rows_per = ceil_div(m, CL);
    if ((size_t)rows_per * 16 * sizeof(double) <= 160 * 1024)
        return qr_impl<16>(ctx, st, m, n, A, lda, Q, ldq, R, ldr, maxabs_bits);
    TN_REQUIRE((size_t)rows_per * 8 * sizeof(double) <= 160 * 1024, "matrix too tall for the shared-memory panel");
    return qr_impl<8>(ctx, st, m, n, A, lda, Q, ldq, R, ldr, maxabs_bits);
}
