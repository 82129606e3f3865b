// Economic Householder QR with non-negative diag(R) -- replaces mps.qr (mps.py:43-59).
//
// Two-level blocked compact-WY algorithm.
//   * Inner panel (m x 16): ONE thread-block cluster of 8 CTAs.  Every thread owns one matrix row in registers
//     (16 doubles); the per-column dot products  x^T A[:, c]  of all 16 columns are reduced with a transposing
//     warp butterfly (16 shuffle steps for 16 values), across warps through shared memory and across the 8 CTAs
//     through distributed shared memory -- one cluster barrier per column.  The same reduction yields the Gram
//     entries needed for the compact-WY factor T, so V, R and T come out of a single pass.
//     (Panels taller than 8 x 1024 rows fall back to a shared-memory variant of the same scheme.)
//   * The inner panels of an outer block (128 columns) update only that block; the block reflector
//     (V_outer, T_outer) is then applied to the trailing matrix and, at the end, to Q with K = 128 DMMA GEMMs.
#include <cooperative_groups.h>

#include <stdlib.h>
#include <utility>

#include "common.cuh"

namespace cg = cooperative_groups;

int tn_throughput_mode();      // svd.cu: many solver instances share the GPU -> small grids for the latency-bound helpers

int tn_gemm_impl(tn_ctx* ctx, cudaStream_t st, int tA, int tB, int M, int N, int K, double alpha, const double* A,
                 int lda, int64_t sA, const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC,
                 int batch);

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int PT = 1024;       // threads per CTA
constexpr int JB = 16;         // inner panel width
constexpr int NB = 128;        // outer block width

// ---------------------------------------------------------------------------------------------------------------
// register-resident panel factorisation: each of the RT threads of a CTA owns up to RPT rows (RT * RPT >= rows_per)
constexpr int RT = 256;        // threads per CTA of the register kernel
constexpr int RPT = 4;         // rows per thread

// Optional phase timers (tools/microbench builds with -DTN_PHASES): cycles spent in each phase of the column loop,
// accumulated by thread 0 of CTA 0 into a global array read back by tn_debug_phases().
#ifdef TN_PHASES
__device__ long long g_phase[16];
#define PH_DECL long long ph_t = clock64(), ph_n;
#define PH(k) do { if (tid == 0 && rank == 0) { ph_n = clock64(); g_phase[k] += ph_n - ph_t; ph_t = ph_n; } } while (0)
#else
#define PH_DECL
#define PH(k) do { } while (0)
#endif

// The cluster size is a launch attribute: 1, 2, 4 or 8 CTAs, the smallest that gives every row a register slot
// (RT * RPT rows per CTA).  A single CTA needs no cluster traffic at all.
__global__ void __launch_bounds__(RT, 1)
qr_panel_reg_kernel(double* __restrict__ A, int lda, int m, int j0, int jb, double* __restrict__ Vall, int ldv,
                    double* __restrict__ Tout) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int ncl = (int)cluster.num_blocks();
    __shared__ double red[RT / 32][JB];
    __shared__ double cw[2][CL][JB];          // per-CTA partial dots, double-buffered by column parity
    __shared__ double prow[2][JB];            // pivot row broadcast
    __shared__ double sc[JB];                 // tau * v^T a_c
    __shared__ double par[4];                 // beta, tau, scale
    __shared__ double Ts[JB * JB];
    __shared__ double stage[JB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m_rem = m - j0;
    const int rows_per = (m_rem + ncl - 1) / ncl;
    int rel[RPT];
    bool have[RPT];
    double row[RPT][JB];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int loc = tid + q * RT;
        rel[q] = rank * rows_per + loc;                          // my row, relative to j0
        have[q] = (loc < rows_per) && (rel[q] < m_rem);
#pragma unroll
        for (int c = 0; c < JB; ++c) row[q][c] = (have[q] && c < jb) ? A[(int64_t)(j0 + rel[q]) * lda + j0 + c] : 0.0;
    }
    if (tid < JB * JB) Ts[tid] = 0.0;
    // every CTA of the cluster must have started before anyone writes into its shared memory
    if (ncl > 1) cluster.sync(); else __syncthreads();

    // The register file of every thread is ROTATED by one column per step, so the current column is always slot 0,
    // the columns still to be updated are slots 1 .. 15-j and the finished reflectors are slots 16-j .. 15.  The loop
    // body is therefore the same code for every column.  (Measured in round 2: a fully unrolled variant with
    // compile-time column indices -- no rotation, no per-element predicates, 2.3x fewer instructions per column -- is
    // 330 KB of SASS and 2-12 % SLOWER: 4461 vs 4359 us for 8192 x 512, profiles/r2b_time_qr_lean_vs_rotating.txt.)
    PH_DECL
    PH(0);
#pragma unroll 1
    for (int j = 0; j < jb; ++j) {
        const int buf = j & 1;
        const int owner = j / rows_per;
        // ---- products of column j with every slot over the rows strictly below the pivot, summed over my rows
        double v[JB];
#pragma unroll
        for (int c = 0; c < JB; ++c) v[c] = 0.0;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const double x = (have[q] && rel[q] > j) ? row[q][0] : 0.0;
#pragma unroll
            for (int c = 0; c < JB; ++c) v[c] += x * row[q][c];
        }
        PH(1);
        // transposing butterfly: after the halving steps lane l holds the warp total of slot idx(l)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bool up = lane & 16;
            double send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool up = lane & 8;
            double send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            bool up = lane & 4;
            double send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        {
            bool up = lane & 2;
            double send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
            v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        if ((lane & 1) == 0) {
            int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            red[warp][idx] = v[0];
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            if (have[q] && rel[q] == j) {
#pragma unroll
                for (int c = 0; c < JB; ++c) stage[c] = row[q][c];     // stage the pivot row (owner CTA only)
            }
        }
        PH(2);
        __syncthreads();
        PH(3);
        if (tid < 32) {
            // lanes 0-15 and 16-31 hold the same column sums; the two half-warps share the remote stores
            const int c = tid & 15;
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < RT / 32; ++w) s += red[w][c];
            for (int r = (tid >> 4); r < ncl; r += 2) *cluster.map_shared_rank(&cw[buf][rank][c], r) = s;
            if (rank == owner) {
                const double pv = stage[c];
                for (int r = (tid >> 4); r < ncl; r += 2) *cluster.map_shared_rank(&prow[buf][c], r) = pv;
            }
        }
        PH(4);
        if (ncl > 1) cluster.sync(); else __syncthreads();
        PH(5);
        // ---- reflector parameters, tau * v^T a_slot and column j of T: 16 lanes, redundant scalar work
        if (tid < JB) {
            double wc = 0.0, w0 = 0.0;
            for (int r = 0; r < ncl; ++r) { wc += cw[buf][r][tid]; w0 += cw[buf][r][0]; }
            const double alpha = prow[buf][0];
            double beta, tau, scale;
            if (w0 == 0.0) { beta = alpha; tau = 0.0; scale = 0.0; }
            else {
                beta = -copysign(sqrt(alpha * alpha + w0), alpha);
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            // v^T a_slot = a_slot[pivot] + scale * w_slot  (v = e_pivot + scale * x below the pivot); for the finished
            // reflectors (slots 16-j ..) this is the Gram entry (V^T V)[c][j] needed by
            // T[0:j, j] = -tau * T[0:j, 0:j] * (V[:, 0:j]^T v_j)
            const double vta = prow[buf][tid] + scale * wc;
            sc[tid] = tau * vta;
            double s = 0.0;
            for (int k = 0; k < j; ++k) {
                const double gk = __shfl_sync(0xffffu, vta, JB - j + k);      // column k sits in slot 16 - j + k
                if (k >= tid) s += Ts[tid * JB + k] * gk;
            }
            if (tid < j) Ts[tid * JB + j] = -tau * s;
            if (tid == j) { Ts[j * JB + j] = tau; par[0] = beta; par[1] = tau; par[2] = scale; }
        }
        PH(6);
        __syncthreads();
        PH(7);
        const double beta = par[0], scale = par[2];
        const int last = JB - 1 - j;                 // slots 1 .. last are the columns still to be updated
        // ---- apply the reflector to my rows, then rotate the slots
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            if (have[q]) {
                if (rel[q] > j) {
                    const double vi = row[q][0] * scale;
#pragma unroll
                    for (int c = 1; c < JB; ++c) if (c <= last) row[q][c] -= vi * sc[c];
                    row[q][0] = vi;
                } else if (rel[q] == j) {
#pragma unroll
                    for (int c = 1; c < JB; ++c) if (c <= last) row[q][c] -= sc[c];
                    row[q][0] = beta;
                }
            }
            const double t0 = row[q][0];
#pragma unroll
            for (int c = 0; c < JB - 1; ++c) row[q][c] = row[q][c + 1];
            row[q][JB - 1] = t0;
        }
        PH(8);
    }
    // ---- write R (upper triangle of the pivot rows) and the explicit V; slot s holds column (s + jb) mod 16
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        if (have[q]) {
#pragma unroll
            for (int sl = 0; sl < JB; ++sl) {
                const int c = (sl + jb) & (JB - 1);
                if (c < jb) {
                    if (rel[q] <= c) A[(int64_t)(j0 + rel[q]) * lda + j0 + c] = row[q][sl];
                    Vall[(int64_t)(j0 + rel[q]) * ldv + j0 + c] = (rel[q] > c) ? row[q][sl] : (rel[q] == c ? 1.0 : 0.0);
                }
            }
        }
    }
    __syncthreads();
    if (rank == 0 && tid < JB * JB) Tout[tid] = Ts[tid];
    PH(9);
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory variant for very tall panels (rows_per > PT); same algorithm, panel rows in dynamic shared memory
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(PT, 1)
qr_panel_smem_kernel(double* __restrict__ A, int lda, int m, int j0, int jb, double* __restrict__ Vall, int ldv,
                     double* __restrict__ Tout) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) double P[];          // [rows_per][JB]
    __shared__ double cw[2][CL][JB];
    __shared__ double prow[2][JB];
    __shared__ double red[PT / 32][JB];
    __shared__ double tau_s[JB];
    __shared__ double Gpart[CL][JB * JB];
    __shared__ double Ts[JB * JB];
    __shared__ double gred[PT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid % JB, rg = tid / JB;
    constexpr int RG = PT / JB;
    const int m_rem = m - j0;
    const int rows_per = (m_rem + CL - 1) / CL;
    const int r_begin = rank * rows_per;
    const int nloc = max(0, min(rows_per, m_rem - r_begin));
    for (int idx = tid; idx < nloc * JB; idx += PT) {
        int i = idx / JB, cc = idx % JB;
        P[idx] = (cc < jb) ? A[(int64_t)(j0 + r_begin + i) * lda + j0 + cc] : 0.0;
    }
    cluster.sync();      // all CTAs started (required before the first distributed-shared-memory access)
    for (int j = 0; j < jb; ++j) {
        const int buf = j & 1;
        const int owner = j / rows_per;
        double acc = 0.0;
        for (int i = rg; i < nloc; i += RG)
            if (r_begin + i > j) acc += P[i * JB + j] * P[i * JB + c];
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < JB) red[warp][lane] = acc;
        __syncthreads();
        if (tid < JB) {
            double s = 0.0;
            for (int w = 0; w < PT / 32; ++w) s += red[w][tid];
            for (int r = 0; r < CL; ++r) *cluster.map_shared_rank(&cw[buf][rank][tid], r) = s;
            if (rank == owner) {
                double pv = P[(j - r_begin) * JB + tid];
                for (int r = 0; r < CL; ++r) *cluster.map_shared_rank(&prow[buf][tid], r) = pv;
            }
        }
        cluster.sync();
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
        const double scv = tau * (prow[buf][c] + scale * wc);
        for (int i = rg; i < nloc; i += RG) {
            int rel = r_begin + i;
            if (rel > j) { if (c > j) P[i * JB + c] -= (P[i * JB + j] * scale) * scv; }
            else if (rel == j) { if (c > j) P[i * JB + c] -= scv; }
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
    for (int idx = tid; idx < nloc * JB; idx += PT) {
        int i = idx / JB, cc = idx % JB;
        int rel = r_begin + i;
        double x = P[idx];
        if (cc < jb) {
            if (rel <= cc) A[(int64_t)(j0 + rel) * lda + j0 + cc] = x;
            double v = (rel > cc) ? x : (rel == cc ? 1.0 : 0.0);
            P[idx] = v;
            Vall[(int64_t)(j0 + rel) * ldv + j0 + cc] = v;
        } else P[idx] = 0.0;
    }
    __syncthreads();
    {
        constexpr int NP = JB * JB;
        constexpr int SL = PT / NP;
        int pair = tid % NP, sl = tid / NP;
        int a = pair / JB, b = pair % JB;
        double g = 0.0;
        for (int i = sl; i < nloc; i += SL) g += P[i * JB + a] * P[i * JB + b];
        gred[tid] = g;
        __syncthreads();
        if (tid < NP) {
            double s = 0.0;
            for (int q = 0; q < SL; ++q) s += gred[q * NP + tid];
            *cluster.map_shared_rank(&Gpart[rank][tid], 0) = s;
        }
    }
    cluster.sync();
    if (rank == 0) {
        constexpr int NP = JB * JB;
        if (tid < NP) {
            double s = 0.0;
            for (int r = 0; r < CL; ++r) s += Gpart[r][tid];
            Gpart[0][tid] = s;
            Ts[tid] = 0.0;
        }
        __syncthreads();
        if (tid == 0) {
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

// ---------------------------------------------------------------------------------------------------------------
// T of an outer block from the inner T_p and the Gram matrix G = V^T V of the block's reflectors:
//   T[0:J, Jblk] = -T[0:J, 0:J] * G[0:J, Jblk] * T_Jblk        (merging compact-WY factors)
__global__ void __launch_bounds__(1024, 1)
build_outer_T_kernel(const double* __restrict__ G, int ldg, const double* __restrict__ Tin, int nbw, double* __restrict__ Tout) {
    extern __shared__ __align__(16) double sm[];
    double* T = sm;                 // [NB][NB]
    double* Y = sm + NB * NB;       // [NB][JB]
    const int tid = threadIdx.x;
    for (int i = tid; i < NB * NB; i += 1024) T[i] = 0.0;
    __syncthreads();
    const int nblk = (nbw + JB - 1) / JB;
    for (int q = 0; q < nblk; ++q) {
        const int J = q * JB, jb = min(JB, nbw - J);
        const double* Tq = Tin + (size_t)q * JB * JB;
        // diagonal block
        for (int i = tid; i < JB * JB; i += 1024) {
            int r = i / JB, c = i % JB;
            if (r < jb && c < jb) T[(J + r) * NB + J + c] = Tq[r * JB + c];
        }
        // Y = T[0:J, 0:J] * G[0:J, Jblk]
        for (int i = tid; i < J * JB; i += 1024) {
            int r = i / JB, c = i % JB;
            double s = 0.0;
            if (c < jb)
                for (int k = r; k < J; ++k) s += T[r * NB + k] * G[(size_t)k * ldg + J + c];
            Y[i] = s;
        }
        __syncthreads();
        // T[0:J, Jblk] = -Y * T_q
        for (int i = tid; i < J * JB; i += 1024) {
            int r = i / JB, c = i % JB;
            if (c < jb) {
                double s = 0.0;
                for (int k = 0; k <= c; ++k) s += Y[r * JB + k] * Tq[k * JB + c];
                T[r * NB + J + c] = -s;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < NB * NB; i += 1024) Tout[i] = T[i];
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

int launch_panel(tn_ctx* ctx, cudaStream_t st, double* A, int lda, int m, int j0, int jb, double* Vall, int ldv, double* T) {
    const int m_rem = m - j0;
    if (m_rem <= CL * RT * RPT) {
        // smallest cluster that gives every row a register slot: short panels keep to one or two SMs
        int ncl = 1;
        while (ncl * RT * RPT < m_rem) ncl *= 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncl, 1, 1);
        cfg.blockDim = dim3(RT, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ncl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TN_CUDA(cudaLaunchKernelEx(&cfg, qr_panel_reg_kernel, A, lda, m, j0, jb, Vall, ldv, T));
    } else {
        const int rows_per = ceil_div(m_rem, CL);
        size_t smem = (size_t)rows_per * JB * sizeof(double);
        if (smem > 150 * 1024) {
            tn_set_error("QR panel of %d rows is too tall for the cluster panel kernels", m - j0);
            return TN_ERR_ARG;
        }
        TN_CUDA(cudaFuncSetAttribute(qr_panel_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024));
        qr_panel_smem_kernel<<<CL, PT, smem, st>>>(A, lda, m, j0, jb, Vall, ldv, T);
    }
    TN_LAUNCHED(ctx);
    return TN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Block reflector of ONE inner panel applied to the other columns of its outer block:
//   C2 <- (I - V T^T V^T) C2,   V (mr x kb, kb <= 16), C2 (mr x nc, nc <= 112).
// As library GEMMs this is W = V^T C2 (split-K + reduction), W2 = T^T W, C2 -= V W2: four launches of ~15 us each for a
// few MFLOP.  Here: (1) wy_w_kernel -- the whole grid streams V and C2 once, every CTA reduces its row chunk to a
// 16 x 32 partial, and the LAST CTA to finish a column group (ticket counter) sums the partials in a fixed order and
// multiplies by T^T; (2) wy_update_kernel -- C2 -= V W2 with W2's column in registers.  Two launches, ~10 us.
constexpr int WY_COLS = 32;        // columns of C2 per CTA
constexpr int WY_SL = 8;           // row slices per CTA (one warp each)
constexpr size_t WY_SCRATCH = (size_t)(256 + 1) * JB * NB;      // partials of up to 256 row splits + W2, nc <= NB

__global__ void __launch_bounds__(WY_COLS * WY_SL)
wy_w_kernel(int mr, int kb, int nc, const double* __restrict__ V, int ldv, const double* __restrict__ T, int ldt,
            const double* __restrict__ C2, int ldc, int rows_per_split, double* __restrict__ Wpart, int ncpad,
            unsigned int* __restrict__ counters, double* __restrict__ W2) {
    __shared__ double red[WY_SL][JB][WY_COLS + 1];
    __shared__ double wfull[JB][WY_COLS + 1];
    __shared__ int is_last;
    const int c = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int col = blockIdx.x * WY_COLS + c;
    const int split = blockIdx.y, S = gridDim.y;
    const int r0 = split * rows_per_split, r1 = min(mr, r0 + rows_per_split);
    double acc[JB];
#pragma unroll
    for (int k = 0; k < JB; ++k) acc[k] = 0.0;
    for (int i = r0 + sl; i < r1; i += WY_SL) {
        const double x = (col < nc) ? C2[(int64_t)i * ldc + col] : 0.0;
        const double* v = V + (int64_t)i * ldv;            // the same row for the whole warp: broadcast loads
#pragma unroll
        for (int k = 0; k < JB; ++k) acc[k] += ((k < kb) ? v[k] : 0.0) * x;
    }
#pragma unroll
    for (int k = 0; k < JB; ++k) red[sl][k][c] = acc[k];
    __syncthreads();
    for (int o = threadIdx.x; o < JB * WY_COLS; o += WY_COLS * WY_SL) {
        const int k = o / WY_COLS, cc = o % WY_COLS;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < WY_SL; ++q) s += red[q][k][cc];
        Wpart[((int64_t)split * JB + k) * ncpad + blockIdx.x * WY_COLS + cc] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&counters[blockIdx.x], 1u) == (unsigned)(S - 1));
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int o = threadIdx.x; o < JB * WY_COLS; o += WY_COLS * WY_SL) {
        const int k = o / WY_COLS, cc = o % WY_COLS;
        double s = 0.0;
        for (int q = 0; q < S; ++q) s += Wpart[((int64_t)q * JB + k) * ncpad + blockIdx.x * WY_COLS + cc];
        wfull[k][cc] = s;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < JB * WY_COLS; o += WY_COLS * WY_SL) {
        const int j = o / WY_COLS, cc = o % WY_COLS;
        double s = 0.0;
        if (j < kb)
            for (int k = 0; k <= j; ++k) s += T[k * ldt + j] * wfull[k][cc];      // (T^T W)[j] : T is upper triangular
        W2[(int64_t)j * ncpad + blockIdx.x * WY_COLS + cc] = s;
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;          // ready for the next call on this stream
}

__global__ void __launch_bounds__(WY_COLS * WY_SL)
wy_update_kernel(int mr, int kb, int nc, const double* __restrict__ V, int ldv, const double* __restrict__ W2, int ncpad,
                 double* __restrict__ C2, int ldc, int rows_per_cta) {
    const int c = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int col = blockIdx.x * WY_COLS + c;
    double w[JB];
#pragma unroll
    for (int k = 0; k < JB; ++k) w[k] = (k < kb) ? W2[(int64_t)k * ncpad + col] : 0.0;      // padded columns hold zeros
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(mr, r0 + rows_per_cta);
    for (int i = r0 + sl; i < r1; i += WY_SL) {
        const double* v = V + (int64_t)i * ldv;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < JB; ++k) s += ((k < kb) ? v[k] : 0.0) * w[k];
        if (col < nc) C2[(int64_t)i * ldc + col] -= s;
    }
}

// W / W2 scratch of the caller must hold (splits * JB + JB) * ncpad doubles
int apply_panel(tn_ctx* ctx, cudaStream_t st, int mr, int kb, int nc, const double* V, int ldv, const double* T, int ldt,
                double* C2, int ldc, double* scratch, size_t scratch_doubles) {
    const int groups = ceil_div(nc, WY_COLS), ncpad = groups * WY_COLS;
    // row splits: ~2 CTAs per SM when the factorisation runs alone; TN_WY_CTAS caps the grid (experiments with many
    // concurrent instances)
    static const int cta_cap = [] { const char* e = getenv("TN_WY_CTAS"); return e ? atoi(e) : 0; }();
    const int want_ctas = cta_cap > 0 ? cta_cap : (tn_throughput_mode() ? ctx->sm_count / 2 : 2 * ctx->sm_count);
    int S = ceil_div(want_ctas, groups);
    const int max_by_rows = ceil_div(mr, 2 * WY_SL);                 // at least two rows per warp
    if (S > max_by_rows) S = max_by_rows;
    if (S > 256) S = 256;
    if (S < 1) S = 1;
    const int rows_per_split = ceil_div(mr, S);
    S = ceil_div(mr, rows_per_split);
    if ((size_t)(S + 1) * JB * ncpad > scratch_doubles) {
        tn_set_error("apply_panel: scratch too small");
        return TN_ERR_ARG;
    }
    double* Wpart = scratch;
    double* W2 = scratch + (size_t)S * JB * ncpad;
    wy_w_kernel<<<dim3(groups, S), WY_COLS * WY_SL, 0, st>>>(mr, kb, nc, V, ldv, T, ldt, C2, ldc, rows_per_split, Wpart, ncpad,
                                                            (unsigned int*)ctx->counters, W2);
    TN_LAUNCHED(ctx);
    const int rows_per_cta = 64;
    wy_update_kernel<<<dim3(groups, ceil_div(mr, rows_per_cta)), WY_COLS * WY_SL, 0, st>>>(mr, kb, nc, V, ldv, W2, ncpad, C2, ldc,
                                                                                          rows_per_cta);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

// C2 <- (I - V T' V^T) C2 with T' = T or T^T; V (mr x kb, ldv), T (kb x kb, ldt), C2 (mr x nc, ldc); W, W2 scratch (kb x nc)
int apply_block(tn_ctx* ctx, cudaStream_t st, int mr, int kb, int nc, const double* V, int ldv, const double* T, int ldt,
                int transT, double* C2, int ldc, double* W, double* W2) {
    int rc;
    if ((rc = tn_gemm_impl(ctx, st, 1, 0, kb, nc, mr, 1.0, V, ldv, 0, C2, ldc, 0, 0.0, W, nc, 0, 1))) return rc;
    if ((rc = tn_gemm_impl(ctx, st, transT, 0, kb, nc, kb, 1.0, T, ldt, 0, W, nc, 0, 0.0, W2, nc, 0, 1))) return rc;
    return tn_gemm_impl(ctx, st, 0, 0, mr, nc, kb, -1.0, V, ldv, 0, W2, nc, 0, 1.0, C2, ldc, 0, 1);
}

}  // namespace

static bool fused_apply() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TN_QR_APPLY"); on = (e && e[0] == 'g') ? 0 : 1; }
    return on == 1;
}

static size_t qr_scratch_need(int m, int n) {
    const int k = m < n ? m : n;
    const int nouter = ceil_div(k, NB), npan = ceil_div(k, JB), wcols = n > k ? n : k;
    return ((size_t)m * k + (size_t)npan * JB * JB + (size_t)nouter * NB * NB + (size_t)NB * NB + 2 * (size_t)NB * wcols +
            WY_SCRATCH) * sizeof(double);
}

static int qr_body(tn_ctx* ctx, cudaStream_t st, int m, int n, double* A, int lda, double* Q, int ldq, double* R, int ldr,
                   unsigned long long* maxabs_bits) {
    const int k = m < n ? m : n;
    const int nouter = ceil_div(k, NB);
    const int wcols = n > k ? n : k;
    // scratch: V (m x k) | inner T's (JB x JB each) | outer T's (NB x NB each) | G (NB x NB) | W, W2 (NB x wcols)
    const int npan = ceil_div(k, JB);
    double* ws = (double*)tn_scratch(ctx, TN_SLOT_QR, qr_scratch_need(m, n));
    if (!ws) return TN_ERR_NOMEM;
    double* Vall = ws;
    double* Tin = Vall + (size_t)m * k;
    double* Tout = Tin + (size_t)npan * JB * JB;
    double* G = Tout + (size_t)nouter * NB * NB;
    double* W = G + (size_t)NB * NB;
    double* W2 = W + (size_t)NB * wcols;
    double* WY = W2 + (size_t)NB * wcols;
    const size_t tsmem = ((size_t)NB * NB + (size_t)NB * JB) * sizeof(double);
    TN_FUNC_ATTR_ONCE(ctx, build_outer_T_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
    int rc;
    for (int ob = 0; ob < nouter; ++ob) {
        const int J0 = ob * NB, nbw = (k - J0) < NB ? (k - J0) : NB;
        const int pan0 = J0 / JB;
        // V_outer has zeros above the diagonal of the block: the panel kernels only write rows >= their own first row
        TN_CUDA(cudaMemset2DAsync(Vall + (size_t)J0 * k + J0, (size_t)k * sizeof(double), 0, (size_t)nbw * sizeof(double), nbw, st));
        for (int jj = J0; jj < J0 + nbw; jj += JB) {
            const int jb = (J0 + nbw - jj) < JB ? (J0 + nbw - jj) : JB;
            double* Tp = Tin + (size_t)(jj / JB) * JB * JB;
            if ((rc = launch_panel(ctx, st, A, lda, m, jj, jb, Vall, k, Tp))) return rc;
            const int rem = J0 + nbw - (jj + jb);       // remaining columns of this outer block
            if (rem > 0) {
                // two fused kernels on the FP64 vector pipe (fewer launches: better when many instances share the GPU) or
                // three DMMA GEMMs + split-K reduction (shorter when the factorisation runs alone); TN_QR_APPLY=gemm|fused
                if (fused_apply()) {
                    if ((rc = apply_panel(ctx, st, m - jj, jb, rem, Vall + (size_t)jj * k + jj, k, Tp, JB,
                                          A + (size_t)jj * lda + jj + jb, lda, WY, WY_SCRATCH))) return rc;
                } else {
                    if ((rc = apply_block(ctx, st, m - jj, jb, rem, Vall + (size_t)jj * k + jj, k, Tp, JB, 1,
                                          A + (size_t)jj * lda + jj + jb, lda, W, W2))) return rc;
                }
            }
        }
        double* To = Tout + (size_t)ob * NB * NB;
        const double* Vo = Vall + (size_t)J0 * k + J0;
        const int mr = m - J0;
        if (nbw > JB) {
            // block reflector of the whole outer block
            if ((rc = tn_gemm_impl(ctx, st, 1, 0, nbw, nbw, mr, 1.0, Vo, k, 0, Vo, k, 0, 0.0, G, NB, 0, 1))) return rc;
            build_outer_T_kernel<<<1, 1024, tsmem, st>>>(G, NB, Tin + (size_t)pan0 * JB * JB, nbw, To);
            TN_LAUNCHED(ctx);
        } else {
            // single inner panel: T_outer = T_inner (embedded with leading dimension NB)
            TN_CUDA(cudaMemsetAsync(To, 0, (size_t)NB * NB * sizeof(double), st));
            TN_CUDA(cudaMemcpy2DAsync(To, NB * sizeof(double), Tin + (size_t)pan0 * JB * JB, JB * sizeof(double),
                                      JB * sizeof(double), JB, cudaMemcpyDeviceToDevice, st));
        }
        const int n2 = n - (J0 + nbw);
        if (n2 > 0) {
            if ((rc = apply_block(ctx, st, mr, nbw, n2, Vo, k, To, NB, 1, A + (size_t)J0 * lda + J0 + nbw, lda, W, W2))) return rc;
        }
    }
    // accumulate Q = H_1 ... H_p [I; 0], outer blocks backwards
    {
        int64_t total = (int64_t)m * k;
        int blocks = (int)((total + 255) / 256 < 8 * ctx->sm_count ? (total + 255) / 256 : 8 * ctx->sm_count);
        set_identity_kernel<<<blocks, 256, 0, st>>>(Q, ldq, m, k);
        TN_LAUNCHED(ctx);
    }
    for (int ob = nouter - 1; ob >= 0; --ob) {
        const int J0 = ob * NB, nbw = (k - J0) < NB ? (k - J0) : NB;
        if ((rc = apply_block(ctx, st, m - J0, nbw, k - J0, Vall + (size_t)J0 * k + J0, k, Tout + (size_t)ob * NB * NB, NB, 0,
                              Q + (size_t)J0 * ldq + J0, ldq, W, W2))) return rc;
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

extern "C" int tn_qr_pos(tn_ctx* ctx, void* stream, int m, int n, double* A, int lda, double* Q, int ldq, double* R,
                         int ldr, unsigned long long* maxabs_bits) {
    TN_REQUIRE(ctx != nullptr, "null context");
    TN_REQUIRE(m >= 1 && n >= 1, "empty matrix");
    TN_REQUIRE(lda >= n && ldq >= (m < n ? m : n) && ldr >= n, "bad leading dimension");
    // (Replaying the launch sequence as a CUDA graph was tried in both rounds: -6 % latency for a single stream, but 3x
    // LOWER throughput with 24 solver streams -- 1.59 vs 0.48 s per instance, profiles/r2c_bench_batch_sweep.txt -- so the
    // kernels are launched directly.)
    return qr_body(ctx, as_stream(stream), m, n, A, lda, Q, ldq, R, ldr, maxabs_bits);
}

#ifdef TN_PHASES
extern "C" int tn_debug_phases(long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_phase, sizeof(long long) * 16);
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof(z)); }
    return 0;
}
#endif
