// Thin SVD by one-sided (Hestenes) Jacobi with block round-robin ordering -- replaces mps.svd / mps.svd_S
// (mps.py:24-40, 62-73: LAPACK gesdd with gesvd fall-back).
//
// The matrix is held as "extended vectors" E[k] = [ w_k (length a) | j_k (length ext) ]: w_k are the rows or
// columns of C being orthogonalised, j_k the accumulated rotations (identity at start, ext = 0 when only
// singular values are wanted).  A CTA owns a pair of vector blocks in shared memory, runs cyclic Jacobi among
// them with one warp per pair (dot products reduced with warp shuffles) and writes them back; block pairs
// follow a round-robin tournament so that all pairs meet once per sweep.  Small problems are resident in one
// CTA and iterate to convergence inside a single launch.  Rotations are computed from the 2x2 Gram of the two
// vectors only, which keeps small singular values accurate to high relative precision; a Gram/eigh formulation
// of the whole matrix is deliberately NOT used (hard part 1-iii).
//
// Deflation.  The centre matrices of the boundary MPS are triangular QR factors whose singular values span
// 1 ... 1e-40 (measured on the reference at L = 2048).  Everything below eps * S0 is discarded by the caller
// (mps.py:805-806), so vectors whose norm is below DEAD_FLOOR * ||C||_F (1e-2 * eps) are never rotated: they are
// reported as exact zero singular values.  The orientation (rows or columns of C) with fewer live vectors is
// orthogonalised -- for a triangular factor that is the graded side, typically 100-200 of 512.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

// x mod n for 0 <= x < 2n without an integer division (the tournament indices are computed per pair and per phase)
__device__ __forceinline__ int wrap(int x, int n) { return x >= n ? x - n : x; }

constexpr int JT = 512;              // threads per CTA
constexpr int JW = JT / 32;
constexpr size_t SMEM_LIMIT = 200 * 1024;
constexpr double DEAD_FLOOR = 2.2e-18;
constexpr int MAX_SWEEPS = 60;

struct SvdMeta {
    int use_rows;      // 1: the rows of C are orthogonalised, 0: the columns
    int nlive;
    double fro2;
};

// squared norms of all rows (blocks 0..m-1) and columns (blocks m..m+n-1)
__global__ void svd_norms_kernel(const double* __restrict__ C, int ldc, int m, int n, double* __restrict__ norms2) {
    __shared__ double red[8];
    int b = blockIdx.x;
    double s = 0.0;
    if (b < m) {
        for (int j = threadIdx.x; j < n; j += blockDim.x) { double v = C[(int64_t)b * ldc + j]; s += v * v; }
    } else {
        int j = b - m;
        for (int i = threadIdx.x; i < m; i += blockDim.x) { double v = C[(int64_t)i * ldc + j]; s += v * v; }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        norms2[b] = t;
    }
}

// single thread block: Frobenius norm, live flags, orientation, compact index list
__global__ void svd_select_kernel(const double* __restrict__ norms2, int m, int n, int force_rows, int deflate,
                                  SvdMeta* meta, int* __restrict__ live_idx) {
    __shared__ double fro2_s;
    __shared__ int cnt_s[2];
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < m; ++i) t += norms2[i];
        fro2_s = t;
        cnt_s[0] = cnt_s[1] = 0;
    }
    __syncthreads();
    const double floor_cnt = DEAD_FLOOR * DEAD_FLOOR * fro2_s;      // orientation: the side with fewer live vectors
    const double floor2 = deflate ? floor_cnt : -1.0;               // liveness: everything when not deflating
    int cr = 0, cc = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) cr += (norms2[i] > floor_cnt);
    for (int j = threadIdx.x; j < n; j += blockDim.x) cc += (norms2[m + j] > floor_cnt);
    atomicAdd(&cnt_s[0], cr);
    atomicAdd(&cnt_s[1], cc);
    __syncthreads();
    if (threadIdx.x == 0) {
        int use_rows = (force_rows >= 0) ? force_rows : (cnt_s[0] <= cnt_s[1]);
        int cnt = 0;
        if (use_rows) { for (int i = 0; i < m; ++i) if (norms2[i] > floor2) live_idx[cnt++] = i; }
        else { for (int j = 0; j < n; ++j) if (norms2[m + j] > floor2) live_idx[cnt++] = j; }
        meta->use_rows = use_rows;
        meta->nlive = cnt;
        meta->fro2 = fro2_s;
    }
}

// E[e] = [ live vector e of C | unit vector e ]
__global__ void jacobi_init_kernel(const double* __restrict__ C, int ldc, const SvdMeta* __restrict__ meta,
                                   const int* __restrict__ live_idx, double* __restrict__ E, int ldw, int a, int ext, int nc) {
    const int use_rows = meta->use_rows;
    int64_t total = (int64_t)nc * ldw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = (int)(i / ldw), e = (int)(i % ldw);
        int src = live_idx[k];
        double v;
        if (e < a) v = use_rows ? C[(int64_t)src * ldc + e] : C[(int64_t)e * ldc + src];
        else v = (e - a == k) ? 1.0 : 0.0;
        E[i] = v;
    }
}

// one tournament round over block pairs; CTA b handles the pair given by the circle method
__global__ void __launch_bounds__(JT, 1)
jacobi_round_kernel(double* __restrict__ E, int ldw, int a, int nc, int bsz, int nblocks, int round, int inner_max,
                    double tol, const SvdMeta* __restrict__ meta, unsigned int* rot_count) {
    extern __shared__ __align__(16) double S[];      // [vectors here][ldw]
    __shared__ int col_of[512];
    __shared__ unsigned int sweep_rot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npe = nblocks + (nblocks & 1), n1 = npe - 1;
    int b0, b1;
    if (npe <= 2) { b0 = 0; b1 = 1; }
    else {
        int i = blockIdx.x;
        b0 = (round + i) % n1;
        b1 = (i == 0) ? n1 : (round + n1 - i) % n1;
    }
    if (b0 >= nblocks || b1 >= nblocks) return;      // partner is the phantom block of an odd tournament
    const double floor2 = DEAD_FLOOR * DEAD_FLOOR * meta->fro2;
    int c0 = b0 * bsz, n0 = min(bsz, nc - c0);
    int c1 = b1 * bsz, n1c = min(bsz, nc - c1);
    const int ncol = n0 + n1c;
    for (int i = tid; i < ncol; i += JT) col_of[i] = (i < n0) ? c0 + i : c1 + (i - n0);
    __syncthreads();
    for (int64_t i = tid; i < (int64_t)ncol * ldw; i += JT) {
        int k = (int)(i / ldw), e = (int)(i % ldw);
        S[i] = E[(int64_t)col_of[k] * ldw + e];
    }
    __syncthreads();

    const int nce = ncol + (ncol & 1), r1 = nce - 1, half = nce / 2;
    unsigned int my_rot_total = 0;
    for (int sweep = 0; sweep < inner_max; ++sweep) {
        if (tid == 0) sweep_rot = 0;
        __syncthreads();
        unsigned int my_rot = 0;
        for (int r = 0; r < (nce > 1 ? r1 : 0); ++r) {
            for (int i = warp; i < half; i += JW) {
                int p = wrap(r + i, r1);
                int q = (i == 0) ? r1 : wrap(r + r1 - i, r1);
                if (p >= ncol || q >= ncol) continue;
                if (p > q) { int t = p; p = q; q = t; }
                double* xp = S + (int64_t)p * ldw;
                double* xq = S + (int64_t)q * ldw;
                double app = 0.0, aqq = 0.0, apq = 0.0;
                for (int e = lane; e < a; e += 32) {
                    double u = xp[e], v = xq[e];
                    app += u * u; aqq += v * v; apq += u * v;
                }
                app = warp_sum(app); aqq = warp_sum(aqq); apq = warp_sum(apq);
                if (app > floor2 && aqq > floor2 && apq * apq > tol * tol * app * aqq) {
                    // rotation of the smaller angle from two reciprocal square roots (see jacobi_cluster_kernel)
                    const double d = aqq - app, h = 2.0 * apq;
                    const double rr = rsqrt(d * d + h * h);
                    const double x = 0.5 + 0.5 * (fabs(d) * rr);
                    const double y = rsqrt(x);
                    const double cs = x * y, sn = 0.5 * (copysign(1.0, d) * h * rr) * y;
                    for (int e = lane; e < ldw; e += 32) {
                        double u = xp[e], v = xq[e];
                        xp[e] = cs * u - sn * v;
                        xq[e] = sn * u + cs * v;
                    }
                    if (lane == 0) my_rot++;
                }
            }
            __syncthreads();
        }
        if (lane == 0 && my_rot) atomicAdd(&sweep_rot, my_rot);
        my_rot_total += my_rot;
        __syncthreads();
        unsigned int done = sweep_rot;
        __syncthreads();
        if (done == 0) break;
    }
    for (int64_t i = tid; i < (int64_t)ncol * ldw; i += JT) {
        int k = (int)(i / ldw), e = (int)(i % ldw);
        E[(int64_t)col_of[k] * ldw + e] = S[i];
    }
    if (lane == 0 && my_rot_total) atomicAdd(rot_count, my_rot_total);
}

// ------------------------------------------------------------------------------------------------
// Cluster-resident Jacobi: ALL vectors stay in the shared memory of one 8-CTA cluster for the whole iteration;
// every vector is split along its length, CTA r holding slice r of the w-part and slice r of the j-part of every
// vector.  One round of the tournament = (1) every CTA computes the partial 2x2 Gram of every pair on its slice and
// sends it to the pair's owner CTA through DSMEM, (2) the owner sums the 8 partials in a fixed order, decides the
// rotation and broadcasts (c, s), (3) every CTA rotates its slices.  Two cluster barriers per round and no kernel
// launch or host read-back until convergence.
constexpr int CLJ = 8;                                // portable cluster size
constexpr int CLJ_MAX = 16;                           // opt-in (non-portable) size: twice the shared memory, half the traffic per CTA
constexpr int CJ_MAXPAIRS = 256;                      // nc <= 512
constexpr int CJ_MAXOWN = CJ_MAXPAIRS / CLJ;

__global__ void __launch_bounds__(JT, 1)
jacobi_cluster_kernel(double* __restrict__ E, int ldw, int a, int ext, int nc, int max_sweeps, double tol,
                      const SvdMeta* __restrict__ meta, int* __restrict__ status /* [0] = sweeps, [1] = converged */,
                      int vglob /* 1: the accumulated right vectors stay in global memory (L2), only the w-parts are resident */) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int ncl = (int)cluster.num_blocks();        // 8 or 16 (launch attribute)
    const int ncl_log2 = (ncl == 16) ? 4 : (ncl == 8 ? 3 : 2);   // 4-CTA clusters: throughput mode (half the SMs per SVD)
    extern __shared__ __align__(16) double Sl[];      // [nc][ll]
    __shared__ double part[CLJ_MAX][CJ_MAXOWN][3];
    __shared__ double rot[CJ_MAXPAIRS][2];
    __shared__ unsigned int cnt[2][CLJ_MAX];
    __shared__ unsigned int myrot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int aw = (a + ncl - 1) / ncl, ej = (ext + ncl - 1) / ncl, ll = vglob ? aw : aw + ej;
    const int lls = ll | 1;                                       // odd row stride: pair rows fall on different banks
    const int wlo = rank * aw, jlo = rank * ej;
    const int ejv = max(0, min(ej, ext - jlo));                   // my valid part of the j-slice (vglob path)
    const double floor2 = DEAD_FLOOR * DEAD_FLOOR * meta->fro2;
    for (int64_t i = tid; i < (int64_t)nc * ll; i += JT) {
        int k = (int)(i / ll), e = (int)(i % ll);
        double v = 0.0;
        if (e < aw) { if (wlo + e < a) v = E[(int64_t)k * ldw + wlo + e]; }
        else { int f = e - aw; if (jlo + f < ext) v = E[(int64_t)k * ldw + a + jlo + f]; }
        Sl[(int64_t)k * lls + e] = v;
    }
    if (tid == 0) myrot = 0;
    cluster.sync();      // all CTAs started (required before the first distributed-shared-memory access)
    const int nce = nc + (nc & 1), r1 = nce - 1, half = nce / 2;
    const int sub = lane & 7, grp = lane >> 3;                    // 8 lanes per pair, 4 pairs per warp
    int sweep = 0, converged = 0;
    for (; sweep < max_sweeps && r1 >= 1; ++sweep) {
        for (int r = 0; r < r1; ++r) {
            // ---- (1) partial Gram triples on my w-slice
            for (int i0 = warp * 4; i0 < half; i0 += JW * 4) {
                int i = i0 + grp;
                double app = 0.0, aqq = 0.0, apq = 0.0;
                int p = 0, q = 0;
                bool valid = (i < half);
                if (valid) {
                    p = wrap(r + i, r1);
                    q = (i == 0) ? r1 : wrap(r + r1 - i, r1);
                    valid = (p < nc) && (q < nc);
                }
                if (valid) {
                    const double* xp = Sl + (int64_t)p * lls;
                    const double* xq = Sl + (int64_t)q * lls;
                    for (int e = sub; e < aw; e += 8) {
                        double u = xp[e], v = xq[e];
                        app += u * u; aqq += v * v; apq += u * v;
                    }
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    app += __shfl_xor_sync(0xffffffffu, app, o);
                    aqq += __shfl_xor_sync(0xffffffffu, aqq, o);
                    apq += __shfl_xor_sync(0xffffffffu, apq, o);
                }
                if (sub == 0 && i < half) {
                    double* dst = cluster.map_shared_rank(&part[rank][i >> ncl_log2][0], i & (ncl - 1));
                    dst[0] = app; dst[1] = aqq; dst[2] = apq;
                }
            }
            cluster.sync();
            // ---- (2) owners decide the rotations and broadcast them
            {
                int i = tid * ncl + rank;                          // pairs owned by this CTA: i % ncl == rank
                if (tid < CJ_MAXOWN && i < half) {
                    // partial sums added pairwise (a three-level tree for 8 CTAs): the FP64 pipe of this part is slow,
                    // so the length of the dependent chain is what costs
                    double app = 0.0, aqq = 0.0, apq = 0.0;
                    for (int src = 0; src < ncl; src += 4) {
                        const double a0 = part[src][tid][0] + part[src + 1][tid][0], a1 = part[src + 2][tid][0] + part[src + 3][tid][0];
                        const double b0 = part[src][tid][1] + part[src + 1][tid][1], b1 = part[src + 2][tid][1] + part[src + 3][tid][1];
                        const double c0 = part[src][tid][2] + part[src + 1][tid][2], c1 = part[src + 2][tid][2] + part[src + 3][tid][2];
                        app += a0 + a1; aqq += b0 + b1; apq += c0 + c1;
                    }
                    double cs = 1.0, sn = 0.0;
                    if (app > floor2 && aqq > floor2 && apq * apq > tol * tol * app * aqq) {
                        // Jacobi rotation of the smaller angle from two reciprocal square roots (no division):
                        //   cos 2t = |d| / sqrt(d^2 + h^2),  sin 2t = sgn(d) h / sqrt(d^2 + h^2),   d = aqq - app, h = 2 apq
                        //   cos t = sqrt((1 + cos 2t) / 2) = x rsqrt(x),   sin t = sin 2t / (2 cos t) = sin 2t rsqrt(x) / 2
                        // identical in exact arithmetic to t = sgn(z) / (|z| + sqrt(1 + z^2)), z = d / h (the textbook form)
                        const double d = aqq - app, h = 2.0 * apq;
                        const double r = rsqrt(d * d + h * h);
                        const double x = 0.5 + 0.5 * (fabs(d) * r);
                        const double y = rsqrt(x);
                        cs = x * y;
                        sn = 0.5 * (copysign(1.0, d) * h * r) * y;
                        atomicAdd(&myrot, 1u);
                    }
                    for (int dstc = 0; dstc < ncl; ++dstc) {
                        double* dst = cluster.map_shared_rank(&rot[i][0], dstc);
                        dst[0] = cs; dst[1] = sn;
                    }
                }
            }
            cluster.sync();
            // ---- (3) rotate my slices of every pair
            for (int i0 = warp * 4; i0 < half; i0 += JW * 4) {
                int i = i0 + grp;
                if (i >= half) continue;
                const double cs = rot[i][0], sn = rot[i][1];
                if (sn == 0.0) continue;
                int p = wrap(r + i, r1);
                int q = (i == 0) ? r1 : wrap(r + r1 - i, r1);
                double* xp = Sl + (int64_t)p * lls;
                double* xq = Sl + (int64_t)q * lls;
                if (!vglob) {
                    for (int e = sub; e < ll; e += 8) {
                        double u = xp[e], v = xq[e];
                        xp[e] = cs * u - sn * v;
                        xq[e] = sn * u + cs * v;
                    }
                } else {
                    // my slice of the right-vector parts lives in global memory (L2): no other CTA touches it, and the
                    // block barrier below orders it for the warp that rotates these vectors next round.  The loads
                    // of a chunk are issued before the shared-memory rotation so that the L2 round trip overlaps it.
                    double* gp = E + (int64_t)p * ldw + a + jlo;
                    double* gq = E + (int64_t)q * ldw + a + jlo;
                    double gu[4], gv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = sub + 8 * k;
                        if (e < ejv) { gu[k] = gp[e]; gv[k] = gq[e]; }
                    }
                    for (int e = sub; e < ll; e += 8) {
                        double u = xp[e], v = xq[e];
                        xp[e] = cs * u - sn * v;
                        xq[e] = sn * u + cs * v;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = sub + 8 * k;
                        if (e < ejv) { gp[e] = cs * gu[k] - sn * gv[k]; gq[e] = sn * gu[k] + cs * gv[k]; }
                    }
                    for (int e = sub + 32; e < ejv; e += 8) {          // slices longer than 32 (not reached for k <= 512)
                        double u = gp[e], v = gq[e];
                        gp[e] = cs * u - sn * v;
                        gq[e] = sn * u + cs * v;
                    }
                }
            }
            __syncthreads();
        }
        // ---- end of sweep: all CTAs learn the total number of rotations
        const int buf = sweep & 1;
        if (tid < ncl) {
            unsigned int* dst = cluster.map_shared_rank(&cnt[buf][rank], tid);
            *dst = myrot;
        }
        cluster.sync();
        unsigned int total = 0;
        for (int src = 0; src < ncl; ++src) total += cnt[buf][src];
        __syncthreads();
        if (tid == 0) myrot = 0;
        __syncthreads();
        if (total == 0) { converged = 1; ++sweep; break; }
    }
    if (r1 < 1) converged = 1;
    for (int64_t i = tid; i < (int64_t)nc * ll; i += JT) {
        int k = (int)(i / ll), e = (int)(i % ll);
        const double x = Sl[(int64_t)k * lls + e];
        if (e < aw) { if (wlo + e < a) E[(int64_t)k * ldw + wlo + e] = x; }
        else { int f = e - aw; if (jlo + f < ext) E[(int64_t)k * ldw + a + jlo + f] = x; }
    }
    if (rank == 0 && tid == 0) { status[0] = sweep; status[1] = converged; }
}


// ------------------------------------------------------------------------------------------------
// Cluster-resident Jacobi, second generation: ONE cluster barrier per tournament round.
// The vectors are split along their length over the ncl CTAs of the cluster exactly as above (ncl = 1, 2, 4, 8 or 16,
// the smallest size whose shared memory holds the problem; ncl = 1 runs without any cluster traffic).  Per round every
// CTA computes the partial 2x2 Gram of every pair on its slice and sends the triple to ALL CTAs (all-gather through
// DSMEM, double-buffered by round parity); after the barrier every CTA sums the ncl partials in the same fixed order
// and derives the rotation itself -- bit-identical decisions everywhere, so there is no owner, no broadcast, no
// second barrier and no exchange of rotation counts.  Eight lanes work on a pair (four pairs per warp); the lanes
// that computed a pair's partial Gram also rotate it, so the only block barrier of a round is the one that
// publishes the rotated vectors to the warps that meet them next.
constexpr int JT2 = 1024;
constexpr int J2_SLOTS = JT2 / 8;                     // pairs processed concurrently by one CTA

__global__ void __launch_bounds__(JT2, 1)
jacobi_cluster2_kernel(double* __restrict__ E, int ldw, int a, int ext, int nc, int max_sweeps, double tol,
                       const SvdMeta* __restrict__ meta, int* __restrict__ status /* [0] = sweeps, [1] = converged */) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int ncl = (int)cluster.num_blocks();
    extern __shared__ __align__(16) double Sl[];      // [nc][lls] | part[2][ncl][half][3]
    __shared__ unsigned int rotcnt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int aw = (a + ncl - 1) / ncl, ej = (ext + ncl - 1) / ncl, ll = aw + ej;
    const int lls = ll | 1;                                       // odd row stride: pair rows fall on different banks
    const int wlo = rank * aw, jlo = rank * ej;
    const int nce = nc + (nc & 1), r1 = nce - 1, half = nce / 2;
    double* part = Sl + (size_t)nc * lls;
    const double floor2 = DEAD_FLOOR * DEAD_FLOOR * meta->fro2;
    const double tol2 = tol * tol;
    for (int64_t i = tid; i < (int64_t)nc * ll; i += JT2) {
        int k = (int)(i / ll), e = (int)(i % ll);
        double v = 0.0;
        if (e < aw) { if (wlo + e < a) v = E[(int64_t)k * ldw + wlo + e]; }
        else { int f = e - aw; if (jlo + f < ext) v = E[(int64_t)k * ldw + a + jlo + f]; }
        Sl[(int64_t)k * lls + e] = v;
    }
    if (tid == 0) rotcnt = 0;
    // every CTA of the cluster must have started before anyone writes into its shared memory
    if (ncl > 1) cluster.sync(); else __syncthreads();
    const int sub = lane & 7, slot = warp * 4 + (lane >> 3);
    int sweep = 0, converged = (r1 < 1), it = 0;
    for (; sweep < max_sweeps && r1 >= 1; ++sweep) {
        unsigned int my_rot = 0;
        for (int r = 0; r < r1; ++r, ++it) {
            double* pbuf = part + (size_t)(it & 1) * ncl * half * 3;
            if (ncl > 1) {
                // ---- (1) partial Gram triples on my w-slice, all-gathered
                for (int i0 = 0; i0 < half; i0 += J2_SLOTS) {
                    const int i = i0 + slot;
                    bool valid = (i < half);
                    int p = 0, q = 0;
                    if (valid) {
                        p = wrap(r + i, r1);
                        q = (i == 0) ? r1 : wrap(r + r1 - i, r1);
                        valid = (p < nc) && (q < nc);
                    }
                    double app = 0.0, aqq = 0.0, apq = 0.0;
                    if (valid) {
                        const double* xp = Sl + (int64_t)p * lls;
                        const double* xq = Sl + (int64_t)q * lls;
                        for (int e = sub; e < aw; e += 8) {
                            const double u = xp[e], v = xq[e];
                            app += u * u; aqq += v * v; apq += u * v;
                        }
                    }
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) {
                        app += __shfl_xor_sync(0xffffffffu, app, o);
                        aqq += __shfl_xor_sync(0xffffffffu, aqq, o);
                        apq += __shfl_xor_sync(0xffffffffu, apq, o);
                    }
                    if (i < half) {
                        for (int d = sub; d < ncl; d += 8) {
                            double* dst = cluster.map_shared_rank(pbuf + ((size_t)rank * half + i) * 3, d);
                            dst[0] = app; dst[1] = aqq; dst[2] = apq;
                        }
                    }
                }
                cluster.sync();
            }
            // ---- (2) every CTA sums the partials in the same order, decides the rotation and applies it to its slices
            for (int i0 = 0; i0 < half; i0 += J2_SLOTS) {
                const int i = i0 + slot;
                bool valid = (i < half);
                int p = 0, q = 0;
                if (valid) {
                    p = wrap(r + i, r1);
                    q = (i == 0) ? r1 : wrap(r + r1 - i, r1);
                    valid = (p < nc) && (q < nc);
                }
                double* xp = Sl + (int64_t)p * lls;
                double* xq = Sl + (int64_t)q * lls;
                double app = 0.0, aqq = 0.0, apq = 0.0;
                if (ncl > 1) {
                    if (valid) {
                        for (int src = sub; src < ncl; src += 8) {       // ncl = 16: lane adds sources sub and sub + 8
                            const double* t3 = pbuf + ((size_t)src * half + i) * 3;
                            app += t3[0]; aqq += t3[1]; apq += t3[2];
                        }
                    }
                } else if (valid) {
                    for (int e = sub; e < aw; e += 8) {
                        const double u = xp[e], v = xq[e];
                        app += u * u; aqq += v * v; apq += u * v;
                    }
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    app += __shfl_xor_sync(0xffffffffu, app, o);
                    aqq += __shfl_xor_sync(0xffffffffu, aqq, o);
                    apq += __shfl_xor_sync(0xffffffffu, apq, o);
                }
                if (valid && app > floor2 && aqq > floor2 && apq * apq > tol2 * app * aqq) {
                    // rotation of the smaller angle from two reciprocal square roots (see jacobi_cluster_kernel)
                    const double d = aqq - app, h = 2.0 * apq;
                    const double rr = rsqrt(d * d + h * h);
                    const double x = 0.5 + 0.5 * (fabs(d) * rr);
                    const double y = rsqrt(x);
                    const double cs = x * y, sn = 0.5 * (copysign(1.0, d) * h * rr) * y;
                    for (int e = sub; e < ll; e += 8) {
                        const double u = xp[e], v = xq[e];
                        xp[e] = cs * u - sn * v;
                        xq[e] = sn * u + cs * v;
                    }
                    if (sub == 0) my_rot++;
                }
            }
            __syncthreads();
        }
        // ---- end of sweep: every CTA took the same decisions, so the block-local count is the global one
        if (my_rot) atomicAdd(&rotcnt, my_rot);
        __syncthreads();
        const unsigned int total = rotcnt;
        __syncthreads();
        if (tid == 0) rotcnt = 0;
        if (total == 0) { converged = 1; ++sweep; break; }
    }
    __syncthreads();
    for (int64_t i = tid; i < (int64_t)nc * ll; i += JT2) {
        int k = (int)(i / ll), e = (int)(i % ll);
        const double x = Sl[(int64_t)k * lls + e];
        if (e < aw) { if (wlo + e < a) E[(int64_t)k * ldw + wlo + e] = x; }
        else { int f = e - aw; if (jlo + f < ext) E[(int64_t)k * ldw + a + jlo + f] = x; }
    }
    if (rank == 0 && tid == 0) { status[0] = sweep; status[1] = converged; }
    // no CTA may exit while another one can still write partials into its shared memory
    if (ncl > 1) cluster.sync();
}

// norms -> S (sorted descending, dead vectors = exact zeros at the end), singular vectors with the sign rule of
// mps.svd (mps.py:35-39).  U and Vt must be zero-filled by the caller.
__global__ void __launch_bounds__(JT, 1)
jacobi_finish_kernel(const double* __restrict__ E, int ldw, int a, int ext, int nc, int kfull, int has_dead,
                     const SvdMeta* __restrict__ meta, const int* __restrict__ live_idx, double* __restrict__ U, int ldu,
                     double* __restrict__ Sout, double* __restrict__ Vt, int ldvt, double* sv_tmp, int* rank_tmp) {
    // Every CTA of the grid recomputes all norms and ranks (cheap, and it avoids a second launch); the CTAs write
    // identical values to sv_tmp / rank_tmp / Sout and each one reads back only what it wrote itself.  The singular
    // vectors -- the bulk of the output traffic -- are then split over the CTAs.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int use_rows = meta->use_rows;
    for (int k = warp; k < nc; k += JW) {
        const double* w = E + (int64_t)k * ldw;
        double s = 0.0;
        for (int e = lane; e < a; e += 32) s += w[e] * w[e];     // entries are O(1) after nfactor: no scaling needed
        s = warp_sum(s);
        if (lane == 0) sv_tmp[k] = sqrt(s);
    }
    for (int k = nc + tid; k < kfull; k += JT) Sout[k] = 0.0;
    __syncthreads();
    for (int k = tid; k < nc; k += JT) {
        double sk = sv_tmp[k];
        int r = 0;
        for (int j = 0; j < nc; ++j) {
            double sj = sv_tmp[j];
            r += (sj > sk) || (sj == sk && j < k);
        }
        rank_tmp[k] = r;
        Sout[r] = sk;
    }
    __syncthreads();
    if (ext == 0) return;
    for (int k = blockIdx.x * JW + warp; k < nc; k += gridDim.x * JW) {
        const double* w = E + (int64_t)k * ldw;
        const double* jv = w + a;
        const double sk = sv_tmp[k];
        const double inv = (sk > 0.0) ? 1.0 / sk : 0.0;
        const int r = rank_tmp[k];
        // sign rule: flip when |min| > max holds for both the left and the right vector (zeros of dead entries count)
        double mxw = -INFINITY, mnw = INFINITY, mxj = has_dead ? 0.0 : -INFINITY, mnj = -mxj;
        for (int e = lane; e < a; e += 32) { double v = w[e] * inv; mxw = fmax(mxw, v); mnw = fmin(mnw, v); }
        for (int e = lane; e < ext; e += 32) { double v = jv[e]; mxj = fmax(mxj, v); mnj = fmin(mnj, v); }
        mxw = warp_max(mxw); mnw = warp_min(mnw); mxj = warp_max(mxj); mnj = warp_min(mnj);
        const double sg = ((fabs(mnw) > mxw) && (fabs(mnj) > mxj)) ? -1.0 : 1.0;
        if (use_rows) {
            // C = J S W^T: U[:, r] = j-part scattered through the live row list, Vt[r, :] = w / s
            for (int e = lane; e < ext; e += 32) U[(int64_t)live_idx[e] * ldu + r] = sg * jv[e];
            for (int e = lane; e < a; e += 32) Vt[(int64_t)r * ldvt + e] = sg * w[e] * inv;
        } else {
            for (int e = lane; e < a; e += 32) U[(int64_t)e * ldu + r] = sg * w[e] * inv;
            for (int e = lane; e < ext; e += 32) Vt[(int64_t)r * ldvt + live_idx[e]] = sg * jv[e];
        }
    }
}

__global__ void truncation_rank_kernel(const double* __restrict__ S, int k, double tol, int Dmax, int* keep_out,
                                       double* disc_out) {
    // single warp; k is at most a few thousand
    int lane = threadIdx.x;
    double s0 = S[0];
    int cnt = 0;
    for (int i = lane; i < k; i += 32) cnt += (S[i] > s0 * tol);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    int keep = min(cnt, Dmax);
    double ss = 0.0;
    for (int i = keep + lane; i < k; i += 32) ss += S[i] * S[i];
    ss = warp_sum(ss);
    if (lane == 0) { *keep_out = keep; *disc_out = sqrt(ss) / s0; }
}

}  // namespace

static int g_throughput_mode = -1;
int tn_throughput_mode() {
    if (g_throughput_mode < 0) { const char* e = getenv("TN_THROUGHPUT"); g_throughput_mode = (e && e[0] == '1') ? 1 : 0; }
    return g_throughput_mode;
}
extern "C" int tn_set_throughput_mode(int on) { g_throughput_mode = on ? 1 : 0; return TN_OK; }

// 0 = portable 8-CTA clusters only, 1 = 16-CTA clusters where 8 do not fit (default), 2 = also prefer 16 CTAs for
// every problem with >= 96 live vectors (environment TN_SVD_WIDE = 0 / 1 / 2, read once)
static int wide_clusters() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("TN_SVD_WIDE");
        mode = (e && e[0] >= '0' && e[0] <= '2') ? (e[0] - '0') : 1;
    }
    return mode;
}

// can this device co-schedule one 16-CTA cluster of the Jacobi kernel at its largest shared-memory footprint?
static bool wide_launchable() {
    static int ok = -1;
    if (ok < 0) {
        const size_t CL_SMEM = 210 * 1024;
        cudaFuncSetAttribute(jacobi_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMEM);
        cudaFuncSetAttribute(jacobi_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16, 1, 1);
        cfg.blockDim = dim3(JT, 1, 1);
        cfg.dynamicSmemBytes = CL_SMEM;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 16;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, jacobi_cluster_kernel, &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
        ok = n > 0 ? 1 : 0;
    }
    return ok == 1;
}

extern "C" int tn_svd(tn_ctx* ctx, void* stream, int m, int n, const double* C, int ldc, double* U, int ldu, double* S,
                      double* Vt, int ldvt, int want_vectors, int* h_sweeps) {
    TN_REQUIRE(ctx != nullptr, "null context");
    TN_REQUIRE(m >= 1 && n >= 1, "empty matrix");
    TN_REQUIRE(m <= 4096 && n <= 4096, "matrix too large for the Jacobi SVD");
    cudaStream_t st = as_stream(stream);
    const int kfull = m < n ? m : n;

    // ---- norms, orientation, live vectors
    size_t head_bytes = ((size_t)(m + n) * sizeof(double) + (size_t)(m + n) * sizeof(int) + sizeof(SvdMeta) + 64 + 15) & ~(size_t)15;
    char* head = (char*)tn_scratch(ctx, TN_SLOT_MISC, head_bytes + 256);
    if (!head) return TN_ERR_NOMEM;
    double* norms2 = (double*)head;
    int* live_idx = (int*)(norms2 + (m + n));
    SvdMeta* meta = (SvdMeta*)(((uintptr_t)(live_idx + (m + n)) + 15) & ~(uintptr_t)15);
    unsigned int* rot = (unsigned int*)(meta + 1);
    svd_norms_kernel<<<m + n, 128, 0, st>>>(C, ldc, m, n, norms2);
    TN_LAUNCHED(ctx);
    // a wide or tall matrix is always handled on its short side (k vectors of the long length); for a square one
    // the side with fewer live vectors is chosen on the device
    const int force = (m < n) ? 1 : (m > n ? 0 : -1);
    const int a = (m < n) ? n : m;
    // small problems: everything resident in one CTA or one cluster without deflation -> no host read-back
    const size_t CL_SMEM = 210 * 1024;
    auto cluster_fits = [&](int nvec, int e, int ncl = CLJ) {
        size_t ll = ((size_t)ceil_div(a, ncl) + (size_t)ceil_div(e, ncl)) | 1;
        return nvec >= 2 && nvec <= 2 * CJ_MAXPAIRS && (size_t)nvec * ll * sizeof(double) <= CL_SMEM;
    };
    // second-generation kernel: smallest cluster (1, 2, 4, 8, 16 CTAs) whose shared memory holds nvec vectors + partials
    static const bool v1_only = [] { const char* e = getenv("TN_SVD_V1"); return e && e[0] == '1'; }();
    auto fit2 = [&](int nvec, int e) -> int {
        if (nvec < 2 || v1_only) return 0;
        const int halfp = (nvec + 1) / 2;
        // measured on B200 (profiles/r2a_*): the non-tensor FP64 pipe issues ~15 FMA per clock and SM, so the redundant
        // per-pair scalar work of this kernel only pays off while there is no cluster traffic at all (one CTA, <= 64
        // vectors); larger problems use the owner-computes cluster kernel above
        if (nvec > 64) return 0;
        for (int ncl = 1; ncl <= 1; ncl *= 2) {
            size_t ll = ((size_t)ceil_div(a, ncl) + (size_t)ceil_div(e, ncl)) | 1;
            size_t bytes = (size_t)nvec * ll * sizeof(double) + (ncl > 1 ? (size_t)2 * ncl * halfp * 3 * sizeof(double) : 0);
            if (bytes <= CL_SMEM) return ncl;
        }
        return 0;
    };
    const int ext_full = want_vectors ? kfull : 0;
    const int ldw_full = a + ext_full;
    const bool tiny = kfull <= 32 && (size_t)kfull * ldw_full * sizeof(double) <= SMEM_LIMIT;
    // above 64 vectors the read-back for deflation pays for itself (graded factors have 3-5x fewer live vectors)
    const bool small = tiny || (kfull <= 64 && (fit2(kfull, ext_full) > 0 || cluster_fits(kfull, ext_full)));
    svd_select_kernel<<<1, 256, 0, st>>>(norms2, m, n, force, small ? 0 : 1, meta, live_idx);
    TN_LAUNCHED(ctx);
    int nc = kfull;
    if (!small) {
        // the launch geometry depends on the number of live vectors: one host read-back
        SvdMeta* hmeta = (SvdMeta*)((char*)ctx->pinned + 256);
        TN_CUDA(cudaMemcpyAsync(hmeta, meta, sizeof(SvdMeta), cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaStreamSynchronize(st));
        nc = hmeta->nlive;
    }
    const int ext = want_vectors ? nc : 0;
    const int ldw = a + ext;
    const int has_dead = nc < kfull;

    if (want_vectors) {
        TN_CUDA(cudaMemset2DAsync(U, (size_t)ldu * sizeof(double), 0, (size_t)kfull * sizeof(double), m, st));
        TN_CUDA(cudaMemset2DAsync(Vt, (size_t)ldvt * sizeof(double), 0, (size_t)n * sizeof(double), kfull, st));
    }
    int sweeps = 0;
    if (nc == 0) {
        TN_CUDA(cudaMemsetAsync(S, 0, (size_t)kfull * sizeof(double), st));
        if (h_sweeps) *h_sweeps = 0;
        return TN_OK;
    }
    TN_REQUIRE((size_t)2 * ldw * sizeof(double) <= SMEM_LIMIT, "vectors too long for the shared-memory Jacobi kernel");
    size_t bytes = (size_t)nc * ldw * sizeof(double) + (size_t)nc * (sizeof(double) + sizeof(int)) + 64;
    char* ws = (char*)tn_scratch(ctx, TN_SLOT_SVD, bytes);
    if (!ws) return TN_ERR_NOMEM;
    double* E = (double*)ws;
    double* sv_tmp = E + (size_t)nc * ldw;
    int* rank_tmp = (int*)(sv_tmp + nc);
    {
        int64_t total = (int64_t)nc * ldw;
        int blocks = (int)((total + 255) / 256 < 8 * ctx->sm_count ? (total + 255) / 256 : 8 * ctx->sm_count);
        jacobi_init_kernel<<<blocks, 256, 0, st>>>(C, ldc, meta, live_idx, E, ldw, a, ext, nc);
        TN_LAUNCHED(ctx);
    }
    int cmax = (int)(SMEM_LIMIT / ((size_t)ldw * sizeof(double)));
    cmax -= (cmax & 1);
    if (cmax > 512) cmax = 512;
    const double tol = sqrt((double)a) * 2.220446049250313e-16;
    unsigned int* h_rot = (unsigned int*)ctx->pinned;
    int* status = (int*)(rot + 2);
    if (nc <= 32 && nc <= cmax) {
        // tiny: one CTA iterates to convergence
        if (nc > 1) {
            int bsz = (nc + 1) / 2;
            size_t smem = (size_t)nc * ldw * sizeof(double);
            TN_FUNC_ATTR_ONCE(ctx, jacobi_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
            TN_CUDA(cudaMemsetAsync(rot, 0, sizeof(unsigned int), st));
            jacobi_round_kernel<<<1, JT, smem, st>>>(E, ldw, a, nc, bsz, 2, 0, MAX_SWEEPS, tol, meta, rot);
            TN_LAUNCHED(ctx);
        }
        sweeps = 1;
    } else if (const int ncl2 = fit2(nc, ext)) {
        // second-generation cluster kernel: one barrier per round, smallest cluster that holds the problem
        const int halfp = (nc + 1) / 2;
        size_t ll = ((size_t)ceil_div(a, ncl2) + (size_t)ceil_div(ext, ncl2)) | 1;
        size_t smem = (size_t)nc * ll * sizeof(double) + (ncl2 > 1 ? (size_t)2 * ncl2 * halfp * 3 * sizeof(double) : 0);
        TN_FUNC_ATTR_ONCE(ctx, jacobi_cluster2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMEM);
        if (ncl2 > CLJ) TN_CUDA(cudaFuncSetAttribute(jacobi_cluster2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncl2, 1, 1);
        cfg.blockDim = dim3(JT2, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ncl2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TN_CUDA(cudaLaunchKernelEx(&cfg, jacobi_cluster2_kernel, E, ldw, a, ext, nc, (int)MAX_SWEEPS, tol, (const SvdMeta*)meta, status));
        TN_LAUNCHED(ctx);
        sweeps = -1;
        if (!small) {
            int* hs = (int*)((char*)ctx->pinned + 320);
            TN_CUDA(cudaMemcpyAsync(hs, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            TN_CUDA(cudaStreamSynchronize(st));
            sweeps = hs[0];
            if (!hs[1]) {
                tn_set_error("Jacobi SVD of a %d x %d matrix (%d live vectors) did not converge in %d sweeps", m, n, nc, sweeps);
                return TN_ERR_NOCONV;
            }
        }
    } else if (cluster_fits(nc, ext) || (wide_clusters() && wide_launchable() && cluster_fits(nc, 0, CLJ_MAX))) {
        // cluster-resident: one launch, no host read-back inside the iteration.  8 CTAs when the live vectors fit;
        // 16 CTAs (non-portable cluster size) for larger problems; and when even that is too small for vectors plus
        // accumulated right vectors, 16 CTAs with only the w-parts resident and the right vectors rotated in L2
        const bool fits8 = cluster_fits(nc, ext), fits16 = wide_clusters() && wide_launchable() && cluster_fits(nc, ext, CLJ_MAX);
        // Throughput mode (tn_set_throughput_mode, used when many solver instances share the GPU): a 4-CTA cluster when
        // the vectors fit its shared memory and every CTA owns at most CJ_MAXOWN pairs -- ~1.3x the time on half the SMs,
        // and at most two 8-CTA clusters fit one GPC anyway.
        const bool fits4 = tn_throughput_mode() && cluster_fits(nc, ext, 4) && (nc + 1) / 2 <= 4 * CJ_MAXOWN;
        const int ncl = fits4 ? 4 : ((fits8 && !(wide_clusters() == 2 && fits16 && nc >= 96)) ? CLJ : CLJ_MAX);
        const int vglob = (ncl == CLJ_MAX && !fits16) ? 1 : 0;
        size_t ll = ((size_t)ceil_div(a, ncl) + (vglob ? 0 : (size_t)ceil_div(ext, ncl))) | 1;
        size_t smem = (size_t)nc * ll * sizeof(double);
        TN_FUNC_ATTR_ONCE(ctx, jacobi_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMEM);
        TN_FUNC_ATTR_ONCE(ctx, jacobi_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncl, 1, 1);
        cfg.blockDim = dim3(JT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ncl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TN_CUDA(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel, E, ldw, a, ext, nc, (int)MAX_SWEEPS, tol, (const SvdMeta*)meta, status,
                                   vglob));
        TN_LAUNCHED(ctx);
        sweeps = -1;
        if (!small) {
            // a host read-back already happened for this (large) matrix: also verify convergence.  The fall-back of the
            // reference is a second LAPACK driver (gesdd -> gesvd, mps.py:31-34); here an unconverged iteration is simply
            // continued -- the kernel wrote its partially orthogonalised vectors back -- for up to three more launches
            int* hs = (int*)((char*)ctx->pinned + 320);
            int total_sweeps = 0;
            for (int attempt = 0; attempt < 4; ++attempt) {
                TN_CUDA(cudaMemcpyAsync(hs, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
                TN_CUDA(cudaStreamSynchronize(st));
                total_sweeps += hs[0];
                if (hs[1] || attempt == 3) break;
                TN_CUDA(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel, E, ldw, a, ext, nc, (int)MAX_SWEEPS, tol, (const SvdMeta*)meta,
                                           status, vglob));
                TN_LAUNCHED(ctx);
            }
            sweeps = total_sweeps;
            if (!hs[1]) {
                tn_set_error("Jacobi SVD of a %d x %d matrix (%d live vectors) did not converge in %d sweeps", m, n, nc, sweeps);
                return TN_ERR_NOCONV;
            }
        }
    } else {
        // multi-block fall-back: block pairs per launch, one host read-back per sweep
        TN_CUDA(cudaFuncSetAttribute(jacobi_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        int bsz = cmax / 2;
        while (bsz > 8 && ceil_div(nc, bsz) < 16) bsz = (bsz + 1) / 2;
        int nblocks = ceil_div(nc, bsz);
        int npe = nblocks + (nblocks & 1);
        size_t smem = (size_t)2 * bsz * ldw * sizeof(double);
        bool converged = false;
        for (sweeps = 0; sweeps < MAX_SWEEPS && !converged;) {
            TN_CUDA(cudaMemsetAsync(rot, 0, sizeof(unsigned int), st));
            for (int r = 0; r < npe - 1; ++r) {
                jacobi_round_kernel<<<npe / 2, JT, smem, st>>>(E, ldw, a, nc, bsz, nblocks, r, 2, tol, meta, rot);
                TN_LAUNCHED(ctx);
            }
            ++sweeps;
            TN_CUDA(cudaMemcpyAsync(h_rot, rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            TN_CUDA(cudaStreamSynchronize(st));
            converged = (*h_rot == 0);
        }
        if (!converged) {
            tn_set_error("Jacobi SVD of a %d x %d matrix (%d live vectors) did not converge in %d sweeps", m, n, nc, sweeps);
            return TN_ERR_NOCONV;
        }
    }
    {
        int fin_blocks = ext ? ceil_div(nc, JW) : 1;
        if (fin_blocks > 16) fin_blocks = 16;
        jacobi_finish_kernel<<<fin_blocks, JT, 0, st>>>(E, ldw, a, ext, nc, kfull, has_dead, meta, live_idx, U, ldu, S, Vt, ldvt, sv_tmp, rank_tmp);
    }
    TN_LAUNCHED(ctx);
    if (h_sweeps) *h_sweeps = sweeps;
    return TN_OK;
}

extern "C" int tn_truncation_rank(tn_ctx* ctx, void* stream, const double* S, int k, double tol, int Dmax, int* h_keep,
                                  double* h_discarded) {
    TN_REQUIRE(ctx != nullptr && k >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    char* ws = (char*)tn_scratch(ctx, TN_SLOT_SORT, 64);
    if (!ws) return TN_ERR_NOMEM;
    int* d_keep = (int*)ws;
    double* d_disc = (double*)(ws + 8);
    truncation_rank_kernel<<<1, 32, 0, st>>>(S, k, tol, Dmax, d_keep, d_disc);
    TN_LAUNCHED(ctx);
    char* hp = (char*)ctx->pinned + 64;
    TN_CUDA(cudaMemcpyAsync(hp, ws, 16, cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    *h_keep = *(int*)hp;
    *h_discarded = *(double*)(hp + 8);
    return TN_OK;
}
