// Branch-and-bound kernels: right/left environments, batched conditional marginals fused with the
// negative-probability rule and log2 accumulation, relative cut-off selection, boundary merge, top-M,
// branch materialisation, Gibbs inverse-CDF sampling.  Restates tnac4o.py:381-650, 1506-1531, 1768-1807.
#include "common.cuh"

int tn_sort3_impl(tn_ctx* ctx, cudaStream_t st, unsigned long long* hi, unsigned long long* lo, unsigned long long* tie, int n);
int tn_scan_impl(tn_ctx* ctx, cudaStream_t st, const int* in, int* out, int n, int* tmp);
int tn_gemm_grouped_impl(tn_ctx* ctx, cudaStream_t st, int tile_rows, int ntiles, int N, int K, const double* X, int ldx,
                         const double* B, int ldb, int64_t strideB, const int* bmap, double* C, int ldc);

namespace {

__device__ __forceinline__ double block_reduce(double v, int op, double* sh) {   // op 0 sum, 1 max, 2 min; result to all
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = (op == 0) ? warp_sum(v) : (op == 1 ? warp_max(v) : warp_min(v));
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = sh[0];
    for (int w = 1; w < nw; ++w) r = (op == 0) ? r + sh[w] : (op == 1 ? fmax(r, sh[w]) : fmin(r, sh[w]));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Right environments of one level (tnac4o.py:1776-1782), one CTA per row-start branch:
//   Y[p][b'][l] = sum_r RRin[b'][r] WtrU[u][l][p][r];   RRout[a][l] = sum_{p,b'} A[a][p][b'] Y[p][b'][l]; /nfactor
__global__ void __launch_bounds__(256)
rr_level_kernel(int nb, int Dl, int Dr, int nl, int nd, int nr, int nu, const double* __restrict__ A,
                const double* __restrict__ WtrU, const double* __restrict__ RRin, const uint8_t* __restrict__ up,
                int up_stride, double* __restrict__ RRout) {
    extern __shared__ __align__(16) double sm[];
    double* Y = sm;                                  // [nd*Dr][nl + 1]
    double* Rin = Y + (size_t)nd * Dr * (nl + 1);    // [Dr][nr]
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int K = nd * Dr;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int u = up[(int64_t)b * up_stride];
        const double* W = WtrU + (int64_t)u * nl * nd * nr;
        __syncthreads();
        for (int i = tid; i < Dr * nr; i += blockDim.x) Rin[i] = RRin[(int64_t)b * Dr * nr + i];
        __syncthreads();
        for (int i = tid; i < K * nl; i += blockDim.x) {
            int l = i % nl, pb = i / nl;
            int p = pb / Dr, bp = pb % Dr;
            const double* w = W + ((int64_t)l * nd + p) * nr;
            const double* rr = Rin + bp * nr;
            double s = 0.0;
            for (int r = 0; r < nr; ++r) s += rr[r] * w[r];
            Y[pb * (nl + 1) + l] = s;
        }
        __syncthreads();
        double mx = 0.0;
        double vals[4];
        int cnt = 0;
        for (int o = tid; o < Dl * nl; o += blockDim.x) {
            int a = o / nl, l = o % nl;
            const double* arow = A + (int64_t)a * K;
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += arow[k] * Y[k * (nl + 1) + l];
            if (cnt < 4) vals[cnt] = s;
            else RRout[(int64_t)b * Dl * nl + o] = s;      // (Dl*nl > 1024: spill path, rescaled below)
            cnt++;
            mx = fmax(mx, fabs(s));
        }
        mx = block_reduce(mx, 1, red);
        const double inv = ldexp(1.0, 1023 - (int)((((unsigned long long)__double_as_longlong(mx)) >> 52) & 0x7ff));
        cnt = 0;
        for (int o = tid; o < Dl * nl; o += blockDim.x) {
            double s = (cnt < 4) ? vals[cnt] : RRout[(int64_t)b * Dl * nl + o];
            RRout[(int64_t)b * Dl * nl + o] = s * inv;
            cnt++;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same level as ONE contraction on the DMMA path.  With the MPS tensor and the traced PEPS tensor contracted
// first (once per level, not per branch),
//   AW_u[(b', r)][(a, l)] = sum_p A[a][p][b'] Wtr[l][p][r][u],
// the level is  RRout_b[(a, l)] = sum_{(b', r)} RRin_b[(b', r)] AW_{u_b}[(b', r)][(a, l)]: a row-vector times one of nu
// matrices.  Branches are grouped by their up index u_b (rows padded to whole GEMM tiles), multiplied by their
// group's matrix with the grouped GEMM, and scattered back with the per-branch nfactor scaling.
// 2 Dr nr Dl nl flop per branch-level (524 288 at D = 32) instead of 786 432, at tensor-core rates.
__global__ void rr_aw_kernel(int Dl, int Dr, int nl, int nd, int nr, int nu, const double* __restrict__ A,
                             const double* __restrict__ WtrU, double* __restrict__ AW) {
    const int K = Dr * nr, N = Dl * nl;
    const int64_t total = (int64_t)nu * K * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N), k = (int)((i / N) % K), u = (int)(i / ((int64_t)N * K));
        const int a = n / nl, l = n % nl, bp = k / nr, r = k % nr;
        const double* ap = A + (int64_t)a * nd * Dr + bp;
        const double* wp = WtrU + ((int64_t)u * nl + l) * nd * nr + r;
        double s = 0.0;
        for (int p = 0; p < nd; ++p) s += ap[(int64_t)p * Dr] * wp[(int64_t)p * nr];
        AW[i] = s;
    }
}

// one CTA: group sizes by up index, tile-padded group offsets, tile -> group map, destination row of every branch
// (the order inside a group is arbitrary: a row's product does not depend on where the row sits)
__global__ void __launch_bounds__(1024)
rr_group_kernel(int nb, int nu, int tile, int ntiles, const uint8_t* __restrict__ up, int up_stride, int* __restrict__ pos,
                int* __restrict__ bmap) {
    __shared__ int cnt[256], off[256], cur[256];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += blockDim.x) { cnt[i] = 0; cur[i] = 0; off[i] = 0; }
    __syncthreads();
    for (int b = tid; b < nb; b += blockDim.x) atomicAdd(&cnt[up[(int64_t)b * up_stride]], 1);
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int u = 0; u < nu; ++u) { off[u] = o; o += (cnt[u] + tile - 1) / tile * tile; }
    }
    __syncthreads();
    for (int t = tid; t < ntiles; t += blockDim.x) {
        const int row = t * tile;
        int g = -1;
        for (int u = 0; u < nu; ++u)
            if (cnt[u] > 0 && row >= off[u] && row < off[u] + (cnt[u] + tile - 1) / tile * tile) g = u;
        bmap[t] = g;
    }
    for (int b = tid; b < nb; b += blockDim.x) {
        const int u = up[(int64_t)b * up_stride];
        pos[b] = off[u] + atomicAdd(&cur[u], 1);
    }
}

__global__ void rr_gather_kernel(int nb, int K, const double* __restrict__ RRin, const int* __restrict__ pos,
                                 double* __restrict__ X) {
    const int64_t total = (int64_t)nb * K;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / K), k = (int)(i % K);
        X[(int64_t)pos[b] * K + k] = RRin[i];
    }
}

// RRout[b] = C[pos[b]] / nfactor(C[pos[b]]) -- one warp per branch, same exponent rule as rr_level_kernel
__global__ void __launch_bounds__(256)
rr_scatter_kernel(int nb, int N, const double* __restrict__ C, const int* __restrict__ pos, double* __restrict__ RRout) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < nb; b += gridDim.x * wpb) {
        const double* row = C + (int64_t)pos[b] * N;
        double mx = 0.0;
        for (int i = lane; i < N; i += 32) mx = fmax(mx, fabs(row[i]));
        mx = warp_max(mx);
        const double inv = ldexp(1.0, 1023 - (int)((((unsigned long long)__double_as_longlong(mx)) >> 52) & 0x7ff));
        for (int i = lane; i < N; i += 32) RRout[(int64_t)b * N + i] = row[i] * inv;
    }
}

// ------------------------------------------------------------------------------------------------
// Marginals of one cell for all branches, one CTA (256 threads) per branch (tnac4o.py:1786-1807, 450-453).
__global__ void __launch_bounds__(256)
marginals_kernel(int nb, int nS, int nl, int nd, int nr, int nu, int Dr, const double* __restrict__ Wlu,
                 const uint8_t* __restrict__ dmap, const uint8_t* __restrict__ rmap, const double* __restrict__ T1,
                 const double* __restrict__ RR, const int32_t* __restrict__ root, const uint8_t* __restrict__ vind,
                 int vstride, int nx, const double* __restrict__ prob, double* __restrict__ cand, double* __restrict__ flag,
                 unsigned long long* max_bits, double* __restrict__ P_out) {
    extern __shared__ __align__(16) double sm[];
    double* t1 = sm;                     // [nd][Dr]
    double* rr = t1 + nd * Dr;           // [Dr][nr]
    double* t2 = rr + Dr * nr;           // [nd][nr]
    __shared__ double red[8];
    const int tid = threadIdx.x;
    double cta_max = -INFINITY;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        __syncthreads();
        const double* rrg = RR + (int64_t)root[b] * Dr * nr;
        for (int i = tid; i < nd * Dr; i += blockDim.x) t1[i] = T1[(int64_t)b * nd * Dr + i];
        for (int i = tid; i < Dr * nr; i += blockDim.x) rr[i] = rrg[i];
        __syncthreads();
        for (int o = tid; o < nd * nr; o += blockDim.x) {
            int d = o / nr, r = o % nr;
            double s = 0.0;
            for (int k = 0; k < Dr; ++k) s += t1[d * Dr + k] * rr[k * nr + r];
            t2[o] = s;
        }
        __syncthreads();
        const int l = vind[(int64_t)b * vstride + nx], u = vind[(int64_t)b * vstride + nx + 1];
        const double* w = Wlu + ((int64_t)l * nu + u) * nS;
        // blockDim.x >= nS is not assumed: each thread owns states tid, tid + 256, ...
        double pmin = INFINITY;
        for (int s = tid; s < nS; s += blockDim.x) {
            double p = w[s] * t2[dmap[s] * nr + rmap[s]];
            pmin = fmin(pmin, p);
        }
        pmin = block_reduce(pmin, 2, red);
        double fl = pmin;
        double cntneg = 0.0, tot = 0.0;
        for (int s = tid; s < nS; s += blockDim.x) {
            double p = w[s] * t2[dmap[s] * nr + rmap[s]];
            if (pmin < 0.0 && p < fabs(pmin)) { p = fabs(pmin); cntneg += 1.0; }
            tot += p;
        }
        tot = block_reduce(tot, 0, red);
        if (pmin < 0.0) { cntneg = block_reduce(cntneg, 0, red); fl = pmin * cntneg; }
        double inv = 0.0;
        if (tot > 0.0) { inv = 1.0 / tot; fl *= inv; } else fl = -1.0;
        const double pb = prob ? prob[b] : 0.0;
        for (int s = tid; s < nS; s += blockDim.x) {
            double p = w[s] * t2[dmap[s] * nr + rmap[s]];
            if (pmin < 0.0 && p < fabs(pmin)) p = fabs(pmin);
            p = (tot > 0.0) ? p * inv : p + 1.0 / (double)nS;
            if (P_out) P_out[(int64_t)b * nS + s] = p;
            if (cand) {
                double c = log2(p) + pb;
                cand[(int64_t)b * nS + s] = c;
                cta_max = fmax(cta_max, c);
            }
        }
        if (tid == 0) flag[b] = fl;
    }
    if (cand && max_bits) {
        cta_max = warp_max(cta_max);
        if ((tid & 31) == 0 && cta_max > -INFINITY) atomicMax(max_bits, ordered_bits(cta_max));
    }
}

// ------------------------------------------------------------------------------------------------
// The same marginals with ONE WARP per branch and the (d x Dr) . (Dr x r) product on the DMMA path, for the interior
// chimera shape nd = nr = 16 (every site that is not on the right / bottom edge).  The one-CTA-per-branch kernel above
// spends its time in four block-wide reductions and recomputes w[s] * t2[d(s), r(s)] three times (13 % of the HBM peak at
// 10^5 branches, profiles/r1c_*); here a warp loads its branch's T1 and RR fragments straight from global memory (each
// element exactly once: 8 KB per branch), accumulates the 16 x 16 result in four m8n8k4 tiles, parks it in 2 KB of
// shared memory for the indirect look-up t2[d(s), r(s)], keeps its 8 of the 256 state weights in registers, and reduces
// min / sum / max with warp shuffles.  No block barrier; a branch's result does not depend on its neighbours in the
// launch (the sharded search relies on that).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

constexpr int MW_WARPS = 8;

__global__ void __launch_bounds__(MW_WARPS * 32)
marginals_warp_kernel(int nb, int nS, int nu, int Dr, const double* __restrict__ Wlu, const uint8_t* __restrict__ dmap,
                      const uint8_t* __restrict__ rmap, const double* __restrict__ T1, const double* __restrict__ RR,
                      const int32_t* __restrict__ root, const uint8_t* __restrict__ vind, int vstride, int nx,
                      const double* __restrict__ prob, double* __restrict__ cand, double* __restrict__ flag,
                      unsigned long long* max_bits, double* __restrict__ P_out) {
    __shared__ double t2s[MW_WARPS][16 * 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* t2 = t2s[warp];
    double wmax = -INFINITY;
    for (int b = blockIdx.x * MW_WARPS + warp; b < nb; b += gridDim.x * MW_WARPS) {
        const double* t1 = T1 + (int64_t)b * 16 * Dr;
        const double* rr = RR + (int64_t)root[b] * Dr * 16;
        double acc[2][2][2] = {};
        for (int kk = 0; kk < Dr; kk += 4) {
            const double a0 = t1[(int64_t)g * Dr + kk + t], a1 = t1[(int64_t)(8 + g) * Dr + kk + t];
            const double b0 = rr[(int64_t)(kk + t) * 16 + g], b1 = rr[(int64_t)(kk + t) * 16 + 8 + g];
            dmma_m8n8k4(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma_m8n8k4(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma_m8n8k4(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma_m8n8k4(acc[1][1][0], acc[1][1][1], a1, b1);
        }
        __syncwarp();                                   // the previous branch's look-ups are done
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                t2[(i * 8 + g) * 16 + j * 8 + 2 * t] = acc[i][j][0];
                t2[(i * 8 + g) * 16 + j * 8 + 2 * t + 1] = acc[i][j][1];
            }
        __syncwarp();
        const int l = vind[(int64_t)b * vstride + nx], u = vind[(int64_t)b * vstride + nx + 1];
        const double* w = Wlu + ((int64_t)l * nu + u) * nS;
        double p[8];
        double pmin = INFINITY;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int s = lane + 32 * q;
            p[q] = (s < nS) ? w[s] * t2[dmap[s] * 16 + rmap[s]] : INFINITY;
            pmin = fmin(pmin, p[q]);
        }
        pmin = warp_min(pmin);
        double fl = pmin, cntneg = 0.0, tot = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (lane + 32 * q < nS) {
                if (pmin < 0.0 && p[q] < fabs(pmin)) { p[q] = fabs(pmin); cntneg += 1.0; }
                tot += p[q];
            }
        }
        tot = warp_sum(tot);
        if (pmin < 0.0) { cntneg = warp_sum(cntneg); fl = pmin * cntneg; }
        double inv = 0.0;
        if (tot > 0.0) { inv = 1.0 / tot; fl *= inv; } else fl = -1.0;
        const double pb = prob ? prob[b] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int s = lane + 32 * q;
            if (s < nS) {
                const double pn = (tot > 0.0) ? p[q] * inv : p[q] + 1.0 / (double)nS;
                if (P_out) P_out[(int64_t)b * nS + s] = pn;
                if (cand) {
                    const double c = log2(pn) + pb;
                    cand[(int64_t)b * nS + s] = c;
                    wmax = fmax(wmax, c);
                }
            }
        }
        if (lane == 0) flag[b] = fl;
    }
    if (cand && max_bits) {
        wmax = warp_max(wmax);
        if (lane == 0 && wmax > -INFINITY) atomicMax(max_bits, ordered_bits(wmax));
    }
}

// ------------------------------------------------------------------------------------------------
// relative cut-off (tnac4o.py:456-465): keep cand > max + log2cut; record the largest discarded value
__global__ void select_kernel(const double* __restrict__ cand, int64_t n, const unsigned long long* max_bits, double log2cut,
                              int use_cut, int32_t* __restrict__ surv, int* count, unsigned long long* pd_bits) {
    const double mx = from_ordered_bits(*max_bits);
    const double thr = mx + log2cut;
    double lost = -INFINITY;
    bool any_lost = false;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = i0 + threadIdx.x;
        bool keep = false;
        double c = 0.0;
        if (i < n) {
            c = cand[i];
            keep = !use_cut || (c > thr) || (c == mx);
            if (!keep && !(c != c)) { lost = fmax(lost, c); any_lost = true; }
        }
        unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (mask) {
            int lane = threadIdx.x & 31;
            int basepos = 0;
            if (lane == 0) basepos = atomicAdd(count, __popc(mask));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (keep) surv[basepos + __popc(mask & ((1u << lane) - 1))] = (int32_t)i;
        }
    }
    if (any_lost) atomicMax(pd_bits, ordered_bits(lost));
}

struct KeyLayout {
    int npos;
    uint8_t off[64];     // bit offset of every position inside the 128-bit key
};

// new boundary row + energy of every surviving candidate (tnac4o.py:469-478, 1506-1531)
__global__ void expand_kernel(int K, int nS, int nx, int has_left, int has_up, int nl, int nu, KeyLayout lay,
                              const int32_t* __restrict__ surv, const uint8_t* __restrict__ vind, int vstride,
                              const uint8_t* __restrict__ dmap, const uint8_t* __restrict__ rmap,
                              const double* __restrict__ Es, const double* __restrict__ Esl, const double* __restrict__ Esu,
                              const double* __restrict__ Eng, const double* __restrict__ cand,
                              unsigned long long* __restrict__ khi, unsigned long long* __restrict__ klo,
                              unsigned long long* __restrict__ ktie, int32_t* __restrict__ parent, int32_t* __restrict__ cell,
                              double* __restrict__ Enew, double* __restrict__ Pnew) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const int id = surv[i];
    const int b = id / nS, s = id % nS;
    const uint8_t* row = vind + (int64_t)b * vstride;
    unsigned long long hi = 0, lo = 0;
    for (int j = 0; j < lay.npos; ++j) {
        unsigned long long v = (j == nx) ? dmap[s] : (j == nx + 1 ? rmap[s] : row[j]);
        int o = lay.off[j];
        if (o < 64) {
            lo |= v << o;
            if (o > 56) hi |= v >> (64 - o);
        } else hi |= v << (o - 64);
    }
    khi[i] = hi; klo[i] = lo;
    ktie[i] = ((unsigned long long)(unsigned)id << 32) | (unsigned)i;
    double dE = Es[s];
    if (has_left) dE += Esl[(int64_t)s * nl + row[nx]];
    if (has_up) dE += Esu[(int64_t)s * nu + row[nx + 1]];
    Enew[i] = Eng[b] + dE;
    Pnew[i] = cand[id];
    parent[i] = b;
    cell[i] = s;
}

__global__ void heads_kernel(int K, const unsigned long long* __restrict__ khi, const unsigned long long* __restrict__ klo,
                             int* __restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    head[i] = (i == 0) || (khi[i] != khi[i - 1]) || (klo[i] != klo[i - 1]);
}

// one thread per group: representative = first minimum of the energy in sorted order, degeneracy and mean
// log-probability over members within min_dEng of the minimum (tnac4o.py:493-509)
__global__ void group_reduce_kernel(int K, const int* __restrict__ head, const int* __restrict__ incl,
                                    const unsigned long long* __restrict__ ktie, const double* __restrict__ Enew,
                                    const double* __restrict__ Pnew, const int32_t* __restrict__ parent,
                                    const long long* __restrict__ deg, double min_dEng, int32_t* __restrict__ g_rep,
                                    long long* __restrict__ g_deg, double* __restrict__ g_prob, double* __restrict__ g_E,
                                    int32_t* __restrict__ g_start, int32_t* __restrict__ g_size) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K || !head[i]) return;
    const int g = incl[i] - 1;
    int j = i;
    int best = (int)(ktie[i] & 0xffffffffu);
    double Emin = Enew[best];
    for (j = i + 1; j < K && !head[j]; ++j) {
        int m = (int)(ktie[j] & 0xffffffffu);
        double e = Enew[m];
        if (e < Emin) { Emin = e; best = m; }
    }
    const int end = j;
    long long dsum = 0;
    double psum = 0.0;
    int ntied = 0;
    for (j = i; j < end; ++j) {
        int m = (int)(ktie[j] & 0xffffffffu);
        if (Enew[m] - Emin <= min_dEng) { dsum += deg[parent[m]]; psum += Pnew[m]; ntied++; }
    }
    g_rep[g] = best;
    g_deg[g] = dsum;
    g_prob[g] = (ntied > 1) ? psum / (double)ntied : Pnew[best];
    g_E[g] = Emin;
    g_start[g] = i;
    g_size[g] = end - i;
}

__global__ void topm_keys_kernel(int G, const double* __restrict__ g_prob, unsigned long long* khi, unsigned long long* klo,
                                 unsigned long long* ktie) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    khi[g] = ~ordered_bits(g_prob[g]);      // descending probability
    klo[g] = 0;
    ktie[g] = (unsigned long long)g;
}

__global__ void topm_take_kernel(int G, int M, const unsigned long long* __restrict__ ktie, const double* __restrict__ g_prob,
                                 int32_t* __restrict__ sel, unsigned long long* pd_bits) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) sel[j] = (int32_t)ktie[j];
    if (j == M && M < G) atomicMax(pd_bits, ordered_bits(g_prob[(int)ktie[M]]));
}

__global__ void iota_kernel(int n, int32_t* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// new branch arrays + left environment RL' = RL[parent] . A[:, d, :] / nfactor  (tnac4o.py:470-477, 528-535)
// one warp per new branch
__global__ void materialise_kernel(int B, int nx, int pos, int nsites, int vstride, int Dl, int nd, int Dr,
                                   const int32_t* __restrict__ sel, const int32_t* __restrict__ g_rep,
                                   const long long* __restrict__ g_deg, const double* __restrict__ g_prob,
                                   const int32_t* __restrict__ parent, const int32_t* __restrict__ cell,
                                   const double* __restrict__ Enew, const uint8_t* __restrict__ dmap,
                                   const uint8_t* __restrict__ rmap, const uint8_t* __restrict__ vind_in,
                                   const uint8_t* __restrict__ states_in, const int32_t* __restrict__ root_in,
                                   const double* __restrict__ RL_in, const double* __restrict__ A,
                                   uint8_t* __restrict__ vind_out, uint8_t* __restrict__ states_out,
                                   int32_t* __restrict__ root_out, double* __restrict__ Eng_out, double* __restrict__ prob_out,
                                   long long* __restrict__ deg_out, double* __restrict__ RL_out) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= B) return;
    const int g = sel ? sel[j] : j;
    const int i = g_rep ? g_rep[g] : g;
    const int b = parent[i], s = cell[i];
    const int dd = dmap[s], rr = rmap[s];
    for (int k = lane; k < vstride; k += 32) {
        uint8_t v = vind_in[(int64_t)b * vstride + k];
        if (k == nx) v = (uint8_t)dd;
        if (k == nx + 1) v = (uint8_t)rr;
        vind_out[(int64_t)j * vstride + k] = v;
    }
    for (int k = lane; k < nsites; k += 32) {
        uint8_t v = states_in[(int64_t)b * nsites + k];
        if (k == pos) v = (uint8_t)s;
        states_out[(int64_t)j * nsites + k] = v;
    }
    if (lane == 0) {
        root_out[j] = root_in[b];
        Eng_out[j] = Enew[i];
        if (prob_out) prob_out[j] = g_prob[g];
        if (deg_out) deg_out[j] = g_deg[g];
    }
    // left environment
    const double* rl = RL_in + (int64_t)b * Dl;
    double mx = 0.0;
    double vals[8];
    int cnt = 0;
    for (int o = lane; o < Dr; o += 32) {
        double acc = 0.0;
        for (int a = 0; a < Dl; ++a) acc += rl[a] * A[((int64_t)a * nd + dd) * Dr + o];
        if (cnt < 8) vals[cnt] = acc;
        cnt++;
        mx = fmax(mx, fabs(acc));
    }
    mx = warp_max(mx);
    const double inv = ldexp(1.0, 1023 - (int)((((unsigned long long)__double_as_longlong(mx)) >> 52) & 0x7ff));
    cnt = 0;
    for (int o = lane; o < Dr; o += 32) {
        RL_out[(int64_t)j * Dr + o] = vals[cnt] * inv;
        cnt++;
    }
}

__global__ void row_shift_kernel(int B, int vstride, uint8_t* vind) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint8_t* row = vind + (int64_t)b * vstride;
    for (int k = vstride - 1; k >= 1; --k) row[k] = row[k - 1];
    row[0] = 0;
}

// Gibbs step: inverse-CDF draw per sample with a sequential cumulative sum (np.cumsum + searchsorted,
// tnac4o.py:616-621), one thread per sample; writes (parent = b, cell) so that materialise_kernel finishes the step.
__global__ void sample_kernel(int B, int nS, const double* __restrict__ P, const double* __restrict__ uni,
                              int has_left, int has_up, int nl, int nu, int nx, const uint8_t* __restrict__ vind,
                              int vstride, const double* __restrict__ Es, const double* __restrict__ Esl,
                              const double* __restrict__ Esu, const double* __restrict__ Eng, int32_t* __restrict__ parent,
                              int32_t* __restrict__ cell, double* __restrict__ Enew) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double u = uni[b];
    const double* p = P + (int64_t)b * nS;
    double c = 0.0;
    int s = nS;
    for (int k = 0; k < nS; ++k) {
        c += p[k];
        if (c >= u) { s = k; break; }
    }
    if (s >= nS) s = nS - 1;
    const uint8_t* row = vind + (int64_t)b * vstride;
    double dE = Es[s];
    if (has_left) dE += Esl[(int64_t)s * nl + row[nx]];
    if (has_up) dE += Esu[(int64_t)s * nu + row[nx + 1]];
    Enew[b] = Eng[b] + dE;
    parent[b] = b;
    cell[b] = s;
}

}  // namespace

static inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads < 1 ? 1 : (n + threads - 1) / threads); }

extern "C" {

int tn_rr_level(tn_ctx* ctx, void* stream, const tn_site* site, int nb, int Dl, int Dr, const double* A,
                const double* RRin, const uint8_t* up, int up_stride, double* RRout) {
    TN_REQUIRE(ctx && site && nb >= 0, "bad arguments");
    if (nb == 0) return TN_OK;
    {
        const int K = Dr * site->nr, N = Dl * site->nl, nu = site->nu;
        const size_t aw_bytes = (size_t)nu * K * N * sizeof(double);
        if (nu <= 256 && aw_bytes <= ((size_t)512 << 20)) {
            cudaStream_t st = as_stream(stream);
            const int tile = nb >= 4096 ? 128 : 64;
            const int ntiles = ceil_div(nb, tile) + nu;
            const size_t rows = (size_t)ntiles * tile;
            // one stream-ordered block: AW | X | C | pos | bmap
            const size_t nd_ = ((size_t)nu * K * N + rows * K + rows * N) * sizeof(double);
            const size_t ni_ = ((size_t)nb + ntiles) * sizeof(int);
            void* blk = nullptr;
            TN_CUDA(tn_malloc_async(ctx, &blk, nd_ + ni_, st));
            double* AW = (double*)blk;
            double* X = AW + (size_t)nu * K * N;
            double* C = X + rows * K;
            int* pos = (int*)(C + rows * N);
            int* bmap = pos + nb;
            const int cap = 8 * ctx->sm_count;
            int64_t want = ((int64_t)nu * K * N + 255) / 256;
            rr_aw_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(Dl, Dr, site->nl, site->nd, site->nr, nu, A, site->Wtr, AW);
            TN_LAUNCHED(ctx);
            rr_group_kernel<<<1, 1024, 0, st>>>(nb, nu, tile, ntiles, up, up_stride, pos, bmap);
            TN_LAUNCHED(ctx);
            want = ((int64_t)nb * K + 255) / 256;
            rr_gather_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(nb, K, RRin, pos, X);
            TN_LAUNCHED(ctx);
            int rc = tn_gemm_grouped_impl(ctx, st, tile, ntiles, N, K, X, K, AW, N, (int64_t)K * N, bmap, C, N);
            if (rc) { cudaFreeAsync(blk, st); return rc; }
            want = (nb + 7) / 8;
            rr_scatter_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(nb, N, C, pos, RRout);
            TN_LAUNCHED(ctx);
            TN_CUDA(cudaFreeAsync(blk, st));
            return TN_OK;
        }
    }
    // fall-back for very wide up legs: one CTA per branch on the CUDA cores
    size_t smem = ((size_t)site->nd * Dr * (site->nl + 1) + (size_t)Dr * site->nr) * sizeof(double);
    TN_REQUIRE(smem <= 220 * 1024, "bond dimension too large for rr_level shared memory");
    // a fixed maximum: concurrent host threads must not lower the attribute under each other's launches
    TN_CUDA(cudaFuncSetAttribute(rr_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int grid = nb < 4 * ctx->sm_count ? nb : 4 * ctx->sm_count;
    rr_level_kernel<<<grid, 256, smem, as_stream(stream)>>>(nb, Dl, Dr, site->nl, site->nd, site->nr, site->nu, A,
                                                            site->Wtr, RRin, up, up_stride, RRout);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_marginals(tn_ctx* ctx, void* stream, const tn_site* site, int nb, int Dr, const double* T1, const double* RR,
                 const int32_t* root, const uint8_t* vind, int vstride, int nx, const double* prob, double* cand,
                 double* flag, unsigned long long* max_bits, double* P_out) {
    TN_REQUIRE(ctx && site && nb >= 0, "bad arguments");
    if (nb == 0) return TN_OK;
    cudaStream_t st = as_stream(stream);
    if (max_bits) TN_CUDA(cudaMemsetAsync(max_bits, 0, sizeof(unsigned long long), st));
    static const bool warp_path = [] { const char* e = getenv("TN_MARGINALS"); return !(e && e[0] == 'c'); }();
    if (warp_path && site->nd == 16 && site->nr == 16 && site->nS <= 256 && Dr % 4 == 0) {
        // interior chimera shape: one warp per branch, (d x Dr) . (Dr x r) on DMMA (TN_MARGINALS=cta selects the old kernel)
        const int want = ceil_div(nb, MW_WARPS);
        const int grid = want < 16 * ctx->sm_count ? want : 16 * ctx->sm_count;
        marginals_warp_kernel<<<grid, MW_WARPS * 32, 0, st>>>(nb, site->nS, site->nu, Dr, site->Wlu, site->dmap, site->rmap, T1, RR,
                                                             root, vind, vstride, nx, prob, cand, flag, max_bits, P_out);
        TN_LAUNCHED(ctx);
        return TN_OK;
    }
    size_t smem = ((size_t)site->nd * Dr + (size_t)Dr * site->nr + (size_t)site->nd * site->nr) * sizeof(double);
    TN_REQUIRE(smem <= 220 * 1024, "bond dimension too large for marginals shared memory");
    TN_FUNC_ATTR_ONCE(ctx, marginals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    int grid = nb < 8 * ctx->sm_count ? nb : 8 * ctx->sm_count;
    marginals_kernel<<<grid, 256, smem, st>>>(nb, site->nS, site->nl, site->nd, site->nr, site->nu, Dr, site->Wlu,
                                              site->dmap, site->rmap, T1, RR, root, vind, vstride, nx, prob, cand, flag,
                                              max_bits, P_out);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_select(tn_ctx* ctx, void* stream, const double* cand, int64_t n, const unsigned long long* max_bits,
              double relative_P_cutoff, int32_t* surv, int* count, unsigned long long* pd_bits, int* h_count) {
    TN_REQUIRE(ctx && n >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    TN_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
    int use_cut = relative_P_cutoff > 0.0;
    double log2cut = use_cut ? log2(relative_P_cutoff) : 0.0;
    int grid = (int)((n + 255) / 256 < 8 * ctx->sm_count ? (n + 255) / 256 : 8 * ctx->sm_count);
    select_kernel<<<grid, 256, 0, st>>>(cand, n, max_bits, log2cut, use_cut, surv, count, pd_bits);
    TN_LAUNCHED(ctx);
    if (h_count) {
        int* hp = (int*)((char*)ctx->pinned + 128);
        TN_CUDA(cudaMemcpyAsync(hp, count, sizeof(int), cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaStreamSynchronize(st));
        *h_count = *hp;
    }
    return TN_OK;
}

int tn_expand(tn_ctx* ctx, void* stream, const tn_site* site, int K, int nx, int has_left, int has_up, int npos,
              const uint8_t* h_bit_offsets, const int32_t* surv, const uint8_t* vind, int vstride, const double* Eng,
              const double* cand, unsigned long long* khi, unsigned long long* klo, unsigned long long* ktie,
              int32_t* parent, int32_t* cell, double* Enew, double* Pnew) {
    TN_REQUIRE(ctx && site && K >= 0 && npos <= 64, "bad arguments");
    if (K == 0) return TN_OK;
    KeyLayout lay;
    lay.npos = npos;
    for (int j = 0; j < npos; ++j) lay.off[j] = h_bit_offsets[j];
    expand_kernel<<<blocks_for(K, 256), 256, 0, as_stream(stream)>>>(K, site->nS, nx, has_left, has_up, site->nl, site->nu,
                                                                     lay, surv, vind, vstride, site->dmap, site->rmap,
                                                                     site->Es, site->Esl, site->Esu, Eng, cand, khi, klo,
                                                                     ktie, parent, cell, Enew, Pnew);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

/* sort + group: after the call g_* describe G groups (host value *h_G); scratch ints: head (K), incl (K), tmp (K/1024+1) */
int tn_merge(tn_ctx* ctx, void* stream, int K, unsigned long long* khi, unsigned long long* klo, unsigned long long* ktie,
             const double* Enew, const double* Pnew, const int32_t* parent, const long long* deg, double min_dEng,
             int32_t* g_rep, long long* g_deg, double* g_prob, double* g_E, int32_t* g_start, int32_t* g_size, int* h_G) {
    TN_REQUIRE(ctx && K >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    int rc = tn_sort3_impl(ctx, st, khi, klo, ktie, K);
    if (rc) return rc;
    size_t bytes = ((size_t)2 * K + K / 1024 + 8) * sizeof(int);
    int* head = (int*)tn_scratch(ctx, TN_SLOT_SEARCH, bytes);
    if (!head) return TN_ERR_NOMEM;
    int* incl = head + K;
    int* tmp = incl + K;
    heads_kernel<<<blocks_for(K, 256), 256, 0, st>>>(K, khi, klo, head);
    TN_LAUNCHED(ctx);
    if ((rc = tn_scan_impl(ctx, st, head, incl, K, tmp))) return rc;
    group_reduce_kernel<<<blocks_for(K, 128), 128, 0, st>>>(K, head, incl, ktie, Enew, Pnew, parent, deg, min_dEng, g_rep,
                                                            g_deg, g_prob, g_E, g_start, g_size);
    TN_LAUNCHED(ctx);
    int* hp = (int*)((char*)ctx->pinned + 192);
    TN_CUDA(cudaMemcpyAsync(hp, incl + (K - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    *h_G = *hp;
    return TN_OK;
}

/* top-M over merged groups (tnac4o.py:518-526): sel[0..B) lists the kept groups, B = min(G, M) */
int tn_topm(tn_ctx* ctx, void* stream, int G, int M, const double* g_prob, unsigned long long* khi, unsigned long long* klo,
            unsigned long long* ktie, int32_t* sel, unsigned long long* pd_bits) {
    TN_REQUIRE(ctx && G >= 1 && M >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    if (G <= M) {
        iota_kernel<<<blocks_for(G, 256), 256, 0, st>>>(G, sel);
        TN_LAUNCHED(ctx);
        return TN_OK;
    }
    topm_keys_kernel<<<blocks_for(G, 256), 256, 0, st>>>(G, g_prob, khi, klo, ktie);
    TN_LAUNCHED(ctx);
    int rc = tn_sort3_impl(ctx, st, khi, klo, ktie, G);
    if (rc) return rc;
    topm_take_kernel<<<blocks_for(M + 1, 256), 256, 0, st>>>(G, M, ktie, g_prob, sel, pd_bits);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_materialise(tn_ctx* ctx, void* stream, const tn_site* site, int B, int nx, int pos, int nsites, int vstride, int Dl,
                   int Dr, const int32_t* sel, const int32_t* g_rep, const long long* g_deg, const double* g_prob,
                   const int32_t* parent, const int32_t* cell, const double* Enew, const uint8_t* vind_in,
                   const uint8_t* states_in, const int32_t* root_in, const double* RL_in, const double* A, uint8_t* vind_out,
                   uint8_t* states_out, int32_t* root_out, double* Eng_out, double* prob_out, long long* deg_out,
                   double* RL_out) {
    TN_REQUIRE(ctx && site && B >= 1, "bad arguments");
    TN_REQUIRE(Dr <= 256, "bond dimension above 256 is not supported by materialise");
    materialise_kernel<<<blocks_for((int64_t)B * 32, 128), 128, 0, as_stream(stream)>>>(
        B, nx, pos, nsites, vstride, Dl, site->nd, Dr, sel, g_rep, g_deg, g_prob, parent, cell, Enew, site->dmap, site->rmap,
        vind_in, states_in, root_in, RL_in, A, vind_out, states_out, root_out, Eng_out, prob_out, deg_out, RL_out);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_row_shift(tn_ctx* ctx, void* stream, int B, int vstride, uint8_t* vind) {
    TN_REQUIRE(ctx && B >= 1, "bad arguments");
    row_shift_kernel<<<blocks_for(B, 256), 256, 0, as_stream(stream)>>>(B, vstride, vind);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_sample(tn_ctx* ctx, void* stream, const tn_site* site, int B, int nx, int has_left, int has_up, const double* P,
              const double* uniforms, const uint8_t* vind, int vstride, const double* Eng, int32_t* parent, int32_t* cell,
              double* Enew) {
    TN_REQUIRE(ctx && site && B >= 1, "bad arguments");
    sample_kernel<<<blocks_for(B, 128), 128, 0, as_stream(stream)>>>(B, site->nS, P, uniforms, has_left, has_up, site->nl,
                                                                     site->nu, nx, vind, vstride, site->Es, site->Esl,
                                                                     site->Esu, Eng, parent, cell, Enew);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

}  // extern "C"
