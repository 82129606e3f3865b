// Small memory-bound helpers of the boundary-MPS path: transpose, MPO application, power-of-two
// normalisation (mps.nfactor), Schmidt-spectrum distance.
#include "common.cuh"

namespace {

__global__ void transpose_kernel(int m, int n, const double* __restrict__ in, int ldin, double* __restrict__ out, int ldout) {
    __shared__ double tile[32][33];
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;          // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int r = by + j, c = bx + tx;
        if (r < m && c < n) tile[j][tx] = in[(int64_t)r * ldin + c];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int r = bx + j, c = by + tx;                  // out is n x m
        if (r < n && c < m) out[(int64_t)r * ldout + c] = tile[tx][j];
    }
}

// conj: out[(a,l), u, (b,r)] = sum_p A[a,p,b] W[l,p,r,u]      (mps.py:755-757)
// else: out[(l,a), o, (r,b)] = sum_p W[l,o,r,p] A[a,p,b]      (mps.py:759-760)
// One thread per output element; the p-sum (<= 256 terms) runs in ascending p like the dgemm it replaces is
// free to; W (<= 512 KiB) and A stay in L2.  The kernel is bound by the 8 B/element output stream.
__global__ void mpo_apply_kernel(int conj, int Dl, int dp, int Dr, int wl, int wr, int du, const double* __restrict__ A,
                                 const double* __restrict__ W, double* __restrict__ out) {
    const int64_t ncol = (int64_t)Dr * wr;
    const int64_t total = (int64_t)Dl * wl * du * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t col = i % ncol;
        int64_t rest = i / ncol;
        int u = (int)(rest % du);
        int row = (int)(rest / du);
        int a, l, b, r;
        if (conj) { a = row / wl; l = row % wl; b = (int)(col / wr); r = (int)(col % wr); }
        else { l = row / Dl; a = row % Dl; r = (int)(col / Dr); b = (int)(col % Dr); }
        double s = 0.0;
        if (conj) {
            // W[l,p,r,u]: ((l*dp + p)*wr + r)*du + u
            const double* w = W + ((int64_t)l * dp * wr + r) * du + u;
            const double* x = A + (int64_t)a * dp * Dr + b;
            for (int p = 0; p < dp; ++p) s += x[(int64_t)p * Dr] * w[(int64_t)p * wr * du];
        } else {
            // W[l,o,r,p] with o = u: ((l*du + u)*wr + r)*dp + p
            const double* w = W + (((int64_t)l * du + u) * wr + r) * dp;
            const double* x = A + (int64_t)a * dp * Dr + b;
            for (int p = 0; p < dp; ++p) s += w[p] * x[(int64_t)p * Dr];
        }
        out[i] = s;
    }
}

__global__ void maxabs_kernel(const double* __restrict__ x, int64_t n, unsigned long long* maxabs_bits) {
    double m = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(m));
}

__global__ void pow2_scale_kernel(double* __restrict__ x, int64_t n, const unsigned long long* maxabs_bits,
                                  double* log2_accum) {
    const unsigned long long bits = *maxabs_bits;
    const int e = (int)((bits >> 52) & 0x7ff);
    if (n == 1) {
        // mps.py:778-780 / 793-795: a 1 x 1 centre matrix becomes exactly 1 (its sign is folded into Q by the caller;
        // with the non-negative-diagonal QR the entry is already >= 0)
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            x[0] = 1.0;
            if (log2_accum) *log2_accum += (double)(e - 1023);
        }
        return;
    }
    const double inv = ldexp(1.0, 1023 - e);      // exact reciprocal of 2^(e-1023); e = 0 -> 2^1023
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= inv;
    if (log2_accum && blockIdx.x == 0 && threadIdx.x == 0) *log2_accum += (double)(e - 1023);
}

__global__ void diff_norm_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) { double d = a[i] - b[i]; s += d * d; }
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = sqrt(s);
}

}  // namespace

static inline int grid_for(tn_ctx* ctx, int64_t total, int threads) {
    int64_t b = (total + threads - 1) / threads;
    int64_t cap = (int64_t)16 * ctx->sm_count;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

extern "C" int tn_transpose(tn_ctx* ctx, void* stream, int m, int n, const double* in, int ldin, double* out, int ldout) {
    TN_REQUIRE(ctx != nullptr && m >= 1 && n >= 1, "bad arguments");
    dim3 grid(ceil_div(n, 32), ceil_div(m, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, as_stream(stream)>>>(m, n, in, ldin, out, ldout);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

extern "C" int tn_mpo_apply(tn_ctx* ctx, void* stream, int conj, int Dl, int dp, int Dr, int wl, int wr, int du,
                            const double* A, const double* W, double* out) {
    TN_REQUIRE(ctx != nullptr && Dl >= 1 && dp >= 1 && Dr >= 1 && wl >= 1 && wr >= 1 && du >= 1, "bad arguments");
    int64_t total = (int64_t)Dl * wl * du * Dr * wr;
    mpo_apply_kernel<<<grid_for(ctx, total, 256), 256, 0, as_stream(stream)>>>(conj, Dl, dp, Dr, wl, wr, du, A, W, out);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

extern "C" int tn_maxabs(tn_ctx* ctx, void* stream, const double* x, int64_t n, unsigned long long* maxabs_bits) {
    TN_REQUIRE(ctx != nullptr && n >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    TN_CUDA(cudaMemsetAsync(maxabs_bits, 0, sizeof(unsigned long long), st));
    maxabs_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(x, n, maxabs_bits);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

extern "C" int tn_pow2_scale(tn_ctx* ctx, void* stream, double* x, int64_t n, const unsigned long long* maxabs_bits,
                             double* log2_accum) {
    TN_REQUIRE(ctx != nullptr && n >= 1, "bad arguments");
    pow2_scale_kernel<<<grid_for(ctx, n, 256), 256, 0, as_stream(stream)>>>(x, n, maxabs_bits, log2_accum);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

extern "C" int tn_diff_norm(tn_ctx* ctx, void* stream, const double* a, const double* b, int n, double* out) {
    TN_REQUIRE(ctx != nullptr && n >= 1, "bad arguments");
    diff_norm_kernel<<<1, 32, 0, as_stream(stream)>>>(a, b, n, out);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-site PEPS tables built on the device from the small exponent tables (tnac4o.py:1566-1607 without ever
// forming the dense 5-leg tensor, and without shipping 1.5 MiB per site over PCIe):
//   Wlu[l][u][s]     = exp((E0[s] + E1[s][l]) + E4[s][u]) * Xu[u] * Xl[l] * Xr[r(s)] * Xd[d(s)]   (same product order)
//   WtrU[u][l][d][r] = sum over s with (d(s), r(s)) = (d, r), ascending s                         (np.sum(axis=0))
//   Wmpo[l][d][r][u] = the same numbers in the MPO leg order
namespace {

__global__ void site_wlu_kernel(int nS, int nl, int nu, const double* __restrict__ E0, const double* __restrict__ E1,
                                const double* __restrict__ E4, const double* __restrict__ Xu, const double* __restrict__ Xl,
                                const double* __restrict__ Xr, const double* __restrict__ Xd, const uint8_t* __restrict__ dmap,
                                const uint8_t* __restrict__ rmap, double* __restrict__ Wlu) {
    int64_t total = (int64_t)nl * nu * nS;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(i % nS);
        int u = (int)((i / nS) % nu);
        int l = (int)(i / ((int64_t)nS * nu));
        double w = exp((E0[s] + E1[(int64_t)s * nl + l]) + E4[(int64_t)s * nu + u]);
        w = w * Xu[u];
        w = w * Xl[l];
        w = w * Xr[rmap[s]];
        w = w * Xd[dmap[s]];
        Wlu[i] = w;
    }
}

__global__ void site_trace_kernel(int nS, int nl, int nd, int nr, int nu, const double* __restrict__ Wlu,
                                  const uint8_t* __restrict__ dmap, const uint8_t* __restrict__ rmap, double* __restrict__ WtrU,
                                  double* __restrict__ Wmpo) {
    int64_t total = (int64_t)nu * nl * nd * nr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i % nr);
        int d = (int)((i / nr) % nd);
        int l = (int)((i / ((int64_t)nr * nd)) % nl);
        int u = (int)(i / ((int64_t)nr * nd * nl));
        const double* w = Wlu + ((int64_t)l * nu + u) * nS;
        double acc = 0.0;
        for (int s = 0; s < nS; ++s)
            if (dmap[s] == d && rmap[s] == r) acc += w[s];
        WtrU[i] = acc;
        Wmpo[(((int64_t)l * nd + d) * nr + r) * nu + u] = acc;
    }
}

}  // namespace

extern "C" int tn_build_site_tables(tn_ctx* ctx, void* stream, int nS, int nl, int nd, int nr, int nu, const double* E0,
                                    const double* E1, const double* E4, const double* Xu, const double* Xl, const double* Xr,
                                    const double* Xd, const uint8_t* dmap, const uint8_t* rmap, double* Wlu, double* WtrU,
                                    double* Wmpo) {
    TN_REQUIRE(ctx && nS >= 1 && nl >= 1 && nd >= 1 && nr >= 1 && nu >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    site_wlu_kernel<<<grid_for(ctx, (int64_t)nl * nu * nS, 256), 256, 0, st>>>(nS, nl, nu, E0, E1, E4, Xu, Xl, Xr, Xd, dmap, rmap, Wlu);
    TN_LAUNCHED(ctx);
    site_trace_kernel<<<grid_for(ctx, (int64_t)nu * nl * nd * nr, 256), 256, 0, st>>>(nS, nl, nd, nr, nu, Wlu, dmap, rmap, WtrU, Wmpo);
    TN_LAUNCHED(ctx);
    return TN_OK;
}
