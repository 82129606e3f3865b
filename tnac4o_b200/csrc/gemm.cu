// FP64 GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a; tcgen05 has no
// f64 kind).  Operands are staged global -> shared with a 3-stage cp.async ring; shared tiles are padded so
// that both fragment patterns (k-contiguous / mn-contiguous) are bank-conflict free for 8-byte loads.
//
// Replaces the np.tensordot -> dgemm calls of the reference's boundary-MPS code (mps.py:655-769) and of
// tnac4o.py:1779-1794.  Row-major, arbitrary M, N, K and leading dimensions, strided batch, deterministic
// split-K (partials in context scratch, reduced in a fixed order).
#include <stdlib.h>

#include "common.cuh"

int tn_throughput_mode();      // svd.cu: many solver instances share the GPU

namespace {

constexpr int BK = 16;
constexpr int STAGES = 3;

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = valid ? 8 : 0;   // src-size 0 -> the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Shared-memory tile of one operand: R "outer" rows (M or N direction) by BK.
//  KC = true : stored [outer][BK + 4]   (global operand is k-contiguous)
//  KC = false: stored [BK][outer + 4]   (global operand is outer-contiguous)
template <int OUTER, bool KC>
struct Tile {
    static constexpr int STRIDE = KC ? (BK + 4) : (OUTER + 4);
    static constexpr int SIZE = KC ? OUTER * STRIDE : BK * STRIDE;
    __device__ static __forceinline__ int at(int o, int k) { return KC ? o * STRIDE + k : k * STRIDE + o; }
};

// global -> shared copy of one OUTER x BK tile.  `outer_n`/`k_n` are the remaining valid extents.
template <int OUTER, bool KC, int THREADS>
__device__ __forceinline__ void load_tile(double* s, const double* g, int ld, int outer_n, int k_n, int tid) {
    constexpr int TOTAL = OUTER * BK;
#pragma unroll
    for (int i = 0; i < TOTAL / THREADS; ++i) {
        int idx = tid + i * THREADS;
        int o, k;
        if (KC) { o = idx / BK; k = idx % BK; } else { k = idx / OUTER; o = idx % OUTER; }
        bool ok = (o < outer_n) && (k < k_n);
        const double* src = ok ? (KC ? g + (int64_t)o * ld + k : g + (int64_t)k * ld + o) : g;
        cp_async8(s + Tile<OUTER, KC>::at(o, k), src, ok);
    }
}

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int lda, int64_t strideA,
            const double* __restrict__ B, int ldb, int64_t strideB, double beta, double* __restrict__ C, int ldc,
            int64_t strideC, int splitk, int kchunk, double* __restrict__ partial, const int* __restrict__ bmap) {
    constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    constexpr int THREADS = WARPS_M * WARPS_N * 32;
    constexpr int TM = WM / 8, TN = WN / 8;
    using TA = Tile<BM, AKC>;
    using TB = Tile<BN, BKC>;
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + STAGES * TA::SIZE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int g = lane >> 2, t = lane & 3;
    const int batch = blockIdx.z / splitk, ks = blockIdx.z % splitk;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = ks * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const int ktiles = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

    // grouped mode: batch entry z multiplies by B operand number bmap[z]; negative = unused entry
    int bsel = batch;
    if (bmap) {
        bsel = bmap[batch];
        if (bsel < 0) return;
    }
    const double* Ab = A + batch * strideA + (AKC ? (int64_t)m0 * lda : (int64_t)m0);
    const double* Bb = B + bsel * strideB + (BKC ? (int64_t)n0 * ldb : (int64_t)n0);

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int kt) {
        if (kt < ktiles) {
            int k0 = kbeg + kt * BK;
            int st = kt % STAGES;
            const double* ga = AKC ? Ab + k0 : Ab + (int64_t)k0 * lda;
            const double* gb = BKC ? Bb + k0 : Bb + (int64_t)k0 * ldb;
            load_tile<BM, AKC, THREADS>(sA + st * TA::SIZE, ga, lda, M - m0, kend - k0, tid);
            load_tile<BN, BKC, THREADS>(sB + st * TB::SIZE, gb, ldb, N - n0, kend - k0, tid);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);
        const double* a = sA + (kt % STAGES) * TA::SIZE;
        const double* b = sB + (kt % STAGES) * TB::SIZE;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double fa[TM], fb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) fa[i] = a[TA::at(wm * WM + i * 8 + g, kk + t)];
#pragma unroll
            for (int j = 0; j < TN; ++j) fb[j] = b[TB::at(wn * WN + j * 8 + g, kk + t)];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue
    const bool to_partial = (splitk > 1);
    double* out = to_partial ? partial + ((int64_t)blockIdx.z) * M * N : C + batch * strideC;
    const int ldo = to_partial ? N : ldc;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int row = m0 + wm * WM + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int col = n0 + wn * WN + j * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < N) {
                    double* p = out + (int64_t)row * ldo + col + e;
                    double v = acc[i][j][e];
                    if (to_partial) *p = v;
                    else *p = (beta == 0.0) ? alpha * v : alpha * v + beta * (*p);
                }
            }
        }
    }
}

__global__ void splitk_reduce_kernel(int M, int N, int splitk, int batch, const double* __restrict__ partial,
                                     double alpha, double beta, double* __restrict__ C, int ldc, int64_t strideC) {
    int64_t total = (int64_t)batch * M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(i / ((int64_t)M * N));
        int64_t r = i % ((int64_t)M * N);
        int row = (int)(r / N), col = (int)(r % N);
        double s = 0.0;
        for (int k = 0; k < splitk; ++k) s += partial[((int64_t)(b * splitk + k)) * M * N + r];
        double* p = C + b * strideC + (int64_t)row * ldc + col;
        *p = (beta == 0.0) ? alpha * s : alpha * s + beta * (*p);
    }
}

template <int BM, int BN, int WM, int WN>
int launch_cfg(tn_ctx* ctx, cudaStream_t st, int tA, int tB, int M, int N, int K, double alpha, const double* A, int lda,
               int64_t sA, const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC, int batch,
               const int* bmap = nullptr) {
    constexpr int THREADS = (BM / WM) * (BN / WN) * 32;
    const bool akc = !tA, bkc = tB;
    size_t smem = 0;
    {
        size_t a = akc ? Tile<BM, true>::SIZE : Tile<BM, false>::SIZE;
        size_t b = bkc ? Tile<BN, true>::SIZE : Tile<BN, false>::SIZE;
        smem = STAGES * (a + b) * sizeof(double);
    }
    int tiles = ceil_div(M, BM) * ceil_div(N, BN) * batch;
    int splitk = 1;
    static const int cta_cap = [] { const char* e = getenv("TN_GEMM_SPLITK_CTAS"); return e ? atoi(e) : 0; }();
    // CTAs a split-K product may spread over: the whole GPU when it runs alone, a quarter of it in throughput mode
    // (measured with 32 concurrent instances: 0.440 vs 0.476 s per instance, profiles/r2c_bench_batch_sweep.txt)
    const int target = cta_cap > 0 ? cta_cap : (tn_throughput_mode() ? ctx->sm_count / 4 : ctx->sm_count);
    if (tiles * 2 <= target && K >= 512 && !bmap) {
        splitk = min(min(target / tiles, K / 256), 32);
        if (splitk < 1) splitk = 1;
    }
    int kchunk = ceil_div(ceil_div(K, splitk), BK) * BK;
    splitk = ceil_div(K, kchunk);
    double* partial = nullptr;
    if (splitk > 1) {
        partial = (double*)tn_scratch(ctx, TN_SLOT_GEMM, (size_t)splitk * batch * M * N * sizeof(double));
        if (!partial) return TN_ERR_NOMEM;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), batch * splitk);
#define TN_GEMM_LAUNCH(AK, BKc)                                                                                        \
    do {                                                                                                               \
        auto kern = gemm_kernel<BM, BN, WM, WN, AK, BKc>;                                                              \
        TN_FUNC_ATTR_ONCE(ctx, kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
        kern<<<grid, THREADS, smem, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, splitk, kchunk,    \
                                          partial, bmap);                                                              \
    } while (0)
    if (akc && bkc) TN_GEMM_LAUNCH(true, true);
    else if (akc && !bkc) TN_GEMM_LAUNCH(true, false);
    else if (!akc && bkc) TN_GEMM_LAUNCH(false, true);
    else TN_GEMM_LAUNCH(false, false);
#undef TN_GEMM_LAUNCH
    TN_LAUNCHED(ctx);
    if (splitk > 1) {
        int64_t total = (int64_t)batch * M * N;
        int64_t want_blocks = (total + 255) / 256;
        int blocks = (int)(want_blocks < 4 * ctx->sm_count ? want_blocks : 4 * ctx->sm_count);
        splitk_reduce_kernel<<<blocks, 256, 0, st>>>(M, N, splitk, batch, partial, alpha, beta, C, ldc, sC);
        TN_LAUNCHED(ctx);
    }
    return TN_OK;
}

}  // namespace

int tn_gemm_tma_try(tn_ctx* ctx, cudaStream_t st, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                    int ldb, double beta, double* C, int ldc);

// internal entry used by the other translation units (qr.cu)
int tn_gemm_impl(tn_ctx* ctx, cudaStream_t st, int tA, int tB, int M, int N, int K, double alpha, const double* A,
                 int lda, int64_t sA, const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC,
                 int batch) {
    if (M <= 0 || N <= 0 || batch <= 0) return TN_OK;
    // big tile when it still fills the machine, small tile otherwise
    int big_tiles = ceil_div(M, 128) * ceil_div(N, 128) * batch;
    if (big_tiles >= ctx->sm_count && M >= 128 && N >= 128) {
        if (!tA && !tB && batch == 1) {
            // TMA-staged operands (gemm_tma.cu) when the operands are 16-byte aligned; 1 = launched, 0 = not eligible
            const int rc = tn_gemm_tma_try(ctx, st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
            if (rc != 0) return rc < 0 ? rc : TN_OK;
        }
        return launch_cfg<128, 128, 32, 64>(ctx, st, tA, tB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
    }
    return launch_cfg<64, 64, 32, 32>(ctx, st, tA, tB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
}

// Grouped GEMM (search.cu): ntiles row tiles of `tile_rows` rows each; tile z computes
//   C[z * tile_rows ..][0:N] = X[z * tile_rows ..][0:K] . B_{bmap[z]} (K x N, row-major),  bmap[z] < 0 = skip.
// No split-K: every output row is accumulated in ascending k whatever the tile configuration, so a row's result does
// not depend on how many other rows are in the call.
int tn_gemm_grouped_impl(tn_ctx* ctx, cudaStream_t st, int tile_rows, int ntiles, int N, int K, const double* X, int ldx,
                         const double* B, int ldb, int64_t strideB, const int* bmap, double* C, int ldc) {
    if (ntiles <= 0 || N <= 0) return TN_OK;
    if (tile_rows == 128)
        return launch_cfg<128, 128, 32, 64>(ctx, st, 0, 0, 128, N, K, 1.0, X, ldx, (int64_t)128 * ldx, B, ldb, strideB, 0.0, C,
                                            ldc, (int64_t)128 * ldc, ntiles, bmap);
    if (tile_rows == 64)
        return launch_cfg<64, 64, 32, 32>(ctx, st, 0, 0, 64, N, K, 1.0, X, ldx, (int64_t)64 * ldx, B, ldb, strideB, 0.0, C, ldc,
                                          (int64_t)64 * ldc, ntiles, bmap);
    tn_set_error("grouped GEMM: tile_rows must be 64 or 128");
    return TN_ERR_ARG;
}

extern "C" int tn_gemm(tn_ctx* ctx, void* stream, int transA, int transB, int M, int N, int K, double alpha,
                       const double* A, int lda, int64_t strideA, const double* B, int ldb, int64_t strideB, double beta,
                       double* C, int ldc, int64_t strideC, int batch) {
    TN_REQUIRE(ctx != nullptr, "null context");
    TN_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "negative dimension");
    TN_REQUIRE(lda >= 1 && ldb >= 1 && ldc >= 1, "bad leading dimension");
    return tn_gemm_impl(ctx, as_stream(stream), transA, transB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C,
                        ldc, strideC, batch);
}
