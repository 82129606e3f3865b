// FP64 DMMA GEMM with TMA-staged operands (cp.async.bulk.tensor + mbarrier) for the large row-major products of the
// boundary-MPS path (attach_AC / attach_CA, mps.py:368-380, 740-746: (8192 x 512) . (512 x 512) and its mirror).
//
// Same math as gemm.cu (mma.sync.m8n8k4.f64 -> DMMA.8x8x4, 128 x 128 x 16 CTA tiles, 8 warps of 32 x 64), but the operand
// tiles are fetched by the TMA engine: one elected thread arms an mbarrier with the byte count of a stage and issues
// nine bulk tensor copies (A: one 128 x 16 box, B: eight 16 x 16 boxes); all warps wait on the barrier's phase bit.  No
// thread spends registers or issue slots on address arithmetic for the copies, out-of-range rows / columns / k are
// zero-filled by the hardware, and the tiles land in shared memory in the 128-byte swizzle so that the 8-byte fragment
// loads of the two operands are at most two-way bank-conflicted without any padding.
//
// Eligible calls: no transposition, single batch, no split-K, operand base addresses and row pitches multiples of 16
// bytes; everything else goes through gemm.cu.  SASS evidence (UTMALDG, SYNCS, DMMA) is kept under profiles/.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int TBM = 128, TBN = 128, TBK = 16, TST = 6;
constexpr int TWM = 32, TWN = 64;
constexpr int A_BYTES = TBM * TBK * 8;            // 16 KiB
constexpr int B_SUB_BYTES = TBK * 16 * 8;         // one 16 (k) x 16 (n) box: 2 KiB
constexpr int B_BYTES = (TBN / 16) * B_SUB_BYTES; // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(256, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                double alpha, double beta, double* __restrict__ C, int ldc) {
    extern __shared__ unsigned char smraw[];
    __shared__ __align__(8) unsigned long long full[TST];
    unsigned char* base = (unsigned char*)(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);      // 128-byte swizzle atoms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % (TBM / TWM), wn = warp / (TBM / TWM);
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
    const int ktiles = (K + TBK - 1) / TBK;
    constexpr int TM = TWM / 8, TN = TWN / 8;

    if (tid == 0) {
        for (int s = 0; s < TST; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int st, int kt) {           // elected thread only
        unsigned char* sa = base + (size_t)st * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        mbar_expect_tx(&full[st], STAGE_BYTES);
        tma_load_2d(sa, &tmA, kt * TBK, m0, &full[st]);
#pragma unroll
        for (int s = 0; s < TBN / 16; ++s) tma_load_2d(sb + s * B_SUB_BYTES, &tmB, n0 + 16 * s, kt * TBK, &full[st]);
    };
    if (tid == 0)
        for (int s = 0; s < TST && s < ktiles; ++s) issue(s, s);

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // Swizzled fragment addresses, hoisted out of the k loop.  128-byte swizzle: the 16-byte chunk index of an element is
    // XORed with (row & 7) of its 128-byte row.
    //   A tile [128 rows m][16 k]: element (r, k), r = wm*32 + i*8 + g, k = kk + t  ->  r*128 + (((k>>1) ^ g) << 4) + (k&1)*8;
    //     (k>>1) ^ g = ((kk/2) ^ (g&6)) | ((t>>1) ^ (g&1)): one thread-constant per kk, the i-dependence is the immediate i*1024.
    //   B tile: eight [16 rows k][16 n] boxes; element (k, n), n = wn*64 + j*8 + g  ->
    //     (n>>4)*2048 + k*128 + ((((n&15)>>1) ^ (k&7)) << 4) + (n&1)*8,  ((n&15)>>1) ^ (k&7) = (((j&1)*4) ^ (kk&4)) | ((g>>1) ^ t):
    //     the (j, kk) part is a compile-time constant, the rest one thread-constant.
    const int baseA = (wm * TWM + g) * 128 + ((((t >> 1) ^ (g & 1))) << 4) + ((t & 1) << 3);
    int xa[TBK / 4];
#pragma unroll
    for (int q = 0; q < TBK / 4; ++q) xa[q] = ((2 * q) ^ (g & 6)) << 4;
    const int baseB = wn * (TWN / 16) * B_SUB_BYTES + t * 128 + ((((g >> 1) ^ t)) << 4) + ((g & 1) << 3);

    for (int kt = 0; kt < ktiles; ++kt) {
        const int st = kt % TST;
        const unsigned parity = (unsigned)((kt / TST) & 1);
        {
            unsigned spins = 0;
            while (!mbar_try_wait(&full[st], parity)) {
                if (++spins > (1u << 28)) __trap();          // a lost transaction must fail the launch, not hang the GPU
            }
        }
        // every warp has finished tile kt - 1: its stage is refilled (tile kt - 1 + TST) while tile kt is multiplied
        __syncthreads();
        if (tid == 0 && kt >= 1 && kt - 1 + TST < ktiles) issue((kt - 1) % TST, kt - 1 + TST);
        const unsigned char* sa = base + (size_t)st * STAGE_BYTES + baseA;
        const unsigned char* sb = base + (size_t)st * STAGE_BYTES + A_BYTES + baseB;
#pragma unroll
        for (int kk = 0; kk < TBK; kk += 4) {
            double fa[TM], fb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) fa[i] = *(const double*)(sa + xa[kk >> 2] + i * 1024);
#pragma unroll
            for (int j = 0; j < TN; ++j)
                fb[j] = *(const double*)(sb + (j >> 1) * B_SUB_BYTES + kk * 128 + (((((j & 1) << 2) ^ (kk & 4))) << 4));
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + wm * TWM + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + wn * TWN + j * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < N) {
                    double* p = C + (int64_t)row * ldc + col + e;
                    const double v = acc[i][j][e];
                    *p = (beta == 0.0) ? alpha * v : alpha * v + beta * (*p);
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// row-major (rows x cols, pitch ld doubles) -> tensor map with boxes of box_rows x 16 doubles, 128-byte swizzle
bool make_map(CUtensorMap* map, const double* ptr, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool tma_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TN_GEMM_TMA"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

}  // namespace

// returns 1 when the product was launched here, 0 when the caller must use the cp.async kernel, < 0 on error
int tn_gemm_tma_try(tn_ctx* ctx, cudaStream_t st, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                    int ldb, double beta, double* C, int ldc) {
    if (!tma_enabled() || K < 2 * TBK) return 0;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return 0;
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, A, M, K, lda, TBM) || !make_map(&tmB, B, K, N, ldb, TBK)) return 0;
    const size_t smem = (size_t)TST * STAGE_BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        TN_CUDA(cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid(ceil_div(N, TBN), ceil_div(M, TBM));
    gemm_tma_kernel<<<grid, 256, smem, st>>>(tmA, tmB, M, N, K, alpha, beta, C, ldc);
    TN_LAUNCHED(ctx);
    return 1;
}
