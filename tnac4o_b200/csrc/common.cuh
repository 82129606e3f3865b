// Shared declarations of the tnac4o_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/tnac4o_b200.h"

// ---- optional per-primitive timing of the native drivers (tn_profile / tn_profile_read): CUDA events on the launching
// stream around every primitive call of csrc/mps_native.cu and csrc/search_native.cu, with the algorithmic flops and
// bytes of the call -- the numbers bench.py's roofline is computed from
enum { TN_P_GEMM = 0, TN_P_QR, TN_P_SVD, TN_P_MPS_OTHER, TN_P_RR, TN_P_MARGINALS, TN_P_SELECT_MERGE, TN_P_COUNT };
struct tn_prof_rec {
    int cat;
    double flops, bytes;
    cudaEvent_t e0, e1;
};

struct tn_ctx {
    bool prof_on = false;
    std::vector<tn_prof_rec> prof_recs;
    double prof_acc[TN_P_COUNT][4] = {};      // seconds, flops, bytes, calls
    int device = 0;
    int sm_count = 148;
    static constexpr int SLOTS = 7;   // independent grow-only device scratch areas
    void* scratch[SLOTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[SLOTS] = {0, 0, 0, 0, 0, 0, 0};
    uint64_t scratch_gen = 0;         // bumped whenever a slot is reallocated (captured graphs hold slot pointers)
    bool capturing = false;           // a stream capture is in progress: slots must not be reallocated
    void* pinned = nullptr;           // small pinned host buffer for scalar read-backs
    void* counters = nullptr;         // 1024 zero-initialised device words: "last CTA reduces" tickets (reset by their user)
    cudaMemPool_t pool = nullptr;     // private stream-ordered pool: no cross-stream reuse, hence no hidden dependencies
                                      // between the streams of concurrent solver instances
    int64_t launches = 0;
};

void tn_set_error(const char* fmt, ...);
int tn_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
enum { TN_SLOT_GEMM = 0, TN_SLOT_QR = 1, TN_SLOT_SVD = 2, TN_SLOT_SEARCH = 3, TN_SLOT_SORT = 4, TN_SLOT_MISC = 5, TN_SLOT_STAGE = 6 };
// stream-ordered allocation from the context's private pool (temporaries of the native drivers)
cudaError_t tn_malloc_async(tn_ctx* ctx, void** p, size_t bytes, cudaStream_t st);
void* tn_scratch(tn_ctx* ctx, int slot, size_t bytes);   // returns nullptr (and sets the error) on failure
// Process-wide lock that serialises stream captures against device-wide synchronising calls of other host threads
// (scratch reallocation): cudaDeviceSynchronize / cudaFree fail while another thread's capture is open.
void tn_capture_lock();
void tn_capture_unlock();

#define TN_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return tn_cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// cudaFuncSetAttribute once per call site and device instead of once per launch (the call costs a few microseconds of
// host time; the boundary-MPS build issues ~10^5 launches per instance and the host side is what limits 8-GPU scaling)
#define TN_FUNC_ATTR_ONCE(ctx, func, attr, value)                                                    \
    do {                                                                                             \
        static unsigned long long done__ = 0;                                                        \
        const unsigned long long bit__ = 1ull << ((ctx)->device & 63);                               \
        if (!(done__ & bit__)) {                                                                     \
            TN_CUDA(cudaFuncSetAttribute(func, attr, value));                                        \
            done__ |= bit__;                                                                         \
        }                                                                                            \
    } while (0)

#define TN_LAUNCHED(ctx)                                                       \
    do {                                                                       \
        (ctx)->launches++;                                                     \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess) return tn_cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define TN_REQUIRE(cond, msg)                                                  \
    do {                                                                       \
        if (!(cond)) {                                                         \
            tn_set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);    \
            return TN_ERR_ARG;                                                 \
        }                                                                      \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

struct tn_prof_scope {
    tn_ctx* ctx;
    cudaStream_t st;
    size_t idx = 0;
    bool on;
    tn_prof_scope(tn_ctx* c, cudaStream_t s, int cat, double flops, double bytes) : ctx(c), st(s), on(c && c->prof_on) {
        if (!on) return;
        tn_prof_rec r{cat, flops, bytes, nullptr, nullptr};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        idx = ctx->prof_recs.size();
        ctx->prof_recs.push_back(r);
    }
    ~tn_prof_scope() {
        if (on) cudaEventRecord(ctx->prof_recs[idx].e1, st);
    }
};
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device helpers ---------------------------------------------------------------------------
// Order-preserving map double -> uint64 (total order on non-NaN values), used for atomicMax/Min.
__host__ __device__ inline unsigned long long ordered_bits(double x) {
    unsigned long long u;
#ifdef __CUDA_ARCH__
    u = (unsigned long long)__double_as_longlong(x);
#else
    memcpy(&u, &x, 8);
#endif
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double from_ordered_bits(unsigned long long k) {
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// 2^floor(log2 |x|) by exponent-field extraction, the rule of mps.nfactor (mps.py:83-85):
// biased exponent e -> 2^(e-1023); e = 0 (zero / subnormal) gives 2^-1023.
__device__ __forceinline__ double pow2_floor_from_bits(unsigned long long abs_bits) {
    int e = (int)((abs_bits >> 52) & 0x7ff);
    return ldexp(1.0, e - 1023);
}
#endif
