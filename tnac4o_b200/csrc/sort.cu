// Device-wide bitonic sort of (hi, lo, tie) 3-word keys and an int32 inclusive scan.
// Used by the boundary merge (np.unique(vind, axis=0) + argsort, tnac4o.py:481-485) and the top-M selection
// (argpartition, tnac4o.py:518-526).  Keys are made unique by the tie word, so the result is a total order and
// bit-reproducible from run to run.
#include "common.cuh"

namespace {

constexpr int TILE = 2048;     // elements sorted inside shared memory by one CTA of 1024 threads

__device__ __forceinline__ bool key_less(unsigned long long h0, unsigned long long l0, unsigned long long t0,
                                         unsigned long long h1, unsigned long long l1, unsigned long long t1) {
    if (h0 != h1) return h0 < h1;
    if (l0 != l1) return l0 < l1;
    return t0 < t1;
}

__global__ void sort_pad_kernel(unsigned long long* hi, unsigned long long* lo, unsigned long long* tie, int n, int npad) {
    for (int i = n + blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += gridDim.x * blockDim.x) {
        hi[i] = ~0ull; lo[i] = ~0ull; tie[i] = ~0ull;
    }
}

// all steps (k, j) with j < TILE, for k from k_first up to k_last (k_first == k_last > TILE: only the tail of stage k)
__global__ void __launch_bounds__(1024, 1)
bitonic_local_kernel(unsigned long long* hi, unsigned long long* lo, unsigned long long* tie, int k_first, int k_last) {
    __shared__ unsigned long long sh[TILE], sl[TILE], stt[TILE];
    const int base = blockIdx.x * TILE, tid = threadIdx.x;
    for (int i = tid; i < TILE; i += 1024) { sh[i] = hi[base + i]; sl[i] = lo[base + i]; stt[i] = tie[base + i]; }
    __syncthreads();
    for (int k = k_first; k <= k_last; k <<= 1) {
        int jstart = (k > TILE) ? TILE / 2 : k / 2;
        for (int j = jstart; j > 0; j >>= 1) {
            int i = 2 * tid - (tid & (j - 1));           // index with bit j clear
            int p = i + j;
            bool up = (((base + i) & k) == 0);
            bool less = key_less(sh[p], sl[p], stt[p], sh[i], sl[i], stt[i]);
            if (less == up) {
                unsigned long long a = sh[i], b = sl[i], c = stt[i];
                sh[i] = sh[p]; sl[i] = sl[p]; stt[i] = stt[p];
                sh[p] = a; sl[p] = b; stt[p] = c;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < TILE; i += 1024) { hi[base + i] = sh[i]; lo[base + i] = sl[i]; tie[base + i] = stt[i]; }
}

__global__ void bitonic_global_kernel(unsigned long long* hi, unsigned long long* lo, unsigned long long* tie, int npad,
                                      int k, int j) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npad / 2) return;
    int i = 2 * t - (t & (j - 1));
    int p = i + j;
    bool up = ((i & k) == 0);
    unsigned long long h0 = hi[i], l0 = lo[i], t0 = tie[i], h1 = hi[p], l1 = lo[p], t1 = tie[p];
    bool less = key_less(h1, l1, t1, h0, l0, t0);
    if (less == up) {
        hi[i] = h1; lo[i] = l1; tie[i] = t1;
        hi[p] = h0; lo[p] = l0; tie[p] = t0;
    }
}

// ---- inclusive scan of int32 ------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) scan_block_kernel(const int* __restrict__ in, int* __restrict__ out, int n,
                                                             int* __restrict__ block_sums) {
    __shared__ int warp_tot[32];
    int i = blockIdx.x * 1024 + threadIdx.x;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v = (i < n) ? in[i] : 0;
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane];
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v += warp_tot[warp - 1];
    if (i < n) out[i] = v;
    if (threadIdx.x == 1023) block_sums[blockIdx.x] = v;
}

__global__ void scan_sums_kernel(int* block_sums, int nblocks) {
    // single thread block, sequential over chunks: nblocks <= a few thousand
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nblocks; ++b) { int v = block_sums[b]; block_sums[b] = run; run += v; }
    }
}

__global__ void scan_add_kernel(int* out, int n, const int* __restrict__ block_sums) {
    int i = blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
}

}  // namespace

static int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// Sorts n keys ascending by (hi, lo, tie).  The arrays must have room for next_pow2(max(n, TILE)) elements.
int tn_sort3_impl(tn_ctx* ctx, cudaStream_t st, unsigned long long* hi, unsigned long long* lo, unsigned long long* tie,
                  int n) {
    if (n <= 1) return TN_OK;
    int npad = next_pow2(n);
    if (npad < TILE) npad = TILE;
    if (npad > n) {
        sort_pad_kernel<<<ceil_div(npad - n, 256) < 1024 ? ceil_div(npad - n, 256) : 1024, 256, 0, st>>>(hi, lo, tie, n, npad);
        TN_LAUNCHED(ctx);
    }
    int tiles = npad / TILE;
    bitonic_local_kernel<<<tiles, 1024, 0, st>>>(hi, lo, tie, 2, TILE);
    TN_LAUNCHED(ctx);
    for (int k = 2 * TILE; k <= npad; k <<= 1) {
        for (int j = k / 2; j >= TILE; j >>= 1) {
            bitonic_global_kernel<<<ceil_div(npad / 2, 256), 256, 0, st>>>(hi, lo, tie, npad, k, j);
            TN_LAUNCHED(ctx);
        }
        bitonic_local_kernel<<<tiles, 1024, 0, st>>>(hi, lo, tie, k, k);
        TN_LAUNCHED(ctx);
    }
    return TN_OK;
}

int tn_sort_capacity(int n) {
    int p = next_pow2(n);
    return p < TILE ? TILE : p;
}

// inclusive scan; `tmp` needs ceil(n / 1024) ints
int tn_scan_impl(tn_ctx* ctx, cudaStream_t st, const int* in, int* out, int n, int* tmp) {
    if (n <= 0) return TN_OK;
    int nblocks = ceil_div(n, 1024);
    scan_block_kernel<<<nblocks, 1024, 0, st>>>(in, out, n, tmp);
    TN_LAUNCHED(ctx);
    if (nblocks > 1) {
        scan_sums_kernel<<<1, 32, 0, st>>>(tmp, nblocks);
        TN_LAUNCHED(ctx);
        scan_add_kernel<<<nblocks, 1024, 0, st>>>(out, n, tmp);
        TN_LAUNCHED(ctx);
    }
    return TN_OK;
}

extern "C" int tn_sort_keys(tn_ctx* ctx, void* stream, unsigned long long* hi, unsigned long long* lo,
                            unsigned long long* tie, int n) {
    TN_REQUIRE(ctx != nullptr && n >= 0, "bad arguments");
    return tn_sort3_impl(ctx, as_stream(stream), hi, lo, tie, n);
}

extern "C" int tn_sort_capacity_for(int n) { return tn_sort_capacity(n); }
