// Microbenchmark: cost of cluster.sync() with and without distributed-shared-memory stores (sm_100a).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void __cluster_dims__(8, 1, 1) k(int iters, long long* out, double* sink) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double buf[2][8][16];
    int rank = cluster.block_rank();
    cluster.sync();
    long long t0 = clock64();
    double acc = 0;
    for (int i = 0; i < iters; ++i) {
        if (MODE >= 1 && threadIdx.x < 16) {
            for (int r = 0; r < 8; ++r) *cluster.map_shared_rank(&buf[i & 1][rank][threadIdx.x], r) = (double)i;
        }
        if (MODE == 3) __syncthreads();
        if (MODE != 4) cluster.sync();
        else __syncthreads();
        if (MODE >= 1) acc += buf[i & 1][(rank + 1) & 7][threadIdx.x & 15];
        if (MODE == 2) { __syncthreads(); __syncthreads(); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 12345.678) sink[0] = acc;
}

int main() {
    long long* d; double* s;
    cudaMalloc(&d, 8); cudaMalloc(&s, 8);
    const int iters = 2000;
    for (int threads : {256, 1024}) {
        long long h;
#define RUN(M, name) k<M><<<8, threads>>>(iters, d, s); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); \
        printf("threads %4d  %-42s %8.1f cycles/iter\n", threads, name, (double)h / iters);
        RUN(0, "cluster.sync only")
        RUN(1, "16x8 DSMEM stores + cluster.sync + read")
        RUN(2, "same + 2 __syncthreads")
        RUN(3, "stores + __syncthreads + cluster.sync")
        RUN(4, "stores + __syncthreads only (no cluster)")
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
