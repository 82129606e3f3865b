// Integer kernels of the droplet bookkeeping and the independent energy check:
//   - XOR difference of two cell-state rows -> (positions, xor patterns)      (tnac4o.py:859-861)
//   - states from a ground state and lists of droplet shapes                  (tnac4o.py:1380-1385)
//   - Ising energy of 0/1 encoded spin states from a CSR coupling matrix      (auxx.py:82-107, without the
//     dense L x L product: 2 nnz + L integer-spin MACs per state)
#include "common.cuh"

namespace {

// one warp per (winner, loser) pair; rows are parent rows with one cell overridden
__global__ void xor_diff_kernel(int npairs, int nsites, int pos, const uint8_t* __restrict__ states,
                                const int32_t* __restrict__ row_a, const int32_t* __restrict__ cell_a,
                                const int32_t* __restrict__ row_b, const int32_t* __restrict__ cell_b,
                                int16_t* __restrict__ out_pos, uint8_t* __restrict__ out_xor, int32_t* __restrict__ out_len) {
    const int lane = threadIdx.x & 31;
    const int pidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pidx >= npairs) return;
    const uint8_t* ra = states + (int64_t)row_a[pidx] * nsites;
    const uint8_t* rb = states + (int64_t)row_b[pidx] * nsites;
    int n = 0;
    for (int k0 = 0; k0 < nsites; k0 += 32) {
        int k = k0 + lane;
        uint8_t x = 0;
        if (k < nsites) {
            uint8_t va = (k == pos) ? (uint8_t)cell_a[pidx] : ra[k];
            uint8_t vb = (k == pos) ? (uint8_t)cell_b[pidx] : rb[k];
            x = va ^ vb;
        }
        unsigned mask = __ballot_sync(0xffffffffu, x != 0);
        if (x) {
            int o = n + __popc(mask & ((1u << lane) - 1));
            out_pos[(int64_t)pidx * nsites + o] = (int16_t)k;
            out_xor[(int64_t)pidx * nsites + o] = x;
        }
        n += __popc(mask);
    }
    if (lane == 0) out_len[pidx] = n;
}

// states[i] = ground XOR (all droplets of flip list i); CSR flip lists and CSR droplet dictionary; one warp per state
__global__ void apply_droplets_kernel(int nstates, int nsites, const uint8_t* __restrict__ ground,
                                      const int32_t* __restrict__ flip_ptr, const int32_t* __restrict__ flip_key,
                                      const int32_t* __restrict__ drop_ptr, const int16_t* __restrict__ drop_pos,
                                      const uint8_t* __restrict__ drop_xor, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nstates) return;
    uint8_t* row = out + (int64_t)i * nsites;
    for (int k = lane; k < nsites; k += 32) row[k] = ground[k];
    __syncwarp();
    for (int f = flip_ptr[i]; f < flip_ptr[i + 1]; ++f) {
        int key = flip_key[f];
        for (int e = drop_ptr[key] + lane; e < drop_ptr[key + 1]; e += 32) row[drop_pos[e]] ^= drop_xor[e];
        __syncwarp();
    }
}

// E[k] = sum_{(i,j,v) in CSR, i<j} v s_i s_j + sum_i v_ii s_i, s = 2 bit - 1; one warp per state, the warp's
// lanes split the coupling list and the partial sums are combined in a fixed butterfly order.
__global__ void energy_ising_kernel(int nstates, int L, const int8_t* __restrict__ bits, int64_t nnz,
                                    const int32_t* __restrict__ ci, const int32_t* __restrict__ cj,
                                    const double* __restrict__ cv, double* __restrict__ E) {
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= nstates) return;
    const int8_t* s = bits + (int64_t)k * L;
    double acc = 0.0;
    for (int64_t e = lane; e < nnz; e += 32) {
        int i = ci[e], j = cj[e];
        int si = 2 * s[i] - 1;
        int sj = (i == j) ? 1 : 2 * s[j] - 1;
        acc += cv[e] * (double)(si * sj);
    }
    acc = warp_sum(acc);
    if (lane == 0) E[k] = acc;
}

}  // namespace

extern "C" {

int tn_xor_diff(tn_ctx* ctx, void* stream, int npairs, int nsites, int pos, const uint8_t* states, const int32_t* row_a,
                const int32_t* cell_a, const int32_t* row_b, const int32_t* cell_b, int16_t* out_pos, uint8_t* out_xor,
                int32_t* out_len) {
    TN_REQUIRE(ctx && npairs >= 0 && nsites >= 1, "bad arguments");
    if (npairs == 0) return TN_OK;
    int blocks = (int)(((int64_t)npairs * 32 + 127) / 128);
    xor_diff_kernel<<<blocks, 128, 0, as_stream(stream)>>>(npairs, nsites, pos, states, row_a, cell_a, row_b, cell_b, out_pos,
                                                           out_xor, out_len);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_apply_droplets(tn_ctx* ctx, void* stream, int nstates, int nsites, const uint8_t* ground, const int32_t* flip_ptr,
                      const int32_t* flip_key, const int32_t* drop_ptr, const int16_t* drop_pos, const uint8_t* drop_xor,
                      uint8_t* out) {
    TN_REQUIRE(ctx && nstates >= 0 && nsites >= 1, "bad arguments");
    if (nstates == 0) return TN_OK;
    int blocks = (int)(((int64_t)nstates * 32 + 127) / 128);
    apply_droplets_kernel<<<blocks, 128, 0, as_stream(stream)>>>(nstates, nsites, ground, flip_ptr, flip_key, drop_ptr,
                                                                 drop_pos, drop_xor, out);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_energy_ising(tn_ctx* ctx, void* stream, int nstates, int L, const int8_t* bits, int64_t nnz, const int32_t* ci,
                    const int32_t* cj, const double* cv, double* E) {
    TN_REQUIRE(ctx && nstates >= 0 && L >= 1, "bad arguments");
    if (nstates == 0) return TN_OK;
    int blocks = (int)(((int64_t)nstates * 32 + 127) / 128);
    energy_ising_kernel<<<blocks, 128, 0, as_stream(stream)>>>(nstates, L, bits, nnz, ci, cj, cv, E);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

}  // extern "C"
