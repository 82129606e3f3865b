// Level-synchronous expansion of the droplet tree into low-energy states -- replaces the Python enumeration
// _exc_unpack_v1 (tnac4o.py:2295-2335) and the per-state XOR loop of decode_low_energy_states (tnac4o.py:1360-1389).
//
// The reference walks the lattice sites nn = N-1 .. 0 and keeps, for every partial combination of droplets, its
// accumulated excitation energy, its list of droplet keys and a stack of tree nodes whose children may still be
// added.  At site nn every combination (including the ones created at this very site) spawns one new combination
// per child `ee` of its stack top with  ee.last == nn  and  E + ee.dE <= max_dEng  (children are scanned in order
// and the scan stops at the first child with last > nn); afterwards the lowest max_states combinations are kept and
// every stack is popped while its top has first >= nn.
//
// Here the tree is a flat array of nodes (dE, key, first, last, child range) in HBM and a combination is three words:
// its energy, the id of its creation record, and the id of the record that is its current stack top.  Records are
// immutable and shared: record r = (tree node, record of the combination it was created from, stack frame below
// it), so the droplet list of a combination is the chain of `from` links and its stack is the chain of `below`
// links -- nothing is copied when a combination spawns another.  One site = waves of {count, scan, emit} kernels
// over the frontier (first all combinations, then only the ones the previous wave created), an optional top-K cut
// by a device-wide sort, and one pop kernel.  Sites at which no tree node ends are skipped (the pops they would have
// made are implied by the pop threshold of the next processed site).  Energies are accumulated exactly as the
// reference does (parent energy + dE, in creation order), so they are bit-identical to it.
#include <algorithm>
#include <vector>

#include "common.cuh"

int tn_sort3_impl(tn_ctx* ctx, cudaStream_t st, unsigned long long* hi, unsigned long long* lo, unsigned long long* tie,
                  int n);
int tn_sort_capacity(int n);
int tn_scan_impl(tn_ctx* ctx, cudaStream_t st, const int* in, int* out, int n, int* tmp);

namespace {

struct Tree {
    const double* dE;
    const int32_t* key;
    const int32_t* first;
    const int32_t* last;
    const int32_t* child_ptr;
    const int32_t* child_idx;
};

// children of the stack top that end at site nn and fit the energy bound (scan order and stop rule of the reference)
__global__ void decode_count_kernel(Tree t, int nn, double max_dE, int lo, int hi, const double* __restrict__ E,
                                    const int32_t* __restrict__ frame, const int32_t* __restrict__ rec_node,
                                    int* __restrict__ cnt) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const int node = rec_node[frame[i]];
    const double e = E[i];
    int c = 0;
    for (int k = t.child_ptr[node]; k < t.child_ptr[node + 1]; ++k) {
        const int ch = t.child_idx[k];
        const int last = t.last[ch];
        if (last == nn && e + t.dE[ch] <= max_dE) ++c;
        else if (last > nn) break;
    }
    cnt[i - lo] = c;
}

__global__ void decode_emit_kernel(Tree t, int nn, double max_dE, int lo, int hi, int n, int nrec,
                                   const int* __restrict__ incl /* inclusive scan of cnt */, double* __restrict__ E,
                                   int32_t* __restrict__ rec, int32_t* __restrict__ frame, int32_t* __restrict__ rec_node,
                                   int32_t* __restrict__ rec_from, int32_t* __restrict__ rec_below) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const int node = rec_node[frame[i]];
    const double e = E[i];
    int o = (i == lo) ? 0 : incl[i - lo - 1];
    for (int k = t.child_ptr[node]; k < t.child_ptr[node + 1]; ++k) {
        const int ch = t.child_idx[k];
        const int last = t.last[ch];
        if (last == nn && e + t.dE[ch] <= max_dE) {
            const int r = nrec + o, c = n + o;
            rec_node[r] = ch;
            rec_from[r] = rec[i];
            rec_below[r] = frame[i];
            E[c] = e + t.dE[ch];
            rec[c] = r;
            frame[c] = r;
            ++o;
        } else if (last > nn) break;
    }
}

// pop every stack while its top starts at or after site `thr`
__global__ void decode_pop_kernel(Tree t, int thr, int n, int32_t* __restrict__ frame, const int32_t* __restrict__ rec_node,
                                  const int32_t* __restrict__ rec_below) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = frame[i];
    while (t.first[rec_node[f]] >= thr) f = rec_below[f];
    frame[i] = f;
}

__global__ void decode_keys_kernel(int n, const double* __restrict__ E, unsigned long long* hi, unsigned long long* lo,
                                   unsigned long long* tie) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hi[i] = ordered_bits(E[i]);
    lo[i] = (unsigned long long)i;
    tie[i] = 0ull;
}

__global__ void decode_gather_kernel(int n, const unsigned long long* __restrict__ order, const double* __restrict__ E,
                                     const int32_t* __restrict__ rec, const int32_t* __restrict__ frame,
                                     double* __restrict__ E2, int32_t* __restrict__ rec2, int32_t* __restrict__ frame2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = (int)order[i];
    E2[i] = E[s];
    rec2[i] = rec[s];
    frame2[i] = frame ? frame[s] : 0;
}

// state i = ground XOR every droplet on the `from` chain of combination i; one warp per state
__global__ void decode_states_kernel(int nstates, int nsites, const uint8_t* __restrict__ ground,
                                     const int32_t* __restrict__ rec, const int32_t* __restrict__ rec_node,
                                     const int32_t* __restrict__ rec_from, const int32_t* __restrict__ node_key,
                                     const int32_t* __restrict__ drop_ptr, const int16_t* __restrict__ drop_pos,
                                     const uint8_t* __restrict__ drop_xor, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= nstates) return;
    uint8_t* row = out + (int64_t)i * nsites;
    for (int k = lane; k < nsites; k += 32) row[k] = ground[k];
    __syncwarp();
    for (int r = rec[i]; r > 0; r = rec_from[r]) {           // record 0 is the root (no droplet)
        const int key = node_key[rec_node[r]];
        for (int e = drop_ptr[key] + lane; e < drop_ptr[key + 1]; e += 32) row[drop_pos[e]] ^= drop_xor[e];
        __syncwarp();
    }
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int ensure(size_t want, cudaStream_t st, size_t keep_bytes) {
        if (want <= bytes) return TN_OK;
        size_t cap = std::max(want, bytes * 2);
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, cap);
        if (e != cudaSuccess) return tn_cuda_fail(e, "cudaMalloc(decode)", __FILE__, __LINE__);
        if (p && keep_bytes) {
            e = cudaMemcpyAsync(q, p, keep_bytes, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) { cudaFree(q); return tn_cuda_fail(e, "cudaMemcpyAsync(decode)", __FILE__, __LINE__); }
        }
        if (p) { cudaStreamSynchronize(st); cudaFree(p); }
        p = q; bytes = cap;
        return TN_OK;
    }
    template <typename T> T* as() { return (T*)p; }
};

}  // namespace

struct tn_decode {
    DevBuf E, rec, frame, rec_node, rec_from, rec_below, cnt, scan, tmp, khi, klo, ktie, E2, rec2, frame2;
    DevBuf tree_dE, tree_key, tree_first, tree_last, tree_cptr, tree_cidx;
    int n = 0, nrec = 0, nsites = 0;
    int64_t launches_waves = 0;
    cudaStream_t st = nullptr;
};

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

static int decode_sort_by_energy(tn_ctx* ctx, tn_decode* d, int keep, bool with_frames) {
    // combinations sorted by (energy, creation slot): the first `keep` survive, in this order
    cudaStream_t st = d->st;
    const int n = d->n;
    const size_t cap = (size_t)tn_sort_capacity(n) * sizeof(unsigned long long);
    TRY(d->khi.ensure(cap, st, 0)); TRY(d->klo.ensure(cap, st, 0)); TRY(d->ktie.ensure(cap, st, 0));
    decode_keys_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, d->E.as<double>(), d->khi.as<unsigned long long>(),
                                                         d->klo.as<unsigned long long>(), d->ktie.as<unsigned long long>());
    TN_LAUNCHED(ctx);
    TRY(tn_sort3_impl(ctx, st, d->khi.as<unsigned long long>(), d->klo.as<unsigned long long>(), d->ktie.as<unsigned long long>(), n));
    TRY(d->E2.ensure((size_t)keep * sizeof(double), st, 0));
    TRY(d->rec2.ensure((size_t)keep * sizeof(int32_t), st, 0));
    TRY(d->frame2.ensure((size_t)keep * sizeof(int32_t), st, 0));
    decode_gather_kernel<<<ceil_div(keep, 256), 256, 0, st>>>(keep, d->klo.as<unsigned long long>(), d->E.as<double>(),
                                                              d->rec.as<int32_t>(), with_frames ? d->frame.as<int32_t>() : nullptr,
                                                              d->E2.as<double>(), d->rec2.as<int32_t>(), d->frame2.as<int32_t>());
    TN_LAUNCHED(ctx);
    TN_CUDA(cudaMemcpyAsync(d->E.p, d->E2.p, (size_t)keep * sizeof(double), cudaMemcpyDeviceToDevice, st));
    TN_CUDA(cudaMemcpyAsync(d->rec.p, d->rec2.p, (size_t)keep * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (with_frames) TN_CUDA(cudaMemcpyAsync(d->frame.p, d->frame2.p, (size_t)keep * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    d->n = keep;
    return TN_OK;
}

extern "C" {

/* Enumerates all droplet combinations with excitation energy <= max_dEng (at most max_states, the lowest ones), sorted by
 * energy.  Tree in host arrays: node 0 is the root (dE 0, first -1, last nsites - 1, children = the first layer);
 * child_ptr (nnodes + 1) / child_idx list the children of every node in the reference's order.  Returns the handle and
 * the number of combinations; tn_decode_fetch writes energies and states. */
int tn_decode_enumerate(tn_ctx* ctx, void* stream, int nsites, int nnodes, const double* h_dE, const int32_t* h_key,
                        const int32_t* h_first, const int32_t* h_last, const int32_t* h_child_ptr, const int32_t* h_child_idx,
                        double max_dEng, int64_t max_states, tn_decode** out, int64_t* h_count) {
    TN_REQUIRE(ctx && out && h_count && nsites >= 1 && nnodes >= 1 && max_states >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    tn_decode* d = new tn_decode();
    d->st = st; d->nsites = nsites;
    const int nchild = h_child_ptr[nnodes];
    auto upload = [&](DevBuf& b, const void* src, size_t bytes) -> int {
        TRY(b.ensure(bytes ? bytes : 8, st, 0));
        if (bytes) TN_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
        return TN_OK;
    };
    int rc = 0;
    if (!rc) rc = upload(d->tree_dE, h_dE, (size_t)nnodes * sizeof(double));
    if (!rc) rc = upload(d->tree_key, h_key, (size_t)nnodes * sizeof(int32_t));
    if (!rc) rc = upload(d->tree_first, h_first, (size_t)nnodes * sizeof(int32_t));
    if (!rc) rc = upload(d->tree_last, h_last, (size_t)nnodes * sizeof(int32_t));
    if (!rc) rc = upload(d->tree_cptr, h_child_ptr, (size_t)(nnodes + 1) * sizeof(int32_t));
    if (!rc) rc = upload(d->tree_cidx, h_child_idx, (size_t)nchild * sizeof(int32_t));
    if (rc) { cudaStreamSynchronize(st); delete d; return rc; }
    Tree t{d->tree_dE.as<double>(), d->tree_key.as<int32_t>(), d->tree_first.as<int32_t>(), d->tree_last.as<int32_t>(),
           d->tree_cptr.as<int32_t>(), d->tree_cidx.as<int32_t>()};
    // sites at which some node ends, descending
    std::vector<char> ends(nsites, 0);
    for (int k = 1; k < nnodes; ++k) if (h_last[k] >= 0 && h_last[k] < nsites) ends[h_last[k]] = 1;

    auto body = [&]() -> int {
        const size_t cap0 = 1 << 16;
        TRY(d->E.ensure(cap0 * sizeof(double), st, 0));
        TRY(d->rec.ensure(cap0 * sizeof(int32_t), st, 0));
        TRY(d->frame.ensure(cap0 * sizeof(int32_t), st, 0));
        TRY(d->rec_node.ensure(cap0 * sizeof(int32_t), st, 0));
        TRY(d->rec_from.ensure(cap0 * sizeof(int32_t), st, 0));
        TRY(d->rec_below.ensure(cap0 * sizeof(int32_t), st, 0));
        // combination 0 / record 0: the empty combination on the root node
        TN_CUDA(cudaMemsetAsync(d->E.p, 0, sizeof(double), st));
        TN_CUDA(cudaMemsetAsync(d->rec.p, 0, sizeof(int32_t), st));
        TN_CUDA(cudaMemsetAsync(d->frame.p, 0, sizeof(int32_t), st));
        TN_CUDA(cudaMemsetAsync(d->rec_node.p, 0, sizeof(int32_t), st));
        TN_CUDA(cudaMemsetAsync(d->rec_from.p, 0, sizeof(int32_t), st));
        TN_CUDA(cudaMemsetAsync(d->rec_below.p, 0, sizeof(int32_t), st));
        d->n = 1; d->nrec = 1;
        int* h_total = (int*)((char*)ctx->pinned + 512);
        for (int nn = nsites - 1; nn >= 0; --nn) {
            if (!ends[nn]) continue;
            // pops of all sites above nn (tnac4o.py:2331-2333, applied lazily)
            decode_pop_kernel<<<ceil_div(d->n, 256), 256, 0, st>>>(t, nn + 1, d->n, d->frame.as<int32_t>(), d->rec_node.as<int32_t>(),
                                                                   d->rec_below.as<int32_t>());
            TN_LAUNCHED(ctx);
            int lo = 0, hi = d->n;
            while (hi > lo) {
                const int m = hi - lo;
                TRY(d->cnt.ensure((size_t)m * sizeof(int), st, 0));
                TRY(d->scan.ensure((size_t)m * sizeof(int), st, 0));
                TRY(d->tmp.ensure((size_t)(ceil_div(m, 1024) + 1) * sizeof(int), st, 0));
                decode_count_kernel<<<ceil_div(m, 256), 256, 0, st>>>(t, nn, max_dEng, lo, hi, d->E.as<double>(), d->frame.as<int32_t>(),
                                                                      d->rec_node.as<int32_t>(), d->cnt.as<int>());
                TN_LAUNCHED(ctx);
                TRY(tn_scan_impl(ctx, st, d->cnt.as<int>(), d->scan.as<int>(), m, d->tmp.as<int>()));
                TN_CUDA(cudaMemcpyAsync(h_total, d->scan.as<int>() + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
                TN_CUDA(cudaStreamSynchronize(st));
                const int total = *h_total;
                d->launches_waves++;
                if (total == 0) break;
                TN_REQUIRE((int64_t)d->nrec + total < (int64_t)1 << 31, "more than 2^31 droplet combinations");
                const size_t nc = (size_t)d->n + total, nr = (size_t)d->nrec + total;
                TRY(d->E.ensure(nc * sizeof(double), st, (size_t)d->n * sizeof(double)));
                TRY(d->rec.ensure(nc * sizeof(int32_t), st, (size_t)d->n * sizeof(int32_t)));
                TRY(d->frame.ensure(nc * sizeof(int32_t), st, (size_t)d->n * sizeof(int32_t)));
                TRY(d->rec_node.ensure(nr * sizeof(int32_t), st, (size_t)d->nrec * sizeof(int32_t)));
                TRY(d->rec_from.ensure(nr * sizeof(int32_t), st, (size_t)d->nrec * sizeof(int32_t)));
                TRY(d->rec_below.ensure(nr * sizeof(int32_t), st, (size_t)d->nrec * sizeof(int32_t)));
                decode_emit_kernel<<<ceil_div(m, 256), 256, 0, st>>>(t, nn, max_dEng, lo, hi, d->n, d->nrec, d->scan.as<int>(),
                                                                     d->E.as<double>(), d->rec.as<int32_t>(), d->frame.as<int32_t>(),
                                                                     d->rec_node.as<int32_t>(), d->rec_from.as<int32_t>(),
                                                                     d->rec_below.as<int32_t>());
                TN_LAUNCHED(ctx);
                lo = d->n; hi = d->n + total;
                d->n += total; d->nrec += total;
            }
            if ((int64_t)d->n > max_states) TRY(decode_sort_by_energy(ctx, d, (int)max_states, true));
        }
        // final order: ascending energy (tnac4o.py:1374-1375)
        const int keep = (int)std::min<int64_t>(d->n, max_states);
        TRY(decode_sort_by_energy(ctx, d, keep, false));
        return TN_OK;
    };
    rc = body();
    if (rc) { cudaStreamSynchronize(st); delete d; return rc; }
    *out = d;
    *h_count = d->n;
    return TN_OK;
}

/* energies (count doubles, excitation energies relative to the ground state) and states (count x nsites bytes) of the
 * first `count` combinations; the droplet dictionary is CSR (drop_ptr / drop_pos / drop_xor, indexed by node key) */
int tn_decode_fetch(tn_ctx* ctx, tn_decode* d, int64_t count, const uint8_t* ground, const int32_t* drop_ptr,
                    const int16_t* drop_pos, const uint8_t* drop_xor, double* E_out, uint8_t* states_out) {
    TN_REQUIRE(ctx && d && count >= 0 && count <= d->n, "bad arguments");
    if (count == 0) return TN_OK;
    cudaStream_t st = d->st;
    TN_CUDA(cudaMemcpyAsync(E_out, d->E.p, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st));
    const int blocks = (int)((count * 32 + 127) / 128);
    decode_states_kernel<<<blocks, 128, 0, st>>>((int)count, d->nsites, ground, d->rec.as<int32_t>(), d->rec_node.as<int32_t>(),
                                                 d->rec_from.as<int32_t>(), d->tree_key.as<int32_t>(), drop_ptr, drop_pos, drop_xor,
                                                 states_out);
    TN_LAUNCHED(ctx);
    return TN_OK;
}

int tn_decode_free(tn_decode* d) {
    if (d) {
        cudaStreamSynchronize(d->st);
        delete d;
    }
    return TN_OK;
}

}  // extern "C"
