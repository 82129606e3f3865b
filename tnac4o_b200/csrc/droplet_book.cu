// Device-side droplet bookkeeping of search_low_energy_spectrum, excitations_encoding = 1 -- replaces the per-site host
// loops of _search_low_energy_spectrum_v1 (tnac4o.py:844-893) with integer kernels over flat arrays kept in HBM for
// the whole search; nothing but two counters per site travels to the host.
//
// Reference data structures and their device form
//   * d / invd (tnac4o.py:2051-2069): dictionary of droplet shapes (touched cells + XOR patterns).  Here: an open-addressing
//     hash table keyed by a 64-bit hash of the shape (a second, independent 64-bit hash is stored and compared, and a
//     mismatch raises an error instead of silently merging two shapes); shapes live in a CSR pool, ids are assigned in
//     pair order (deterministic).  The reference garbage-collects unused shapes after every site (_exc_clear_d,
//     2249-2268), which only renumbers keys; the host export keeps the shapes the final tree references.
//   * el (844-875): per live branch a list of excitations ((dE, key, first, last, dlogP), sub-excitations).  Here: an
//     append-only node pool (dE, dP, key, first, last, child range), a children pool of (node, energy budget) pairs, and
//     per branch a (pointer, length) into a list pool that is rebuilt every site (old list of the winner + one new node
//     per merged branch).
//   * _exc_cut_energy (2071-2079), the recursive pruning of the loser's excitations, is LAZY here: a child entry carries
//     the energy budget it was pruned with, and nested prunings compose as min(stored budget, outer budget - dE) with the
//     same floating-point subtractions the recursion performs; the host export materialises the nested tuples the
//     reference's save format needs once, at the end of the search.
#include <algorithm>

#include "common.cuh"

int tn_scan_impl(tn_ctx* ctx, cudaStream_t st, const int* in, int* out, int n, int* tmp);

namespace {

constexpr unsigned long long EMPTY = 0ull;
constexpr int TABLE_BITS = 21;                       // 2 M slots
constexpr unsigned int TABLE_MASK = (1u << TABLE_BITS) - 1;

struct GBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~GBuf() { if (p) cudaFree(p); }
    int ensure(size_t want, cudaStream_t st, size_t keep) {
        if (want <= bytes) return TN_OK;
        size_t cap = std::max(want, bytes * 2);
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, cap);
        if (e != cudaSuccess) return tn_cuda_fail(e, "cudaMalloc(book)", __FILE__, __LINE__);
        if (p && keep) cudaMemcpyAsync(q, p, keep, cudaMemcpyDeviceToDevice, st);
        if (p) { cudaStreamSynchronize(st); cudaFree(p); }
        p = q; bytes = cap;
        return TN_OK;
    }
    template <typename T> T* as() const { return (T*)p; }
};

// ---- pairs (winner, merged member) of the kept groups ------------------------------------------------------------
__global__ void book_pair_count_kernel(int Bn, const int* __restrict__ sel, const int* __restrict__ g_rep,
                                       const int* __restrict__ g_start, const int* __restrict__ g_size,
                                       const double* __restrict__ g_E, const int* __restrict__ order,
                                       const double* __restrict__ Enew, double max_dE, int* __restrict__ cnt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Bn) return;
    const int g = sel[j];
    int c = 0;
    if (g_size[g] > 1) {
        const int rep = g_rep[g];
        const double e0 = g_E[g];
        for (int q = g_start[g]; q < g_start[g] + g_size[g]; ++q) {
            const int m = order[q];
            c += (m != rep) && (Enew[m] - e0 <= max_dE);
        }
    }
    cnt[j] = c;
}

__global__ void book_pair_fill_kernel(int Bn, const int* __restrict__ sel, const int* __restrict__ g_rep,
                                      const int* __restrict__ g_start, const int* __restrict__ g_size,
                                      const double* __restrict__ g_E, const int* __restrict__ order,
                                      const double* __restrict__ Enew, double max_dE, const int* __restrict__ incl,
                                      int* __restrict__ pw, int* __restrict__ pm, int* __restrict__ pg) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Bn) return;
    const int g = sel[j];
    if (g_size[g] <= 1) return;
    int o = (j == 0) ? 0 : incl[j - 1];
    const int rep = g_rep[g];
    const double e0 = g_E[g];
    for (int q = g_start[g]; q < g_start[g] + g_size[g]; ++q) {
        const int m = order[q];
        if ((m != rep) && (Enew[m] - e0 <= max_dE)) { pw[o] = rep; pm[o] = m; pg[o] = j; ++o; }
    }
}

// ---- XOR difference of winner and member rows (tnac4o.py:859-861), hashes, Hamming filter; one warp per pair -----
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return h;
}

__global__ void book_diff_kernel(int npairs, int nsites, int pos, const uint8_t* __restrict__ states,
                                 const int* __restrict__ parent, const int* __restrict__ cell, const int* __restrict__ pw,
                                 const int* __restrict__ pm, int lim_hd, int16_t* __restrict__ out_pos,
                                 uint8_t* __restrict__ out_xor, int* __restrict__ out_len, unsigned long long* __restrict__ h1,
                                 unsigned long long* __restrict__ h2, int* __restrict__ accept) {
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= npairs) return;
    const int a = pw[k], b = pm[k];
    const uint8_t* ra = states + (int64_t)parent[a] * nsites;
    const uint8_t* rb = states + (int64_t)parent[b] * nsites;
    int n = 0;
    for (int k0 = 0; k0 < nsites; k0 += 32) {
        const int i = k0 + lane;
        uint8_t x = 0;
        if (i < nsites) {
            const uint8_t va = (i == pos) ? (uint8_t)cell[a] : ra[i];
            const uint8_t vb = (i == pos) ? (uint8_t)cell[b] : rb[i];
            x = va ^ vb;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, x != 0);
        if (x) {
            const int o = n + __popc(mask & ((1u << lane) - 1));
            out_pos[(int64_t)k * nsites + o] = (int16_t)i;
            out_xor[(int64_t)k * nsites + o] = x;
        }
        n += __popc(mask);
    }
    __syncwarp();
    if (lane == 0) {
        unsigned long long a1 = 0x243f6a8885a308d3ull, a2 = 0x13198a2e03707344ull;
        int bits = 0;
        for (int e = 0; e < n; ++e) {
            bits += __popc((unsigned)out_xor[(int64_t)k * nsites + e]);
            const unsigned long long v = ((unsigned long long)(unsigned short)out_pos[(int64_t)k * nsites + e] << 8) |
                                         out_xor[(int64_t)k * nsites + e];
            a1 = mix64(a1, v);
            a2 = mix64(a2 ^ 0xa4093822299f31d0ull, v * 0x9fb21c651e98df25ull + 1);
        }
        a1 = mix64(a1, (unsigned long long)n);
        if (a1 == EMPTY) a1 = 1;
        out_len[k] = n;
        h1[k] = a1;
        h2[k] = a2;
        // _exc_hd (tnac4o.py:2143-2150): touched cells in Ising mode; set bits of the patterns in RMF mode (lim_hd < 0 here)
        accept[k] = (lim_hd >= 0) ? ((lim_hd <= 1) || (n >= lim_hd)) : ((-lim_hd <= 1) || (bits >= -lim_hd));
    }
}

// ---- dictionary: claim / find the slot of every accepted pair; remember the first pair that touched a fresh slot ---
__global__ void book_dict_probe_kernel(int npairs, const int* __restrict__ accept, const unsigned long long* __restrict__ h1,
                                       unsigned long long* __restrict__ tkey, int* __restrict__ tfirst,
                                       int* __restrict__ slot_of) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs || !accept[k]) return;
    const unsigned long long key = h1[k];
    unsigned int s = (unsigned int)(key >> 17) & TABLE_MASK;
    for (int probe = 0; probe < (1 << TABLE_BITS); ++probe) {
        const unsigned long long old = atomicCAS(&tkey[s], EMPTY, key);
        if (old == EMPTY || old == key) {
            atomicMin(&tfirst[s], k);                        // fresh slots start at INT_MAX, known shapes at -1
            slot_of[k] = (int)s;
            return;
        }
        s = (s + 1) & TABLE_MASK;
    }
    slot_of[k] = -1;                                         // table full
}

// isnew[k] = 1 for the pair that introduces a shape (the lowest pair index on a fresh slot)
__global__ void book_dict_new_kernel(int npairs, const int* __restrict__ accept, const int* __restrict__ slot_of,
                                     const int* __restrict__ tfirst, const int* __restrict__ out_len,
                                     int* __restrict__ isnew, int* __restrict__ newlen, int* __restrict__ err) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    int f = 0;
    if (accept[k]) {
        if (slot_of[k] < 0) atomicExch(err, 1);
        else f = (tfirst[slot_of[k]] == k);
    }
    isnew[k] = f;
    newlen[k] = f ? out_len[k] : 0;
}

// new shapes get ids nshapes + rank and their content appended to the pool; the slot publishes the id
__global__ void book_dict_publish_kernel(int npairs, int nsites, int nshapes, int nelems, const int* __restrict__ isnew,
                                         const int* __restrict__ incl_new, const int* __restrict__ incl_len,
                                         const int* __restrict__ slot_of, const int* __restrict__ out_len,
                                         const int16_t* __restrict__ out_pos, const uint8_t* __restrict__ out_xor,
                                         const unsigned long long* __restrict__ h2, int* __restrict__ tid_of_slot,
                                         unsigned long long* __restrict__ tkey2, int* __restrict__ sptr,
                                         int16_t* __restrict__ spos, uint8_t* __restrict__ sxor) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs || !isnew[k]) return;
    const int id = nshapes + incl_new[k] - 1;
    const int off = nelems + incl_len[k] - out_len[k];
    for (int e = 0; e < out_len[k]; ++e) {
        spos[off + e] = out_pos[(int64_t)k * nsites + e];
        sxor[off + e] = out_xor[(int64_t)k * nsites + e];
    }
    sptr[id + 1] = off + out_len[k];
    tid_of_slot[slot_of[k]] = id;
    tkey2[slot_of[k]] = h2[k];
}

// key of every accepted pair; the second hash must agree with the slot's (else two shapes collided on the first hash)
__global__ void book_dict_key_kernel(int npairs, const int* __restrict__ accept, const int* __restrict__ slot_of,
                                     const unsigned long long* __restrict__ h2, const int* __restrict__ tid_of_slot,
                                     const unsigned long long* __restrict__ tkey2, int* __restrict__ tfirst,
                                     int* __restrict__ key, int* __restrict__ err) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs || !accept[k]) return;
    const int s = slot_of[k];
    if (s < 0) return;
    if (tkey2[s] != h2[k]) atomicExch(err, 2);
    key[k] = tid_of_slot[s];
    tfirst[s] = -1;                                           // the slot is a known shape from now on
}

// ---- sub-excitations of the new nodes: entries of the loser's list that end at or after the droplet's first cell and
//      fit the energy bound (tnac4o.py:866-868) -------------------------------------------------------------------
__global__ void book_child_count_kernel(int npairs, int nsites, const int* __restrict__ accept, const int* __restrict__ pm,
                                        const int* __restrict__ pg, const int* __restrict__ sel, const int* __restrict__ parent,
                                        const double* __restrict__ Enew, const double* __restrict__ g_E,
                                        const int16_t* __restrict__ out_pos, const int* __restrict__ el_ptr,
                                        const int* __restrict__ el_cnt, const int* __restrict__ lnode,
                                        const double* __restrict__ node_dE, const int* __restrict__ node_last, double max_dE,
                                        int* __restrict__ ccnt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    int c = 0;
    if (accept[k]) {
        const int m = pm[k];
        const double gap = Enew[m] - g_E[sel[pg[k]]];
        const int dfirst = out_pos[(int64_t)k * nsites];
        const int ob = parent[m];
        for (int e = el_ptr[ob]; e < el_ptr[ob] + el_cnt[ob]; ++e) {
            const int nd = lnode[e];
            c += (node_last[nd] >= dfirst) && (node_dE[nd] + gap <= max_dE);
        }
    }
    ccnt[k] = c;
}

// new nodes (one per accepted pair, id = nnodes + rank) with their children; budgets as _exc_cut_energy receives them
__global__ void book_node_fill_kernel(int npairs, int nsites, int site, int nnodes, int nchildren, const int* __restrict__ accept,
                                      const int* __restrict__ incl_acc, const int* __restrict__ incl_child,
                                      const int* __restrict__ ccnt, const int* __restrict__ pm, const int* __restrict__ pg,
                                      const int* __restrict__ sel, const int* __restrict__ parent, const double* __restrict__ Enew,
                                      const double* __restrict__ Pnew, const double* __restrict__ g_E,
                                      const double* __restrict__ g_prob, const int16_t* __restrict__ out_pos,
                                      const int* __restrict__ key, const int* __restrict__ el_ptr, const int* __restrict__ el_cnt,
                                      const int* __restrict__ lnode, double max_dE, double* __restrict__ node_dE,
                                      double* __restrict__ node_dP, int* __restrict__ node_key, int* __restrict__ node_first,
                                      int* __restrict__ node_last, int* __restrict__ node_cptr, int* __restrict__ node_ccnt,
                                      int* __restrict__ cnode, double* __restrict__ cbud) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs || !accept[k]) return;
    const int id = nnodes + incl_acc[k] - 1;
    const int m = pm[k], g = sel[pg[k]];
    const double gap = Enew[m] - g_E[g];
    const int dfirst = out_pos[(int64_t)k * nsites];
    int o = nchildren + incl_child[k] - ccnt[k];
    node_dE[id] = gap;
    node_dP[id] = Pnew[m] - g_prob[g];
    node_key[id] = key[k];
    node_first[id] = dfirst;
    node_last[id] = site;
    node_cptr[id] = o;
    node_ccnt[id] = ccnt[k];
    const int ob = parent[m];
    for (int e = el_ptr[ob]; e < el_ptr[ob] + el_cnt[ob]; ++e) {
        const int nd = lnode[e];
        const double dE = node_dE[nd];              // (older node: id < nnodes, never written by this launch)
        if ((node_last[nd] >= dfirst) && (dE + gap <= max_dE)) {
            cnode[o] = nd;
            cbud[o] = max_dE - (dE + gap);
            ++o;
        }
    }
}

// ---- excitation lists of the new branches: winner's old list + its new nodes (tnac4o.py:856, 871-875) --------------
__global__ void book_list_len_kernel(int Bn, const int* __restrict__ sel, const int* __restrict__ g_rep,
                                     const int* __restrict__ parent, const int* __restrict__ el_cnt,
                                     const int* __restrict__ pair_incl, const int* __restrict__ accept, int* __restrict__ len) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Bn) return;
    const int ob = parent[g_rep[sel[j]]];
    int c = el_cnt[ob];
    if (pair_incl) {
        const int lo = (j == 0) ? 0 : pair_incl[j - 1], hi = pair_incl[j];
        for (int k = lo; k < hi; ++k) c += accept[k];
    }
    len[j] = c;
}

__global__ void book_list_fill_kernel(int Bn, int nnodes, const int* __restrict__ sel, const int* __restrict__ g_rep,
                                      const int* __restrict__ parent, const int* __restrict__ el_ptr,
                                      const int* __restrict__ el_cnt, const int* __restrict__ lnode,
                                      const int* __restrict__ pair_incl, const int* __restrict__ accept,
                                      const int* __restrict__ incl_acc, const int* __restrict__ len_incl,
                                      const int* __restrict__ len, int* __restrict__ new_ptr, int* __restrict__ new_cnt,
                                      int* __restrict__ new_lnode) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Bn) return;
    const int ob = parent[g_rep[sel[j]]];
    int o = len_incl[j] - len[j];
    new_ptr[j] = o;
    new_cnt[j] = len[j];
    for (int e = el_ptr[ob]; e < el_ptr[ob] + el_cnt[ob]; ++e) new_lnode[o++] = lnode[e];
    if (pair_incl) {
        const int lo = (j == 0) ? 0 : pair_incl[j - 1], hi = pair_incl[j];
        for (int k = lo; k < hi; ++k)
            if (accept[k]) new_lnode[o++] = nnodes + incl_acc[k] - 1;
    }
}

__global__ void book_fill_int_kernel(int* x, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

}  // namespace

struct tn_book {
    cudaStream_t st = nullptr;
    int nsites = 0, M = 0;
    int nnodes = 0, nchildren = 0, nshapes = 0, nelems = 0;
    int64_t pairs_total = 0;
    // pools
    GBuf node_dE, node_dP, node_key, node_first, node_last, node_cptr, node_ccnt, cnode, cbud;
    GBuf sptr, spos, sxor, tkey, tkey2, tfirst, tid_of_slot;
    GBuf el_ptr[2], el_cnt[2], lnode[2];
    int cur = 0;
    // per-site scratch
    GBuf cnt, incl, tmp, pw, pm, pg, out_pos, out_xor, out_len, h1, h2, accept, slot_of, isnew, newlen, incl_new, incl_len,
        key, ccnt, incl_child, incl_acc, len, len_incl, err;
};

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

static int scan(tn_ctx* ctx, tn_book* b, const int* in, int* out, int n) {
    TRY(b->tmp.ensure((size_t)(n / 1024 + 2) * sizeof(int), b->st, 0));
    return tn_scan_impl(ctx, b->st, in, out, n, b->tmp.as<int>());
}

extern "C" {

int tn_book_create(tn_ctx* ctx, void* stream, int nsites, int M, tn_book** out) {
    TN_REQUIRE(ctx && out && nsites >= 1 && M >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    tn_book* b = new tn_book();
    b->st = st; b->nsites = nsites; b->M = M;
    auto body = [&]() -> int {
        const size_t T = (size_t)1 << TABLE_BITS;
        TRY(b->tkey.ensure(T * 8, st, 0)); TRY(b->tkey2.ensure(T * 8, st, 0));
        TRY(b->tfirst.ensure(T * 4, st, 0)); TRY(b->tid_of_slot.ensure(T * 4, st, 0));
        TN_CUDA(cudaMemsetAsync(b->tkey.p, 0, T * 8, st));
        book_fill_int_kernel<<<ceil_div((int64_t)T, 256), 256, 0, st>>>(b->tfirst.as<int>(), (int)T, 0x7fffffff);
        TN_LAUNCHED(ctx);
        for (int i = 0; i < 2; ++i) {
            TRY(b->el_ptr[i].ensure((size_t)M * 4, st, 0));
            TRY(b->el_cnt[i].ensure((size_t)M * 4, st, 0));
            TRY(b->lnode[i].ensure(4096, st, 0));
            TN_CUDA(cudaMemsetAsync(b->el_ptr[i].p, 0, (size_t)M * 4, st));
            TN_CUDA(cudaMemsetAsync(b->el_cnt[i].p, 0, (size_t)M * 4, st));
        }
        TRY(b->sptr.ensure(4096, st, 0));
        TN_CUDA(cudaMemsetAsync(b->sptr.p, 0, 4, st));
        TRY(b->err.ensure(4, st, 0));
        TN_CUDA(cudaMemsetAsync(b->err.p, 0, 4, st));
        return TN_OK;
    };
    int rc = body();
    if (rc) { cudaStreamSynchronize(st); delete b; return rc; }
    *out = b;
    return TN_OK;
}

/* One site of the spectrum search, after tn_merge / tn_topm / tn_materialise: `order` is the sorted member list of the
 * merge (low words of the merge keys, copied before tn_topm reuses them), the g_* arrays describe the groups, sel the Bn
 * kept groups, Enew / Pnew / parent / cell the K survivors, old_states the state rows of the OLD branches. */
int tn_book_site(tn_ctx* ctx, tn_book* b, int site, int K, int Bn, const int32_t* order, const int32_t* g_rep,
                 const int32_t* g_start, const int32_t* g_size, const double* g_E, const double* g_prob, const int32_t* sel,
                 const double* Enew, const double* Pnew, const int32_t* parent, const int32_t* cell, const uint8_t* old_states,
                 double max_dEng, int lim_hd) {
    TN_REQUIRE(ctx && b && K >= 1 && Bn >= 1 && Bn <= b->M, "bad arguments");
    cudaStream_t st = b->st;
    const int ns = b->nsites;
    const int cur = b->cur, nxt = cur ^ 1;
    int* hp = (int*)((char*)ctx->pinned + 640);
    // ---- pairs
    TRY(b->cnt.ensure((size_t)Bn * 4, st, 0)); TRY(b->incl.ensure((size_t)Bn * 4, st, 0));
    book_pair_count_kernel<<<ceil_div(Bn, 128), 128, 0, st>>>(Bn, sel, g_rep, g_start, g_size, g_E, order, Enew, max_dEng,
                                                             b->cnt.as<int>());
    TN_LAUNCHED(ctx);
    TRY(scan(ctx, b, b->cnt.as<int>(), b->incl.as<int>(), Bn));
    TN_CUDA(cudaMemcpyAsync(hp, b->incl.as<int>() + (Bn - 1), 4, cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    const int np = hp[0];
    b->pairs_total += np;
    int nacc = 0, nnew = 0, nnewel = 0, nch = 0;
    if (np > 0) {
        const size_t n4 = (size_t)np * 4;
        TRY(b->pw.ensure(n4, st, 0)); TRY(b->pm.ensure(n4, st, 0)); TRY(b->pg.ensure(n4, st, 0));
        TRY(b->out_pos.ensure((size_t)np * ns * 2, st, 0)); TRY(b->out_xor.ensure((size_t)np * ns, st, 0));
        TRY(b->out_len.ensure(n4, st, 0)); TRY(b->h1.ensure((size_t)np * 8, st, 0)); TRY(b->h2.ensure((size_t)np * 8, st, 0));
        TRY(b->accept.ensure(n4, st, 0)); TRY(b->slot_of.ensure(n4, st, 0)); TRY(b->isnew.ensure(n4, st, 0));
        TRY(b->newlen.ensure(n4, st, 0)); TRY(b->incl_new.ensure(n4, st, 0)); TRY(b->incl_len.ensure(n4, st, 0));
        TRY(b->key.ensure(n4, st, 0)); TRY(b->ccnt.ensure(n4, st, 0)); TRY(b->incl_child.ensure(n4, st, 0));
        TRY(b->incl_acc.ensure(n4, st, 0));
        book_pair_fill_kernel<<<ceil_div(Bn, 128), 128, 0, st>>>(Bn, sel, g_rep, g_start, g_size, g_E, order, Enew, max_dEng,
                                                                b->incl.as<int>(), b->pw.as<int>(), b->pm.as<int>(), b->pg.as<int>());
        TN_LAUNCHED(ctx);
        book_diff_kernel<<<ceil_div((int64_t)np * 32, 128), 128, 0, st>>>(np, ns, site, old_states, parent, cell, b->pw.as<int>(),
                                                                         b->pm.as<int>(), lim_hd, b->out_pos.as<int16_t>(),
                                                                         b->out_xor.as<uint8_t>(), b->out_len.as<int>(),
                                                                         b->h1.as<unsigned long long>(), b->h2.as<unsigned long long>(),
                                                                         b->accept.as<int>());
        TN_LAUNCHED(ctx);
        const int gp = ceil_div(np, 128);
        book_dict_probe_kernel<<<gp, 128, 0, st>>>(np, b->accept.as<int>(), b->h1.as<unsigned long long>(),
                                                   b->tkey.as<unsigned long long>(), b->tfirst.as<int>(), b->slot_of.as<int>());
        TN_LAUNCHED(ctx);
        book_dict_new_kernel<<<gp, 128, 0, st>>>(np, b->accept.as<int>(), b->slot_of.as<int>(), b->tfirst.as<int>(),
                                                 b->out_len.as<int>(), b->isnew.as<int>(), b->newlen.as<int>(), b->err.as<int>());
        TN_LAUNCHED(ctx);
        TRY(scan(ctx, b, b->isnew.as<int>(), b->incl_new.as<int>(), np));
        TRY(scan(ctx, b, b->newlen.as<int>(), b->incl_len.as<int>(), np));
        TRY(scan(ctx, b, b->accept.as<int>(), b->incl_acc.as<int>(), np));
        book_child_count_kernel<<<gp, 128, 0, st>>>(np, ns, b->accept.as<int>(), b->pm.as<int>(), b->pg.as<int>(), sel, parent, Enew,
                                                    g_E, b->out_pos.as<int16_t>(), b->el_ptr[cur].as<int>(), b->el_cnt[cur].as<int>(),
                                                    b->lnode[cur].as<int>(), b->node_dE.as<double>(), b->node_last.as<int>(), max_dEng,
                                                    b->ccnt.as<int>());
        TN_LAUNCHED(ctx);
        TRY(scan(ctx, b, b->ccnt.as<int>(), b->incl_child.as<int>(), np));
        TN_CUDA(cudaMemcpyAsync(hp + 1, b->incl_new.as<int>() + (np - 1), 4, cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaMemcpyAsync(hp + 2, b->incl_len.as<int>() + (np - 1), 4, cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaMemcpyAsync(hp + 3, b->incl_acc.as<int>() + (np - 1), 4, cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaMemcpyAsync(hp + 4, b->incl_child.as<int>() + (np - 1), 4, cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaMemcpyAsync(hp + 5, b->err.p, 4, cudaMemcpyDeviceToHost, st));
        TN_CUDA(cudaStreamSynchronize(st));
        nnew = hp[1]; nnewel = hp[2]; nacc = hp[3]; nch = hp[4];
        if (hp[5]) { tn_set_error("droplet dictionary: %s", hp[5] == 1 ? "hash table full" : "hash collision between two shapes"); return TN_ERR_ARG; }
        // ---- grow the pools, publish the new shapes, fill the new nodes
        TRY(b->sptr.ensure((size_t)(b->nshapes + nnew + 2) * 4, st, (size_t)(b->nshapes + 1) * 4));
        TRY(b->spos.ensure((size_t)(b->nelems + nnewel + 1) * 2, st, (size_t)b->nelems * 2));
        TRY(b->sxor.ensure((size_t)(b->nelems + nnewel + 1), st, (size_t)b->nelems));
        const size_t nn = (size_t)b->nnodes + nacc + 1, no = (size_t)b->nnodes;
        TRY(b->node_dE.ensure(nn * 8, st, no * 8)); TRY(b->node_dP.ensure(nn * 8, st, no * 8));
        TRY(b->node_key.ensure(nn * 4, st, no * 4)); TRY(b->node_first.ensure(nn * 4, st, no * 4));
        TRY(b->node_last.ensure(nn * 4, st, no * 4)); TRY(b->node_cptr.ensure(nn * 4, st, no * 4));
        TRY(b->node_ccnt.ensure(nn * 4, st, no * 4));
        TRY(b->cnode.ensure((size_t)(b->nchildren + nch + 1) * 4, st, (size_t)b->nchildren * 4));
        TRY(b->cbud.ensure((size_t)(b->nchildren + nch + 1) * 8, st, (size_t)b->nchildren * 8));
        book_dict_publish_kernel<<<gp, 128, 0, st>>>(np, ns, b->nshapes, b->nelems, b->isnew.as<int>(), b->incl_new.as<int>(),
                                                     b->incl_len.as<int>(), b->slot_of.as<int>(), b->out_len.as<int>(),
                                                     b->out_pos.as<int16_t>(), b->out_xor.as<uint8_t>(), b->h2.as<unsigned long long>(),
                                                     b->tid_of_slot.as<int>(), b->tkey2.as<unsigned long long>(), b->sptr.as<int>(),
                                                     b->spos.as<int16_t>(), b->sxor.as<uint8_t>());
        TN_LAUNCHED(ctx);
        book_dict_key_kernel<<<gp, 128, 0, st>>>(np, b->accept.as<int>(), b->slot_of.as<int>(), b->h2.as<unsigned long long>(),
                                                 b->tid_of_slot.as<int>(), b->tkey2.as<unsigned long long>(), b->tfirst.as<int>(),
                                                 b->key.as<int>(), b->err.as<int>());
        TN_LAUNCHED(ctx);
        book_node_fill_kernel<<<gp, 128, 0, st>>>(np, ns, site, b->nnodes, b->nchildren, b->accept.as<int>(), b->incl_acc.as<int>(),
                                                  b->incl_child.as<int>(), b->ccnt.as<int>(), b->pm.as<int>(), b->pg.as<int>(), sel,
                                                  parent, Enew, Pnew, g_E, g_prob, b->out_pos.as<int16_t>(), b->key.as<int>(),
                                                  b->el_ptr[cur].as<int>(), b->el_cnt[cur].as<int>(), b->lnode[cur].as<int>(), max_dEng,
                                                  b->node_dE.as<double>(), b->node_dP.as<double>(), b->node_key.as<int>(),
                                                  b->node_first.as<int>(), b->node_last.as<int>(), b->node_cptr.as<int>(),
                                                  b->node_ccnt.as<int>(), b->cnode.as<int>(), b->cbud.as<double>());
        TN_LAUNCHED(ctx);
    }
    // ---- lists of the new branches
    TRY(b->len.ensure((size_t)Bn * 4, st, 0)); TRY(b->len_incl.ensure((size_t)Bn * 4, st, 0));
    book_list_len_kernel<<<ceil_div(Bn, 128), 128, 0, st>>>(Bn, sel, g_rep, parent, b->el_cnt[cur].as<int>(),
                                                           np > 0 ? b->incl.as<int>() : nullptr, b->accept.as<int>(), b->len.as<int>());
    TN_LAUNCHED(ctx);
    TRY(scan(ctx, b, b->len.as<int>(), b->len_incl.as<int>(), Bn));
    TN_CUDA(cudaMemcpyAsync(hp + 6, b->len_incl.as<int>() + (Bn - 1), 4, cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    const int total_len = hp[6];
    TRY(b->lnode[nxt].ensure((size_t)(total_len + 1) * 4, st, 0));
    book_list_fill_kernel<<<ceil_div(Bn, 128), 128, 0, st>>>(Bn, b->nnodes, sel, g_rep, parent, b->el_ptr[cur].as<int>(),
                                                            b->el_cnt[cur].as<int>(), b->lnode[cur].as<int>(),
                                                            np > 0 ? b->incl.as<int>() : nullptr, b->accept.as<int>(),
                                                            b->incl_acc.as<int>(), b->len_incl.as<int>(), b->len.as<int>(),
                                                            b->el_ptr[nxt].as<int>(), b->el_cnt[nxt].as<int>(), b->lnode[nxt].as<int>());
    TN_LAUNCHED(ctx);
    b->nnodes += nacc; b->nchildren += nch; b->nshapes += nnew; b->nelems += nnewel;
    b->cur = nxt;
    return TN_OK;
}

/* h_sizes[0..5] = nodes, children, shapes, shape elements, length of branch 0's list, pairs examined */
int tn_book_sizes(tn_ctx* ctx, tn_book* b, int64_t* h_sizes) {
    TN_REQUIRE(ctx && b && h_sizes, "bad arguments");
    int* hp = (int*)((char*)ctx->pinned + 640);
    TN_CUDA(cudaMemcpyAsync(hp, b->el_cnt[b->cur].p, 4, cudaMemcpyDeviceToHost, b->st));
    TN_CUDA(cudaStreamSynchronize(b->st));
    h_sizes[0] = b->nnodes; h_sizes[1] = b->nchildren; h_sizes[2] = b->nshapes; h_sizes[3] = b->nelems; h_sizes[4] = hp[0];
    h_sizes[5] = b->pairs_total;
    return TN_OK;
}

/* copies the pools into HOST arrays sized by tn_book_sizes (synchronises) */
int tn_book_export(tn_ctx* ctx, tn_book* b, double* h_dE, double* h_dP, int32_t* h_key, int32_t* h_first, int32_t* h_last,
                   int32_t* h_cptr, int32_t* h_ccnt, int32_t* h_cnode, double* h_cbud, int32_t* h_sptr, int16_t* h_spos,
                   uint8_t* h_sxor, int32_t* h_list0) {
    TN_REQUIRE(ctx && b, "bad arguments");
    cudaStream_t st = b->st;
    int* hp = (int*)((char*)ctx->pinned + 640);
    TN_CUDA(cudaMemcpyAsync(hp, b->el_cnt[b->cur].p, 4, cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaMemcpyAsync(hp + 1, b->el_ptr[b->cur].p, 4, cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    const int n0 = hp[0], p0 = hp[1];
    auto dl = [&](void* dst, const GBuf& src, size_t bytes, size_t off = 0) -> int {
        if (bytes) TN_CUDA(cudaMemcpyAsync(dst, (const char*)src.p + off, bytes, cudaMemcpyDeviceToHost, st));
        return TN_OK;
    };
    const size_t n = b->nnodes;
    TRY(dl(h_dE, b->node_dE, n * 8)); TRY(dl(h_dP, b->node_dP, n * 8)); TRY(dl(h_key, b->node_key, n * 4));
    TRY(dl(h_first, b->node_first, n * 4)); TRY(dl(h_last, b->node_last, n * 4)); TRY(dl(h_cptr, b->node_cptr, n * 4));
    TRY(dl(h_ccnt, b->node_ccnt, n * 4));
    TRY(dl(h_cnode, b->cnode, (size_t)b->nchildren * 4)); TRY(dl(h_cbud, b->cbud, (size_t)b->nchildren * 8));
    TRY(dl(h_sptr, b->sptr, (size_t)(b->nshapes + 1) * 4)); TRY(dl(h_spos, b->spos, (size_t)b->nelems * 2));
    TRY(dl(h_sxor, b->sxor, (size_t)b->nelems));
    TRY(dl(h_list0, b->lnode[b->cur], (size_t)n0 * 4, (size_t)p0 * 4));
    TN_CUDA(cudaStreamSynchronize(st));
    return TN_OK;
}

int tn_book_free(tn_book* b) {
    if (b) {
        cudaStreamSynchronize(b->st);
        delete b;
    }
    return TN_OK;
}

}  // extern "C"
