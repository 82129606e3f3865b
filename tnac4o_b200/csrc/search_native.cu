// Native (C++) driver of the whole branch-and-bound of search_ground_state (tnac4o.py:417-551): the same kernel
// sequence as tnac4o_b200/solver.py::_setup_RR/_site_marginals/_site_step, issued without the interpreter in the loop.
// Two host read-backs per site (number of survivors, number of groups) size the next launches.
#include <vector>

#include "common.cuh"

int tn_gemm_impl(tn_ctx* ctx, cudaStream_t st, int tA, int tB, int M, int N, int K, double alpha, const double* A,
                 int lda, int64_t sA, const double* B, int ldb, int64_t sB, double beta, double* C, int ldc, int64_t sC,
                 int batch);
extern "C" int tn_sort_capacity_for(int n);

namespace {

thread_local tn_ctx* g_ctx = nullptr;

struct DBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t bytes, cudaStream_t s) {
        if (p) cudaFreeAsync(p, st);
        st = s;
        cudaError_t e = tn_malloc_async(g_ctx, &p, bytes > 0 ? bytes : 1, s);
        if (e != cudaSuccess) { p = nullptr; return tn_cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__); }
        return TN_OK;
    }
    template <class T> T* as() const { return (T*)p; }
};

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

__global__ void fill_f64_kernel(double* x, int64_t n, double v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void iota_i32_kernel(int32_t* x, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = i;
}
__global__ void min_flag_kernel(const double* __restrict__ flag, int n, double* acc) {
    double m = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmin(m, flag[i]);
    m = warp_min(m);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmin(m, sh[w]);
        if (m < *acc) *acc = m;
    }
}

struct Branches {
    DBuf vind, states, root, Eng, prob, deg, RL;
    int n = 1;
    int alloc(int cap, int Nx, int nsites, int Dcap, cudaStream_t st) {
        TRY(vind.alloc((size_t)cap * (Nx + 1), st));
        TRY(states.alloc((size_t)cap * nsites, st));
        TRY(root.alloc((size_t)cap * sizeof(int32_t), st));
        TRY(Eng.alloc((size_t)cap * sizeof(double), st));
        TRY(prob.alloc((size_t)cap * sizeof(double), st));
        TRY(deg.alloc((size_t)cap * sizeof(long long), st));
        TRY(RL.alloc((size_t)cap * Dcap * sizeof(double), st));
        return TN_OK;
    }
};

}  // namespace

extern "C" int tn_search_ground_state(tn_ctx* ctx, void* stream, int Nx, int Ny, const tn_site* sites,
                                      const double* const* A, const int* D, const uint8_t* key_offsets, int M,
                                      double relative_P_cutoff, double min_dEng, uint8_t* states_out, double* Eng_out,
                                      double* prob_out, long long* deg_out, int* h_count, double* h_pd_max,
                                      double* h_neg_min, long long* h_marginals) {
    TN_REQUIRE(ctx && sites && A && D && key_offsets && M >= 1 && Nx >= 1 && Ny >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    g_ctx = ctx;
    const int nsites = Nx * Ny, vs = Nx + 1;
    int nsmax = 1, Dcap = 1, ndmax = 1;
    for (int i = 0; i < nsites; ++i) {
        nsmax = sites[i].nS > nsmax ? sites[i].nS : nsmax;
        ndmax = sites[i].nd > ndmax ? sites[i].nd : ndmax;
    }
    for (int i = 0; i < (Ny + 1) * (Nx + 1); ++i) Dcap = D[i] > Dcap ? D[i] : Dcap;
    const int64_t ncand = (int64_t)M * nsmax;
    const int kcap = tn_sort_capacity_for((int)ncand);

    Branches buf[2];
    TRY(buf[0].alloc(M, Nx, nsites, Dcap, st));
    TRY(buf[1].alloc(M, Nx, nsites, Dcap, st));
    DBuf cand, flag, surv, khi, klo, ktie, parent, cell, g_rep, g_start, g_size, sel, Enew, Pnew, g_prob, g_E, g_deg, scal, T1;
    TRY(cand.alloc(ncand * 8, st)); TRY(flag.alloc((size_t)M * 8, st)); TRY(surv.alloc(ncand * 4, st));
    TRY(khi.alloc((size_t)kcap * 8, st)); TRY(klo.alloc((size_t)kcap * 8, st)); TRY(ktie.alloc((size_t)kcap * 8, st));
    TRY(parent.alloc(ncand * 4, st)); TRY(cell.alloc(ncand * 4, st)); TRY(g_rep.alloc(ncand * 4, st));
    TRY(g_start.alloc(ncand * 4, st)); TRY(g_size.alloc(ncand * 4, st)); TRY(sel.alloc(ncand * 4, st));
    TRY(Enew.alloc(ncand * 8, st)); TRY(Pnew.alloc(ncand * 8, st)); TRY(g_prob.alloc(ncand * 8, st));
    TRY(g_E.alloc(ncand * 8, st)); TRY(g_deg.alloc(ncand * 8, st));
    TRY(scal.alloc(64, st));                                   // [0] count, [8] maxbits, [16] pdbits, [24] gmin
    TRY(T1.alloc((size_t)M * ndmax * Dcap * 8, st));
    int* d_count = (int*)scal.p;
    unsigned long long* d_maxbits = (unsigned long long*)((char*)scal.p + 8);
    unsigned long long* d_pdbits = (unsigned long long*)((char*)scal.p + 16);
    double* d_gmin = (double*)((char*)scal.p + 24);
    {
        unsigned long long init_pd = ordered_bits(-INFINITY);
        double zero = 0.0;
        unsigned long long* hp = (unsigned long long*)((char*)ctx->pinned + 448);
        hp[0] = init_pd;
        memcpy(&hp[1], &zero, 8);
        TN_CUDA(cudaMemcpyAsync(d_pdbits, &hp[0], 8, cudaMemcpyHostToDevice, st));
        TN_CUDA(cudaMemcpyAsync(d_gmin, &hp[1], 8, cudaMemcpyHostToDevice, st));
    }
    // one branch: empty configuration
    int cur = 0;
    Branches* br = &buf[cur];
    TN_CUDA(cudaMemsetAsync(br->vind.p, 0, (size_t)M * vs, st));
    TN_CUDA(cudaMemsetAsync(br->states.p, 0, (size_t)M * nsites, st));
    TN_CUDA(cudaMemsetAsync(br->Eng.p, 0, (size_t)M * 8, st));
    TN_CUDA(cudaMemsetAsync(br->prob.p, 0, (size_t)M * 8, st));
    {
        long long one = 1;
        long long* hp = (long long*)((char*)ctx->pinned + 480);
        *hp = one;
        TN_CUDA(cudaMemcpyAsync(br->deg.p, hp, 8, cudaMemcpyHostToDevice, st));
    }
    br->n = 1;
    long long marginals = 0;
    std::vector<DBuf> RRat(Nx + 1);

    for (int ny = 0; ny < Ny; ++ny) {
        br = &buf[cur];
        const int nb0 = br->n;
        const double* const* Arow = A + (size_t)(ny + 1) * Nx;              // tensors of rhoT[ny + 1]
        const int* Drow = D + (size_t)(ny + 1) * (Nx + 1);
        // ---- right environments of the row for every row-start branch (tnac4o.py:1768-1784)
        TRY(RRat[Nx].alloc((size_t)nb0 * 8, st));
        fill_f64_kernel<<<ceil_div(nb0, 256), 256, 0, st>>>(RRat[Nx].as<double>(), nb0, 1.0);
        TN_LAUNCHED(ctx);
        for (int nx = Nx - 1; nx >= 1; --nx) {
            const tn_site* s = &sites[ny * Nx + nx];
            TRY(RRat[nx].alloc((size_t)nb0 * Drow[nx] * s->nl * 8, st));
            // per branch-level 2 Dr nr Dl nl flop, plus the once-per-level A x Wtr contraction
            tn_prof_scope prof(ctx, st, TN_P_RR,
                               2.0 * nb0 * Drow[nx + 1] * s->nr * Drow[nx] * s->nl +
                                   2.0 * s->nu * Drow[nx + 1] * s->nr * Drow[nx] * s->nl * s->nd,
                               8.0 * nb0 * ((double)Drow[nx + 1] * s->nr + (double)Drow[nx] * s->nl));
            TRY(tn_rr_level(ctx, st, s, nb0, Drow[nx], Drow[nx + 1], Arow[nx], RRat[nx + 1].as<double>(),
                            br->vind.as<uint8_t>() + nx + 1, vs, RRat[nx].as<double>()));
        }
        iota_i32_kernel<<<ceil_div(nb0, 256), 256, 0, st>>>(br->root.as<int32_t>(), nb0);
        TN_LAUNCHED(ctx);
        fill_f64_kernel<<<ceil_div(nb0, 256), 256, 0, st>>>(br->RL.as<double>(), nb0, 1.0);
        TN_LAUNCHED(ctx);
        for (int nx = 0; nx < Nx; ++nx) {
            br = &buf[cur];
            Branches* nxt = &buf[cur ^ 1];
            const tn_site* s = &sites[ny * Nx + nx];
            const int Dl = Drow[nx], Dr = Drow[nx + 1], B = br->n;
            // ---- marginals: T1 = RL . A on the DMMA path, then the fused kernel (tnac4o.py:1786-1807, 450-453)
            // structure-aware work per branch marginal: 2 Dl nd Dr + 2 nd Dr nr + 4 nS flop and
            // 8 (Dl + Dr nr + 2 nS) + 40 bytes
            {
            tn_prof_scope pm(ctx, st, TN_P_MARGINALS,
                             (double)B * (2.0 * Dl * s->nd * Dr + 2.0 * s->nd * Dr * s->nr + 4.0 * s->nS),
                             (double)B * (8.0 * (Dl + (double)Dr * s->nr + 2.0 * s->nS) + 40.0));
            TRY(tn_gemm_impl(ctx, st, 0, 0, B, s->nd * Dr, Dl, 1.0, br->RL.as<double>(), Dl, 0, Arow[nx], s->nd * Dr, 0, 0.0,
                             T1.as<double>(), s->nd * Dr, 0, 1));
            TRY(tn_marginals(ctx, st, s, B, Dr, T1.as<double>(), RRat[nx + 1].as<double>(), br->root.as<int32_t>(),
                             br->vind.as<uint8_t>(), vs, nx, br->prob.as<double>(), cand.as<double>(), flag.as<double>(),
                             d_maxbits, nullptr));
            min_flag_kernel<<<1, 256, 0, st>>>(flag.as<double>(), B, d_gmin);
            TN_LAUNCHED(ctx);
            }
            marginals += B;
            // selection / merge / top-M / materialise: 8 B per candidate read + ~48 B per survivor record
            tn_prof_scope psel(ctx, st, TN_P_SELECT_MERGE, 0.0, 8.0 * B * s->nS);
            // ---- select -> expand -> merge -> top-M -> materialise (tnac4o.py:456-535)
            int K = 0, G = 0;
            TRY(tn_select(ctx, st, cand.as<double>(), (int64_t)B * s->nS, d_maxbits, relative_P_cutoff, surv.as<int32_t>(),
                          d_count, d_pdbits, &K));
            if (K < 1) { tn_set_error("no candidate survived the cut-off at site (%d, %d)", ny, nx); return TN_ERR_ARG; }
            TRY(tn_expand(ctx, st, s, K, nx, nx > 0, ny > 0, vs, key_offsets + (size_t)(ny * Nx + nx) * vs, surv.as<int32_t>(),
                          br->vind.as<uint8_t>(), vs, br->Eng.as<double>(), cand.as<double>(), khi.as<unsigned long long>(),
                          klo.as<unsigned long long>(), ktie.as<unsigned long long>(), parent.as<int32_t>(), cell.as<int32_t>(),
                          Enew.as<double>(), Pnew.as<double>()));
            TRY(tn_merge(ctx, st, K, khi.as<unsigned long long>(), klo.as<unsigned long long>(), ktie.as<unsigned long long>(),
                         Enew.as<double>(), Pnew.as<double>(), parent.as<int32_t>(), br->deg.as<long long>(), min_dEng,
                         g_rep.as<int32_t>(), g_deg.as<long long>(), g_prob.as<double>(), g_E.as<double>(),
                         g_start.as<int32_t>(), g_size.as<int32_t>(), &G));
            TRY(tn_topm(ctx, st, G, M, g_prob.as<double>(), khi.as<unsigned long long>(), klo.as<unsigned long long>(),
                        ktie.as<unsigned long long>(), sel.as<int32_t>(), d_pdbits));
            const int Bn = G < M ? G : M;
            TRY(tn_materialise(ctx, st, s, Bn, nx, ny * Nx + nx, nsites, vs, Dl, Dr, sel.as<int32_t>(), g_rep.as<int32_t>(),
                               g_deg.as<long long>(), g_prob.as<double>(), parent.as<int32_t>(), cell.as<int32_t>(),
                               Enew.as<double>(), br->vind.as<uint8_t>(), br->states.as<uint8_t>(), br->root.as<int32_t>(),
                               br->RL.as<double>(), Arow[nx], nxt->vind.as<uint8_t>(), nxt->states.as<uint8_t>(),
                               nxt->root.as<int32_t>(), nxt->Eng.as<double>(), nxt->prob.as<double>(), nxt->deg.as<long long>(),
                               nxt->RL.as<double>()));
            nxt->n = Bn;
            cur ^= 1;
        }
        br = &buf[cur];
        TRY(tn_row_shift(ctx, st, br->n, vs, br->vind.as<uint8_t>()));
    }
    br = &buf[cur];
    const int n = br->n;
    TN_CUDA(cudaMemcpyAsync(states_out, br->states.p, (size_t)n * nsites, cudaMemcpyDeviceToDevice, st));
    TN_CUDA(cudaMemcpyAsync(Eng_out, br->Eng.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    TN_CUDA(cudaMemcpyAsync(prob_out, br->prob.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    TN_CUDA(cudaMemcpyAsync(deg_out, br->deg.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    unsigned long long* hp = (unsigned long long*)((char*)ctx->pinned + 512);
    TN_CUDA(cudaMemcpyAsync(hp, d_pdbits, 16, cudaMemcpyDeviceToHost, st));
    TN_CUDA(cudaStreamSynchronize(st));
    *h_count = n;
    *h_pd_max = from_ordered_bits(hp[0]);
    memcpy(h_neg_min, &hp[1], 8);
    *h_marginals = marginals;
    return TN_OK;
}
