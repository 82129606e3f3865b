/*
 * tnac4o_b200 -- C ABI of the B200 (sm_100a) contraction hot path of tnac4o.
 *
 * The reference (marekrams/tnac4o) is pure Python and has no FFI of its own; the boundary it
 * exposes is the object API  tnac4o(mode, Nx, Ny, Nc, J, beta)  with search_ground_state /
 * gibbs_sampling / search_low_energy_spectrum / decode_low_energy_states.  The entry points
 * below are what a ctypes binding *underneath* the reference's private methods binds; every
 * declaration cites the reference lines it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure; tn_last_error()
 *     returns a human-readable message for the calling thread's last failure;
 *   - all array arguments are DEVICE pointers to caller-owned memory (torch tensors on the
 *     Python side) unless the name starts with h_ (host pointer); the library never frees or
 *     retains them beyond the call;
 *   - all matrices are float64, row-major, leading dimension in elements;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream unless
 *     they return a host value (documented per function);
 *   - the library's own scratch memory lives in the context and only grows.
 */
#ifndef TNAC4O_B200_H
#define TNAC4O_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tn_ctx tn_ctx;

#define TN_OK 0
#define TN_ERR_CUDA (-1)
#define TN_ERR_ARG (-2)
#define TN_ERR_NOCONV (-3)
#define TN_ERR_NOMEM (-4)

/* ---------------------------------------------------------------- context */
int tn_version(void);
const char* tn_last_error(void);
int tn_create(int device, tn_ctx** out);
int tn_destroy(tn_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t tn_launch_count(const tn_ctx* ctx);
/* Per-primitive timing of the native drivers (tn_row_compress, tn_search_ground_state): while enabled, every primitive
 * call is bracketed by CUDA events on the launching stream and tagged with its algorithmic flops / bytes.
 * tn_profile_read synchronises and returns, per category c < ncat (0 gemm, 1 qr, 2 svd, 3 other boundary-MPS kernels,
 * 4 right environments, 5 marginals, 6 select/merge/top-M), h_out[4c..4c+3] = seconds, flops, bytes, calls.  The
 * reference's counterpart is the time.time() deltas it logs per phase (tnac4o.py:407-415, 430-431). */
int tn_profile(tn_ctx* ctx, int on);
int tn_profile_read(tn_ctx* ctx, double* h_out, int ncat);
/* cudaDeviceScheduleBlockingSync for the current device: waiting host threads sleep instead of spinning (the reference is
 * single-threaded; this is for running many solver instances per GPU from fewer host cores) */
int tn_set_blocking_sync(int on);
/* throughput mode (process-wide): kernels that can trade latency for SM occupancy do so -- the cluster Jacobi SVD uses 4
 * instead of 8 CTAs when its vectors fit; for runs where many solver instances share one GPU (TN_THROUGHPUT=1 does the same) */
int tn_set_throughput_mode(int on);

/* ---------------------------------------------------------------- boundary-MPS primitives (tnac4o/mps.py) */

/* C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b], FP64 on the DMMA tensor path.
 * Replaces every np.tensordot of mps.py:655-769 (environment updates, attach_AC/CA, RAR) and
 * tnac4o.py:1779-1794.  op(X) = X or X^T (trans flag), op(A) is M x K, op(B) is K x N. */
int tn_gemm(tn_ctx* ctx, void* stream, int transA, int transB, int M, int N, int K, double alpha,
            const double* A, int lda, int64_t strideA, const double* B, int ldb, int64_t strideB, double beta,
            double* C, int ldc, int64_t strideC, int batch);

/* out[n x m] = in[m x n]^T */
int tn_transpose(tn_ctx* ctx, void* stream, int m, int n, const double* in, int ldin, double* out, int ldout);

/* Economic Householder QR with non-negative diagonal of R: A (m x n, overwritten) = Q (m x k) R (k x n),
 * k = min(m, n).  Replaces mps.qr (mps.py:43-59: scipy dgeqrf + dorgqr + sign fix).  If maxabs_bits
 * is not NULL the bit pattern of max|R| is written there (device uint64) for tn_pow2_scale. */
int tn_qr_pos(tn_ctx* ctx, void* stream, int m, int n, double* A, int lda, double* Q, int ldq, double* R, int ldr,
              unsigned long long* maxabs_bits);

/* max|x| over n elements -> *maxabs_bits (device, bit pattern of the double; zero-initialised by the call). */
int tn_maxabs(tn_ctx* ctx, void* stream, const double* x, int64_t n, unsigned long long* maxabs_bits);

/* x *= 1 / 2^floor(log2(max|x|)) where the maximum was left in *maxabs_bits; the exponent is added to
 * *log2_accum (device double) when not NULL.  Replaces mps.nfactor (mps.py:76-85) and the divisions at
 * mps.py:782, 797.  A 1x1 input is set to exactly 1 and its sign written to *sign_out (mps.py:778-780). */
int tn_pow2_scale(tn_ctx* ctx, void* stream, double* x, int64_t n, const unsigned long long* maxabs_bits,
                  double* log2_accum);

/* Thin SVD by one-sided (Hestenes) Jacobi: C (m x n) = U (m x k) diag(S) Vt (k x n), k = min(m, n), S sorted
 * descending, signs fixed as in mps.svd (mps.py:24-40).  want_vectors = 0 computes S only (mps.svd_S,
 * mps.py:62-73).  Synchronises the stream (the sweep loop reads a convergence counter); *h_sweeps receives the
 * number of sweeps.  Returns TN_ERR_NOCONV if 60 sweeps did not converge (the reference's gesdd -> gesvd
 * fall-back, mps.py:31-34, has the same role). */
int tn_svd(tn_ctx* ctx, void* stream, int m, int n, const double* C, int ldc, double* U, int ldu, double* S, double* Vt,
           int ldvt, int want_vectors, int* h_sweeps);

/* keep = min(#{S > S[0] * tol}, Dmax); discarded = ||S[keep:]|| / S[0]   (mps.py:805-809).
 * Host outputs; synchronises the stream. */
int tn_truncation_rank(tn_ctx* ctx, void* stream, const double* S, int k, double tol, int Dmax, int* h_keep,
                       double* h_discarded);

/* MPO application on one site, Hconj=True form (mps.py:753-757):
 *   out[(a,l), u, (b,r)] = sum_p A[a,p,b] * W[l,p,r,u],  A (Dl,dp,Dr), W (wl,dp,wr,du), out (Dl*wl, du, Dr*wr).
 * conj = 0 gives the Hconj=False form (mps.py:759-760): out[(l,a), o, (r,b)] = sum_p W[l,o,r,p] A[a,p,b]. */
int tn_mpo_apply(tn_ctx* ctx, void* stream, int conj, int Dl, int dp, int Dr, int wl, int wr, int du, const double* A,
                 const double* W, double* out);

/* sqrt(sum (a_i - b_i)^2) with b padded by the unit vector convention of mps.py:555-558 handled by the caller.
 * Result in *out (device). */
int tn_diff_norm(tn_ctx* ctx, void* stream, const double* a, const double* b, int n, double* out);

/* ---------------------------------------------------------------- one boundary-MPS row (tnac4o.py:1683-1694) */

/* Native driver of  psi <- compress_mps( MPO . psi )  = MPS.copy + apply_mpo + compress_mps (mps.py:159-200, 353-359)
 * with the reference's fixed truncation schedule; issues the primitives above back to back on `stream` without a
 * host interpreter in the loop.  A_in / W are HOST arrays of L device pointers: A_in[n] is the previous row's tensor
 * (Dl[n], dphys[n], Dr[n]); W[n] the MPO tensor with legs (wl[n], dphys[n], wr[n], du[n]) for conj = 1 (Hconj=True) or
 * (wl[n], du[n], wr[n], dphys[n]) for conj = 0.  The compressed row stays on the device inside *out. */
typedef struct tn_row tn_row;
int tn_row_compress(tn_ctx* ctx, void* stream, int L, const double* const* A_in, const int* Dl, const int* dphys,
                    const int* Dr, const double* const* W, const int* wl, const int* wr, const int* du, int conj,
                    double Dmax, double tolS, double tolV, int max_sweeps, int graduate, tn_row** out);
/* bond dimensions D[0..L] and physical dimensions d[0..L-1] of the compressed row */
int tn_row_shapes(const tn_row* row, int* D, int* d);
/* copy the tensors into caller-owned device buffers (sizes from tn_row_shapes); host outputs: overlap with the
 * uncompressed state (rhoT_overlap), per-bond discarded weights (L + 1), log2 of the accumulated norm */
int tn_row_fetch(tn_row* row, double* const* A_out, double* h_overlap, double* h_discarded, double* h_log2norm);
int tn_row_free(tn_row* row);

/* ---------------------------------------------------------------- branch-and-bound (tnac4o/tnac4o.py) */

/* Per-site constant tables, built by the host with the reference's numpy expressions and uploaded once per
 * (model, gauge) (tnac4o.py:1562-1607, 1506-1531). */
typedef struct tn_site {
    int nS;               /* number of cell states 2^n */
    int nl, nd, nr, nu;   /* leg dimensions left, down, right, up */
    const double* Wlu;    /* [nl][nu][nS]: exp(-beta E) * gauges, the slice W[:, l, d(s), r(s), u] of _peps_tensor */
    const double* Wtr;    /* [nu][nl][nd][nr]: PEPS tensor traced over the cell state, u-major */
    const uint8_t* dmap;  /* [nS] down bond index of a cell state  (_ind_bond_down, tnac4o.py:1469) */
    const uint8_t* rmap;  /* [nS] right bond index                 (_ind_bond_right, tnac4o.py:1480) */
    const double* Es;     /* [nS]      in-cell energy        (tnac4o.py:1513) */
    const double* Esl;    /* [nS][nl]  coupling to the left  (tnac4o.py:1522) */
    const double* Esu;    /* [nS][nu]  coupling upwards      (tnac4o.py:1529) */
} tn_site;

/* Builds the per-site tables on the device from the exponent tables E0[s] = beta (min Es - Es[s]), E1[s][l], E4[s][u]
 * (tnac4o.py:1571-1583), the gauge vectors Xu/Xl/Xr/Xd and the bond maps -- the compact form of _peps_tensor
 * (tnac4o.py:1586-1607) and of its trace over the cell state (tnac4o.py:1686).  Wmpo has the MPO leg order (l, d, r, u). */
int tn_build_site_tables(tn_ctx* ctx, void* stream, int nS, int nl, int nd, int nr, int nu, const double* E0,
                         const double* E1, const double* E4, const double* Xu, const double* Xl, const double* Xr,
                         const double* Xd, const uint8_t* dmap, const uint8_t* rmap, double* Wlu, double* WtrU, double* Wmpo);

/* Right environments of one row level for nb row-start branches (tnac4o.py:1776-1782):
 *   RRout[b][a][l] = sum_{p,b',r} A[a,p,b'] RRin[b][b'][r] Wtr[l,p,r,u_b] / nfactor,  u_b = up[b * up_stride].
 * A (Dl, nd, Dr); RRin (nb, Dr, nr); RRout (nb, Dl, nl).  Evaluated as AW_u = A.Wtr[..u] once per level, then one
 * grouped DMMA GEMM over the branches bucketed by u_b (temporaries from the context's stream-ordered pool); a branch's
 * result does not depend on which other branches are in the call. */
int tn_rr_level(tn_ctx* ctx, void* stream, const tn_site* site, int nb, int Dl, int Dr, const double* A,
                const double* RRin, const uint8_t* up, int up_stride, double* RRout);

/* Conditional marginals of one cell for all live branches, fused with the negative/zero rule, log2 and
 * accumulation (tnac4o.py:1786-1807, 450-453):
 *   T1 (nb, nd*Dr) = RL (nb, Dl) . A (Dl, nd*Dr)            -- computed by the caller with tn_gemm
 *   T2[b] = T1[b] (nd, Dr) . RR[root[b]] (Dr, nr)
 *   P[b][s] ~ Wlu[l_b][u_b][s] * T2[b][dmap[s]][rmap[s]], l_b = vind[b][nx], u_b = vind[b][nx+1]
 *   clamp / normalise as _calculate_Pn; cand[b][s] = log2 P + prob[b]; flag[b] = the returned negativity.
 * *max_bits receives the maximum of cand in the order-preserving uint64 encoding.  cand/prob/max_bits may be
 * NULL (Gibbs sampling needs only P_out); P_out may be NULL. */
int tn_marginals(tn_ctx* ctx, void* stream, const tn_site* site, int nb, int Dr, const double* T1, const double* RR,
                 const int32_t* root, const uint8_t* vind, int vstride, int nx, const double* prob, double* cand,
                 double* flag, unsigned long long* max_bits, double* P_out);

/* Relative cut-off (tnac4o.py:456-465): survivors are the candidates with cand > max + log2(relative_P_cutoff)
 * (all of them when the cut-off is <= 0); their flat ids go to surv (capacity n), the count to *count and, when
 * h_count is not NULL, to the host (synchronises).  The largest discarded value is max-ed into *pd_bits. */
int tn_select(tn_ctx* ctx, void* stream, const double* cand, int64_t n, const unsigned long long* max_bits,
              double relative_P_cutoff, int32_t* surv, int* count, unsigned long long* pd_bits, int* h_count);

/* New boundary row (packed into a 128-bit key with the host-supplied bit offsets), energy and log-probability of
 * every survivor (tnac4o.py:469-478, 1506-1531).  Energy is Eng[parent] + ((Es + Esl) + Esu), the reference's
 * float64 addition order.  ktie = (candidate id << 32) | survivor index makes the sort order total. */
int tn_expand(tn_ctx* ctx, void* stream, const tn_site* site, int K, int nx, int has_left, int has_up, int npos,
              const uint8_t* h_bit_offsets, const int32_t* surv, const uint8_t* vind, int vstride, const double* Eng,
              const double* cand, unsigned long long* khi, unsigned long long* klo, unsigned long long* ktie,
              int32_t* parent, int32_t* cell, double* Enew, double* Pnew);

/* Boundary merge (tnac4o.py:481-515): sort survivors by key, one group per distinct key; representative = first
 * energy minimum, degeneracy = sum over members within min_dEng, log-probability = their mean.  Key arrays need
 * tn_sort_capacity_for(K) elements.  *h_G = number of groups (synchronises). */
int tn_merge(tn_ctx* ctx, void* stream, int K, unsigned long long* khi, unsigned long long* klo, unsigned long long* ktie,
             const double* Enew, const double* Pnew, const int32_t* parent, const long long* deg, double min_dEng,
             int32_t* g_rep, long long* g_deg, double* g_prob, double* g_E, int32_t* g_start, int32_t* g_size, int* h_G);

/* Top-M over merged groups (tnac4o.py:518-526): sel[0 .. min(G, M)) lists the kept groups (descending
 * log-probability when truncating); the largest dropped value is max-ed into *pd_bits. */
int tn_topm(tn_ctx* ctx, void* stream, int G, int M, const double* g_prob, unsigned long long* khi, unsigned long long* klo,
            unsigned long long* ktie, int32_t* sel, unsigned long long* pd_bits);

/* Branch arrays of the next site and the left environments RL' = RL[parent] . A[:, d, :] / nfactor
 * (tnac4o.py:470-477, 528-535).  sel / g_rep / g_deg / g_prob may be NULL (Gibbs: branch j = sample j). */
int tn_materialise(tn_ctx* ctx, void* stream, const tn_site* site, int B, int nx, int pos, int nsites, int vstride, int Dl,
                   int Dr, const int32_t* sel, const int32_t* g_rep, const long long* g_deg, const double* g_prob,
                   const int32_t* parent, const int32_t* cell, const double* Enew, const uint8_t* vind_in,
                   const uint8_t* states_in, const int32_t* root_in, const double* RL_in, const double* A, uint8_t* vind_out,
                   uint8_t* states_out, int32_t* root_out, double* Eng_out, double* prob_out, long long* deg_out,
                   double* RL_out);

/* vind[:, 1:] = vind[:, :-1]; vind[:, 0] = 0   (tnac4o.py:540-542) */
int tn_row_shift(tn_ctx* ctx, void* stream, int B, int vstride, uint8_t* vind);

/* Gibbs step (tnac4o.py:616-627): cell = searchsorted(cumsum(P[b]), uniforms[b]); energy increment. */
int tn_sample(tn_ctx* ctx, void* stream, const tn_site* site, int B, int nx, int has_left, int has_up, const double* P,
              const double* uniforms, const uint8_t* vind, int vstride, const double* Eng, int32_t* parent, int32_t* cell,
              double* Enew);

/* Native driver of the whole branch-and-bound of search_ground_state (tnac4o.py:417-551): the calls above in the
 * reference's order, row by row and site by site, without a host interpreter in the loop.  HOST inputs: sites
 * (Ny * Nx descriptors, row-major), A ((Ny + 1) * Nx device pointers, A[ny * Nx + nx] = rhoT[ny].A[nx]; row 0 unused),
 * D ((Ny + 1) * (Nx + 1) bond dimensions), key_offsets (Ny * Nx * (Nx + 1) bit offsets of the merge key).  DEVICE
 * outputs sized for M branches; host outputs: number of final branches, largest discarded log2 P, smallest
 * negativity flag, number of branch marginals evaluated.  Synchronises the stream. */
int tn_search_ground_state(tn_ctx* ctx, void* stream, int Nx, int Ny, const tn_site* sites, const double* const* A,
                           const int* D, const uint8_t* key_offsets, int M, double relative_P_cutoff, double min_dEng,
                           uint8_t* states_out, double* Eng_out, double* prob_out, long long* deg_out, int* h_count,
                           double* h_pd_max, double* h_neg_min, long long* h_marginals);

/* ascending sort of n (hi, lo, tie) keys; arrays need tn_sort_capacity_for(n) elements */
int tn_sort_keys(tn_ctx* ctx, void* stream, unsigned long long* hi, unsigned long long* lo, unsigned long long* tie, int n);
int tn_sort_capacity_for(int n);

/* ---------------------------------------------------------------- droplets / energies (tnac4o.py:859-861, 1380-1385; auxx.py:82-107) */

/* XOR difference of pairs of cell-state rows (row of `states` with cell `pos` overridden): positions and xor
 * patterns of the differing cells, padded to nsites per pair, and their count. */
int tn_xor_diff(tn_ctx* ctx, void* stream, int npairs, int nsites, int pos, const uint8_t* states, const int32_t* row_a,
                const int32_t* cell_a, const int32_t* row_b, const int32_t* cell_b, int16_t* out_pos, uint8_t* out_xor,
                int32_t* out_len);

/* out[i] = ground XOR all droplets listed for state i (CSR flip lists, CSR droplet dictionary). */
int tn_apply_droplets(tn_ctx* ctx, void* stream, int nstates, int nsites, const uint8_t* ground, const int32_t* flip_ptr,
                      const int32_t* flip_key, const int32_t* drop_ptr, const int16_t* drop_pos, const uint8_t* drop_xor,
                      uint8_t* out);

/* Droplet bookkeeping of the spectrum search with excitations_encoding = 1 on the device -- replaces the per-site host
 * loops of _search_low_energy_spectrum_v1 (tnac4o.py:844-893): dictionary of droplet shapes (_exc_add_to_d, 2051-2069) as
 * a hash table + CSR pool, excitation lists `el` as node / children / list pools, energy pruning (_exc_cut_energy,
 * 2071-2079) as lazily composed budgets.  tn_book_site is called once per lattice site after tn_merge / tn_topm /
 * tn_materialise with the arrays those calls produced (device pointers; `order` = low words of the sorted merge keys,
 * copied before tn_topm; old_states = state rows of the branches BEFORE the site; lim_hd > 1 drops droplets touching
 * fewer cells, lim_hd < -1 droplets with fewer than -lim_hd set pattern bits -- the RMF rule of _exc_hd, 2143-2150); it
 * synchronises twice (pair count, pool growth).  tn_book_sizes / tn_book_export hand the pools to the host once, after the last site. */
typedef struct tn_book tn_book;
int tn_book_create(tn_ctx* ctx, void* stream, int nsites, int M, tn_book** out);
int tn_book_site(tn_ctx* ctx, tn_book* book, int site, int K, int Bn, const int32_t* order, const int32_t* g_rep,
                 const int32_t* g_start, const int32_t* g_size, const double* g_E, const double* g_prob, const int32_t* sel,
                 const double* Enew, const double* Pnew, const int32_t* parent, const int32_t* cell, const uint8_t* old_states,
                 double max_dEng, int lim_hd);
int tn_book_sizes(tn_ctx* ctx, tn_book* book, int64_t* h_sizes);
int tn_book_export(tn_ctx* ctx, tn_book* book, double* h_dE, double* h_dP, int32_t* h_key, int32_t* h_first, int32_t* h_last,
                   int32_t* h_cptr, int32_t* h_ccnt, int32_t* h_cnode, double* h_cbud, int32_t* h_sptr, int16_t* h_spos,
                   uint8_t* h_sxor, int32_t* h_list0);
int tn_book_free(tn_book* book);

/* Enumeration of all droplet combinations of an excitation tree (excitations_encoding = 1) with excitation energy
 * <= max_dEng, at most max_states (the lowest), sorted by energy -- replaces the Python loop _exc_unpack_v1
 * (tnac4o.py:2295-2335) as a level-synchronous expansion over a flattened tree.  HOST arrays describe the tree: node 0
 * is the root (dE 0, first -1, last nsites - 1); h_child_ptr (nnodes + 1) / h_child_idx list every node's children in
 * the reference's order; h_key[k] indexes the droplet dictionary passed to tn_decode_fetch.  Synchronises (the number of
 * combinations created per wave sizes the next launch); *h_count = number of combinations kept. */
typedef struct tn_decode tn_decode;
int tn_decode_enumerate(tn_ctx* ctx, void* stream, int nsites, int nnodes, const double* h_dE, const int32_t* h_key,
                        const int32_t* h_first, const int32_t* h_last, const int32_t* h_child_ptr, const int32_t* h_child_idx,
                        double max_dEng, int64_t max_states, tn_decode** out, int64_t* h_count);
/* Energies (excitation energy of combination i, ascending) and states (ground XOR the droplets of combination i) of the
 * first `count` combinations -- replaces the per-state loop of decode_low_energy_states (tnac4o.py:1377-1385).  Device
 * pointers; droplet dictionary in CSR form (drop_ptr / drop_pos / drop_xor). */
int tn_decode_fetch(tn_ctx* ctx, tn_decode* dec, int64_t count, const uint8_t* ground, const int32_t* drop_ptr,
                    const int16_t* drop_pos, const uint8_t* drop_xor, double* E_out, uint8_t* states_out);
int tn_decode_free(tn_decode* dec);

/* E[k] = sum_{i<j} J_ij s_i s_j + sum_i J_ii s_i for 0/1 encoded states (L per row), couplings in coordinate form. */
int tn_energy_ising(tn_ctx* ctx, void* stream, int nstates, int L, const int8_t* bits, int64_t nnz, const int32_t* ci,
                    const int32_t* cj, const double* cv, double* E);

#ifdef __cplusplus
}
#endif
#endif /* TNAC4O_B200_H */
