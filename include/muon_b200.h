/*
 * muon_b200.h -- C ABI of the B200-native sparse hot path of scverse/muon.
 *
 * muon itself is pure Python and has no FFI: the reference boundary for this path is the
 * Python signature of mu.atac.pp.tfidf / mu.atac.tl.lsi / mu.tl.mofa.
 * This header is the boundary *below* that Python shim: every entry point is what a
 * ctypes / cffi / pybind binding of the reference would bind for the arithmetic the
 * reference delegates to scipy / ARPACK / mofapy2.  Each declaration cites the reference
 * lines whose arithmetic it replaces (paths relative to the muon checkout).
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - return 0 on success, <0 on error; mub_last_error() returns a thread-local message.
 *   - every pointer is a DEVICE pointer unless its name ends in _h (host).
 *   - the library never allocates or frees caller memory; scratch space is passed in and
 *     sized by the matching *_workspace_bytes() query.
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*); nothing
 *     synchronises unless documented.  Re-entrant, no global mutable state.
 *   - CSR layout: indptr int64[n_rows+1], indices int32[nnz], values float32[nnz]
 *     (nnz may exceed 2^31: 6e9 at BASELINE configs[1]).
 *   - dense operands are row-major with a leading dimension `ld` (floats) that must be one
 *     of the padded widths 32, 64 or 128; columns >= the logical width must be zero.
 */
#ifndef MUON_B200_H
#define MUON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mub_stream_t; /* cudaStream_t */

/* flags of mub_tfidf_* : keyword arguments of tfidf(), muon/_atac/preproc.py:18-21,101 */
#define MUB_TFIDF_LOG_TF 1u    /* log_tf=True      preproc.py:103-104 */
#define MUB_TFIDF_LOG_IDF 2u   /* log_idf=True     preproc.py:107-108 */
#define MUB_TFIDF_LOG_TFIDF 4u /* log_tfidf=True   preproc.py:116-117 */
#define MUB_TFIDF_NO_SCALE 8u  /* scale_factor in {None,0,1}: multiply skipped, preproc.py:101 */
#define MUB_TFIDF_BINARIZE 16u /* treat every stored non-zero as 1: binarize() (preproc.py:132-152) fused in */

int mub_version(void);
const char* mub_last_error(void);
/* sm count, compute capability and L2 size of the current device */
int mub_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);

/* ---- TF-IDF (muon/_atac/preproc.py:92-119) -------------------------------------------- */
/* pass 1: row_sum[i] = sum_j c_ij (preproc.py:93);  col_sum[j] += sum_i c_ij (preproc.py:106).
 * row_sum is overwritten, col_sum is ACCUMULATED (zero it first; allreduce it across
 * cell shards before pass 2).  status (optional int32 word, zeroed by the caller) receives
 * bit0 if some row's column indices are not strictly increasing (unsorted or duplicate entries)
 * and bit1 if an explicit zero is stored: such input must be canonicalised first to reproduce
 * the reference's output pattern (scipy's matmul merges duplicates and drops zeros). */
int mub_tfidf_reduce_f32(const int64_t* indptr, const int32_t* indices, const float* data,
                         int64_t n_rows, int32_t n_cols, float* row_sum, float* col_sum,
                         int32_t* status, uint32_t flags, mub_stream_t stream);
int mub_tfidf_reduce_f64(const int64_t* indptr, const int32_t* indices, const double* data,
                         int64_t n_rows, int32_t n_cols, double* row_sum, double* col_sum,
                         int32_t* status, uint32_t flags, mub_stream_t stream);
/* pass 1, tiled variant for fp32 matrices whose rows hold SORTED column indices: a CTA owns mub_tfidf_tile_rows()
 * rows and sweeps the columns tile by tile with the tile's column sums and entry counts in shared memory (two
 * shared-memory atomics per non-zero, one global flush per tile) instead of one global atomic per non-zero.
 * Same outputs and status bits as mub_tfidf_reduce_f32; status is mandatory (bit0 set = some row is not sorted:
 * the sums are then invalid, use mub_tfidf_reduce_f32).  row_sum / indptr point at the first row of the call,
 * row_base = its absolute row number (a multiple of the tile height).  Optional by-product: col_count
 * [n_chunks x n_cols] int32 (zeroed by the caller, accumulated) = stored entries per (row chunk, column), row
 * chunks given by chunk_bounds[n_chunks+1] (absolute rows, multiples of the tile height): the histogram
 * mub_csr_transpose_count would otherwise compute in a pass of its own.  Second optional by-product: rb_count
 * uint16[ceil(n_rows_total / tile height)][n_cols] = stored entries per (row block, column), written (not accumulated)
 * for the blocks of this call; mub_csr_transpose_fill_tiled turns it into write offsets. */
int mub_tfidf_reduce_tiled_f32(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                               int32_t n_cols, float* row_sum, float* col_sum, int32_t* status, uint32_t flags,
                               int32_t* col_count, const int64_t* chunk_bounds, int32_t n_chunks, int64_t row_base,
                               uint16_t* rb_count, mub_stream_t stream);
int mub_tfidf_tile_rows(void);
/* idf[j] = n_obs_total / col_sum[j], log1p if MUB_TFIDF_LOG_IDF (preproc.py:106-108) */
int mub_tfidf_idf_f32(const float* col_sum, int32_t n_cols, double n_obs_total, uint32_t flags,
                      float* idf, mub_stream_t stream);
int mub_tfidf_idf_f64(const double* col_sum, int32_t n_cols, double n_obs_total, uint32_t flags,
                      double* idf, mub_stream_t stream);
/* pass 2: out_ij = log1p(((1/r_i) * c_ij) * sf) * idf_j   (preproc.py:94-96,101-104,110-112,
 * 116-117), association order as written.  data_out may alias data_in (in place). */
int mub_tfidf_apply_f32(const int64_t* indptr, const int32_t* indices, const float* data_in,
                        float* data_out, int64_t n_rows, int32_t n_cols, const float* row_sum,
                        const float* idf, float scale_factor, uint32_t flags, mub_stream_t stream);
int mub_tfidf_apply_f64(const int64_t* indptr, const int32_t* indices, const double* data_in,
                        double* data_out, int64_t n_rows, int32_t n_cols, const double* row_sum,
                        const double* idf, double scale_factor, uint32_t flags, mub_stream_t stream);

/* ---- CSR x dense SpMM: the operator applications inside svds ------------------------------
 * C[n_rows x ld] (=|+=) A * B[n_cols x ld].  Replaces scipy's csr_matvec / csc_matvec loop
 * driven by ARPACK (scipy _svds.py:428-460 called from muon/_atac/tools.py:53) and the dense
 * Y^T Z / Y W contractions of mofapy2 (called from muon/_core/tools.py:583-585).
 * ld in {32,64,128}.  accumulate: 0 -> C = A*B, 1 -> C += A*B.
 * row_counter: optional device int64 scratch word, zeroed by the caller, enabling dynamic row
 * scheduling (recommended for skewed row lengths, e.g. transposed ATAC matrices); may be NULL. */
int mub_spmm_csr_f32(const int64_t* indptr, const int32_t* indices, const float* data,
                     int64_t n_rows, int64_t n_cols, const float* B, int32_t ld, float* C,
                     int32_t accumulate, unsigned long long* row_counter, mub_stream_t stream);

/* Same product, shared-memory staged variant for matrices whose column indices are SORTED within
 * each row (canonical CSR): column panels of B are TMA-copied (cp.async.bulk) into shared memory
 * behind mbarriers and a persistent CTA sweeps a block of rows over them with register-resident
 * accumulators, cutting the L2->SM gather traffic of the row-warp kernel.  Deterministic. */
int mub_spmm_csr_panel_f32(const int64_t* indptr, const int32_t* indices, const float* data,
                           int64_t n_rows, int64_t n_cols, const float* B, int32_t ld, float* C,
                           int32_t accumulate, mub_stream_t stream);

/* ---- CSR transpose (one-time per matrix): builds the CSR of A^T so that A^T * Y is again a
 * row-gather SpMM (scipy does the same implicitly through csc_matvec, _svds.py:447).
 * Step 1 counts entries per column into t_count[n_cols+1] (int64, zeroed by caller, slot 0 unused
 * so that an inclusive scan of t_count is t_indptr); the caller scans (any device scan), then
 * step 2 scatters.  cursor: int64[n_cols] scratch, overwritten.  Entry order inside a
 * transposed row is not deterministic (atomic slot claim): sums over it differ run to run in
 * the last fp32 bits. */
int mub_csr_transpose_count(const int32_t* indices, int64_t nnz, int32_t n_cols, int64_t* t_count,
                            mub_stream_t stream);
int mub_csr_transpose_fill(const int64_t* indptr, const int32_t* indices, const float* data,
                           int64_t n_rows, int32_t n_cols, int64_t row_offset,
                           const int64_t* t_indptr, int64_t* cursor, int32_t* t_indices,
                           float* t_data, mub_stream_t stream);

/* "Pairs" layout for transposed panels: entry k of row j is the 8-byte pair {int32 index, float32 value bits},
 * so the scatter writes once per non-zero and the product reads one 8-byte word per entry.
 * fill_pairs: like mub_csr_transpose_fill, writing t_pairs[2*nnz] (int32 view of the int2 array).
 * spmm_csrp : C (+)= A B for a pairs-layout matrix (same kernel family and rules as mub_spmm_csr_f32).
 * csrp_row_stats: mub_csr_row_stats_f32 on the pairs layout. */
int mub_csr_transpose_fill_pairs(const int64_t* indptr, const int32_t* indices, const float* data,
                                 int64_t n_rows, int32_t n_cols, int64_t row_offset, const int64_t* t_indptr,
                                 int64_t* cursor, int32_t* t_pairs, mub_stream_t stream);
/* fill_pairs without global atomics, for a row panel that starts on a row-block boundary: rb_count points at the
 * panel's first row block (see mub_tfidf_reduce_tiled_f32), indptr at its first row, t_indptr[n_cols+1] are the panel's
 * transposed row offsets, base is uint32 scratch [ceil(n_rows / tile height)][n_cols] (the scanned write offsets),
 * status (int32, zeroed by the caller) gets bit0 if counts and t_indptr disagree (then the result is invalid). */
int mub_csr_transpose_fill_tiled(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                                 int32_t n_cols, const uint16_t* rb_count, const int64_t* t_indptr, uint32_t* base,
                                 int32_t* t_pairs, int32_t* status, mub_stream_t stream);
int mub_spmm_csrp_f32(const int64_t* indptr, const int32_t* pairs, int64_t n_rows, int64_t n_cols,
                      const float* B, int32_t ld, float* C, int32_t accumulate, unsigned long long* row_counter,
                      mub_stream_t stream);
int mub_csrp_row_stats_f32(const int64_t* indptr, const int32_t* pairs, int64_t n_rows, double* sum,
                           double* sumsq, mub_stream_t stream);

/* per-row sum and sum of squares of a CSR (fp64).  On the CSR of A^T these are the per-feature
 * moments behind MOFA's centring / intercepts (muon/_core/tools.py:283-287, mofapy2 process_data). */
int mub_csr_row_stats_f32(const int64_t* indptr, const float* data, int64_t n_rows, double* sum,
                          double* sumsq, mub_stream_t stream);

/* ---- tall-skinny Gram (the k x k contraction that is allreduced across cell shards) ------
 * G[l x l] = Y^T Y for Y[n x ld] row-major, float64 result (partials are fp32 per CTA slab,
 * combined in fp64 in a fixed order: deterministic).  Replaces the dense Gram inside
 * scipy.linalg.svd(Av) (_svds.py:511-521) and mofapy2's Z^T Z / W^T diag(tau) W.
 * weights: optional per-row weight w_i (G = Y^T diag(w) Y), may be NULL. */
size_t mub_gram_workspace_bytes(int64_t n, int32_t ld);
int mub_gram_f32(const float* Y, const float* weights, int64_t n, int32_t ld, int32_t l, double* G,
                 void* workspace, mub_stream_t stream);

/* ---- MOFA+ variational updates (mofapy2 training loop reached from muon/_core/tools.py:583-585;
 * equations restated in oracle/mofa_ref.py).  One thread per row, Gauss-Seidel over the K factors,
 * fp64 arithmetic on fp32 storage.  Dense operands are row-major with leading dimension ld >= K.
 * Centring is implicit: Praw = Y^T E[Z] of the UN-centred sparse view, mu[D] the feature means
 * (NULL = no centring), zsum[K] = column sums of E[Z]; inv_scale = 1/std for scale_views.
 *   update_w : spike-and-slab weights of one view over G groups of cells (G=1: no grouping).  Arrays
 *              carry a leading group dimension: Praw[G][D x ld], mu[G][D], zsum[G][K], inv_scale[G],
 *              ZZ[G][K x K] (E[Z^T Z] of the group's cells observed in this view, E[z^2] on the
 *              diagonal), tau[G][D] = E[tau]; alpha/lnth/ln1mth[K] = E[alpha], E[ln theta], E[ln(1-theta)].
 *              W is read (current means) and overwritten; WW = E[(sw)^2], S = q(s=1),
 *              What2 = E[what^2] (both branches) are written.
 *   update_z : factors; Q = sum_m Y_m (tau*W_m) un-centred, and per cell class c (group x set of views
 *              the cell is observed in; cls[N] or NULL = one class): qshift[C][K] the centring correction,
 *              GW[C][K x K] = sum_m W^T diag(tau) W, zvar[C][K] = Var[z_k].  Z read and overwritten.
 *   tau      : b_out[D] = b0 + 1/2 E||y_d - Z w_d||^2 from ssq[D] (centred sum of squares), one
 *              (view, group) block per call. */
int mub_mofa_update_w_f32(const float* Praw, const float* mu, const double* zsum, const double* inv_scale,
                          const double* ZZ, const float* tau, const double* alpha, const double* lnth,
                          const double* ln1mth, float* W, float* WW, float* S, float* What2, int64_t D,
                          int32_t ld, int32_t K, int32_t G, int32_t spikeslab, mub_stream_t stream);
int mub_mofa_update_z_f32(const float* Q, const double* qshift, const double* GW, const double* zvar,
                          const int32_t* cls, float* Z, int64_t N, int32_t ld, int32_t K, int32_t C,
                          mub_stream_t stream);
int mub_mofa_tau_f32(const float* Praw, const float* mu, const double* zsum, double inv_scale, const double* ZZ,
                     const double* ssq, const float* W, const float* WW, double b0, double* b_out, int64_t D,
                     int32_t ld, int32_t K, mub_stream_t stream);

/* Non-gaussian likelihoods (mofapy2 guesses "poisson" for integer and "bernoulli" for binary views when muon
 * passes likelihoods=None, muon/_core/tools.py:272-280): Seeger pseudo-data of a DENSE row-major view around
 * zeta = E[Z] E[W]^T at fixed precision kappa[D] (restated in oracle/mofa_ref.py::mofa_ref_general).
 * kind: 1 = poisson  yhat = zeta - sigmoid(zeta) (1 - y / ln(1+e^zeta)) / kappa_d
 *       2 = bernoulli yhat = zeta - (sigmoid(zeta) - y) / kappa_d
 * mofa_pseudo overwrites zeta[n_rows x D] with yhat; mofa_loglik ACCUMULATES sum ln p(y | zeta) into *out (device
 * double; poisson: y ln rate - rate, bernoulli: y zeta - ln(1+e^zeta)). */
int mub_mofa_pseudo_f32(float* zeta, const float* obs, const float* kappa, int64_t n_rows, int32_t D, int32_t kind,
                        mub_stream_t stream);
int mub_mofa_loglik_f32(const float* zeta, const float* obs, int64_t n_rows, int32_t D, int32_t kind, double* out,
                        mub_stream_t stream);


/* ---- exact k-nearest neighbours (groundwork for the WNN row, muon/_core/preproc.py:520-528: the reference asks
 * umap's NN-descent for n_multineighbors+1 = 201 neighbours per cell and modality; this search is exact).
 * For every row of X[nq x ld] (d meaningful columns) the k nearest rows of Y[nc x ld] in Euclidean distance,
 * ascending, ties by lower index; out_idx[nq x k] (-1 if fewer than k candidates), out_dist[nq x k] (not squared).
 * X == Y is allowed (a point is then its own first neighbour at distance exactly 0).  k <= 320. */
int mub_knn_l2_f32(const float* X, int64_t nq, const float* Y, int64_t nc, int32_t d, int32_t ld, int32_t k,
                   int32_t* out_idx, float* out_dist, mub_stream_t stream);

/* Same result (bit-identical indices and distances), tensor-core version: tcgen05 TF32 inner products in TMEM select
 * the candidates (threshold widened by the TF32 error bound so that no true neighbour is lost), an fp32 pass with
 * the arithmetic of mub_knn_l2_f32 ranks them.  d <= 128, k <= 512.  `workspace` of
 * mub_knn_l2_tc_workspace_bytes(nq, nc, d) bytes; *status (zeroed by the caller) gets bit1 if some query had more than
 * ~900 points inside its error band (result then incomplete: use mub_knn_l2_f32), bit2 on an internal timeout. */
size_t mub_knn_l2_tc_workspace_bytes(int64_t nq, int64_t nc, int32_t d);
int mub_knn_l2_tc_f32(const float* X, int64_t nq, const float* Y, int64_t nc, int32_t d, int32_t ld, int32_t k,
                      int32_t* out_idx, float* out_dist, void* workspace, size_t workspace_bytes, int32_t* status,
                      mub_stream_t stream);

/* ---- WNN building blocks (muon/_core/preproc.py:264-640; numba helpers :46-159) ------------------------------
 * bandwidth: for every cell i the n_bw cells minimising the reference's tie-breaking metric
 *   (N - jd*N) + (bbox - e)/bbox  (jd = Jaccard distance of the kNN sets, e = Euclidean distance; preproc.py:51-76)
 *   among cells sharing a neighbour with i, enumerated exactly through the transposed kNN graph
 *   (g_* = kNN graph CSR with sorted rows, t_* = CSR of its transpose); sigma[i] = mean Euclidean distance to
 *   them (preproc.py:462-470).  First pass: cell_list = NULL (tables in shared memory); cells sharing neighbours
 *   with more than 1536 cells get sigma = -1 and status bit0, and are redone by a second call with cell_list
 *   (status bit1 = even the fallback tables were too small).
 * affinity_topk: union of the per-modality candidate lists cands[m][n x n_cand] (-1 = none), affinity
 *   sum_m weight[i,m] * exp(-||x^m_i - x^m_j|| / sigmas[m][i]), distance sqrt(0.5 (1 - affinity)) and the n_out
 *   smallest per cell, ascending (preproc.py:569-604 and _sparse_csr_fast_knn_ :114-135).  reps / dims / lds /
 *   cands / sigmas are HOST arrays of n_mod device pointers / ints. */
int mub_wnn_bandwidth_f32(const int64_t* g_indptr, const int32_t* g_indices, const int64_t* t_indptr,
                          const int32_t* t_indices, const float* X, int64_t n, int32_t d, int32_t ld, int32_t n_bw,
                          double bbox_norm, double* sigma, int32_t* status, const int64_t* cell_list,
                          int64_t n_cells, void* workspace, int32_t table_slots, int32_t n_tables,
                          mub_stream_t stream);
/* fallback pass for hub cells: cell_list[n_cells] = cells with sigma == -1 after the first pass, hash tables of
 * table_slots (power of two) entries in `workspace` (mub_wnn_bandwidth_workspace_bytes(table_slots, n_tables)) */
size_t mub_wnn_bandwidth_workspace_bytes(int32_t table_slots, int32_t n_tables);
int mub_wnn_affinity_topk_f32(int32_t n_mod, const float* const* reps, const int32_t* dims, const int32_t* lds,
                              const int32_t* const* cands, const double* const* sigmas, const double* weight,
                              int64_t n, int32_t n_cand, int32_t n_out, int32_t* out_idx, double* out_dist,
                              int32_t* status, mub_stream_t stream);

/* ---- half-precision dense operand for the early block-Lanczos steps of the svds replacement --------------
 * (scipy _svds.py:428-460 as driven from muon/_atac/tools.py:53).  The operator application is bound by the
 * 4*ld bytes of the dense operand every non-zero pulls through L2 -> L1; with B stored as IEEE half that is
 * 2*ld.  Products and sums stay fp32.  f32_to_f16_scaled: dst[i] = half(src[i] * scale), n a multiple of 4
 * (scale: a power of two that lifts orthonormal columns, |x| <= 1, out of the subnormal range).
 * spmm_csr_h16 / spmm_csrp_h16: C (=|+=) out_scale * A * B_half, same rules as mub_spmm_csr_f32. */
int mub_f32_to_f16_scaled(const float* src, int64_t n, float scale, void* dst, mub_stream_t stream);
int mub_spmm_csr_h16(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                     int64_t n_cols, const void* B_half, int32_t ld, float* C, int32_t accumulate,
                     float out_scale, unsigned long long* row_counter, mub_stream_t stream);
int mub_spmm_csrp_h16(const int64_t* indptr, const int32_t* pairs, int64_t n_rows, int64_t n_cols,
                      const void* B_half, int32_t ld, float* C, int32_t accumulate, float out_scale,
                      unsigned long long* row_counter, mub_stream_t stream);

/* ---- host <-> device staging (ingest side: adata.X is a pageable scipy CSR, muon/_atac/preproc.py:86-129
 * reads it and rebinds a freshly allocated host matrix).  A stager owns a ring of pinned buffers and a pool
 * of host threads; transfers are chunked through the ring so that the host-side copy of chunk i+1 overlaps
 * the DMA of chunk i.  Fused into the host-side copy: int64 -> int32 narrowing of scipy's index arrays
 * (nnz >= 2^31) and a position-dependent 64-bit fingerprint  sum_i (e_i + C1) * (i*C2 + C3)  of the 32-bit
 * element stream, by which a later call can prove a host array still equals its device twin.
 * Pointers ending in _h are HOST pointers.  A stager is not thread-safe: one transfer at a time.
 *   create : n_bufs pinned buffers of buf_bytes each (n_bufs = 0: thread pool only, no CUDA context needed)
 *   h2d    : src_elem_bytes 1, 4 or 8; narrow = 1 (8-byte sources): int64 -> int32; narrow = 2 (4-byte sources):
 *            float32 -> uint8 for count data (widen on the device with mub_u8_to_f32); *overflow_h = 1 if a value
 *            does not survive the narrowing (the destination is then garbage: resend without narrowing); hash_h
 *            (optional; 4-byte sources or narrow = 1) receives the fingerprint.  Chunks are enqueued on `stream`;
 *            returns when the last chunk is enqueued.
 *   d2h    : synchronous; hash_h (optional, n_bytes % 4 == 0) receives the fingerprint of what was written.
 *   host_fingerprint   : fingerprint of a host array (elem_bytes 4, or 8 = int64 read as narrowed int32).
 *   device_fingerprint : the same function of a device array of n 32-bit elements, ACCUMULATED into *out
 *            (device uint64, zeroed by the caller). */
int mub_stager_create(size_t buf_bytes, int32_t n_bufs, int32_t n_threads, void** out);
int mub_stager_destroy(void* stager);
int mub_stager_h2d(void* stager, const void* src_h, void* dst, size_t n_elems, int32_t src_elem_bytes,
                   int32_t narrow, uint64_t* hash_h, int32_t* overflow_h, mub_stream_t stream);
int mub_stager_d2h(void* stager, const void* src, void* dst_h, size_t n_bytes, uint64_t* hash_h,
                   mub_stream_t stream);
int mub_u8_to_f32(const void* src, int64_t n, float* dst, mub_stream_t stream);
int mub_host_fingerprint(void* stager, const void* src_h, size_t n_elems, int32_t elem_bytes, uint64_t* hash_h);
int mub_device_fingerprint(const void* src, int64_t n, uint64_t* out, mub_stream_t stream);

/* ---- synthetic ATAC count generator (benchmark / test input) ---------------
 * Deterministic counter-based planted-topic model; bit-identical to the numpy generator in
 * muon_b200/_synth.py.  Step 1 writes nnz per row; caller scans into indptr; step 2 fills.
 * beta[n_cols], topic[n_topics x n_cols], row_topic[n_rows], row_scale[n_rows] are host-made
 * tables (device copies); row0 is the global index of the first row (shards hash identically). */
int mub_synth_count(int64_t row0, int64_t n_rows, int32_t n_cols, const float* beta, const float* topic,
                    const int32_t* row_topic, const float* row_scale, uint64_t seed, int64_t* row_nnz,
                    mub_stream_t stream);
int mub_synth_fill(int64_t row0, int64_t n_rows, int32_t n_cols, const float* beta, const float* topic,
                   const int32_t* row_topic, const float* row_scale, uint64_t seed, const int64_t* indptr,
                   int32_t* indices, float* data, mub_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MUON_B200_H */
