/*
 * visreps_b200 -- C ABI of the B200-native RSA hot path.
 *
 * The reference (yashsmehta/visreps) has no FFI layer: its boundary for this path is
 * four Python callables in visreps/analysis/rsa.py plus the bootstrap loop inlined in
 * visreps/evals.py.  Each entry point below names the reference statement(s) it replaces
 * (paths relative to the reference root).  The Python mirror of the reference interface
 * (visreps_b200/rsa.py) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - Every function returns an int status (VR_OK == 0).  On failure a thread-local
 *     message is available from vr_last_error_string().  Nothing throws, nothing
 *     allocates device memory: scratch comes from the caller (`ws`, `ws_bytes`), sized
 *     by the matching *_workspace_bytes() query.
 *   - Pointers are DEVICE pointers unless the parameter says "host".
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the
 *     call returns without synchronising.
 *   - Stimulus count n <= VR_MAX_STIMULI.  N = n(n-1)/2 is the strict upper-triangle
 *     length, in the row-major order of torch.triu_indices(n, n, 1) (rsa.py:111).
 */
#ifndef VISREPS_B200_H
#define VISREPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_ABI_VERSION 3
#define VR_MAX_STIMULI 32767

enum {
  VR_OK = 0,
  VR_ERR_INVALID = 1,     /* bad argument (maps to ValueError on the Python side) */
  VR_ERR_WORKSPACE = 2,   /* ws_bytes smaller than the *_workspace_bytes() answer  */
  VR_ERR_CUDA = 3,        /* a CUDA runtime/driver call failed                     */
  VR_ERR_UNSUPPORTED = 4  /* device is not sm_100, or a feature is not built       */
};

enum { VR_SPEARMAN = 0, VR_PEARSON = 1, VR_KENDALL = 2 };

typedef void* vr_stream_t;

int vr_abi_version(void);
/* Host, thread-local, valid until the next failing call on this thread. */
const char* vr_last_error_string(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
unsigned long long vr_launch_count(void);

/* Tuning options, for A/B measurements and for tests that must drive one specific kernel (no reference
 * counterpart).  "syrk_pair": 1 (default) = 256 x 256 pair tiles on CTA pairs for long rows, 0 = single-CTA
 * 128 x 128 tiles only.  "syrk_waves": CTA waves the SYRK's K splits are sized for (default 1).
 * Initial values come from the environment (VR_SYRK_PAIR, VR_SYRK_WAVES).
 * "sweep_tile_a" / "sweep_tile_b" (1..8, default 8): vectors per side of one Spearman sweep launch.
 * "kendall_tree": 1 (default) = bootstrap jobs of 192 or more masks build the merge tree of the RDM pair once and count every
 * replicate's discordant pairs without sorting, 2 = always, 0 = never.  "kendall_complement": 1 (default) = the sorting
 * path counts among the dropped entries of a mask, 0 = among the kept ones.  Every setting gives the same bits. */
int vr_set_option(const char* name, int value);
int vr_get_option(const char* name, int* value);

/* Built-in stage profiler (no reference counterpart; the reference has no tracing, SURVEY.md
 * section 5).  Between start and stop every stage of the library brackets its launches with CUDA
 * events on the caller's stream.  stop() synchronises those events and fills, per family
 * (vr_profile_family_count() entries, names from vr_profile_family_name()): device milliseconds,
 * number of bracketed scopes, number of kernel launches.  All three arrays are host, nullable. */
int vr_profile_family_count(void);
const char* vr_profile_family_name(int family);
int vr_profile_start(void);
int vr_profile_stop(double* ms, unsigned long long* scopes, unsigned long long* launches);

/* ------------------------------------------------------------------------------------
 * RDM construction: replaces compute_rdm(), visreps/analysis/rsa.py:59-93
 *   x -= rowmean; std = sqrt(mean(x^2)+c); std[std<10c] = 1;
 *   rdm = 1 - clamp((x x^T / d) / (std_i std_j + c), -1, 1); diag = 0.
 * X is row-major (n, d) fp32 with leading dimension ldx (elements); it is not modified
 * (rsa.py:76).  rdm is (n, n) fp32, contiguous, exactly symmetric, diagonal exactly 0.
 * zero_var_rows (nullable) receives the number of rows whose std was replaced by 1
 * (the reference logs a warning for them, rsa.py:84-87).
 * ---------------------------------------------------------------------------------- */
size_t vr_rdm_workspace_bytes(int n, int64_t d);
int vr_rdm_pearson_f32(const float* X, int n, int64_t d, int64_t ldx, float correction,
                       float* rdm, int* zero_var_rows,
                       void* ws, size_t ws_bytes, vr_stream_t stream);

/* Same, with the strict upper triangle packed in the row-major order of torch.triu_indices(n, n, 1) written by the epilogue
 * itself (what compute_rdm_correlation gathers, rsa.py:111-113): triu is n(n-1)/2 fp32; rdm or triu may be NULL, not both. */
int vr_rdm_pearson_triu_f32(const float* X, int n, int64_t d, int64_t ldx, float correction,
                            float* rdm, float* triu, int* zero_var_rows,
                            void* ws, size_t ws_bytes, vr_stream_t stream);
/* The gather alone, for an existing (n, n) matrix with leading dimension ld. */
int vr_triu_pack_f32(const float* m, int n, int64_t ld, float* out, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * The same RDM in stages, for rows whose FEATURES are split over several devices (K-split multi-GPU SYRK,
 * the 10,000 x 193,600 stress shape).  Every rank holds a column shard X_r (n, d_r):
 *   1. vr_rdm_moments_f32          per-row sum and sum of squares of the shard (fp64)   -> all-reduce (2n doubles)
 *   2. vr_rdm_shift_from_moments   c_i = fp32(sum_i / d_total), the shift every rank subtracts
 *   3. vr_rdm_gram_f32             G_r[i][j] = sum_k (x_ik - c_i)(x_jk - c_j) over the shard, j >= i (zero below the
 *                                  diagonal), tensor-core SYRK, K splits summed in fp64, fp32 out -> all-reduce (n x n)
 *   4. vr_rdm_finish_f32           rsa.py:80-92 from the reduced Gram and the whole-row moments: mean, std (+guard),
 *                                  cov = G/d - delta_i delta_j, /(std_i std_j + c), clamp, 1 - r, mirror, diag 0.
 * The Gram is reduced BEFORE the non-linear epilogue.  vr_rdm_gram_f32 needs vr_rdm_workspace_bytes(n, d_r).
 * ---------------------------------------------------------------------------------- */
int vr_rdm_moments_f32(const float* X, int n, int64_t d, int64_t ldx, double* sum, double* sumsq, vr_stream_t stream);
int vr_rdm_shift_from_moments(const double* sum, int n, int64_t d_total, float* shift, vr_stream_t stream);
int vr_rdm_gram_f32(const float* X, int n, int64_t d, int64_t ldx, const float* shift, float* gram, int64_t ldg,
                    void* ws, size_t ws_bytes, vr_stream_t stream);
int vr_rdm_finish_f32(const float* gram, int64_t ldg, int n, int64_t d_total, const double* sum, const double* sumsq,
                      const float* shift, float correction, float* rdm, int* zero_var_rows, vr_stream_t stream);

/* Row-wise 0-based ordinal ranks: replaces _rank(), rsa.py:50-52 (double argsort; ties
 * are split by position, lower index first).  out is (n, d) fp32, contiguous. */
size_t vr_rank_rows_workspace_bytes(int n, int64_t d);
int vr_rank_rows_ordinal_f32(const float* X, int n, int64_t d, int64_t ldx, float* out,
                             void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Presort: replaces the per-call `triu_indices` gather + SciPy stable argsort inside
 * compute_rdm_correlation(), rsa.py:111-113,122 (scipy.stats.spearmanr -> rankdata).
 * The strict upper triangle of `rdm` (n x n fp32, leading dimension ld) is sorted ONCE
 * by value; the result is an opaque blob of vr_presort_bytes(n) bytes that every
 * subsample of the same RDM is ranked from without sorting again.
 * -0.0 ties with +0.0; a NaN anywhere marks the blob so that every score computed from
 * it is NaN (rsa.py:123-125).
 * ---------------------------------------------------------------------------------- */
size_t vr_presort_bytes(int n);
size_t vr_presort_workspace_bytes(int n);
int vr_presort_triu_f32(const float* rdm, int n, int64_t ld, void* presorted,
                        void* ws, size_t ws_bytes, vr_stream_t stream);
/* Same, for `count` (<= 64) RDMs of identical n and ld in one batched sort.  rdms_host and
 * blobs_host are HOST arrays of device pointers. */
size_t vr_presort_batch_workspace_bytes(int n, int count);
int vr_presort_triu_batch_f32(const float* const* rdms_host, int count, int n, int64_t ld,
                              void* const* blobs_host, void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Subsample masks: turns the HOST-drawn index sets of the reference bootstrap
 * (`rng.choice(n_test, int(0.9 n_test), replace=False)`, evals.py:362-364, rsa.py:249-251)
 * into one n-bit membership mask per replicate.  idx is a device copy, (R, m) int32
 * row-major.  masks is (R, vr_mask_words(n)) uint32.  Index order inside a draw does not
 * change the multiset of sub-RDM pairs, so it does not change any score.
 * ---------------------------------------------------------------------------------- */
/* HOST function (no device work): B draws of `RandomState.choice(n, m, replace=False)` from numpy's
 * legacy MT19937 stream, bit-identical to the reference's Python loop.  mt_key (624 words) and
 * mt_pos are RandomState.get_state()[1:3], updated in place so the generator can be handed back
 * with set_state().  out is a HOST (B, m) int32 array, draw order preserved. */
int vr_host_draw_subsamples(uint32_t* mt_key, int* mt_pos, int n, int m, int B, int32_t* out);
/* HOST function: the same stream, but emitting only each draw's n-bit membership mask (the kept set is
 * the complement of the shuffle's last n-m slots, final after n-m swap steps; the other steps only
 * consume the generator).  Writes replicates [b_begin, b_end) to the HOST array masks
 * ((b_end-b_begin), vr_mask_words(n)); the generator is advanced through all B draws. */
int vr_host_draw_subsample_masks(uint32_t* mt_key, int* mt_pos, int n, int m, int B, int b_begin, int b_end,
                                 uint32_t* masks);
int vr_mask_words(int n);
/* Device copy of host-built masks; host_masks must be page-locked (it is read by the kernel, not by a
 * copy engine, so it never waits behind bulk host->device transfers). */
int vr_fetch_masks(const uint32_t* host_masks, int R, int n, uint32_t* masks, vr_stream_t stream);
int vr_build_masks(const int32_t* idx, int R, int m, int n, uint32_t* masks, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Tie-averaged ranks of a subsample: replaces `rdm[idx][:, idx]` + triu gather +
 * scipy.stats.rankdata(method="average") (evals.py:365-366 + rsa.py:111-113 + SciPy).
 * For replicate r and every pair p = (i<j) of the FULL upper triangle:
 *     dev[r][p] = 2*rank - (N_r + 1)   if both i and j are in mask r, else 0
 * where rank is the 1-based tie-averaged rank among the N_r kept pairs.  2*rank is an
 * integer, so dev is exact (int32).  dev rows have stride vr_dev_stride(n) elements.
 * ss[r] = sum_p dev^2 (fp64), nkept[r] = N_r.
 * ---------------------------------------------------------------------------------- */
int64_t vr_dev_stride(int n);
size_t vr_subsample_rank_workspace_bytes(int n, int R);
int vr_subsample_rank(const void* presorted, int n, const uint32_t* masks, int R,
                      int32_t* dev, double* ss, int64_t* nkept,
                      void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Staged form of the same computation, for callers that pipeline host->device copies with
 * the ranking (visreps_b200.rsa_grouped does): all (layer, ROI) pairs of a job share the R masks
 * (evals.py:356 re-seeds RandomState(42) per pair), so everything derived from the masks alone is
 * built once (the rank context) and each RDM is ranked as soon as it exists.
 *   vr_rank_context_build  masks (R, vr_mask_words(n)) -> opaque context of vr_rank_context_bytes(n, R)
 *   vr_rank_materialise    dev (nV, R, vr_dev_stride(n)) int32 and ss (nV, R) fp64 of nV presorted RDMs
 *                          (nkept (nV, R) is nullable)
 *   vr_rank_fused          ranks nF presorted RDMs and, without writing their dev vectors, correlates them
 *                          with nM <= vr_rank_fuse_max() materialised ones:
 *                          out[f * out_stride_fused + m * out_stride_mat + r] = Spearman rho (fp64)
 * blobs_host are HOST arrays of device pointers; ws needs vr_rank_workspace_bytes(n, nV or nF, R).
 * ---------------------------------------------------------------------------------- */
size_t vr_rank_context_bytes(int n, int R);
int vr_rank_context_build(const uint32_t* masks, int n, int R, void* ctx, size_t ctx_bytes, vr_stream_t stream);
size_t vr_rank_workspace_bytes(int n, int n_vectors, int R);
int vr_rank_fuse_max(void);
int vr_rank_materialise(const void* const* blobs_host, int nV, int n, const void* ctx, int R,
                        int32_t* dev, double* ss, int64_t* nkept,
                        void* ws, size_t ws_bytes, vr_stream_t stream);
int vr_rank_fused(const void* const* blobs_host, int nF, int n, const void* ctx, int R,
                  const int32_t* mat_dev, const double* mat_ss, int nM,
                  double* out, int64_t out_stride_fused, int64_t out_stride_mat,
                  void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Spearman job: the same grouped comparison with one warp per (canonical pair, 32 replicates) -- csrc/sweep.cu.
 * Replaces evals.py:361-369 for any number of RDMs that share the R masks of a job (evals.py:356).
 *   vr_spearman_job_init   masks -> per-batch transposed masks and N_r inside the opaque job buffer
 *                          (vr_spearman_job_bytes(n, slots, R) bytes; `slots` = RDMs the job can hold)
 *   vr_spearman_job_add    count tables (rows of {kept word, 2 C - N_r + 1} per 32 sorted positions and replicate) and tie
 *                          sums of `count` presorted RDMs into slots slot0 .. slot0 + count - 1; may be called as the RDMs
 *                          become available (the host pipeline overlaps it with host->device copies)
 *   vr_spearman_job_sweep  rho for every (A vector, B vector, replicate):
 *                          out[a * strideA + b * strideB + r], fp64, replicate order = mask order
 * blobs*_host / slots*_host are HOST arrays (device blob pointers, slot numbers); n, R and slots must be the values the
 * job was initialised with.  Sums are exact integers while N^3 < 9e18 (N <= ~2 M pairs), fp64 beyond.
 * ---------------------------------------------------------------------------------- */
size_t vr_spearman_job_bytes(int n, int slots, int R);
int vr_spearman_job_init(const uint32_t* masks, int n, int R, int slots, void* job, size_t job_bytes, vr_stream_t stream);
int vr_spearman_job_add(void* job, int n, int R, int slots, const void* const* blobs_host, int count, int slot0, vr_stream_t stream);
size_t vr_spearman_sweep_workspace_bytes(int n, int nA, int nB, int R);
int vr_spearman_job_sweep(void* job, int n, int R, int slots,
                          const void* const* blobsA_host, const int* slotsA_host, int nA,
                          const void* const* blobsB_host, const int* slotsB_host, int nB,
                          double* out, int64_t strideA, int64_t strideB,
                          void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Grouped bootstrap comparison: replaces the loop at evals.py:361-369 (and its copies at
 * evals.py:512-520, rsa.py:248-256, run_reconstruction.py:148-154) for nA model RDMs x
 * nB neural RDMs x R masks in one call.  A mask with all n bits set gives the point
 * estimate (evals.py:347).  out is (nA, nB, R) fp64, replicate order = mask order.
 *   Spearman: inputs are presorted blobs (host arrays of device pointers).
 *   Pearson : inputs are the RDMs themselves (n x n fp32, leading dimension ld); only the strict
 *             upper triangle is read.  The moments of every replicate are taken by inclusion-exclusion
 *             over its DROPPED stimuli (totals and row sums live in the workspace, which is required).
 * Undefined correlations (constant vector, < 2 kept stimuli, a NaN among the kept pairs) give NaN.
 * ---------------------------------------------------------------------------------- */
size_t vr_bootstrap_spearman_workspace_bytes(int n, int nA, int nB, int R);
int vr_bootstrap_spearman(const void* const* presortedA_host, int nA,
                          const void* const* presortedB_host, int nB,
                          int n, const uint32_t* masks, int R, double* out,
                          void* ws, size_t ws_bytes, vr_stream_t stream);

size_t vr_bootstrap_pearson_workspace_bytes(int n, int nA, int nB, int R);
int vr_bootstrap_pearson(const float* const* rdmA_host, int nA,
                         const float* const* rdmB_host, int nB,
                         int n, int64_t ld, const uint32_t* masks, int R, double* out,
                         void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Kendall tau-a of two RDMs' upper triangles for R subsample masks: replaces _kendall_tau_a,
 * rsa.py:22-40 (SciPy tau-b rescaled to (C - D) / n0) inside compute_rdm_correlation (rsa.py:122)
 * and inside the bootstrap loops when compare_method = kendall (evals.py:229, utils.py:518).
 * out is (R,) fp64; NaN for < 2 kept pairs, a constant vector or a NaN entry.
 * Jobs of 192 or more masks (option "kendall_tree") build the merge tree of the pair once and count every replicate
 * without sorting, 256 replicates per pass; smaller jobs sort the kept (or dropped) entries of each replicate in
 * batches of up to 64, the batch shrinking to what the workspace holds.  vr_kendall_workspace_bytes(n) is the
 * minimum (sorting path, one replicate at a time), vr_kendall_batch_workspace_bytes(n, R) the size at which an
 * R-mask job takes its fastest path (the tree needs (levels + 2) x 4 N bytes, levels = ceil(log2(N / 32))).
 * ---------------------------------------------------------------------------------- */
size_t vr_kendall_workspace_bytes(int n);
size_t vr_kendall_batch_workspace_bytes(int n, int R);
int vr_kendall_tau_a(const float* rdmA, const float* rdmB, int n, int64_t ld,
                     const uint32_t* masks, int R, double* out,
                     void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Bootstrap of the encoding score: replaces the loop at
 * visreps/analysis/encoding_score.py:228-240,
 *     scores[i] = correlation_score(Y_test[boot_idx], pred_test[boot_idx]).mean()
 * with himalaya's correlation_score = per-column Pearson r (population z-scores).  y_true / y_pred are
 * (n, V) fp32 row-major on the device; mask r keeps the rows of draw r (mask of all ones = the point
 * estimate, encoding_score.py:61).  mean_out is (R,) fp64: the mean over the V columns; per_column
 * (nullable) receives the (R, V) fp32 column scores.  A column without variance among the kept rows makes
 * its replicate NaN.
 * ---------------------------------------------------------------------------------- */
size_t vr_bootstrap_column_corr_workspace_bytes(int n, int V, int R);
int vr_bootstrap_column_corr_f32(const float* y_true, const float* y_pred, int n, int V,
                                 int64_t ld_true, int64_t ld_pred, const uint32_t* masks, int R,
                                 double* mean_out, float* per_column, void* ws, size_t ws_bytes, vr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Percentile confidence interval of bootstrap replicates: replaces np.percentile(scores, [2.5, 97.5]) at
 * visreps/evals.py:371-372, :520-521 and visreps/analysis/rsa.py:262-263 (NumPy's default 'linear' method) bit for
 * bit.  scores: rows x B fp64 on the device (row r at scores + r * row_stride); ci_lo / ci_hi: (rows,) fp64.
 * A row holding a NaN gives NaN bounds.  1 <= B <= vr_percentile_ci_max_replicates().
 * ---------------------------------------------------------------------------------- */
int vr_percentile_ci_max_replicates(void);
int vr_percentile_ci_f64(const double* scores, int rows, int B, int64_t row_stride, double q_lo, double q_hi,
                         double* ci_lo, double* ci_hi, vr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VISREPS_B200_H */
