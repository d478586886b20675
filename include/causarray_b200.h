/* causarray_b200 -- C ABI of the B200-native GCATE + doubly-robust LFC hot path.
 *
 * The reference (jaydu1/causarray) is pure Python and has no FFI of its own;
 * its three Python-level seams are what these entry points
 * replace.  Every function returns 0 on success and a negative code on error
 * (-1 CUDA error, -2 invalid argument); ca_last_error() returns the message.
 * No function throws, none is re-entrant on the same handle (like the
 * reference's numba kernels, which are called from one Python thread).
 *
 * Unless a parameter is documented as a device pointer, pointers are HOST
 * pointers to C-contiguous arrays; the library owns all device memory.
 */
#ifndef CAUSARRAY_B200_H
#define CAUSARRAY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_FAMILY_POISSON 0
#define CA_FAMILY_NB 1

int ca_version(void);
const char* ca_last_error(void);
int ca_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem);
int ca_set_device(int device);
/* Device buffers of destroyed handles are cached for reuse (cudaFree of a batch costs ~0.2 s);
 * this returns the cache to the driver.  The cache is bounded by CA_POOL_GB (default 24). */
int ca_release_cached(void);

/* ------------------------------------------------------------------------
 * S1 -- optimiser seam.  Replaces causarray/gcate_opt.py:148-206
 * (update_with_mask), :89-145 (update), :209-458 (alter_min loop) and the
 * numba kernels they call: nll / grad (causarray/gcate_likelihood.py:46-142),
 * line_search (causarray/gcate_opt.py:30-82), project_norm_ball (:10-26).
 * ------------------------------------------------------------------------ */
typedef struct ca_gcate* ca_gcate_t;

/* n cells, p genes, d observed columns ([X | A], causarray/gcate.py:121),
 * r latent factors; family CA_FAMILY_*; thres_disp as gcate_opt.py:212. */
int ca_gcate_create(ca_gcate_t* out, int n, int p, int d, int r, int family, double thres_disp);
int ca_gcate_destroy(ca_gcate_t h);

/* Raw counts Y (n x p), size factors (n) or NULL (= ones), NB sizes nu (p) or
 * NULL (= ones).  Computes Y / size_factor and the constant
 * Tys = -lgamma(Y+1) [+ lgamma(nu+Y) - lgamma(nu)] sums on the device
 * (causarray/gcate_opt.py:343-347) and the per-gene zero fraction (:305). */
int ca_gcate_load_counts(ca_gcate_t h, const double* Y, const double* size_factor, const double* nu);
/* ---- gene shards over ranks (one process per GPU) -------------------
 * A handle may hold p of p_total genes (columns of Y, rows of B, nu, lam, gene masks); A / U, X and
 * the size factors are replicated.  The gene step of update_with_mask (gcate_opt.py:182-200) is
 * gene-local; the cell step sums over genes (grad(Y^T, B, A) :168 and the per-candidate nll of the
 * line search :173-175), as do the returned nll (:204), the L1 term (:440) and the stage-2
 * projector P2^T G (:188): those sums are completed with NCCL all-reduces on the handle's stream.
 * ca_comm_unique_id: 128-byte NCCL id generated on one rank and handed to all of them;
 * ca_gcate_set_shard: collective over the ranks, before the counts are loaded. */
int ca_comm_unique_id(uint8_t* out128);
int ca_gcate_set_shard(ca_gcate_t h, int rank, int world, int64_t p_total, const uint8_t* id128);
/* transport of the cell-step reductions of a sharded handle: 0 none (one rank), 1 NCCL, 2 NVLink peer memory */
int ca_gcate_transport(ca_gcate_t h);
/* The same in three steps, so that the input checks of fit_gcate
 * (causarray/gcate.py:12-17: integer-valued, no all-zero gene) and the
 * median-of-ratios size factors (causarray/utils.py:179-219) run on the device
 * copy instead of as host passes:
 *   upload:  H2D of the raw counts (+ transpose); optionally counts entries
 *            that are not integer-valued / not finite and genes that are empty
 *   size_factor_medians: log_sf[i] = median_j (log Y_ij - mean_i' log Y_i'j) over
 *            expressed genes (the caller centres and exponentiates, utils.py:213)
 *   prepare: normalisation by the size factors + constant-term sums. */
int ca_gcate_upload_counts(ca_gcate_t h, const double* Y, int64_t* n_nonint, int64_t* n_nonfinite,
                           int64_t* n_empty_genes);
/* Rows `rows[0..n)` of a larger host matrix (leading dimension ld elements, element type dtype as in
 * ca_gcate_upload_counts_typed; CA_F64 allowed) become the n x p count matrix of the handle: the batch
 * drivers (causarray/gcate.py:525-560 build Y[cell_idx] on the host) hand over the full matrix and the
 * cell indices of a perturbation batch; the gather runs inside the pinned-buffer staging copy. */
int ca_gcate_upload_counts_rows(ca_gcate_t h, const void* Y, int64_t ld, const int64_t* rows, int dtype,
                                int64_t* n_nonint, int64_t* n_nonfinite, int64_t* n_empty_genes);

/* Same from a CSR matrix of the counts (scipy.sparse layout: indptr n+1 int64, indices nnz int32,
 * data nnz double, no duplicate entries): only nnz * 12 bytes cross the bus, the dense copy is
 * expanded on the device.  Replaces the host-side densification of a batch slice
 * (_maybe_densify, causarray/nb_glm_fast.py:68-84; causarray/gcate.py:451-470). */
int ca_gcate_upload_counts_csr(ca_gcate_t h, const int64_t* indptr, const int32_t* indices, const double* data,
                               int64_t nnz, int64_t* n_nonint, int64_t* n_nonfinite, int64_t* n_empty_genes);
/* Same from a dense matrix of another element type: count matrices usually arrive as int32 /
 * float32 (AnnData) or small unsigned integers; the reference casts them on the host
 * (Y.astype(float), causarray/gcate.py:24-36), here the narrow array crosses the bus and is widened
 * to FP64 on the device.  dtype: CA_F64 0, CA_F32 1, CA_I32 2, CA_I64 3, CA_U16 4, CA_U8 5. */
enum { CA_F64 = 0, CA_F32 = 1, CA_I32 = 2, CA_I64 = 3, CA_U16 = 4, CA_U8 = 5 };
int ca_gcate_upload_counts_typed(ca_gcate_t h, const void* Y, int dtype, int64_t* n_nonint, int64_t* n_nonfinite,
                                 int64_t* n_empty_genes);
int ca_gcate_size_factor_medians(ca_gcate_t h, double* log_sf);
int ca_gcate_prepare(ca_gcate_t h, const double* size_factor, const double* nu);
/* Same, for callers of update_with_mask that already hold the normalised Y
 * and the Ys matrix (both n x p); Ys may be NULL (recomputed from nu). */
int ca_gcate_load_normalized(ca_gcate_t h, const double* Yn, const double* Ys, const double* nu);

/* A (n x (d+r)), B (p x (d+r)) */
int ca_gcate_set_factors(ca_gcate_t h, const double* A, const double* B);
int ca_gcate_get_factors(ca_gcate_t h, double* A, double* B);

/* Stage configuration.  P1: thin Q of [X | A] (n x d1) or NULL; P2: thin Q of
 * Gamma-hat (p x r2) or NULL; lam: L1 thresholds (p x a) or NULL (= 0);
 * a: number of penalised columns (the block [d-a, d)). */
int ca_gcate_set_stage(ca_gcate_t h, const double* P1, int d1, const double* P2, int r2, const double* lam, int a);
/* gene_active (p), cell_active (n) as uint8 or NULL (= all active);
 * alpha_gene (p) or NULL (= alpha of the call). */
int ca_gcate_set_masks(ca_gcate_t h, const uint8_t* gene_active, const uint8_t* cell_active,
                       const double* alpha_gene);
int ca_gcate_get_masks(ca_gcate_t h, uint8_t* gene_active, uint8_t* cell_active);

/* One alternating iteration == update_with_mask(...) ; A and B are updated on
 * the device (fetch with ca_gcate_get_factors).  func_val = nll(Y, A, B). */
int ca_gcate_update(ca_gcate_t h, double C, double alpha, double beta, int max_iters, double tol,
                    double* func_val);

/* nll(Y, A, B) with the current factors (gcate_likelihood.py:46-94). */
int ca_gcate_nll(ca_gcate_t h, double* out);
/* grad: wrt_A = 0 -> grad(Y, A, B) (p x k); wrt_A = 1 -> grad(Y^T, B, A) (n x k)
 * (gcate_likelihood.py:98-142; call sites gcate_opt.py:168,182). */
int ca_gcate_grad(ca_gcate_t h, int wrt_A, double* out);

typedef struct {
  /* kwargs_ls (gcate_opt.py:265-274) */
  double alpha, beta, tol, C;
  int ls_max_iters;
  double tol_gene, tol_cell;
  int recheck_interval;
  double sparsity_threshold, sparsity_boost;
  int warmup_iters;
  /* kwargs_es (gcate_opt.py:280) */
  int es_max_iters, es_warmup, es_patience;
  double es_tolerance, es_rel_tol;
  /* penalty (gcate_opt.py:352-353): weights = lam / (|B[:, d-a:d]| + 1e-6) */
  double lam;
  int a;
} ca_altmin_params;

typedef struct {
  int n_iter;
  double func_val, resid;
  int hist_len;
  int64_t total_gene_updates, total_cell_updates;
  int64_t cell_evals, gene_evals; /* line-search evaluations actually run */
  double loop_seconds;
  int64_t gene_evals_onehot;      /* of gene_evals: candidates evaluated by the one-hot stage-2 search, which
                                     touches only the treated cells */
  int64_t onehot_treated_cells;   /* cells with a treatment indicator set (0 when that path is not taken) */
} ca_altmin_result;

/* The optimiser loop of alter_min (gcate_opt.py:337-458) from the factors set
 * with ca_gcate_set_factors and the stage set with ca_gcate_set_stage (lam of
 * set_stage is ignored; the adaptive weights are computed on the device).
 * hist must have room for es_max_iters + 1 doubles. */
int ca_gcate_alter_min(ca_gcate_t h, const ca_altmin_params* prm, ca_altmin_result* res, double* hist);

/* ------------------------------------------------------------------------
 * S2 -- per-gene GLM seam.  Replaces fit_glm / fit_glm_auto
 * (causarray/gcate_glm.py:95-272, 365-457; IRLS arithmetic = statsmodels
 * GLM._fit_irls, see oracle/glm.py), estimate_disp (:275-308) and the
 * deviance residuals (causarray/nb_glm_fast.py:729-777).
 * ------------------------------------------------------------------------ */
typedef struct ca_glm* ca_glm_t;

int ca_glm_create(ca_glm_t* out, int n, int p);
int ca_glm_destroy(ca_glm_t h);
/* Raw counts (n x p). */
int ca_glm_load_counts(ca_glm_t h, const double* Y);
/* Share the counts already resident in a GCATE handle (no copy). */
int ca_glm_share_counts(ca_glm_t h, ca_gcate_t g);

/* Fit p independent GLMs  y_j ~ X beta_j + offset  (log link).
 *   X (n x kx), offset (n) or NULL, nu (p) NB sizes (family NB) or NULL,
 *   genes with nu > thres_disp are fitted as Poisson (gcate_glm.py:194).
 *   d_guard: |beta[:d_guard]| > 50 or |beta[d_guard:]| > 10 marks the gene as
 *   diverged (status 1, gcate_glm.py:205); pass kx to disable the second bound.
 * Outputs: B (p x kx), dev (p), n_iter (p), status (p; 0 ok, 1 guard, 2 numeric).
 * Fitted means stay on the device for ca_glm_resid / ca_glm_disp_moments. */
int ca_glm_fit(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
               double thres_disp, int maxiter, double tol, int d_guard, double* B, double* dev, int32_t* n_iter,
               int32_t* status);
/* Extended fit: ridge (kx) is added to the diagonal of every gene's normal
 * equations (penalised IRLS; the device-side stand-in for the reference's
 * regularised fallback, gcate_glm.py:209); gene_mask (p, uint8) restricts the
 * fit to the flagged genes -- the others keep the coefficients of the previous
 * fit of the same width.  Either may be NULL. */
int ca_glm_fit_ex(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                  double thres_disp, int maxiter, double tol, int d_guard, const double* ridge,
                  const uint8_t* gene_mask, double* B, double* dev, int32_t* n_iter, int32_t* status);
/* Second stage of the reference's two-stage "fast" fit (causarray/nb_glm_fast.py:500-567: for every
 * perturbation a 1-D treatment-effect fit against the covariate offset X B_cov^T with the dispersion
 * held fixed).  X (n, kx) must contain a one-hot block of treatment indicators; B (p, kx) carries the
 * covariate coefficients in, which are kept, and returns them together with the fitted treatment
 * coefficients (all perturbations in one launch: the one-hot columns decouple once the covariate part
 * is an offset).  Same convergence rule and outputs as ca_glm_fit_ex. */
int ca_glm_fit_treatment(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                         double thres_disp, int maxiter, double tol, double* B, double* deviance, int32_t* n_iter,
                         int32_t* status);

/* Lasso refit of the genes flagged in gene_mask (p, uint8), warm-started from the coefficients of
 * the previous fit of the same width: penalised IRLS whose fixed point minimises
 *   -loglike(beta) / n + sum_k alpha_k |beta_k|,   l1[k] = n * alpha_k  (kx entries, caller's order),
 * the objective of the statsmodels fit_regularized(alpha, L1_wt = 1) call the reference falls back
 * to when a gene trips its coefficient guards (causarray/gcate_glm.py:205-209).  Each IRLS step
 * solves its quadratic model by coordinate descent on the kx x kx normal equations.  The other
 * genes keep their coefficients.  Outputs as ca_glm_fit (no guard is applied). */
int ca_glm_refit_l1(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                    double thres_disp, int maxiter, double tol, const double* l1, const uint8_t* gene_mask,
                    double* B, double* dev, int32_t* n_iter, int32_t* status);
/* Signed deviance residuals of the last fit, (n x p). */
int ca_glm_resid(ca_glm_t h, double* resid);
/* Same residual matrix left on the device: *dev_ptr receives a DEVICE pointer
 * to (n x p) row-major doubles owned by the handle (valid until the next
 * call on it).  Used to run the initial top-r SVD (gcate_opt.py:312) without
 * a host round trip. */
int ca_glm_resid_device(ca_glm_t h, void** dev_ptr);
/* Fitted means of the last fit, (n x p). */
/* Residuals of the last fit on the device in both layouts: *dev_ptr (n, p) cell-major and *dev_ptr_t
 * (p, *ldt) gene-major -- the two operands of the subspace iteration below.  Valid until the next fit. */
int ca_glm_resid_device2(ca_glm_t h, void** dev_ptr, void** dev_ptr_t, int64_t* ldt);
int ca_glm_fitted(ca_glm_t h, double* mu);
/* Method-of-moments NB size from the last (Poisson) fit (gcate_glm.py:301-306). */
int ca_glm_disp_moments(ca_glm_t h, const double* offset, double* disp);

/* ------------------------------------------------------------------------
 * S3 -- doubly-robust reduction seam.  Replaces AIPW_mean
 * (causarray/DR_estimation.py:624-659), the normalisation
 * etas /= size_factors (causarray/DR_learner.py:208) and the LFC estimand
 * closure (causarray/DR_learner.py:403-482) without materialising the
 * (n, p, a, 2) tensors: Yhat0 = exp(W Bcov^T + offset), Yhat1_k = Yhat0 *
 * exp(Btrt[:,k]), both clipped at yhat_cap (DR_estimation.py:615).
 * ------------------------------------------------------------------------ */
typedef struct {
  int usevar;          /* 0 pooled, 1 unequal (Welch) */
  double thres_min, thres_diff, eps_var;
  double yhat_cap;     /* 1e5 */
} ca_lfc_params;

/* Y (n x p) raw counts, W (n x kw), Bcov (p x kw), Btrt (p x a), offset (n) or
 * NULL, size_factor (n), A (n x a) 0/1 as double, pi (n x a), cell_mask
 * (n x a) uint8 or NULL (= ctrl | case_j).
 * Outputs, each (a x p) treatment-major: mean0, mean1, tau, var, df,
 * estimable (uint8). */
int ca_aipw_lfc(int n, int p, int kw, int a, const double* Y, const double* W, const double* Bcov,
                const double* Btrt, const double* offset, const double* size_factor, const double* A,
                const double* pi, const uint8_t* cell_mask, const ca_lfc_params* prm, double* mean0,
                double* mean1, double* tau, double* var, double* df, uint8_t* estimable);
/* Same with the counts shared from a GLM handle (already on the device). */
int ca_aipw_lfc_shared(ca_glm_t h, int kw, int a, const double* W, const double* Bcov, const double* Btrt,
                       const double* offset, const double* size_factor, const double* A, const double* pi,
                       const uint8_t* cell_mask, const ca_lfc_params* prm, double* mean0, double* mean1,
                       double* tau, double* var, double* df, uint8_t* estimable);
/* Generic variant with explicit Y_hat (n x p x a x 2) for API compatibility
 * with LFC(..., Y_hat=...). */
int ca_aipw_lfc_yhat(int n, int p, int a, const double* Y, const double* Yhat, const double* size_factor,
                     const double* A, const double* pi, const uint8_t* cell_mask, const ca_lfc_params* prm,
                     double* mean0, double* mean1, double* tau, double* var, double* df, uint8_t* estimable);

/* ---- host-side helpers of the inference tail (CPU, no device work) ------------------------------
 * Replace the per-treatment loop around statsmodels multipletests(method='fdr_bh') and the
 * median / MAD of the empirical-null standardisation in bh_correction
 * (causarray/DR_inference.py:153-201).  p, q, t: (a, m) row-major host arrays, one BH family per row;
 * NaN entries are left out of a family and come back as NaN.  mad is the raw median absolute
 * deviation (the caller divides by Phi^{-1}(3/4)). */
int ca_host_bh_rows(const double* p, int a, int m, double* q);
int ca_host_median_mad_rows(const double* t, int a, int m, double* med, double* mad);

/* The same tail on the device, all treatments in one call (replaces the per-treatment
 * bh_correction loop of causarray/DR_learner.py:225-231 / DR_inference.py:153-201).
 * t, df: (a, p) row-major HOST arrays of test statistics (NaN = not tested) and degrees of freedom
 * (df == NULL: normal tails, scipy.stats.norm.sf; else Student-t, scipy.stats.t.sf); outputs (a, p)
 * host arrays: two-sided p-values, BH q-values within each row, and both after the empirical-null
 * standardisation (t - median) / (MAD / Phi^-1(3/4)) (NaN rows when MAD == 0); med_mad (a, 2) or NULL. */
int ca_inference_tail(const double* t, const double* df, int a, int p, double* pvalue, double* padj,
                      double* pvalue_adj, double* padj_adj, double* med_mad);

/* ------------------------------------------------------------------------
 * K7: top-r singular triplets of the deviance-residual matrix (GCATE initialisation; replaces
 * scipy.sparse.linalg.svds of causarray/gcate_opt.py:312 and gcate.py:296).  Block subspace iteration
 * (width r + oversample <= 64) with Cholesky-QR2 and Rayleigh-Ritz, hand-written FP64 DMMA products.
 * R (n, p) and Rt (p, ldt) are DEVICE pointers (ca_glm_resid_device2); zero_genes (p, host) or NULL;
 * U (n, r), S (r), Vt (r, p) host outputs, singular values DESCENDING; returns -3 on a numerical
 * breakdown (rank-deficient block).
 * ------------------------------------------------------------------------ */
int ca_svd_topr(void* R, void* Rt, int64_t ldt, int n, int p, int r, int oversample, int max_iter, double tol,
                uint32_t seed, const uint8_t* zero_genes, double* U, double* S, double* Vt, int32_t* n_iter);

/* ------------------------------------------------------------------------
 * Measurement hooks (bench.py): per-kernel-class device time from CUDA events
 * recorded on the launching stream, and the number of kernels launched so far.
 * ------------------------------------------------------------------------ */
int ca_profile_enable(int on);
int ca_profile_reset(void);
int ca_profile_num_classes(void);
const char* ca_profile_class_name(int cls);
int ca_profile_get(int cls, double* total_ms, int64_t* launches);
int64_t ca_launch_count(void);
/* NVTX ranges for the host-side phases (size factors, GLM initialisation, alternating minimisation,
 * propensities, AIPW, inference); the kernel classes push their own ranges inside the library. */
int ca_nvtx_push(const char* name);
int ca_nvtx_pop(void);
/* Test hook: out[i] = fm_exp(x[i]) (op 0, -700 <= x <= 100) or fm_log(x[i]) (op 1, x > 0 normal),
 * the table-driven FP64 functions of csrc/fastmath.cuh. */
int ca_test_fastmath(const double* x, int n, int op, double* out);

#ifdef __cplusplus
}
#endif
#endif /* CAUSARRAY_B200_H */
