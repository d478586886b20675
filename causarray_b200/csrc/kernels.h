// Internal launcher interface between the translation units of
// libcausarray_b200.  Everything here takes DEVICE pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ca {

// A "row problem": rows of Y (cells for the A-step, genes for the B-step)
// against the fixed factor M (ncols x kpad).
struct RowProblem {
  const double* Y;   // nrows x ldY, row-major
  long ldY;          // multiple of 8
  int ncols;
  const double* M;   // ncols x kpad
  int kpad;          // kpad % 8 == 4
  const double* nu;  // per column (nu_per_row = 0) or per row (nu_per_row = 1)
  const double* inv_nu;  // 1/nu, same indexing
  const double* log_nu;  // log(nu)
  const void* colparm;   // per column (nu, 1/nu, log nu, poisson flag) as double4, for nu_per_row = 0
  int nu_per_row;
  int any_pois;      // some nu exceeds thres_disp
  int family;
  double thres_disp;
  // optional (gradient passes only): theta of the columns with theta_pos[col] >= 0 is also written to
  // theta_out[row * theta_ld + theta_pos[col]] (stage-2 one-hot line search, search_onehot.cu)
  double* theta_out = nullptr;
  const int* theta_pos = nullptr;
  long theta_ld = 0;
};

// stage-2 gene line search for one-hot treatment indicators (search_onehot.cu)
struct S2Launch {
  const double* Ytr;
  const double* Th;
  const signed char* grp;
  long ldtr;
  int ntr, a, kpad, t_lo, ncols_total;
  double* X;
  const double* g;
  const double* lam;
  const double* f0;
  const double* normg;
  const double* tys;
  const double* ell0;
  const double* alpha_row;
  double alpha, beta, tol;
  int max_iters;
  const int* rowlist;
  const int* count;
  int nrows;
  const double* nu;
  const double* inv_nu;
  const double* log_nu;
  int family;
  double thres;
  double* ell_fin;
  int* neval;
  unsigned long long* eval_total;
};
int launch_s2_onehot(const S2Launch& L, cudaStream_t st);
int launch_s2_gather(const double* Ynt, long ldn, int p, const int* cells, int ntr, double* Ytr, long ldtr,
                     cudaStream_t st);

struct SearchLaunch {
  double* X;            // rows x kpad, updated in place
  const double* g;      // rows x kpad
  const double* f0;
  const double* normg;
  const double* tys;
  const double* lam;    // rows x ldlam or nullptr
  int ldlam;
  int prox_lo, prox_hi;
  int mv_lo, mv_hi;     // columns in which g can be non-zero (0, kpad when unknown)
  const double* alpha_row;
  double alpha, beta, tol;
  int max_iters;
  const int* rowlist;
  const int* count;
  int nrows;            // launch bound on the number of rows
  double* ell_fin;
  int* neval;
  unsigned long long* eval_total;  // optional device counter of candidate evaluations
  // two-phase search (all four set, or all null for the single-kernel sequential walk): rows that
  // fail their first candidate are listed (ordered by `neval` of their previous search when given)
  // and finished by the multi-candidate kernel
  unsigned char* pend;   // nrows
  double* f01s;          // nrows
  double* e01s;          // nrows
  int* list2;            // nrows
  int* count2;
};

// gene-sharded cell line search (altmin_aux.cu)
struct LsShard {
  const int* rowlist;
  const int* count;
  double* X;             // rows x kpad: current point, overwritten with the accepted point
  const double* g;       // rows x kpad
  double* Xc;            // rows x kpad: candidate rows
  double* t;             // per row: current step
  int* ne;               // per row: candidates evaluated
  const double* ell;     // per row: log-likelihood sum of the candidate, reduced over the ranks
  const double* tys;
  const double* f0;
  const double* normg;
  double* f01;
  double* e01;
  double* ell_fin;
  unsigned char* pend;
  int* n_pending;
  unsigned long long* eval_total;
  double alpha, beta, tol, ncols_total;
  int kpad, round, max_iters;
};
int launch_ls_candidates(const LsShard& a, int nrows, cudaStream_t st);
int launch_ls_decide(const LsShard& a, int nrows, cudaStream_t st);

int launch_rowpass(const RowProblem& rp, const double* X, const int* rowlist, const int* count, int nrows,
                   bool want_grad, int gc_lo, int gc_n, double* ell_out, double* G_out, int ldG,
                   cudaStream_t st);
int launch_search(const RowProblem& rp, const SearchLaunch& s, cudaStream_t st);

// ---- altmin_aux.cu ----
int launch_normalize_transpose(const double* Yraw, int n, int p, long ld_in, const double* size_factor,
                               double* Yn, long ldYn, double* Ynt, long ldYnt, cudaStream_t st);
int launch_transpose(const double* src, int rows, int cols, long ld_in, double* dst, long ld_out, cudaStream_t st);
int launch_tys_rowsum(const double* Yr, long ldY, int nrows, int ncols, const double* nu, int nu_per_row,
                      int family, double thres, double* out, cudaStream_t st);
int launch_rowsum(const double* M, long ld, int nrows, int ncols, double* out, cudaStream_t st);
int launch_zero_fraction(const double* Yr, long ldY, int nrows, int ncols, double* out, cudaStream_t st);
int launch_sum(const double* v, int n, double* out, cudaStream_t st);
int launch_div_scalar(double* v, long n, double denom, cudaStream_t st);
// W (cp x cv) = P^T V ; P rows x cp (ldP), V rows x cv (ldV)
int launch_pt_v(const double* P, int ldP, int cp, const double* V, int ldV, int cv, int rows, double* W,
                cudaStream_t st);
// V -= P W
int launch_sub_p_w(const double* P, int ldP, int cp, double* V, int ldV, int cv, int rows, const double* W,
                   cudaStream_t st);

struct PrepArgs {
  int nrows;
  const int* rowlist;
  const int* count;
  int kpad, ncols;
  const double* G;   // scaled gradient block, rows x ldG, columns map to [gc_lo, gc_lo + gc_n)
  int ldG, gc_lo, gc_n;
  int ball_lo, ball_hi;
  double radius;
  const double* X;     // rows x kpad
  const double* lam;   // rows x ldlam or nullptr
  int ldlam, prox_lo, prox_hi;
  const double* ell0;
  const double* tys;
  double* g;           // rows x kpad (out)
  double* f0;
  double* normg;
};
int launch_prep_search(const PrepArgs& a, cudaStream_t st);

// widen a dense count matrix of element type `dtype` (CA_F32 ... CA_U8 of the public header) to FP64
int launch_widen_counts(const void* src, int dtype, long count, double* dst, cudaStream_t st);
int launch_csr_expand(const long long* indptr, const int* indices, const double* data, int n, int p, double* Y,
                      cudaStream_t st);
int launch_compact(const uint8_t* mask, int n, int* list, int* count, cudaStream_t st);
int launch_compact_by_key(const uint8_t* mask, const int* key, int n, int* list, int* count, cudaStream_t st);
// clip X[:, lo:hi) to [-bound, bound]; rows where something changed are appended to list/count
int launch_clip(double* X, int rows, int kpad, int lo, int hi, double bound, int* list, int* count, cudaStream_t st);
// mask[i] = || X[i, lo:hi) - Xprev[i, lo:hi) ||_2 > tol ; n_active written to count
int launch_step_mask(const double* X, const double* Xprev, int rows, int kpad, int lo, int hi, double tol,
                     uint8_t* mask, int* count, cudaStream_t st);
int launch_fill_u8(uint8_t* v, int n, uint8_t val, cudaStream_t st);
// out = sum_{i, c in [lo,hi)} |X[i,c]| * lam[i*ldlam + c-lo]
int l1_penalty_blocks(int rows);
int launch_l1_penalty(const double* X, int rows, int kpad, int lo, int hi, const double* lam, int ldlam,
                      double* partial, double* out, cudaStream_t st);
// lam[i, c] = lam0 / (|X[i, lo + c]| + 1e-6)
int launch_adaptive_weights(const double* X, int rows, int kpad, int lo, int hi, double lam0, double* lam,
                            int ldlam, cudaStream_t st);

int launch_check_counts(const double* Y, long total, unsigned long long* flags, cudaStream_t st);
int launch_size_factor_medians(const double* Yraw, long ldY, const double* Yraw_t, long ldn, int n, int p,
                               double* log_geo, unsigned long long* scratch, double* out, cudaStream_t st);

}  // namespace ca
