/* oracle/gcate_oracle.c -- TEST INFRASTRUCTURE (CPU checker and timed CPU baseline).
 *
 * Plain-C / OpenMP restatement of ONE alternating iteration of the reference:
 *   update_with_mask   causarray/gcate_opt.py:148-206
 *   line_search        causarray/gcate_opt.py:30-82   (sequential candidates, early return)
 *   project_norm_ball  causarray/gcate_opt.py:10-26
 *   nll / grad         causarray/gcate_likelihood.py:46-142
 * with the same work structure as the reference's numba kernels: the gradient is
 * a full pass over the matrix, every active row then runs its own backtracking
 * loop, one log-likelihood evaluation (k-dot + exp/log per entry) per candidate,
 * rows distributed over threads like numba's prange.  It is pinned against the
 * real reference through tests/golden/kernels_*.npz (tests/test_oracle_golden.py).
 * Nothing in causarray_b200/ links or calls this file.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC -o _build/libgcate_oracle.so gcate_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* per-entry log-likelihood kernel y*theta - b and its derivative factor */
static inline double ell_entry(double th, double y, double nu, int pois) {
  double xi = th < 100.0 ? th : 100.0;
  double m = exp(xi);
  if (pois) return y * xi - m;
  double q = clip(1.0 / (1.0 + m / nu), 1e-6, 1.0 - 1e-6);
  return y * log1p(-q) + nu * log(q);
}
static inline double dl_entry(double th, double y, double nu, int pois) {
  double xi = th < 100.0 ? th : 100.0;
  double m = exp(xi);
  if (pois) return -(y - m);
  double q = clip(1.0 / (1.0 + m / nu), 1e-6, 1.0 - 1e-6);
  return -(y - m) * q;
}

/* nll of one row: y (nc), M (nc x k), x (k); nu per column (nu_col) or scalar (nu_row) */
static double row_nll(const double* y, const double* M, const double* x, int nc, int k, int family,
                      const double* nu_col, double nu_row, double thres, const double* tys) {
  double s = 0.0;
  for (int c = 0; c < nc; ++c) {
    const double* m = M + (size_t)c * k;
    double th = 0.0;
    for (int q = 0; q < k; ++q) th += m[q] * x[q];
    double nu = nu_col ? nu_col[c] : nu_row;
    int pois = (family == 0) || (nu > thres);
    s += ell_entry(th, y[c], nu, pois) + tys[c];
  }
  return -s / (double)nc;
}

static inline double soft(double v, double lam) {
  double a = fabs(v) - lam;
  if (a < 0.0) a = 0.0;
  return v > 0.0 ? a : (v < 0.0 ? -a : 0.0);
}

static double pen_mean(const double* x, const double* lam, int start, int d) {
  if (d <= start || !lam) return 0.0;
  double s = 0.0;
  for (int c = start; c < d; ++c) s += lam[c - start] * fabs(x[c]);
  return s / (double)(d - start);
}

/* line_search: returns number of candidate evaluations; x0 is overwritten with the result */
static int line_search(const double* y, const double* M, double* x0, const double* g, int nc, int k, int d,
                       int family, const double* nu_col, double nu_row, const double* tys, double thres, int start,
                       const double* lam, double alpha, double beta, int max_iters, double tol, double* x1) {
  double f0 = row_nll(y, M, x0, nc, k, family, nu_col, nu_row, thres, tys) + pen_mean(x0, lam, start, d);
  double t = alpha, ng = 0.0, f1 = 0.0, f01 = 0.0;
  int evals = 0;
  for (int q = 0; q < k; ++q) ng += g[q] * g[q];
  ng = sqrt(ng);
  for (int i = 0; i < max_iters; ++i) {
    for (int q = 0; q < k; ++q) x1[q] = x0[q] - t * g[q];
    if (lam) for (int c = start; c < d; ++c) x1[c] = soft(x1[c], lam[c - start]);
    f1 = row_nll(y, M, x1, nc, k, family, nu_col, nu_row, thres, tys) + pen_mean(x1, lam, start, d);
    ++evals;
    if (i == 0) f01 = f1;
    if (f1 < f0 - tol * t * ng) {
      memcpy(x0, x1, sizeof(double) * k);
      return evals;
    }
    if (t < 1e-4) break;
    t *= beta;
  }
  if (f01 < 1.1 * f1) {
    for (int q = 0; q < k; ++q) x1[q] = x0[q] - alpha * g[q];
    if (lam) for (int c = start; c < d; ++c) x1[c] = soft(x1[c], lam[c - start]);
  }
  memcpy(x0, x1, sizeof(double) * k);
  return evals;
}

static void ball(double* g, int lo, int hi, double radius) {
  double mx = 0.0;
  for (int c = lo; c < hi; ++c) if (fabs(g[c]) > mx) mx = fabs(g[c]);
  if (mx > radius) for (int c = lo; c < hi; ++c) g[c] *= radius / mx;
}

/* V[:, lo:lo+cv) -= P (P^T V[:, lo:lo+cv)) ; P rows x cp, V rows x ldv */
static void project_out(const double* P, int cp, double* V, int ldv, int lo, int cv, int rows) {
  double* W = (double*)calloc((size_t)cp * cv, sizeof(double));
  for (int i = 0; i < rows; ++i)
    for (int a = 0; a < cp; ++a)
      for (int b = 0; b < cv; ++b) W[a * cv + b] += P[(size_t)i * cp + a] * V[(size_t)i * ldv + lo + b];
#pragma omp parallel for schedule(static)
  for (int i = 0; i < rows; ++i)
    for (int b = 0; b < cv; ++b) {
      double s = 0.0;
      for (int a = 0; a < cp; ++a) s += P[(size_t)i * cp + a] * W[a * cv + b];
      V[(size_t)i * ldv + lo + b] -= s;
    }
  free(W);
}

/* One alternating iteration.  Y (n x p) normalised, Yt (p x n) its transpose, Ys / Yst the
 * constant term in both layouts, A (n x k), B (p x k) updated in place, lam (p x la) with
 * la = a (thresholds on columns [d-a, d)) or NULL, nu (p).  stage: 1 = P1 given, 2 = P2 given,
 * 0 = neither.  Returns nll(Y, A, B); evals[0/1] += cell / gene line-search evaluations. */
double gcate_update_with_mask(int n, int p, int k, int d, int a, const double* Y, const double* Yt, double* A,
                              double* B, const double* lam, const double* P1, int d1, const double* P2, int r2,
                              int family, const double* nu, const double* Ys, const double* Yst, double thres,
                              double C, double alpha, double beta, int max_iters, double tol,
                              const unsigned char* gene_active, const unsigned char* cell_active,
                              const double* alpha_gene, long long* evals) {
  long long ev_c = 0, ev_g = 0;
  /* ---- A-step ---- */
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : ev_c)
  for (int i = 0; i < n; ++i) {
    if (cell_active && !cell_active[i]) continue;
    double g[64], x1[64];
    memset(g, 0, sizeof g);
    const double* y = Y + (size_t)i * p;
    double* x = A + (size_t)i * k;
    for (int j = 0; j < p; ++j) {
      const double* m = B + (size_t)j * k;
      double th = 0.0;
      for (int q = 0; q < k; ++q) th += m[q] * x[q];
      int pois = (family == 0) || (nu[j] > thres);
      double dl = dl_entry(th, y[j], nu[j], pois);
      for (int q = d; q < k; ++q) g[q] += dl * m[q];
    }
    for (int q = d; q < k; ++q) g[q] /= (double)p;
    ball(g, d, k, 2.0 * C);
    ev_c += line_search(y, B, x, g, p, k, d, family, nu, 0.0, Ys + (size_t)i * p, thres, 0, NULL, alpha, beta,
                        max_iters, tol, x1);
  }
  if (P1) project_out(P1, d1, A, k, d, k - d, n);
  /* ---- B-step ---- */
  int glo, ghi, blo, bhi;
  if (P1) { glo = d; ghi = k; blo = d; bhi = k; }
  else if (P2) { glo = d - a; ghi = d; blo = 0; bhi = d; }
  else { glo = 0; ghi = k; blo = d; bhi = k; }
  double* G = (double*)calloc((size_t)p * k, sizeof(double));
#pragma omp parallel for schedule(static)
  for (int j = 0; j < p; ++j) {
    const double* y = Yt + (size_t)j * n;
    const double* x = B + (size_t)j * k;
    double* g = G + (size_t)j * k;
    int pois = (family == 0) || (nu[j] > thres);
    for (int i = 0; i < n; ++i) {
      const double* m = A + (size_t)i * k;
      double th = 0.0;
      for (int q = 0; q < k; ++q) th += m[q] * x[q];
      double dl = dl_entry(th, y[i], nu[j], pois);
      for (int q = glo; q < ghi; ++q) g[q] += dl * m[q];
    }
    for (int q = glo; q < ghi; ++q) g[q] /= (double)n;
  }
  if (P2) project_out(P2, r2, G, k, d - a, a, p);
  double total = 0.0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : ev_g)
  for (int j = 0; j < p; ++j) {
    if (gene_active && !gene_active[j]) continue;
    double x1[64];
    double* g = G + (size_t)j * k;
    ball(g, blo, bhi, 2.0 * C);
    ev_g += line_search(Yt + (size_t)j * n, A, B + (size_t)j * k, g, n, k, d, family, NULL, nu[j],
                        Yst + (size_t)j * n, thres, d - a, lam ? lam + (size_t)j * a : NULL,
                        alpha_gene ? alpha_gene[j] : alpha, beta, max_iters, tol, x1);
  }
  free(G);
  if (!P2)
    for (size_t e = 0; e < (size_t)p; ++e)
      for (int q = d; q < k; ++q) B[e * k + q] = clip(B[e * k + q], -10.0, 10.0);
  /* full nll */
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int j = 0; j < p; ++j) {
    const double* y = Yt + (size_t)j * n;
    const double* x = B + (size_t)j * k;
    const double* ty = Yst + (size_t)j * n;
    int pois = (family == 0) || (nu[j] > thres);
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
      const double* m = A + (size_t)i * k;
      double th = 0.0;
      for (int q = 0; q < k; ++q) th += m[q] * x[q];
      s += ell_entry(th, y[i], nu[j], pois) + ty[i];
    }
    total += s;
  }
  if (evals) { evals[0] += ev_c; evals[1] += ev_g; }
  return -total / (double)n;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is entitled to all host cores */
void gcate_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int gcate_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
