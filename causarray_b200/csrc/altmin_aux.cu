// Small streaming / bookkeeping kernels around the alternating minimisation
// (K3 of SURVEY.md section 2b): size-factor normalisation + transpose, the
// constant log h(y) sums (causarray/gcate_opt.py:343-347), thin-Q projections
// (:176-179, :186-190), norm-ball projection (:10-26), Gamma clip (:202-203),
// convergence masks (:427-438), the L1 objective term (:440) and the adaptive
// lasso weights (:353).  All reductions use fixed-order trees.
#include "common.cuh"
#include <algorithm>
#include "kernels.h"
#include "devbuf.h"

namespace ca {

// ---- normalise + transpose -------------------------------------------------
__global__ void k_normalize_transpose(const double* __restrict__ Yraw, int n, int p, long ld_in,
                                      const double* __restrict__ sf, double* __restrict__ Yn, long ldYn,
                                      double* __restrict__ Ynt, long ldYnt) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, j = j0 + threadIdx.x;
    double v = 0.0;
    if (i < n && j < p) {
      v = Yraw[(size_t)i * ld_in + j];
      if (sf) v = v / sf[i];
      if (Yn) Yn[(size_t)i * ldYn + j] = v;
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  if (Ynt) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int j = j0 + r, i = i0 + threadIdx.x;
      if (j < p && i < n) Ynt[(size_t)j * ldYnt + i] = tile[threadIdx.x][r];
    }
  }
}

int launch_normalize_transpose(const double* Yraw, int n, int p, long ld_in, const double* sf, double* Yn,
                               long ldYn, double* Ynt, long ldYnt, cudaStream_t st) {
  dim3 grid((p + 31) / 32, (n + 31) / 32), block(32, 8);
  k_normalize_transpose<<<grid, block, 0, st>>>(Yraw, n, p, ld_in, sf, Yn, ldYn, Ynt, ldYnt);
  CA_LAUNCH_CHECK();
  return 0;
}

int launch_transpose(const double* src, int rows, int cols, long ld_in, double* dst, long ld_out, cudaStream_t st) {
  return launch_normalize_transpose(src, rows, cols, ld_in, nullptr, nullptr, 0, dst, ld_out, st);
}

// ---- row sums of the constant term Tys -------------------------------------
// Tys_ij = -lgamma(y + 1) [+ lgamma(nu + y) - lgamma(nu) for NB entries].  Counts are small integers, so
// the libm values are looked up instead of recomputed per entry (two lgamma calls per entry made these two
// launches 1.7 ms of a benchmark batch): lgamma(k + 1), k < TYS_K1, from a table every CTA fills with the
// same libm calls, and lgamma(nu + k) - lgamma(nu), k < TYS_K2, per gene -- from shared memory when the row
// is a gene, from a p x TYS_K2 table built by k_tys_nb_table when the columns are genes.  Entries outside
// the tables (or not integers) take the libm path, so every term is bit-identical to the direct evaluation.
constexpr int TYS_K1 = 256;
constexpr int TYS_K2 = 32;

__global__ void k_tys_nb_table(const double* __restrict__ nu, int p, double thres, double* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p * TYS_K2) return;
  const int c = i / TYS_K2, k = i - c * TYS_K2;
  const double nc = nu[c];
  tab[i] = !(nc > thres) ? lgamma(nc + (double)k) - lgamma(nc) : 0.0;
}

__global__ void k_tys_rowsum(const double* __restrict__ Yr, long ldY, int nrows, int ncols,
                             const double* __restrict__ nu, int nu_per_row, int family, double thres,
                             const double* __restrict__ nbtab, double* __restrict__ out) {
  __shared__ double scratch[32];
  __shared__ double s_lg1[TYS_K1];
  __shared__ double s_nb[TYS_K2];
  const int row = blockIdx.x;
  const double* y = Yr + (size_t)row * ldY;
  double nur = 1.0, lgr = 0.0;
  bool nb_row = false;
  if (family == CA_FAMILY_NB && nu_per_row) {
    nur = nu[row];
    nb_row = !(nur > thres);
    lgr = lgamma(nur);
  }
  for (int k = threadIdx.x; k < TYS_K1; k += blockDim.x) s_lg1[k] = lgamma((double)k + 1.0);
  if (nb_row)
    for (int k = threadIdx.x; k < TYS_K2; k += blockDim.x) s_nb[k] = lgamma(nur + (double)k) - lgr;
  __syncthreads();
  double s = 0.0;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    const double v = y[c];
    const int vi = (v >= 0.0 && v < (double)TYS_K1) ? (int)v : -1;
    const bool integral = vi >= 0 && (double)vi == v;
    double t = integral ? -s_lg1[vi] : -lgamma(v + 1.0);
    if (family == CA_FAMILY_NB) {
      if (nu_per_row) {
        if (nb_row) t += (integral && vi < TYS_K2) ? s_nb[vi] : lgamma(nur + v) - lgr;
      } else {
        const double nc = nu[c];
        if (!(nc > thres)) t += (integral && vi < TYS_K2) ? nbtab[(size_t)c * TYS_K2 + vi] : lgamma(nc + v) - lgamma(nc);
      }
    }
    s += t;
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[row] = s;
}

int launch_tys_rowsum(const double* Yr, long ldY, int nrows, int ncols, const double* nu, int nu_per_row,
                      int family, double thres, double* out, cudaStream_t st) {
  if (nrows <= 0) return 0;
  DevBuf<double> tab;
  if (family == CA_FAMILY_NB && !nu_per_row) {
    if (tab.alloc((size_t)ncols * TYS_K2)) return -1;
    k_tys_nb_table<<<(ncols * TYS_K2 + 255) / 256, 256, 0, st>>>(nu, ncols, thres, tab.ptr);
    CA_LAUNCH_CHECK();
  }
  k_tys_rowsum<<<nrows, 256, 0, st>>>(Yr, ldY, nrows, ncols, nu, nu_per_row, family, thres, tab.ptr, out);
  CA_LAUNCH_CHECK();
  if (tab.ptr) CA_CUDA(cudaStreamSynchronize(st));   // the table returns to the caching allocator with this scope
  return 0;
}

__global__ void k_rowsum(const double* __restrict__ M, long ld, int ncols, double* __restrict__ out) {
  __shared__ double scratch[32];
  const double* y = M + (size_t)blockIdx.x * ld;
  double s = 0.0;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) s += y[c];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}
int launch_rowsum(const double* M, long ld, int nrows, int ncols, double* out, cudaStream_t st) {
  if (nrows <= 0) return 0;
  k_rowsum<<<nrows, 256, 0, st>>>(M, ld, ncols, out);
  CA_LAUNCH_CHECK();
  return 0;
}

__global__ void k_zero_fraction(const double* __restrict__ Yr, long ldY, int ncols, double* __restrict__ out) {
  __shared__ double scratch[32];
  const double* y = Yr + (size_t)blockIdx.x * ldY;
  double s = 0.0;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) s += (y[c] == 0.0) ? 1.0 : 0.0;
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s / (double)ncols;
}
int launch_zero_fraction(const double* Yr, long ldY, int nrows, int ncols, double* out, cudaStream_t st) {
  if (nrows <= 0) return 0;
  k_zero_fraction<<<nrows, 256, 0, st>>>(Yr, ldY, ncols, out);
  CA_LAUNCH_CHECK();
  return 0;
}

__global__ void k_sum(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
}
int launch_sum(const double* v, int n, double* out, cudaStream_t st) {
  k_sum<<<1, 1024, 0, st>>>(v, n, out);
  CA_LAUNCH_CHECK();
  return 0;
}

__global__ void k_div_scalar(double* v, long n, double denom) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] / denom;
}
int launch_div_scalar(double* v, long n, double denom, cudaStream_t st) {
  if (n <= 0) return 0;
  k_div_scalar<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, denom);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- thin-Q projections ----------------------------------------------------
__global__ void k_pt_v(const double* __restrict__ P, int ldP, const double* __restrict__ V, int ldV, int cv,
                       int rows, double* __restrict__ W) {
  __shared__ double scratch[32];
  const int a = blockIdx.x / cv, b = blockIdx.x % cv;
  double s = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) s += P[(size_t)i * ldP + a] * V[(size_t)i * ldV + b];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) W[a * cv + b] = s;
}
int launch_pt_v(const double* P, int ldP, int cp, const double* V, int ldV, int cv, int rows, double* W,
                cudaStream_t st) {
  if (cp * cv <= 0) return 0;
  k_pt_v<<<cp * cv, 256, 0, st>>>(P, ldP, V, ldV, cv, rows, W);
  CA_LAUNCH_CHECK();
  return 0;
}

__global__ void k_sub_p_w(const double* __restrict__ P, int ldP, int cp, double* __restrict__ V, int ldV, int cv,
                          int rows, const double* __restrict__ W) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)rows * cv) return;
  const int i = (int)(e / cv), b = (int)(e % cv);
  double s = 0.0;
  for (int a = 0; a < cp; ++a) s += P[(size_t)i * ldP + a] * W[a * cv + b];
  V[(size_t)i * ldV + b] -= s;
}
int launch_sub_p_w(const double* P, int ldP, int cp, double* V, int ldV, int cv, int rows, const double* W,
                   cudaStream_t st) {
  const long tot = (long)rows * cv;
  if (tot <= 0) return 0;
  k_sub_p_w<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(P, ldP, cp, V, ldV, cv, rows, W);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- per-row line-search preparation ---------------------------------------
__global__ void k_prep_search(PrepArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = a.rowlist ? *a.count : a.nrows;
  if (idx >= total) return;
  const int row = a.rowlist ? a.rowlist[idx] : idx;
  double* g = a.g + (size_t)row * a.kpad;
  for (int c = 0; c < a.kpad; ++c) g[c] = 0.0;
  for (int c = 0; c < a.gc_n; ++c) g[a.gc_lo + c] = a.G[(size_t)row * a.ldG + c];
  // inf-norm ball projection on [ball_lo, ball_hi)  (gcate_opt.py:10-26)
  double mx = 0.0;
  for (int c = a.ball_lo; c < a.ball_hi; ++c) mx = fmax(mx, fabs(g[c]));
  if (mx > a.radius) {
    const double sc = a.radius / mx;
    for (int c = a.ball_lo; c < a.ball_hi; ++c) g[c] *= sc;
  }
  double ss = 0.0;
  for (int c = 0; c < a.kpad; ++c) ss += g[c] * g[c];
  a.normg[row] = sqrt(ss);
  double pen = 0.0;
  const int nprox = a.prox_hi - a.prox_lo;
  if (a.lam && nprox > 0) {
    const double* x = a.X + (size_t)row * a.kpad;
    for (int c = 0; c < nprox; ++c) pen += a.lam[(size_t)row * a.ldlam + c] * fabs(x[a.prox_lo + c]);
    pen /= (double)nprox;
  }
  a.f0[row] = -(a.ell0[row] + a.tys[row]) / (double)a.ncols + pen;
}
int launch_prep_search(const PrepArgs& a, cudaStream_t st) {
  if (a.nrows <= 0) return 0;
  k_prep_search<<<(a.nrows + 127) / 128, 128, 0, st>>>(a);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- gene-sharded cell line search (multi-GPU): candidates and decisions as separate steps ----
// With the genes split over ranks a cell's objective is a sum over ranks, so the walk of
// causarray/gcate_opt.py:60-82 becomes: build the candidate rows, evaluate the local partial
// log-likelihoods with the ordinary row pass, all-reduce them, decide per row.  Cells have no prox
// term (lam = 0, gcate_opt.py:173) and a common initial step.
__global__ void k_ls_candidates(LsShard a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = *a.count;
  if (idx >= total) return;
  const int row = a.rowlist[idx];
  const double t = a.round == 0 ? a.alpha : a.t[row];
  if (a.round == 0) {
    a.t[row] = t;
    a.ne[row] = 0;
  }
  const double* x0 = a.X + (size_t)row * a.kpad;
  const double* g = a.g + (size_t)row * a.kpad;
  double* xc = a.Xc + (size_t)row * a.kpad;
  for (int c = 0; c < a.kpad; ++c) xc[c] = x0[c] - t * g[c];
}
__global__ void k_ls_decide(LsShard a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = *a.count;
  if (idx >= total) return;
  const int row = a.rowlist[idx];
  const double ell = a.ell[row];                       // reduced over the ranks
  const double f1 = -(ell + a.tys[row]) / a.ncols_total;
  const double t = a.t[row];
  a.ne[row] += 1;
  if (a.round == 0) {
    a.f01[row] = f1;
    a.e01[row] = ell;
  }
  bool done = false, fallback = false;
  double efin = ell;
  if (f1 < a.f0[row] - a.tol * t * a.normg[row]) {
    done = true;
  } else if (t < 1e-4 || a.round == a.max_iters - 1) {
    done = true;
    if (a.f01[row] < 1.1 * f1) {
      fallback = true;
      efin = a.e01[row];
    }
  }
  if (done) {
    double* x = a.X + (size_t)row * a.kpad;
    const double* g = a.g + (size_t)row * a.kpad;
    const double* xc = a.Xc + (size_t)row * a.kpad;
    for (int c = 0; c < a.kpad; ++c) x[c] = fallback ? x[c] - a.alpha * g[c] : xc[c];
    a.ell_fin[row] = efin;
    a.pend[row] = 0;
    atomicAdd(a.eval_total, (unsigned long long)a.ne[row]);
  } else {
    a.t[row] = t * a.beta;
    a.pend[row] = 1;
    atomicAdd(a.n_pending, 1);
  }
}
int launch_ls_candidates(const LsShard& a, int nrows, cudaStream_t st) {
  if (nrows <= 0) return 0;
  k_ls_candidates<<<(nrows + 127) / 128, 128, 0, st>>>(a);
  CA_LAUNCH_CHECK();
  return 0;
}
int launch_ls_decide(const LsShard& a, int nrows, cudaStream_t st) {
  if (nrows <= 0) return 0;
  k_ls_decide<<<(nrows + 127) / 128, 128, 0, st>>>(a);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- mask compaction (deterministic, single CTA) ---------------------------
__global__ void k_compact(const uint8_t* __restrict__ mask, int n, int* __restrict__ list, int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const int flag = (i < n && mask[i]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int pre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (flag) list[off + pre] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}
int launch_compact(const uint8_t* mask, int n, int* list, int* count, cudaStream_t st) {
  k_compact<<<1, 1024, 0, st>>>(mask, n, list, count);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- mask compaction ordered by a small integer key (deterministic, single CTA) ----
// The gene line search walks its 8 rows of a tile in lockstep, so a tile costs as many row
// passes as its slowest row.  Listing the active rows by the number of candidates each needed in
// the previous alternating iteration (descending: long tiles are scheduled first) groups rows of
// similar cost; the per-row results do not depend on which tile a row sits in.
constexpr int CK_BINS = 16;
__global__ void __launch_bounds__(1024) k_compact_by_key(const uint8_t* __restrict__ mask, const int* __restrict__ key,
                                                         int n, int* __restrict__ list, int* __restrict__ count) {
  __shared__ int s_cnt[CK_BINS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = (n + 31) / 32;
  const int lo = warp * seg, hi = min(n, lo + seg);
  int cnt[CK_BINS];
#pragma unroll
  for (int b = 0; b < CK_BINS; ++b) cnt[b] = 0;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const int kb = (i < hi && mask[i]) ? min(max(key[i], 0), CK_BINS - 1) : -1;
#pragma unroll
    for (int b = 0; b < CK_BINS; ++b) cnt[b] += __popc(__ballot_sync(0xffffffffu, kb == b));
  }
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < CK_BINS; ++b) s_cnt[b][warp] = cnt[b];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = CK_BINS - 1; b >= 0; --b)
      for (int w = 0; w < 32; ++w) {
        const int c = s_cnt[b][w];
        s_cnt[b][w] = run;
        run += c;
      }
    *count = run;
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < CK_BINS; ++b) cnt[b] = s_cnt[b][warp];
  const unsigned lt = (1u << lane) - 1u;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const int kb = (i < hi && mask[i]) ? min(max(key[i], 0), CK_BINS - 1) : -1;
#pragma unroll
    for (int b = 0; b < CK_BINS; ++b) {
      const unsigned bal = __ballot_sync(0xffffffffu, kb == b);
      if (kb == b) list[cnt[b] + __popc(bal & lt)] = i;
      cnt[b] += __popc(bal);
    }
  }
}
int launch_compact_by_key(const uint8_t* mask, const int* key, int n, int* list, int* count, cudaStream_t st) {
  k_compact_by_key<<<1, 1024, 0, st>>>(mask, key, n, list, count);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- clip (records rows that changed) --------------------------------------
__global__ void k_clip(double* X, int rows, int kpad, int lo, int hi, double bound, int* list, int* count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  bool changed = false;
  double* x = X + (size_t)i * kpad;
  for (int c = lo; c < hi; ++c) {
    const double v = x[c];
    const double w = fmin(fmax(v, -bound), bound);
    if (w != v) {
      x[c] = w;
      changed = true;
    }
  }
  if (changed && list) list[atomicAdd(count, 1)] = i;
}
int launch_clip(double* X, int rows, int kpad, int lo, int hi, double bound, int* list, int* count, cudaStream_t st) {
  if (rows <= 0) return 0;
  if (count) CA_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
  k_clip<<<(rows + 127) / 128, 128, 0, st>>>(X, rows, kpad, lo, hi, bound, list, count);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- convergence masks -----------------------------------------------------
__global__ void k_step_mask(const double* __restrict__ X, const double* __restrict__ Xp, int rows, int kpad, int lo,
                            int hi, double tol, uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double ss = 0.0;
  for (int c = lo; c < hi; ++c) {
    const double dlt = X[(size_t)i * kpad + c] - Xp[(size_t)i * kpad + c];
    ss += dlt * dlt;
  }
  mask[i] = sqrt(ss) > tol ? 1 : 0;
}
__global__ void k_count_u8(const uint8_t* __restrict__ m, int n, int* __restrict__ count) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += m[i] ? 1.0 : 0.0;
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) *count = (int)(s + 0.5);
}
int launch_step_mask(const double* X, const double* Xprev, int rows, int kpad, int lo, int hi, double tol,
                     uint8_t* mask, int* count, cudaStream_t st) {
  if (rows <= 0) return 0;
  k_step_mask<<<(rows + 127) / 128, 128, 0, st>>>(X, Xprev, rows, kpad, lo, hi, tol, mask);
  CA_LAUNCH_CHECK();
  if (count) {
    k_count_u8<<<1, 1024, 0, st>>>(mask, rows, count);
    CA_LAUNCH_CHECK();
  }
  return 0;
}
__global__ void k_fill_u8(uint8_t* v, int n, uint8_t val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = val;
}
int launch_fill_u8(uint8_t* v, int n, uint8_t val, cudaStream_t st) {
  if (n <= 0) return 0;
  k_fill_u8<<<(n + 255) / 256, 256, 0, st>>>(v, n, val);
  CA_LAUNCH_CHECK();
  return 0;
}

// ---- L1 term and adaptive weights -------------------------------------------
// two deterministic steps: one partial per block of 64 rows, then the fixed-order sum of the partials
constexpr int L1_ROWS = 64;
__global__ void k_l1_penalty(const double* __restrict__ X, int rows, int kpad, int lo, int hi,
                             const double* __restrict__ lam, int ldlam, double* __restrict__ partial) {
  __shared__ double scratch[32];
  const int w = hi - lo;
  const int r0 = blockIdx.x * L1_ROWS, r1 = min(rows, r0 + L1_ROWS);
  double s = 0.0;
  for (int i = r0 + (threadIdx.x >> 4); i < r1; i += blockDim.x >> 4)
    for (int c = threadIdx.x & 15; c < w; c += 16)
      s += fabs(X[(size_t)i * kpad + lo + c]) * lam[(size_t)i * ldlam + c];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
int l1_penalty_blocks(int rows) { return (rows + L1_ROWS - 1) / L1_ROWS; }
int launch_l1_penalty(const double* X, int rows, int kpad, int lo, int hi, const double* lam, int ldlam,
                      double* partial, double* out, cudaStream_t st) {
  const int nb = l1_penalty_blocks(rows);
  k_l1_penalty<<<nb, 256, 0, st>>>(X, rows, kpad, lo, hi, lam, ldlam, partial);
  CA_LAUNCH_CHECK();
  return launch_sum(partial, nb, out, st);
}

__global__ void k_adaptive_weights(const double* __restrict__ X, int rows, int kpad, int lo, int hi, double lam0,
                                   double* __restrict__ lam, int ldlam) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = hi - lo;
  if (e >= (long)rows * w) return;
  const int i = (int)(e / w), c = (int)(e % w);
  lam[(size_t)i * ldlam + c] = lam0 / (fabs(X[(size_t)i * kpad + lo + c]) + 1e-6);
}
int launch_adaptive_weights(const double* X, int rows, int kpad, int lo, int hi, double lam0, double* lam, int ldlam,
                            cudaStream_t st) {
  const long tot = (long)rows * (hi - lo);
  if (tot <= 0) return 0;
  k_adaptive_weights<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(X, rows, kpad, lo, hi, lam0, lam, ldlam);
  CA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ca

// ---------------------------------------------------------------------------
// Input checks and median-of-ratios size factors on the device
// (causarray/gcate.py:12-17 and causarray/utils.py:179-219).
// ---------------------------------------------------------------------------
namespace ca {

template <typename T>
__global__ void k_widen_counts(const T* __restrict__ src, long count, double* __restrict__ dst) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = (double)src[i];
}

int launch_widen_counts(const void* src, int dtype, long count, double* dst, cudaStream_t st) {
  if (count <= 0) return 0;
  const unsigned blocks = (unsigned)std::min<long>((count + 255) / 256, 148L * 16);
  switch (dtype) {
    case 1: k_widen_counts<<<blocks, 256, 0, st>>>((const float*)src, count, dst); break;
    case 2: k_widen_counts<<<blocks, 256, 0, st>>>((const int*)src, count, dst); break;
    case 3: k_widen_counts<<<blocks, 256, 0, st>>>((const long long*)src, count, dst); break;
    case 4: k_widen_counts<<<blocks, 256, 0, st>>>((const unsigned short*)src, count, dst); break;
    case 5: k_widen_counts<<<blocks, 256, 0, st>>>((const unsigned char*)src, count, dst); break;
    default: set_error("launch_widen_counts: element type %d", dtype); return -1;
  }
  CA_LAUNCH_CHECK();
  return 0;
}

// CSR rows -> dense row-major counts (the destination is zero-filled by the caller); one warp per
// row, entries with a column index outside [0, p) are ignored
__global__ void k_csr_expand(const long long* __restrict__ indptr, const int* __restrict__ indices,
                             const double* __restrict__ data, int n, int p, double* __restrict__ Y) {
  const int row = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const long long lo = indptr[row], hi = indptr[row + 1];
  for (long long e = lo + lane; e < hi; e += 32) {
    const int c = indices[e];
    if (c >= 0 && c < p) Y[(size_t)row * p + c] = data[e];
  }
}
int launch_csr_expand(const long long* indptr, const int* indices, const double* data, int n, int p, double* Y,
                      cudaStream_t st) {
  if (n <= 0) return 0;
  const long threads = (long)n * 32;
  k_csr_expand<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(indptr, indices, data, n, p, Y);
  CA_LAUNCH_CHECK();
  return 0;
}

// flags[0] += #non-integer entries (np.allclose(Y, trunc(Y)) tolerance), flags[1] += #non-finite
__global__ void k_check_counts(const double* __restrict__ Y, long total, unsigned long long* flags) {
  unsigned long long nonint = 0, nonfin = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const double v = Y[i];
    if (!isfinite(v)) {
      ++nonfin;
    } else {
      const double t = trunc(v);
      if (fabs(v - t) > 1e-8 + 1e-5 * fabs(t)) ++nonint;
    }
  }
  if (nonint) atomicAdd(flags + 0, nonint);
  if (nonfin) atomicAdd(flags + 1, nonfin);
}
int launch_check_counts(const double* Y, long total, unsigned long long* flags, cudaStream_t st) {
  CA_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(unsigned long long), st));
  k_check_counts<<<592, 256, 0, st>>>(Y, total, flags);
  CA_LAUNCH_CHECK();
  return 0;
}

// per gene (row of Y^T): mean of log(y) over y > 0, -inf if the gene is empty
__global__ void k_log_geomean(const double* __restrict__ Yt, long ldn, int n, double* __restrict__ out) {
  __shared__ double scratch[32];
  const double* y = Yt + (size_t)blockIdx.x * ldn;
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = y[i];
    if (v > 0.0) {
      s += log(v);
      c += 1.0;
    }
  }
  s = block_sum(s, scratch);
  c = block_sum(c, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = c > 0.0 ? s / c : -INFINITY;
}

__device__ __forceinline__ unsigned long long key_of(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double val_of(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// One CTA per cell: median over expressed genes of log(y) - log_geomean (radix select).
__global__ void __launch_bounds__(256) k_row_median(const double* __restrict__ Y, long ldY, int p,
                                                     const double* __restrict__ lg,
                                                     unsigned long long* __restrict__ scratch_keys,
                                                     double* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned int s_k, s_m;
  __shared__ unsigned long long s_min;
  const int row = blockIdx.x, tid = threadIdx.x;
  const double* y = Y + (size_t)row * ldY;
  unsigned long long* keys = scratch_keys + (size_t)row * p;
  const unsigned long long INVALID = 0xffffffffffffffffull;
  if (tid == 0) s_m = 0;
  __syncthreads();
  unsigned int cnt = 0;
  for (int j = tid; j < p; j += 256) {
    const double v = y[j], g = lg[j];
    unsigned long long k = INVALID;
    if (v > 0.0 && isfinite(g)) {
      k = key_of(log(v) - g);
      ++cnt;
    }
    keys[j] = k;
  }
  atomicAdd(&s_m, cnt);
  __syncthreads();
  const unsigned int m = s_m;
  if (m == 0) {
    if (tid == 0) out[row] = NAN;
    return;
  }
  double med[2];
  const unsigned int ranks[2] = {(m - 1) / 2, m / 2};
  for (int which = 0; which < 2; ++which) {
    if (which == 1 && ranks[1] == ranks[0]) {
      med[1] = med[0];
      break;
    }
    if (tid == 0) {
      s_prefix = 0;
      s_k = ranks[which];
    }
    __syncthreads();
    for (int pass = 7; pass >= 0; --pass) {
      hist[tid] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long himask = (pass == 7) ? 0ull : (~0ull << (8 * (pass + 1)));
      for (int j = tid; j < p; j += 256) {
        const unsigned long long k = keys[j];
        if (k != INVALID && (k & himask) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 0xff], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned int kk = s_k, b = 0;
        for (; b < 256; ++b) {
          if (kk < hist[b]) break;
          kk -= hist[b];
        }
        s_k = kk;
        s_prefix = prefix | ((unsigned long long)b << (8 * pass));
      }
      __syncthreads();
    }
    med[which] = val_of(s_prefix);
    __syncthreads();
  }
  (void)s_min;
  if (tid == 0) out[row] = (ranks[0] == ranks[1]) ? med[0] : 0.5 * (med[0] + med[1]);
}

int launch_size_factor_medians(const double* Yraw, long ldY, const double* Yraw_t, long ldn, int n, int p,
                               double* log_geo /*p*/, unsigned long long* scratch /*n*p*/, double* out /*n*/,
                               cudaStream_t st) {
  k_log_geomean<<<p, 256, 0, st>>>(Yraw_t, ldn, n, log_geo);
  CA_LAUNCH_CHECK();
  k_row_median<<<n, 256, 0, st>>>(Yraw, ldY, p, log_geo, scratch, out);
  CA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ca
