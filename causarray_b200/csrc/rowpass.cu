// K1 / K2 of the GCATE alternating minimisation:
//
//   K1  row pass  : for a tile of 8 "rows" (cells in the A-step, genes in the
//                   B-step) compute Theta = X_rows * M^T on the FP64 tensor
//                   cores (DMMA m8n8k4), and in the epilogue of every 8x8
//                   accumulator tile evaluate the Poisson / NB log-likelihood
//                   terms, accumulate the row log-likelihood sums and -- in the
//                   gradient variant -- feed d ell/d theta straight back into a
//                   second DMMA that contracts it with M (the gradient block).
//                   Theta is never written to memory.
//   K2  line search: one CTA owns a tile of 8 rows and walks the backtracking
//                   candidates t = alpha * beta^j sequentially (the accept /
//                   fallback rule of causarray/gcate_opt.py:30-82), one row
//                   pass per candidate, rows that have accepted drop out.
//
// Replaces: nll / grad (causarray/gcate_likelihood.py:46-142) and line_search
// as driven from update_with_mask (causarray/gcate_opt.py:168-175, 182-200).
//
// Layout: Y is (rows x ldY) row-major for the entity being updated (Y for
// cells, Y^T for genes), M is (ncols x kpad) row-major with kpad % 8 == 4 so
// that the B-fragment loads of a half-warp hit 16 distinct bank pairs.
#include "common.cuh"
#include "kernels.h"
#include "prof.h"
#include "fastmath.cuh"

namespace ca {

constexpr int TILE_ROWS = 8;
constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;

struct PassCfg {
  const double* Y;
  long ldY;
  int ncols;
  const double* M;
  int kpad;
  const double* nu;
  const double* inv_nu;  // 1 / nu
  const double* log_nu;  // log(nu)
  int nu_per_row;
  int family;
  double thres;
  int tc;  // columns per staged chunk (multiple of 8)
  const double* fm_tables;
  NbClip clip;
};

__device__ __forceinline__ void stage_chunk(const PassCfg& P, int c, double* dst) {
  const int col0 = c * P.tc;
  const int nvalid = min(P.tc, P.ncols - col0);
  const int nvec = nvalid * P.kpad / 2;  // 16-byte units
  const double* src = P.M + (size_t)col0 * P.kpad;
  for (int i = threadIdx.x; i < nvec; i += NTHREADS) cp_async16(dst + 2 * i, src + 2 * i);
  // zero-fill the ragged tail of the last chunk up to the next multiple of 8 rows
  const int npad = min(P.tc, round_up(nvalid, 8));
  for (int i = nvalid * P.kpad + threadIdx.x; i < npad * P.kpad; i += NTHREADS) dst[i] = 0.0;
}

// One pass over all columns for the 8 rows listed in s_rows (-1 = empty slot).
// s_X: 8 x kpad candidate rows.  On return (after a __syncthreads) s_rowsum[r]
// holds sum_col ell(y, theta) and, if GRAD, s_G[r*NGT*8 + c] holds
// sum_col dl * M[col, gc_lo + c].
template <bool GRAD, int NGT>
__device__ void tile_pass(const PassCfg& P, const FmTables& FT, const int* s_rows, const double* s_X, double* s_M,
                          double* s_red, double* s_Gred, int gc_lo, const unsigned char* s_skip,
                          double* s_rowsum, double* s_G) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int kpad = P.kpad, KS = kpad >> 2, tc = P.tc;
  const int row = s_rows[lr];
  const bool rowok = (row >= 0) && !(s_skip && s_skip[lr]);
  const double* yrow = P.Y + (size_t)(row >= 0 ? row : 0) * P.ldY;
  const bool row_nu = P.nu_per_row && row >= 0;
  const double nu_r = row_nu ? P.nu[row] : 1.0;
  const double inu_r = row_nu ? P.inv_nu[row] : 1.0;
  const double lnu_r = row_nu ? P.log_nu[row] : 0.0;
  const bool pois_r = (P.family == CA_FAMILY_POISSON) || (P.nu_per_row && nu_r > P.thres);
  const bool fam_nb = P.family == CA_FAMILY_NB;
  double ellsum = 0.0;
  double gacc[NGT][2];
#pragma unroll
  for (int j = 0; j < NGT; ++j) gacc[j][0] = gacc[j][1] = 0.0;

  const int nchunks = (P.ncols + tc - 1) / tc;
  const int chunk_elems = tc * kpad;
  stage_chunk(P, 0, s_M);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      stage_chunk(P, c + 1, s_M + ((c + 1) & 1) * chunk_elems);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Mb = s_M + (c & 1) * chunk_elems;
    const int col0 = c * tc;
    const int ntile = min(tc, round_up(P.ncols - col0, 8)) >> 3;
    for (int t = warp; t < ntile; t += NWARPS) {
      const int col = col0 + t * 8 + 2 * lc;
      double2 yv = make_double2(0.0, 0.0);
      if (rowok) yv = *reinterpret_cast<const double2*>(yrow + col);
      double acc0 = 0.0, acc1 = 0.0;
      const double* bp = Mb + (t * 8 + lr) * kpad + lc;
      const double* ap = s_X + lr * kpad + lc;
#pragma unroll 4
      for (int ks = 0; ks < KS; ++ks) dmma884(acc0, acc1, ap[ks * 4], bp[ks * 4]);
      const bool v0 = rowok && (col < P.ncols), v1 = rowok && (col + 1 < P.ncols);
      double e0, e1, d0 = 0.0, d1 = 0.0;
      if (fam_nb) {
        double nu0 = nu_r, nu1 = nu_r, in0 = inu_r, in1 = inu_r, ln0 = lnu_r, ln1 = lnu_r;
        bool p0 = pois_r, p1 = pois_r;
        if (!P.nu_per_row) {
          const int c0 = min(col, P.ncols - 1), c1 = min(col + 1, P.ncols - 1);
          nu0 = P.nu[c0];
          nu1 = P.nu[c1];
          in0 = P.inv_nu[c0];
          in1 = P.inv_nu[c1];
          ln0 = P.log_nu[c0];
          ln1 = P.log_nu[c1];
          p0 = nu0 > P.thres;
          p1 = nu1 > P.thres;
        }
        fm_entry<GRAD>(acc0, yv.x, nu0, in0, ln0, p0, FT, P.clip, e0, d0);
        fm_entry<GRAD>(acc1, yv.y, nu1, in1, ln1, p1, FT, P.clip, e1, d1);
      } else {
        const double x0 = fmin(acc0, 100.0), x1 = fmin(acc1, 100.0);
        const double m0 = fm_exp(fmax(x0, -700.0), FT.exp_t), m1 = fm_exp(fmax(x1, -700.0), FT.exp_t);
        e0 = fma(yv.x, x0, -m0);
        e1 = fma(yv.y, x1, -m1);
        if (GRAD) {
          d0 = m0 - yv.x;
          d1 = m1 - yv.y;
        }
      }
      ellsum += (v0 ? e0 : 0.0) + (v1 ? e1 : 0.0);
      if (GRAD) {
        if (!v0) d0 = 0.0;
        if (!v1) d1 = 0.0;
        const double* g0 = Mb + (t * 8 + 2 * lc) * kpad + gc_lo + lr;
#pragma unroll
        for (int j = 0; j < NGT; ++j) {
          dmma884(gacc[j][0], gacc[j][1], d0, g0[8 * j]);
          dmma884(gacc[j][0], gacc[j][1], d1, g0[kpad + 8 * j]);
        }
      }
    }
    __syncthreads();
  }
  // fixed-order reduction: 4 lanes of a row, then the 8 warps
  ellsum += __shfl_xor_sync(0xffffffffu, ellsum, 1);
  ellsum += __shfl_xor_sync(0xffffffffu, ellsum, 2);
  if (lc == 0) s_red[warp * 8 + lr] = ellsum;
  if (GRAD) {
#pragma unroll
    for (int j = 0; j < NGT; ++j) {
      s_Gred[(warp * 8 + lr) * (NGT * 8) + 8 * j + 2 * lc] = gacc[j][0];
      s_Gred[(warp * 8 + lr) * (NGT * 8) + 8 * j + 2 * lc + 1] = gacc[j][1];
    }
  }
  __syncthreads();
  if (tid < 8) {
    double s = 0.0;
    for (int w = 0; w < NWARPS; ++w) s += s_red[w * 8 + tid];
    s_rowsum[tid] = s;
  }
  if (GRAD) {
    for (int e = tid; e < 8 * NGT * 8; e += NTHREADS) {
      const int r = e / (NGT * 8), cc = e % (NGT * 8);
      double s = 0.0;
      for (int w = 0; w < NWARPS; ++w) s += s_Gred[(w * 8 + r) * (NGT * 8) + cc];
      s_G[e] = s;
    }
  }
  __syncthreads();
}

struct Smem {
  double* tab;    // FM_TABLE_DOUBLES
  double* M;      // 2 * tc * kpad + 64 slack
  double* X;      // 8 * kpad
  double* X0;     // 8 * kpad
  double* Gd;     // 8 * kpad
  double* red;    // 64
  double* rowsum; // 8
  double* Gred;   // 64 * NGT*8
  double* G;      // 8 * NGT*8
};

__device__ __forceinline__ Smem carve(double* base, int tc, int kpad, int ngt8, bool search) {
  Smem s;
  s.tab = base;
  base += FM_TABLE_DOUBLES;
  s.M = base;
  base += 2 * tc * kpad + 64;
  s.X = base;
  base += 8 * kpad;
  s.X0 = base;
  s.Gd = base + 8 * kpad;
  if (search) base += 16 * kpad;
  s.red = base;
  base += 64;
  s.rowsum = base;
  base += 8;
  s.Gred = base;
  base += 64 * ngt8;
  s.G = base;
  return s;
}

size_t rowpass_smem_bytes(int tc, int kpad, int ngt, bool search) {
  size_t n = FM_TABLE_DOUBLES + 2 * (size_t)tc * kpad + 64 + 8 * kpad + (search ? 16 * kpad : 0) + 64 + 8 + 64 * ngt * 8 + 8 * ngt * 8;
  return n * sizeof(double);
}

// ---------------------------------------------------------------------------
// Gradient / log-likelihood pass over a row list.
//   rowlist == nullptr -> rows 0..nrows-1, else rowlist[0..*count)
//   ell_out[row]                      = sum_col ell
//   G_out[row * ldG + c], c < ngt*8   = sum_col dl * M[col, gc_lo + c]   (GRAD)
// ---------------------------------------------------------------------------
template <bool GRAD, int NGT>
__global__ void __launch_bounds__(NTHREADS) k_rowpass(PassCfg P, const double* X, const int* rowlist,
                                                       const int* count, int nrows, int gc_lo,
                                                       double* ell_out, double* G_out, int ldG) {
  extern __shared__ __align__(16) double smem[];
  Smem S = carve(smem, P.tc, P.kpad, NGT * 8, false);
  __shared__ int s_rows[8];
  const int total = rowlist ? *count : nrows;
  const int tile0 = blockIdx.x * TILE_ROWS;
  if (tile0 >= total) return;
  const FmTables FT = fm_load_tables(S.tab, P.fm_tables);
  if (threadIdx.x < 8) {
    const int i = tile0 + threadIdx.x;
    s_rows[threadIdx.x] = (i < total) ? (rowlist ? rowlist[i] : i) : -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 8 * P.kpad; e += NTHREADS) {
    const int r = e / P.kpad, c = e % P.kpad;
    const int row = s_rows[r];
    S.X[e] = (row >= 0) ? X[(size_t)row * P.kpad + c] : 0.0;
  }
  __syncthreads();
  tile_pass<GRAD, NGT>(P, FT, s_rows, S.X, S.M, S.red, S.Gred, gc_lo, nullptr, S.rowsum, S.G);
  if (threadIdx.x < 8 && s_rows[threadIdx.x] >= 0) ell_out[s_rows[threadIdx.x]] = S.rowsum[threadIdx.x];
  if (GRAD) {
    for (int e = threadIdx.x; e < 8 * NGT * 8; e += NTHREADS) {
      const int r = e / (NGT * 8), cc = e % (NGT * 8);
      if (s_rows[r] >= 0 && cc < ldG) G_out[(size_t)s_rows[r] * ldG + cc] = S.G[e];
    }
  }
}

// ---------------------------------------------------------------------------
// Backtracking line search for a row list (causarray/gcate_opt.py:30-82).
//   g[row, 0..kpad)   projected gradient (zeros outside the moving block)
//   f0[row], normg[row] from k_prep_search
//   prox on columns [prox_lo, prox_hi) with thresholds lam[row*ldlam + c-prox_lo]
//   tys[row] = sum_col Tys (constant term of the row objective)
// On exit X[row] is overwritten with the accepted point, ell_fin[row] holds the
// raw log-likelihood sum at that point and neval[row] the number of candidate
// evaluations.
// ---------------------------------------------------------------------------
struct SearchArgs {
  double* X;
  const double* g;
  const double* f0;
  const double* normg;
  const double* tys;
  const double* lam;
  int ldlam;
  int prox_lo, prox_hi;
  const double* alpha_row;  // nullptr -> alpha
  double alpha, beta, tol;
  int max_iters;
  const int* rowlist;
  const int* count;
  int nrows;
  double* ell_fin;
  int* neval;
  unsigned long long* eval_total;
};

__device__ __forceinline__ double soft_thr(double x, double lam) {
  const double a = fabs(x) - lam;
  const double m = a > 0.0 ? a : 0.0;
  return (x > 0.0) ? m : ((x < 0.0) ? -m : 0.0);
}

__global__ void __launch_bounds__(NTHREADS) k_search(PassCfg P, SearchArgs A) {
  extern __shared__ __align__(16) double smem[];
  Smem S = carve(smem, P.tc, P.kpad, 8, true);
  __shared__ int s_rows[8];
  __shared__ double s_t[8], s_f0[8], s_ng[8], s_f01[8], s_e01[8], s_alpha[8], s_efin[8];
  __shared__ unsigned char s_done[8];
  __shared__ int s_mode[8], s_ne[8], s_alldone;
  const int tid = threadIdx.x;
  const int total = A.rowlist ? *A.count : A.nrows;
  const int tile0 = blockIdx.x * TILE_ROWS;
  if (tile0 >= total) return;
  const FmTables FT = fm_load_tables(S.tab, P.fm_tables);
  const int kpad = P.kpad;
  if (tid < 8) {
    const int i = tile0 + tid;
    const int row = (i < total) ? (A.rowlist ? A.rowlist[i] : i) : -1;
    s_rows[tid] = row;
    s_done[tid] = row < 0;
    s_mode[tid] = 0;
    s_ne[tid] = 0;
    if (row >= 0) {
      s_alpha[tid] = A.alpha_row ? A.alpha_row[row] : A.alpha;
      s_t[tid] = s_alpha[tid];
      s_f0[tid] = A.f0[row];
      s_ng[tid] = A.normg[row];
    }
  }
  __syncthreads();
  for (int e = tid; e < 8 * kpad; e += NTHREADS) {
    const int r = e / kpad, c = e % kpad, row = s_rows[r];
    S.X0[e] = (row >= 0) ? A.X[(size_t)row * kpad + c] : 0.0;
    S.Gd[e] = (row >= 0) ? A.g[(size_t)row * kpad + c] : 0.0;
    S.X[e] = 0.0;
  }
  __syncthreads();
  const double ncols_d = (double)P.ncols;
  const int nprox = A.prox_hi - A.prox_lo;
  for (int it = 0; it < A.max_iters; ++it) {
    for (int e = tid; e < 8 * kpad; e += NTHREADS) {
      const int r = e / kpad, c = e % kpad;
      if (!s_done[r]) {
        double v = S.X0[e] - s_t[r] * S.Gd[e];
        if (A.lam && c >= A.prox_lo && c < A.prox_hi)
          v = soft_thr(v, A.lam[(size_t)s_rows[r] * A.ldlam + (c - A.prox_lo)]);
        S.X[e] = v;
      }
    }
    __syncthreads();
    tile_pass<false, 1>(P, FT, s_rows, S.X, S.M, S.red, S.Gred, 0, s_done, S.rowsum, S.G);
    if (tid < 8 && !s_done[tid]) {
      const int row = s_rows[tid];
      double pen = 0.0;
      if (A.lam && nprox > 0) {
        for (int c = 0; c < nprox; ++c)
          pen += A.lam[(size_t)row * A.ldlam + c] * fabs(S.X[tid * kpad + A.prox_lo + c]);
        pen /= (double)nprox;
      }
      const double ell = S.rowsum[tid];
      const double f1 = -(ell + A.tys[row]) / ncols_d + pen;
      s_ne[tid] += 1;
      if (it == 0) {
        s_f01[tid] = f1;
        s_e01[tid] = ell;
      }
      const double t = s_t[tid];
      if (f1 < s_f0[tid] - A.tol * t * s_ng[tid]) {
        s_done[tid] = 1;
        s_efin[tid] = ell;
      } else if (t < 1e-4 || it == A.max_iters - 1) {
        s_done[tid] = 1;
        if (s_f01[tid] < 1.1 * f1) {
          s_mode[tid] = 1;  // fall back to the full initial step
          s_efin[tid] = s_e01[tid];
        } else {
          s_efin[tid] = ell;
        }
      } else {
        s_t[tid] = t * A.beta;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int all = 1;
      for (int r = 0; r < 8; ++r) all &= (int)s_done[r];
      s_alldone = all;
    }
    __syncthreads();
    if (s_alldone) break;
  }
  for (int e = tid; e < 8 * kpad; e += NTHREADS) {
    const int r = e / kpad, c = e % kpad, row = s_rows[r];
    if (row < 0) continue;
    double v = S.X[e];
    if (s_mode[r] == 1) {
      v = S.X0[e] - s_alpha[r] * S.Gd[e];
      if (A.lam && c >= A.prox_lo && c < A.prox_hi) v = soft_thr(v, A.lam[(size_t)row * A.ldlam + (c - A.prox_lo)]);
    }
    A.X[(size_t)row * kpad + c] = v;
  }
  if (tid < 8 && s_rows[tid] >= 0) {
    A.ell_fin[s_rows[tid]] = s_efin[tid];
    if (A.neval) A.neval[s_rows[tid]] = s_ne[tid];
    if (A.eval_total) atomicAdd(A.eval_total, (unsigned long long)s_ne[tid]);
  }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static int pick_tc(int kpad) {
  // two staged chunks of tc x kpad doubles; keep them under ~64 KB so that
  // three CTAs fit on an SM for kpad <= 28.
  int tc = 128;
  while (tc > 16 && 2 * (size_t)tc * kpad * 8 > 64 * 1024) tc >>= 1;
  return tc;
}

static PassCfg make_cfg(const RowProblem& rp) {
  PassCfg P;
  P.Y = rp.Y;
  P.ldY = rp.ldY;
  P.ncols = rp.ncols;
  P.M = rp.M;
  P.kpad = rp.kpad;
  P.nu = rp.nu;
  P.inv_nu = rp.inv_nu;
  P.log_nu = rp.log_nu;
  P.fm_tables = fastmath_tables();
  P.clip = nb_clip_constants();
  P.nu_per_row = rp.nu_per_row;
  P.family = rp.family;
  P.thres = rp.thres_disp;
  P.tc = pick_tc(rp.kpad);
  return P;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  CA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int launch_rowpass(const RowProblem& rp, const double* X, const int* rowlist, const int* count, int nrows,
                   bool want_grad, int gc_lo, int gc_n, double* ell_out, double* G_out, int ldG,
                   cudaStream_t st) {
  CA_REQUIRE(rp.kpad % 8 == 4, "kpad must be congruent to 4 mod 8");
  CA_REQUIRE(rp.ldY % 8 == 0, "ldY must be a multiple of 8");
  if (nrows <= 0) return 0;
  PassCfg P = make_cfg(rp);
  CA_REQUIRE(P.fm_tables, "fastmath tables unavailable");
  const int grid = (nrows + TILE_ROWS - 1) / TILE_ROWS;
  ProfScope prof(want_grad ? PROF_ROWPASS_GRAD : PROF_ROWPASS_ELL, st);
  if (!want_grad) {
    size_t sm = rowpass_smem_bytes(P.tc, P.kpad, 1, false);
    if (set_smem(k_rowpass<false, 1>, sm)) return -1;
    k_rowpass<false, 1><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, 0, ell_out, nullptr, 0);
  } else {
    CA_REQUIRE(gc_n <= 64, "gradient block wider than 64 columns");
    CA_REQUIRE(ldG >= gc_n, "ldG too small");
    if (gc_n <= 16) {
      size_t sm = rowpass_smem_bytes(P.tc, P.kpad, 2, false);
      if (set_smem(k_rowpass<true, 2>, sm)) return -1;
      k_rowpass<true, 2><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
    } else if (gc_n <= 32) {
      size_t sm = rowpass_smem_bytes(P.tc, P.kpad, 4, false);
      if (set_smem(k_rowpass<true, 4>, sm)) return -1;
      k_rowpass<true, 4><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
    } else {
      size_t sm = rowpass_smem_bytes(P.tc, P.kpad, 8, false);
      if (set_smem(k_rowpass<true, 8>, sm)) return -1;
      k_rowpass<true, 8><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
    }
  }
  CA_LAUNCH_CHECK();
  return 0;
}

int launch_search(const RowProblem& rp, const SearchLaunch& s, cudaStream_t st) {
  CA_REQUIRE(rp.kpad % 8 == 4, "kpad must be congruent to 4 mod 8");
  if (s.nrows <= 0) return 0;
  PassCfg P = make_cfg(rp);
  CA_REQUIRE(P.fm_tables, "fastmath tables unavailable");
  SearchArgs A;
  A.X = s.X;
  A.g = s.g;
  A.f0 = s.f0;
  A.normg = s.normg;
  A.tys = s.tys;
  A.lam = s.lam;
  A.ldlam = s.ldlam;
  A.prox_lo = s.prox_lo;
  A.prox_hi = s.prox_hi;
  A.alpha_row = s.alpha_row;
  A.alpha = s.alpha;
  A.beta = s.beta;
  A.tol = s.tol;
  A.max_iters = s.max_iters;
  A.rowlist = s.rowlist;
  A.count = s.count;
  A.nrows = s.nrows;
  A.ell_fin = s.ell_fin;
  A.neval = s.neval;
  A.eval_total = s.eval_total;
  size_t sm = rowpass_smem_bytes(P.tc, P.kpad, 1, true);
  if (set_smem(k_search, sm)) return -1;
  const int grid = (s.nrows + TILE_ROWS - 1) / TILE_ROWS;
  ProfScope prof(PROF_SEARCH, st);
  k_search<<<grid, NTHREADS, sm, st>>>(P, A);
  CA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ca
