// K1 / K2 of the GCATE alternating minimisation, templated on
// KS = kpad / 4 (number of k-steps of the DMMA chain) so that the hot loop is fully
// unrolled with immediate shared-memory offsets.
//
//   K1  row pass  : for a tile of 8 "rows" (cells in the A-step, genes in the B-step)
//                   Theta = X_rows * M^T on the FP64 tensor cores (DMMA m8n8k4); the
//                   epilogue of every 8x8 accumulator tile evaluates the Poisson / NB
//                   log-likelihood terms (table-driven exp / log, fastmath.cuh),
//                   accumulates the row sums and -- in the gradient variant -- feeds
//                   d ell/d theta straight back into a second DMMA that contracts it with
//                   M (the gradient block).  Theta is never written to memory.
//   K2  line search: one CTA owns a tile of 8 rows and walks the backtracking candidates
//                   t = alpha * beta^j sequentially (the accept / fallback rule of
//                   causarray/gcate_opt.py:30-82), one row pass per candidate.
//
// Replaces: nll / grad (causarray/gcate_likelihood.py:46-142) and line_search as driven
// from update_with_mask (causarray/gcate_opt.py:168-175, 182-200).
//
// Layout: Y is (rows x ldY) row-major for the entity being updated (Y for cells, Y^T for
// genes), M is (ncols x kpad) row-major with kpad % 8 == 4 so that the B-fragment loads
// of a half-warp hit 16 distinct bank pairs; per-column NB parameters are packed as
// (nu, 1/nu, log nu, poisson flag) and staged with the M chunk.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.h"

namespace ca {

constexpr int TILE_ROWS = 8;
constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;
#ifndef CA_MINB
#define CA_MINB 2
#endif
#ifndef CA_TC
#define CA_TC 128
#endif
#ifndef CA_ABL
#define CA_ABL 0
#endif

struct PassCfg {
  const double* Y;
  long ldY;
  int ncols;
  const double* M;
  const double* nu;       // per row (nu_per_row) : nu, inv_nu, log_nu indexed by row
  const double* inv_nu;
  const double* log_nu;
  const double4* colparm; // per column (!nu_per_row): (nu, 1/nu, log nu, pois)
  int nu_per_row;
  int family;
  double thres;
  int tc;  // columns per staged chunk (multiple of 8)
  const double* fm_tables;
  NbClip clip;
};

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

template <int KS>
struct Lay {
  static constexpr int KPAD = 4 * KS;
};

// shared-memory carve-up (doubles)
struct Smem {
  double* tab;     // FM_TABLE_DOUBLES
  double* M;       // 2 * tc * kpad + 64 slack
  double4* cp;     // 2 * tc   (column parameters)
  double* X;       // 8 * kpad
  double* X0;      // 8 * kpad (search)
  double* Gd;      // 8 * kpad (search)
  double* red;     // 64
  double* rowsum;  // 8
  double* Gred;    // 64 * NGT*8
  double* G;       // 8 * NGT*8
};

__host__ __device__ inline size_t smem_doubles(int tc, int kpad, int ngt, bool search) {
  return FM_TABLE_DOUBLES + 2 * (size_t)tc * kpad + 64 + 2 * (size_t)tc * 4 + 8 * kpad + (search ? 16 * kpad : 0) + 64 +
         8 + 64 * (size_t)ngt * 8 + 8 * (size_t)ngt * 8;
}

__device__ __forceinline__ Smem carve(double* base, int tc, int kpad, int ngt8, bool search) {
  Smem s;
  s.tab = base;
  base += FM_TABLE_DOUBLES;
  s.M = base;
  base += 2 * tc * kpad + 64;
  s.cp = reinterpret_cast<double4*>(base);
  base += 2 * tc * 4;
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

template <int KS>
__device__ __forceinline__ void stage_chunk(const PassCfg& P, int c, double* dstM, double4* dstC) {
  constexpr int KPAD = 4 * KS;
  const int col0 = c * P.tc;
  const int nvalid = min(P.tc, P.ncols - col0);
  const int nvec = nvalid * KPAD / 2;  // 16-byte units
  const double* src = P.M + (size_t)col0 * KPAD;
#if CA_ABL == 2
  if (c > 1) return;
#endif
  for (int i = threadIdx.x; i < nvec; i += NTHREADS) cp_async16(dstM + 2 * i, src + 2 * i);
  if (P.colparm) {
    const double* sc = reinterpret_cast<const double*>(P.colparm + col0);
    double* dc = reinterpret_cast<double*>(dstC);
    for (int i = threadIdx.x; i < nvalid * 2; i += NTHREADS) cp_async16(dc + 2 * i, sc + 2 * i);
  }
  // ragged tail of the last chunk: zero rows of M and neutral column parameters
  const int npad = min(P.tc, round_up(nvalid, 8));
  for (int i = nvalid * KPAD + threadIdx.x; i < npad * KPAD; i += NTHREADS) dstM[i] = 0.0;
  if (P.colparm)
    for (int i = nvalid + threadIdx.x; i < npad; i += NTHREADS) dstC[i] = make_double4(1.0, 1.0, 0.0, 0.0);
}

// One pass over all columns for the 8 rows listed in s_rows (-1 = empty slot; its sums are
// garbage and ignored by the callers).  s_X: 8 x kpad candidate rows.  On return (after a
// __syncthreads) s_rowsum[r] holds sum_col ell(y, theta) and, if GRAD, s_G[r*NGT*8 + c] holds
// sum_col dl * M[col, gc_lo + c].
template <int KS, bool GRAD, int NGT>
__device__ __forceinline__ void tile_pass(const PassCfg& P, const FmTables& FT, const int* s_rows, const double* s_X,
                                          const Smem& S, int gc_lo) {
  constexpr int KPAD = 4 * KS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int tc = P.tc;
  const int row = s_rows[lr];
  const double* yrow = P.Y + (size_t)(row >= 0 ? row : 0) * P.ldY + 2 * lc;
  const bool fam_nb = P.family == CA_FAMILY_NB;
  const bool per_row = P.nu_per_row != 0;
  double nu_r = 1.0, inu_r = 1.0, lnu_r = 0.0;
  bool pois_r = !fam_nb;
  if (per_row && row >= 0 && fam_nb) {
    nu_r = P.nu[row];
    inu_r = P.inv_nu[row];
    lnu_r = P.log_nu[row];
    pois_r = nu_r > P.thres;
  }
  double afrag[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) afrag[ks] = s_X[lr * KPAD + ks * 4 + lc];
  double ellsum = 0.0;
  double gacc[NGT][2], gacb[NGT][2];  // two accumulator sets: the two columns of a lane go to separate chains
#pragma unroll
  for (int j = 0; j < NGT; ++j) gacc[j][0] = gacc[j][1] = gacb[j][0] = gacb[j][1] = 0.0;

  const int nchunks = (P.ncols + tc - 1) / tc;
  const int chunk_elems = tc * KPAD;
  stage_chunk<KS>(P, 0, S.M, S.cp);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      stage_chunk<KS>(P, c + 1, S.M + ((c + 1) & 1) * chunk_elems, S.cp + ((c + 1) & 1) * tc);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Mb = S.M + (c & 1) * chunk_elems;
    const double4* Cb = S.cp + (c & 1) * tc;
    const int col0 = c * tc;
    const int nrem = P.ncols - col0;
    const int ntile = min(tc, round_up(nrem, 8)) >> 3;
#pragma unroll 1
    for (int t = warp; t < ntile; t += NWARPS) {
      const int cl = t * 8 + 2 * lc;  // column within the chunk
#if CA_ABL == 3
      const double2 yv = make_double2(1.0, 2.0);
#else
      const double2 yv = *reinterpret_cast<const double2*>(yrow + col0 + t * 8);
#endif
      // The DMMA latency on sm_100a is ~250 cycles (tools/fp64_microbench: 16 chains per SMSP
      // are needed to reach the 36.5 TFLOP/s peak), so the k-steps go to independent
      // accumulators and are summed afterwards instead of forming one dependent chain.
      constexpr int NACC = KS < 7 ? KS : 7;
      double pa[NACC][2];
      const double* bp = Mb + (t * 8 + lr) * KPAD + lc;
#pragma unroll
      for (int ks = 0; ks < NACC; ++ks) {
        pa[ks][0] = pa[ks][1] = 0.0;
        dmma884(pa[ks][0], pa[ks][1], afrag[ks], bp[ks * 4]);
      }
#pragma unroll
      for (int ks = NACC; ks < KS; ++ks) dmma884(pa[ks % NACC][0], pa[ks % NACC][1], afrag[ks], bp[ks * 4]);
#pragma unroll
      for (int st = 1; st < NACC; st <<= 1)
#pragma unroll
        for (int i = 0; i + st < NACC; i += 2 * st) {
          pa[i][0] += pa[i + st][0];
          pa[i][1] += pa[i + st][1];
        }
      const double acc0 = pa[0][0], acc1 = pa[0][1];
      double e0, e1, d0 = 0.0, d1 = 0.0;
#if CA_ABL == 1
      e0 = acc0 * yv.x; e1 = acc1 * yv.y; d0 = acc0; d1 = acc1;
      if (false) {
#else
      if (fam_nb) {
#endif
        double x0, x1, m0, m1;
        if (per_row) {
          fm_entry_nb<GRAD>(acc0, yv.x, nu_r, inu_r, lnu_r, FT, P.clip, e0, d0, x0, m0);
          fm_entry_nb<GRAD>(acc1, yv.y, nu_r, inu_r, lnu_r, FT, P.clip, e1, d1, x1, m1);
          if (pois_r) {
            e0 = fma(yv.x, x0, -m0);
            e1 = fma(yv.y, x1, -m1);
            if (GRAD) {
              d0 = m0 - yv.x;
              d1 = m1 - yv.y;
            }
          }
        } else {
          const double4 c0 = Cb[cl], c1 = Cb[cl + 1];
          fm_entry_nb<GRAD>(acc0, yv.x, c0.x, c0.y, c0.z, FT, P.clip, e0, d0, x0, m0);
          fm_entry_nb<GRAD>(acc1, yv.y, c1.x, c1.y, c1.z, FT, P.clip, e1, d1, x1, m1);
          if (c0.w != 0.0) {
            e0 = fma(yv.x, x0, -m0);
            if (GRAD) d0 = m0 - yv.x;
          }
          if (c1.w != 0.0) {
            e1 = fma(yv.y, x1, -m1);
            if (GRAD) d1 = m1 - yv.y;
          }
        }
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
      if (t * 8 + 8 > nrem) {  // ragged last tile (warp uniform)
        if (cl >= nrem) e0 = d0 = 0.0;
        if (cl + 1 >= nrem) e1 = d1 = 0.0;
      }
      ellsum += e0 + e1;
      if (GRAD) {
        const double* g0 = Mb + (t * 8 + 2 * lc) * KPAD + gc_lo + lr;
#pragma unroll
        for (int j = 0; j < NGT; ++j) {
          dmma884(gacc[j][0], gacc[j][1], d0, g0[8 * j]);
          dmma884(gacb[j][0], gacb[j][1], d1, g0[KPAD + 8 * j]);
        }
      }
    }
    __syncthreads();
  }
  // fixed-order reduction: 4 lanes of a row, then the 8 warps
  ellsum += __shfl_xor_sync(0xffffffffu, ellsum, 1);
  ellsum += __shfl_xor_sync(0xffffffffu, ellsum, 2);
  if (lc == 0) S.red[warp * 8 + lr] = ellsum;
  if (GRAD) {
#pragma unroll
    for (int j = 0; j < NGT; ++j) {
      S.Gred[(warp * 8 + lr) * (NGT * 8) + 8 * j + 2 * lc] = gacc[j][0] + gacb[j][0];
      S.Gred[(warp * 8 + lr) * (NGT * 8) + 8 * j + 2 * lc + 1] = gacc[j][1] + gacb[j][1];
    }
  }
  __syncthreads();
  if (tid < 8) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) s += S.red[w * 8 + tid];
    S.rowsum[tid] = s;
  }
  if (GRAD) {
    for (int e = tid; e < 8 * NGT * 8; e += NTHREADS) {
      const int r = e / (NGT * 8), cc = e % (NGT * 8);
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) s += S.Gred[(w * 8 + r) * (NGT * 8) + cc];
      S.G[e] = s;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Gradient / log-likelihood pass over a row list.
//   rowlist == nullptr -> rows 0..nrows-1, else rowlist[0..*count)
//   ell_out[row]                      = sum_col ell
//   G_out[row * ldG + c], c < ngt*8   = sum_col dl * M[col, gc_lo + c]   (GRAD)
// ---------------------------------------------------------------------------
template <int KS, bool GRAD, int NGT>
__global__ void __launch_bounds__(NTHREADS, CA_MINB) k_rowpass(PassCfg P, const double* X, const int* rowlist,
                                                       const int* count, int nrows, int gc_lo,
                                                       double* ell_out, double* G_out, int ldG) {
  constexpr int KPAD = 4 * KS;
  extern __shared__ __align__(16) double smem[];
  const Smem S = carve(smem, P.tc, KPAD, NGT * 8, false);
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
  for (int e = threadIdx.x; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD;
    const int row = s_rows[r];
    S.X[e] = (row >= 0) ? X[(size_t)row * KPAD + c] : 0.0;
  }
  __syncthreads();
  tile_pass<KS, GRAD, NGT>(P, FT, s_rows, S.X, S, gc_lo);
  if (threadIdx.x < 8 && s_rows[threadIdx.x] >= 0) ell_out[s_rows[threadIdx.x]] = S.rowsum[threadIdx.x];
  if (GRAD) {
    for (int e = threadIdx.x; e < 8 * NGT * 8; e += NTHREADS) {
      const int r = e / (NGT * 8), cc = e % (NGT * 8);
      if (s_rows[r] >= 0 && cc < ldG) G_out[(size_t)s_rows[r] * ldG + cc] = S.G[e];
    }
  }
}

__device__ __forceinline__ double soft_thr(double x, double lam) {
  const double a = fabs(x) - lam;
  const double m = a > 0.0 ? a : 0.0;
  return (x > 0.0) ? m : ((x < 0.0) ? -m : 0.0);
}

// ---------------------------------------------------------------------------
// Backtracking line search for a row list (causarray/gcate_opt.py:30-82).
//   g[row, 0..kpad)   projected gradient (zeros outside the moving block)
//   f0[row], normg[row] from k_prep_search
//   prox on columns [prox_lo, prox_hi) with thresholds lam[row*ldlam + c-prox_lo]
//   tys[row] = sum_col Tys (constant term of the row objective)
// On exit X[row] is overwritten with the accepted point, ell_fin[row] holds the raw
// log-likelihood sum at that point.
// ---------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(NTHREADS, CA_MINB) k_search(PassCfg P, SearchArgs A) {
  constexpr int KPAD = 4 * KS;
  extern __shared__ __align__(16) double smem[];
  const Smem S = carve(smem, P.tc, KPAD, 8, true);
  __shared__ int s_rows[8];
  __shared__ double s_t[8], s_f0[8], s_ng[8], s_f01[8], s_e01[8], s_alpha[8], s_efin[8];
  __shared__ unsigned char s_done[8];
  __shared__ int s_mode[8], s_ne[8], s_alldone;
  const int tid = threadIdx.x;
  const int total = A.rowlist ? *A.count : A.nrows;
  const int tile0 = blockIdx.x * TILE_ROWS;
  if (tile0 >= total) return;
  const FmTables FT = fm_load_tables(S.tab, P.fm_tables);
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
  for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD, row = s_rows[r];
    S.X0[e] = (row >= 0) ? A.X[(size_t)row * KPAD + c] : 0.0;
    S.Gd[e] = (row >= 0) ? A.g[(size_t)row * KPAD + c] : 0.0;
    S.X[e] = 0.0;
  }
  __syncthreads();
  const double ncols_d = (double)P.ncols;
  const int nprox = A.prox_hi - A.prox_lo;
  for (int it = 0; it < A.max_iters; ++it) {
    for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
      const int r = e / KPAD, c = e % KPAD;
      if (!s_done[r]) {
        double v = S.X0[e] - s_t[r] * S.Gd[e];
        if (A.lam && c >= A.prox_lo && c < A.prox_hi)
          v = soft_thr(v, A.lam[(size_t)s_rows[r] * A.ldlam + (c - A.prox_lo)]);
        S.X[e] = v;
      }
    }
    __syncthreads();
    tile_pass<KS, false, 1>(P, FT, s_rows, S.X, S, 0);
    if (tid < 8 && !s_done[tid]) {
      const int row = s_rows[tid];
      double pen = 0.0;
      if (A.lam && nprox > 0) {
        for (int c = 0; c < nprox; ++c)
          pen += A.lam[(size_t)row * A.ldlam + c] * fabs(S.X[tid * KPAD + A.prox_lo + c]);
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
  for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD, row = s_rows[r];
    if (row < 0) continue;
    double v = S.X[e];
    if (s_mode[r] == 1) {
      v = S.X0[e] - s_alpha[r] * S.Gd[e];
      if (A.lam && c >= A.prox_lo && c < A.prox_hi) v = soft_thr(v, A.lam[(size_t)row * A.ldlam + (c - A.prox_lo)]);
    }
    A.X[(size_t)row * KPAD + c] = v;
  }
  if (tid < 8 && s_rows[tid] >= 0) {
    A.ell_fin[s_rows[tid]] = s_efin[tid];
    if (A.neval) A.neval[s_rows[tid]] = s_ne[tid];
    if (A.eval_total) atomicAdd(A.eval_total, (unsigned long long)s_ne[tid]);
  }
}

// per-KS launchers (explicitly instantiated in rowpass_inst.cu, one object file per KS)
template <int KS>
int launch_rowpass_ks(const PassCfg& P, const double* X, const int* rowlist, const int* count, int nrows,
                      bool want_grad, int gc_lo, int gc_n, double* ell_out, double* G_out, int ldG, cudaStream_t st);
template <int KS>
int launch_search_ks(const PassCfg& P, const SearchArgs& A, cudaStream_t st);

}  // namespace ca
