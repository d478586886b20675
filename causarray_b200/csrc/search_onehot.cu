// Stage-2 gene line search for one-hot treatment indicators (causarray/gcate_opt.py:30-82 driven from
// update_with_mask :192-200 with P2 given: only the treatment block beta_T of a gene moves).
//
// With A[:, T] one-hot ({0, 1}, at most one non-zero per cell -- the perturbation indicators of a
// screen) a candidate step changes the linear predictor of a cell by ONE scalar:
//     theta_ij(t) = theta_ij(0) + delta_g(t),   delta_g(t) = soft(beta_g - t g_g, lam_g) - beta_g,  g = group of cell i,
// and not at all for control cells.  So per candidate
//   * control cells (half of a perturbation batch) contribute a constant: ell_ctrl = ell(0) - ell_treated(0),
//   * treated cells need no matrix product and no exponential: m = e^{theta(0)} e^{delta_g},
//     one log1p per entry (NB) -- 19 FP64 instructions against 27 + the DMMA chain of the general pass.
// One warp owns a gene and walks its candidates exactly as the reference's loop does (accept rule,
// t < 1e-4 stop, 1.1 f fallback): first {t = 0, t = alpha} in one pass over the gene's treated cells,
// then the remaining candidates six per pass.  No tiles of 8 genes in lockstep, no second launch, no
// compaction between phases.  theta(0) of the treated cells comes from the gradient pass that precedes
// the search (rowpass_impl.cuh writes it next to the gradient), the counts from a group-sorted copy made
// once per batch.
#include <math.h>
#include <type_traits>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.h"
#include "prof.h"

namespace ca {

constexpr int S2_NC = 6;        // candidates per pass after the first
constexpr int S2_WARPS = 8;

__device__ __forceinline__ double s2_soft(double v, double lam) {
  const double a = fabs(v) - lam;
  return a > 0.0 ? (v > 0.0 ? a : -a) : 0.0;
}

struct S2Args {
  const double* Ytr;     // p x ldtr normalised counts of the treated cells, group-sorted
  const double* Th;      // p x ldtr theta(0) of the same entries
  const signed char* grp;  // ntr: group (column of the treatment block) of every treated cell
  long ldtr;
  int ntr, a, kpad, t_lo;  // treatment block = columns [t_lo, t_lo + a) of a row of X
  int ncols_total;         // n (normaliser of the objective)
  double* X;               // p x kpad coefficients, updated in place
  const double* g;         // p x kpad projected gradient
  const double* lam;       // p x a or nullptr
  const double* f0;
  const double* normg;
  const double* tys;
  const double* ell0;      // log-likelihood sum over ALL cells at the current point
  const double* alpha_row;
  double alpha, beta, tol;
  int max_iters;
  const int* rowlist;
  const int* count;
  int nrows;
  const double* nu;        // per gene (row)
  const double* inv_nu;
  const double* log_nu;
  int family;
  double thres;
  double* ell_fin;
  int* neval;
  unsigned long long* eval_total;  // [0] += candidates, [2] += candidates, [3] += candidates (one-hot path)
  const double* fm_tables;
  NbClip clip;
};

// log-likelihood of one entry at theta0 + delta (m0 = e^{clamp(theta0)}, E = e^{delta})
// The N candidates of one entry, statement by statement across the candidates (fm_log_n: the N dependent chains
// are interleaved in source order, which ptxas keeps -- N consecutive scalar evaluations ran one after the other
// with every dependent FP64 operation exposed).  POIS is uniform over a gene's warp, so the two forms are two
// code paths, not selects.  Comparisons are plain FP64 compares here: the kernel is bound by instruction issue
// (55 % of the slots, FP64 pipe 28 %), and the integer-pipe comparisons of fastmath.cuh cost two ISETPs each.
// (M cells x N candidates per call: M = 4, N = 2 in the first pass, M = 1, N = S2_NC afterwards)
template <bool POIS, int M, int N>
__device__ __forceinline__ void s2_entries(const FmTables& FT, const NbClip& clip, const double* y, const double* th,
                                           const int* g, const double (*dl)[32], const double (*Ee)[32], double nu,
                                           double inu, double lnu, double (&acc)[S2_NC]) {
  double xe[M], m0[M];
#pragma unroll
  for (int u = 0; u < M; ++u) {
    const double t = th[u] < 700.0 ? th[u] : 700.0;
    xe[u] = t > -700.0 ? t : -700.0;
  }
  fm_exp_n<M>(xe, m0, FT.exp_t);
  if (POIS) {
#pragma unroll
    for (int u = 0; u < M; ++u)
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double xr = th[u] + dl[k][g[u]];
        const bool below = xr < 100.0;
        const double xi = below ? xr : 100.0;
        const double m = below ? m0[u] * Ee[k][g[u]] : 2.6881171418161356e43;   // e^100: the reference clips theta from above
        acc[k] += fma(y[u], xi, -m);
      }
    return;
  }
  // NB: a theta clipped at 100 has u = e^100 / nu far above u_hi (nu <= the Poisson threshold), and so has the
  // unclipped product m0 E (or its overflow to +inf): the upper clip of u already gives the reference's value
  double uc[M * N], lu[M * N], w[M * N], L[M * N];
#pragma unroll
  for (int u = 0; u < M; ++u)
#pragma unroll
    for (int k = 0; k < N; ++k) {
      uc[u * N + k] = (m0[u] * Ee[k][g[u]]) * inu;
      lu[u * N + k] = (th[u] + dl[k][g[u]]) - lnu;
    }
#pragma unroll
  for (int e = 0; e < M * N; ++e) {
    const bool lo = uc[e] < clip.u_lo, hi = uc[e] > clip.u_hi;
    uc[e] = lo ? clip.u_lo : uc[e];
    lu[e] = lo ? clip.lu_lo : lu[e];
    uc[e] = hi ? clip.u_hi : uc[e];
    lu[e] = hi ? clip.lu_hi : lu[e];
  }
#pragma unroll
  for (int e = 0; e < M * N; ++e) w[e] = 1.0 + uc[e];
  fm_log_n<M * N>(w, L, FT.log_t);
#pragma unroll
  for (int u = 0; u < M; ++u) {
    const double yn = y[u] + nu;
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] += fma(-yn, L[u * N + k], y[u] * lu[u * N + k]);
  }
}

template <bool NB>
__global__ void __launch_bounds__(S2_WARPS * 32, 2) k_s2_onehot(S2Args A) {
  __shared__ double s_tab[FM_TABLE_DOUBLES];
  __shared__ double s_dl[S2_WARPS][S2_NC][32], s_E[S2_WARPS][S2_NC][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const FmTables FT = fm_load_tables(s_tab, A.fm_tables);
  __syncthreads();
  const int total = A.rowlist ? *A.count : A.nrows;
  const int i = blockIdx.x * S2_WARPS + warp;
  if (i >= total) return;
  const int row = A.rowlist ? A.rowlist[i] : i;
  const int a = A.a;
  const double bt = lane < a ? A.X[(size_t)row * A.kpad + A.t_lo + lane] : 0.0;
  const double gt = lane < a ? A.g[(size_t)row * A.kpad + A.t_lo + lane] : 0.0;
  const double lm = (lane < a && A.lam) ? A.lam[(size_t)row * a + lane] : 0.0;
  const double alpha = A.alpha_row ? A.alpha_row[row] : A.alpha;
  const double f0 = A.f0[row], ng = A.normg[row], tys = A.tys[row], ell0 = A.ell0[row];
  double nu = 1.0, inu = 1.0, lnu = 0.0;
  bool pois = true;
  if (NB) {
    nu = A.nu[row];
    inu = A.inv_nu[row];
    lnu = A.log_nu[row];
    pois = nu > A.thres;
  }
  const NbClip clip = A.clip;
  const double* yrow = A.Ytr + (size_t)row * A.ldtr;
  const double* trow = A.Th + (size_t)row * A.ldtr;
  const double ncd = (double)A.ncols_total;
  double (*dl)[32] = s_dl[warp];
  double (*Ee)[32] = s_E[warp];

  // evaluates nc candidates with steps tc[0..nc): returns their treated-cell log-likelihood sums
  // (all lanes hold them) and their penalties
  // (base: candidate 0 is the CURRENT point, delta = 0 -- not the prox of it)
  // (N = 2: the first pass; N = S2_NC afterwards -- a shorter last batch still evaluates S2_NC candidates, the
  // extra ones at delta = 0, and ignores them)
  auto eval = [&](auto pois_c, auto n_c, const double* tc, int nc, bool base, double* ell_out, double* pen_out) {
    constexpr bool POIS = decltype(pois_c)::value;
    constexpr int N = decltype(n_c)::value;
#pragma unroll
    for (int k = 0; k < S2_NC; ++k) {
      if (k < nc) {
        double x1 = bt - tc[k] * gt;
        if (A.lam) x1 = s2_soft(x1, lm);
        const double d = (base && k == 0) ? 0.0 : x1 - bt;
        dl[k][lane] = lane < a ? d : 0.0;
        Ee[k][lane] = lane < a ? exp(d) : 1.0;
        double pen = (lane < a && A.lam) ? lm * fabs(x1) : 0.0;
        pen = warp_sum(pen);
        pen_out[k] = A.lam ? pen / (double)a : 0.0;
      } else if (k < N) {
        dl[k][lane] = 0.0;
        Ee[k][lane] = 1.0;
      }
    }
    __syncwarp();
    double acc[S2_NC];
#pragma unroll
    for (int k = 0; k < S2_NC; ++k) acc[k] = 0.0;
    // four cells per lane and trip: the twelve loads are issued together (the loop is otherwise bound by
    // the latency of one L2 access per cell)
    for (int c0 = 0; c0 < A.ntr; c0 += 128) {
      double y4[4], th4[4];
      int g4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        const bool ok = c < A.ntr;
        y4[u] = ok ? yrow[c] : 0.0;
        th4[u] = ok ? trow[c] : 0.0;
        g4[u] = ok ? (int)A.grp[c] : -1;
      }
      if (N <= 2 && c0 + 128 <= A.ntr) {
        s2_entries<POIS, 4, N>(FT, clip, y4, th4, g4, dl, Ee, nu, inu, lnu, acc);   // (all 128 cells exist)
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g4[u] < 0) continue;
          s2_entries<POIS, 1, N>(FT, clip, y4 + u, th4 + u, g4 + u, dl, Ee, nu, inu, lnu, acc);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < S2_NC; ++k)
      if (k < nc) ell_out[k] = warp_sum(acc[k]);
    __syncwarp();
  };

  double tc[S2_NC], el[S2_NC], pn[S2_NC];
  // ---- pass 1: t = 0 (baseline of the treated cells) and the first candidate t = alpha ----
  tc[0] = 0.0;
  tc[1] = alpha;
  auto eval_any = [&](const double* tcs, int nc, bool base, double* ell_out, double* pen_out) {
    if (!NB || pois) {
      if (nc <= 2)
        eval(std::true_type{}, std::integral_constant<int, 2>{}, tcs, nc, base, ell_out, pen_out);
      else
        eval(std::true_type{}, std::integral_constant<int, S2_NC>{}, tcs, nc, base, ell_out, pen_out);
    } else {
      if (nc <= 2)
        eval(std::false_type{}, std::integral_constant<int, 2>{}, tcs, nc, base, ell_out, pen_out);
      else
        eval(std::false_type{}, std::integral_constant<int, S2_NC>{}, tcs, nc, base, ell_out, pen_out);
    }
  };
  eval_any(tc, 2, true, el, pn);
  const double ell_ctrl = ell0 - el[0];
  int ne = 1;
  double t = alpha;
  const double e01 = ell_ctrl + el[1];
  const double f01 = -(e01 + tys) / ncd + pn[1];
  double t_sel = alpha, efin = e01;       // chosen step and its log-likelihood sum
  bool done = false;
  if (f01 < f0 - A.tol * t * ng) {
    done = true;
  } else if (t < 1e-4 || A.max_iters <= 1) {
    done = true;                           // (f01 < 1.1 f01: the fallback is the same point)
  }
  int it = 1;
  while (!done) {
    int nc = 0;
    double tt = t;
    for (int k = 0; k < S2_NC && it + k < A.max_iters; ++k) {
      tt *= A.beta;
      tc[k] = tt;
      ++nc;
      if (tt < 1e-4) break;                // the reference stops at the first step below 1e-4
    }
    if (nc == 0) break;
    eval_any(tc, nc, false, el, pn);
    for (int k = 0; k < nc; ++k) {
      const double ell = ell_ctrl + el[k];
      const double f1 = -(ell + tys) / ncd + pn[k];
      const double tk = tc[k];
      ne += 1;
      if (f1 < f0 - A.tol * tk * ng) {
        done = true;
        t_sel = tk;
        efin = ell;
        break;
      } else if (tk < 1e-4 || it + k == A.max_iters - 1) {
        done = true;
        if (f01 < 1.1 * f1) {              // fall back to the full initial step
          t_sel = alpha;
          efin = e01;
        } else {
          t_sel = tk;
          efin = ell;
        }
        break;
      }
    }
    it += nc;
    t = tc[nc - 1];
  }
  if (lane < a) {
    double x1 = bt - t_sel * gt;
    if (A.lam) x1 = s2_soft(x1, lm);
    A.X[(size_t)row * A.kpad + A.t_lo + lane] = x1;
  }
  if (lane == 0) {
    A.ell_fin[row] = efin;
    if (A.neval) A.neval[row] = ne;
    if (A.eval_total) {
      atomicAdd(A.eval_total, (unsigned long long)ne);
      atomicAdd(A.eval_total + 2, (unsigned long long)ne);
      atomicAdd(A.eval_total + 3, (unsigned long long)ne);
    }
  }
}

// Ytr[g][c] = Ynt[g][cells[c]]
__global__ void k_s2_gather(const double* __restrict__ Ynt, long ldn, int p, const int* __restrict__ cells, int ntr,
                            double* __restrict__ Ytr, long ldtr) {
  const int gene = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ntr; c += gridDim.x * blockDim.x)
    Ytr[(size_t)gene * ldtr + c] = Ynt[(size_t)gene * ldn + cells[c]];
}

int launch_s2_gather(const double* Ynt, long ldn, int p, const int* cells, int ntr, double* Ytr, long ldtr,
                     cudaStream_t st) {
  if (ntr <= 0) return 0;
  ProfScope prof(PROF_PREP, st);
  k_s2_gather<<<dim3((ntr + 1023) / 1024, p), 256, 0, st>>>(Ynt, ldn, p, cells, ntr, Ytr, ldtr);
  CA_LAUNCH_CHECK();
  return 0;
}

int launch_s2_onehot(const S2Launch& L, cudaStream_t st) {
  if (L.nrows <= 0) return 0;
  S2Args A;
  A.Ytr = L.Ytr;
  A.Th = L.Th;
  A.grp = L.grp;
  A.ldtr = L.ldtr;
  A.ntr = L.ntr;
  A.a = L.a;
  A.kpad = L.kpad;
  A.t_lo = L.t_lo;
  A.ncols_total = L.ncols_total;
  A.X = L.X;
  A.g = L.g;
  A.lam = L.lam;
  A.f0 = L.f0;
  A.normg = L.normg;
  A.tys = L.tys;
  A.ell0 = L.ell0;
  A.alpha_row = L.alpha_row;
  A.alpha = L.alpha;
  A.beta = L.beta;
  A.tol = L.tol;
  A.max_iters = L.max_iters;
  A.rowlist = L.rowlist;
  A.count = L.count;
  A.nrows = L.nrows;
  A.nu = L.nu;
  A.inv_nu = L.inv_nu;
  A.log_nu = L.log_nu;
  A.family = L.family;
  A.thres = L.thres;
  A.ell_fin = L.ell_fin;
  A.neval = L.neval;
  A.eval_total = L.eval_total;
  A.fm_tables = fastmath_tables();
  CA_REQUIRE(A.fm_tables, "fastmath tables unavailable");
  A.clip = nb_clip_constants();
  CA_REQUIRE(L.a >= 1 && L.a <= 32, "the one-hot search handles at most 32 treatment columns");
  ProfScope prof(PROF_SEARCH_ONEHOT, st);
  const int grid = (L.nrows + S2_WARPS - 1) / S2_WARPS;
  if (L.family == CA_FAMILY_NB)
    k_s2_onehot<true><<<grid, S2_WARPS * 32, 0, st>>>(A);
  else
    k_s2_onehot<false><<<grid, S2_WARPS * 32, 0, st>>>(A);
  CA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ca
