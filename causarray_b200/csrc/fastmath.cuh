// Lean FP64 exp / log for the likelihood epilogues.
//
// ncu on the first version (profiles/r01_*) showed the row-pass kernels issuing ~420
// instructions per matrix entry, only 29 % of them on the FP64 pipe: the CUDA libm
// exp / log1p / log / division paths carry integer and special-case work the kernels
// do not need (arguments are finite, exp argument <= 100, log argument >= 1).  These
// table-driven versions cost ~10 FP64 instructions each:
//
//   exp(x)  = 2^e * T[j] * (1 + P5(r)),  x = (128 e + j) ln2/128 + r,  |r| <= ln2/256
//   log(w)  = e ln2 + log c_i + log1p(r), w = 2^e f, f in [1,2), i = top 7 mantissa bits,
//             r = f * (1/c_i) - 1 (one fma), |r| <= 1/256, log1p by a degree-6 series
//
// Both are accurate to ~1 ulp (tables are rounded from long double on the host;
// tests/test_gpu_fastmath.py checks them against numpy).  Tables live in shared memory
// (3 KB per CTA), copied from a device buffer filled once by ca::fastmath_tables().
#pragma once
#include <cuda_runtime.h>

namespace ca {

constexpr int FM_EXP_N = 128;
constexpr int FM_LOG_N = 128;
constexpr int FM_TABLE_DOUBLES = FM_EXP_N + 2 * FM_LOG_N;  // exp table, then (1/c, log c) pairs

// device pointer to the tables (filled on first use); nullptr on failure
const double* fastmath_tables();

struct FmTables {
  const double* exp_t;   // FM_EXP_N
  const double2* log_t;  // FM_LOG_N : x = 1/c_i, y = log c_i
};

__device__ __forceinline__ FmTables fm_load_tables(double* s_tab, const double* g_tab) {
  for (int i = threadIdx.x; i < FM_TABLE_DOUBLES; i += blockDim.x) s_tab[i] = g_tab[i];
  FmTables t;
  t.exp_t = s_tab;
  t.log_t = reinterpret_cast<const double2*>(s_tab + FM_EXP_N);
  return t;
}

// exp(x) for -700 <= x <= 100 (callers clamp); result is a normal double.
__device__ __forceinline__ double fm_exp(double x, const double* __restrict__ T) {
  const double MAGIC = 6755399441055744.0;             // 2^52 + 2^51
  const double INV = 184.6649652337873;                // 128 / ln 2
  // ln2/128 = HI + LO, HI has its low 21 mantissa bits cleared so that kd * HI is exact
  const double HI = __longlong_as_double(0x3F762E42FEE00000LL);
  const double LO = 1.4907929134926466e-12;
  const double t = fma(x, INV, MAGIC);
  const int k = __double2loint(t);
  const double kd = t - MAGIC;
  double r = fma(kd, -HI, x);
  r = fma(kd, -LO, r);
  double p = fma(r, 8.333333333333333e-03, 4.1666666666666664e-02);
  p = fma(r, p, 1.6666666666666666e-01);
  p = fma(r, p, 0.5);
  p = fma(r, p, 1.0);
  p = p * r;
  const double tj = T[k & (FM_EXP_N - 1)];
  const double v = fma(tj, p, tj);
  const int e = k >> 7;
  return __hiloint2double(__double2hiint(v) + (e << 20), __double2loint(v));
}

// log(w) for normal w >= 1 (more generally any positive normal w).
__device__ __forceinline__ double fm_log(double w, const double2* __restrict__ LT) {
  const int hi = __double2hiint(w);
  const int e = (hi >> 20) - 1023;
  const int i = (hi >> 13) & (FM_LOG_N - 1);
  const double f = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(w));
  const double2 tc = LT[i];
  const double r = fma(f, tc.x, -1.0);
  double q = fma(r, -1.6666666666666666e-01, 0.2);
  q = fma(r, q, -0.25);
  q = fma(r, q, 3.3333333333333331e-01);
  q = fma(r, q, -0.5);
  const double lp = fma(r * r, q, r);
  const double ed = (double)e;
  const double LN2_HI = 6.93147180369123816490e-01;   // low 21 bits clear: ed * LN2_HI exact
  const double LN2_LO = 1.90821492927058770002e-10;
  return fma(ed, LN2_HI, tc.y) + fma(ed, LN2_LO, lp);
}

// Constants of the clipped NB regime  q = clip(1/(1+u), 1e-6, 1-1e-6)
// (causarray/gcate_likelihood.py:81): values of log1p(-q) and log(q) at the two bounds.
struct NbClip {
  double u_hi, u_lo;          // u >= u_hi -> q = 1e-6 ; u <= u_lo -> q = 1 - 1e-6
  double l1p_hi, lq_hi;       // log1p(-1e-6), log(1e-6)
  double l1p_lo, lq_lo;       // log1p(-(1-1e-6)), log(1-1e-6)
};

// Per-entry NB / Poisson terms, branch free.
//   theta: linear predictor, y: normalised count, inv_nu = 1/nu, log_nu = log(nu), pois: Poisson branch
//   ell = y*log1p(-q) + nu*log(q)  (NB)  |  y*xi - m  (Poisson)
//   dl  = (m - y) * q             (NB)  |  m - y     (Poisson)
template <bool WANT_DL>
__device__ __forceinline__ void fm_entry(double theta, double y, double nu, double inv_nu, double log_nu, bool pois,
                                         const FmTables& T, const NbClip& C, double& ell, double& dl) {
  const double xi = fmin(theta, 100.0);                // the reference clips from above only
  const double m = fm_exp(fmax(xi, -700.0), T.exp_t);  // exp(-700) ~ 1e-304 stands in for the underflow to 0
  const double u = m * inv_nu;
  const double w = 1.0 + u;
  const double L = fm_log(w, T.log_t);                 // log1p(m/nu) = -log(q)
  // unclipped: log1p(-q) = xi - log(nu) - L ; log(q) = -L
  double l1p = (xi - log_nu) - L;
  double lq = -L;
  const bool hi = u >= C.u_hi, lo = u <= C.u_lo;
  l1p = hi ? C.l1p_hi : (lo ? C.l1p_lo : l1p);
  lq = hi ? C.lq_hi : (lo ? C.lq_lo : lq);
  const double ell_nb = fma(y, l1p, nu * lq);
  const double ell_p = fma(y, xi, -m);
  ell = pois ? ell_p : ell_nb;
  if (WANT_DL) {
    double q = __drcp_rn(w);
    q = fmin(fmax(q, 1e-6), 1.0 - 1e-6);
    const double dm = m - y;
    dl = pois ? dm : dm * q;
  }
}

NbClip nb_clip_constants();

}  // namespace ca
