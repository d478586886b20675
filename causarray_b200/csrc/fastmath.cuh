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

// Polynomial / reduction constants live in constant memory so that DFMA takes them as
// c[bank][offset] operands (the first version materialised every 64-bit literal with a
// pair of UMOVs inside the hot loop -- 9 % of the issued instructions).
static __constant__ double FMC[24] = {
    6755399441055744.0,          // 0  MAGIC = 2^52 + 2^51
    184.6649652337873,           // 1  128 / ln2
    -5.41521234663378e-03,       // 2  -HI  (ln2/128 with the low 21 mantissa bits cleared; set exactly at load)
    -1.4907929134926466e-12,     // 3  -LO
    8.333333333333333e-03,       // 4  1/120
    4.1666666666666664e-02,      // 5  1/24
    1.6666666666666666e-01,      // 6  1/6
    0.5,                         // 7
    -1.6666666666666666e-01,     // 8  log1p series
    0.2,                         // 9
    -0.25,                       // 10
    3.3333333333333331e-01,      // 11
    -0.5,                        // 12
    6.93147180369123816490e-01,  // 13 LN2_HI (low 21 bits clear)
    1.90821492927058770002e-10,  // 14 LN2_LO
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};

// exp(x) for -700 <= x <= 709 (callers clamp); result is a normal double.
__device__ __forceinline__ double fm_exp(double x, const double* __restrict__ T) {
  const double t = fma(x, FMC[1], FMC[0]);
  const int k = __double2loint(t);
  const double kd = t - FMC[0];
  double r = fma(kd, FMC[2], x);
  r = fma(kd, FMC[3], r);
  double p = fma(r, FMC[4], FMC[5]);
  p = fma(r, p, FMC[6]);
  p = fma(r, p, FMC[7]);
  p = fma(r, p, 1.0);
  p = p * r;
  const double tj = T[k & (FM_EXP_N - 1)];
  const double v = fma(tj, p, tj);
  return __hiloint2double(__double2hiint(v) + ((k >> 7) << 20), __double2loint(v));
}

// log(w) for positive normal w.
__device__ __forceinline__ double fm_log(double w, const double2* __restrict__ LT) {
  const int hi = __double2hiint(w);
  const int e = (hi >> 20) - 1023;
  const double f = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(w));
  const double2 tc = LT[(hi >> 13) & (FM_LOG_N - 1)];
  const double r = fma(f, tc.x, -1.0);
  double q = fma(r, FMC[8], FMC[9]);
  q = fma(r, q, FMC[10]);
  q = fma(r, q, FMC[11]);
  q = fma(r, q, FMC[12]);
  const double lp = fma(r * r, q, r);
  const double ed = (double)e;
  return fma(ed, FMC[13], tc.y) + fma(ed, FMC[14], lp);
}

// Clip of the NB success probability  q = clip(1/(1+u), 1e-6, 1-1e-6), u = e^xi / nu
// (causarray/gcate_likelihood.py:81), restated as a clamp of u (and of log u):
//   q_c = 1/(1+u_c), u_c = clamp(u, u_lo, u_hi);  log(q_c) = -log1p(u_c);
//   log1p(-q_c) = log(u_c) - log1p(u_c),  log(u_c) = clamp(xi - log nu, log u_lo, log u_hi).
struct NbClip {
  double u_lo, u_hi, lu_lo, lu_hi;
};

// 1/w for w in [1, 1e6+1]: MUFU.RCP64H seed + two Newton steps (no special-case branch: the
// libm __drcp_rn carries a slow-path CALL that splits the basic block and stops ptxas from
// interleaving the independent per-entry chains).
__device__ __forceinline__ double fm_rcp(double w) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
  // seed error e0 ~ 2^-20; one third-order step (r0 (1 + e + e^2)) leaves e0^3, one Newton step more
  // takes it far below 2^-53
  double e = fma(-w, r, 1.0);
  r = fma(r, fma(e, e, e), r);
  e = fma(-w, r, 1.0);
  return fma(r, e, r);
}

// N independent evaluations written statement by statement across the N values, so that the generated code
// has the N dependent chains interleaved IN SOURCE ORDER (ptxas does not reorder the far-apart chains of N
// consecutive scalar calls when registers are tight: the per-octet IRLS kernel ran its genes one after the
// other, every dependent DFMA exposed).  Same arithmetic, operation for operation, as the scalar versions.
template <int N>
__device__ __forceinline__ void fm_exp_n(const double (&x)[N], double (&out)[N], const double* __restrict__ T) {
  double t[N], kd[N], r[N], p[N], tj[N];
  int k[N];
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(x[i], FMC[1], FMC[0]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    k[i] = __double2loint(t[i]);
    kd[i] = t[i] - FMC[0];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) tj[i] = T[k[i] & (FM_EXP_N - 1)];
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(kd[i], FMC[2], x[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(kd[i], FMC[3], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(r[i], FMC[4], FMC[5]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(r[i], p[i], FMC[6]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(r[i], p[i], FMC[7]);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(r[i], p[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = p[i] * r[i];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double v = fma(tj[i], p[i], tj[i]);
    out[i] = __hiloint2double(__double2hiint(v) + ((k[i] >> 7) << 20), __double2loint(v));
  }
}

template <int N>
__device__ __forceinline__ void fm_log_n(const double (&w)[N], double (&out)[N], const double2* __restrict__ LT) {
  double f[N], r[N], q[N], ed[N];
  double2 tc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int hi = __double2hiint(w[i]);
    ed[i] = (double)((hi >> 20) - 1023);
    f[i] = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(w[i]));
    tc[i] = LT[(hi >> 13) & (FM_LOG_N - 1)];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(f[i], tc[i].x, -1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(r[i], FMC[8], FMC[9]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(r[i], q[i], FMC[10]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(r[i], q[i], FMC[11]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(r[i], q[i], FMC[12]);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(r[i] * r[i], q[i], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(ed[i], FMC[13], tc[i].y) + fma(ed[i], FMC[14], q[i]);
}

template <int N>
__device__ __forceinline__ void fm_rcp_n(const double (&w)[N], double (&out)[N]) {
  double r[N], e[N];
#pragma unroll
  for (int i = 0; i < N; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(w[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-w[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(r[i], fma(e[i], e[i], e[i]), r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-w[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = fma(r[i], e[i], r[i]);
}

// Comparisons against a constant of known sign done on the integer pipe (IEEE-754 doubles of one
// sign order like their bit patterns): a DSETP costs a slot on the FP64 pipe, which is the pipe the
// likelihood kernels are bound by; ISETP pairs run on the ALU.  Exact for every non-NaN input.
__device__ __forceinline__ bool fm_lt_pos(double a, double c) {  // a < c, c > 0
#ifdef CA_NO_ICMP
  return a < c;
#else
  return __double_as_longlong(a) < __double_as_longlong(c);
#endif
}
__device__ __forceinline__ bool fm_gt_pos(double a, double c) {  // a > c, a >= 0, c > 0
#ifdef CA_NO_ICMP
  return a > c;
#else
  return __double_as_longlong(a) > __double_as_longlong(c);
#endif
}
__device__ __forceinline__ bool fm_lt_neg(double a, double c) {  // a < c, c < 0
#ifdef CA_NO_ICMP
  return a < c;
#else
  return (unsigned long long)__double_as_longlong(a) > (unsigned long long)__double_as_longlong(c);
#endif
}
__device__ __forceinline__ bool fm_gt_neg(double a, double c) {  // a > c, c < 0
#ifdef CA_NO_ICMP
  return a > c;
#else
  return (unsigned long long)__double_as_longlong(a) < (unsigned long long)__double_as_longlong(c);
#endif
}

// Per-entry NB terms, branch free (straight-line code so that ptxas interleaves the four
// independent entries a lane has in flight: every FP64 op has an 8-cycle latency).
//   lu = log(u_c), L = log1p(u_c), w = 1 + u_c, m = e^xi:
//   ell = y*log1p(-q) + nu*log(q) = y*lu - (y + nu)*L,   dl = (m - y) * q = (m - y) / w.
__device__ __forceinline__ void fm_terms_nb(double theta, double inv_nu, double log_nu, const FmTables& T,
                                            const NbClip& C, double& xi, double& m, double& lu, double& L,
                                            double& w) {
  xi = fm_lt_pos(theta, 100.0) ? theta : 100.0;         // the reference clips from above only
  const double xe = fm_gt_neg(xi, -700.0) ? xi : -700.0;  // exp(-700) ~ 1e-304 stands in for the underflow to 0
  m = fm_exp(xe, T.exp_t);
  double uc = m * inv_nu;
  lu = xi - log_nu;
  const bool lo = fm_lt_pos(uc, C.u_lo), hi = fm_gt_pos(uc, C.u_hi);
  uc = lo ? C.u_lo : uc;
  lu = lo ? C.lu_lo : lu;
  uc = hi ? C.u_hi : uc;
  lu = hi ? C.lu_hi : lu;
  w = 1.0 + uc;
  L = fm_log(w, T.log_t);                               // log1p(u_c) = -log(q_c)
}

template <bool WANT_DL>
__device__ __forceinline__ void fm_entry_nb(double theta, double y, double nu, double inv_nu, double log_nu,
                                            const FmTables& T, const NbClip& C, double& ell, double& dl, double& xi,
                                            double& m) {
  double lu, L, w;
  fm_terms_nb(theta, inv_nu, log_nu, T, C, xi, m, lu, L, w);
  ell = fma(y, lu, -(y + nu) * L);                      // y (lu - L) - nu L
  if (WANT_DL) dl = (m - y) * fm_rcp(w);
}

NbClip nb_clip_constants();

}  // namespace ca
