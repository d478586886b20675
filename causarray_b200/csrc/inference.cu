// Inference tail on the device: for the (a, p) matrix of test statistics
// that the AIPW / LFC reduction leaves behind -- one row per treatment --
//   * two-sided p-values, Student-t with per-entry (Welch) degrees of freedom or normal,
//   * Benjamini-Hochberg q-values within each treatment (NaN entries left out of the family),
//   * the empirical-null standardisation (t - median) / (MAD / Phi^{-1}(3/4)) and both again.
// Replaces causarray/DR_inference.py:153-201 (scipy.stats.t.sf / norm.sf, statsmodels
// multipletests(method='fdr_bh'), np.nanmedian, scipy median_abs_deviation(scale='normal')), which the
// reference runs per treatment on the host.  On the host this tail cost 12 ms per benchmark batch and
// was the part of the step that slowed down when 8 ranks shared the host cores.
//
// One CTA per treatment: the row is sorted by a bitonic network in global (L2-resident) scratch --
// sizes are a few thousand to a few ten thousand keys -- which yields the order statistics for the
// median / MAD and the ranks for the BH step-up (reverse running minimum of p m / rank, capped at 1).
#include <math.h>
#include <stdint.h>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "devbuf.h"
#include "prof.h"

namespace ca {

// Regularised incomplete beta I_x(a, b) by the modified Lentz continued fraction, with the symmetry
// I_x(a, b) = 1 - I_{1-x}(b, a) where it converges faster.  `x1` = 1 - x is passed separately because
// the callers know it to full precision (x = df / (df + t^2), 1 - x = t^2 / (df + t^2)).
__device__ double incbeta_cf(double a, double b, double x) {
  const double tiny = 1e-300;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 600; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 2e-16) break;
  }
  return h;
}

__device__ double incbeta(double a, double b, double x, double x1) {
  if (!(x > 0.0)) return 0.0;
  if (!(x1 > 0.0)) return 1.0;
  // log B(a, b).  The only caller has b = 1/2 and a = df / 2 up to a few thousand, where
  // lgamma(a) - lgamma(a + 1/2) cancels ~13 digits: use the asymptotic series of the difference,
  // lgamma(a + 1/2) - lgamma(a) = log(a) / 2 - 1/(8a) + 1/(192 a^3) - 1/(640 a^5) + 17/(14336 a^7) - ...
  double lbeta;
  if (b == 0.5 && a >= 20.0) {
    const double ia = 1.0 / a, ia2 = ia * ia;
    const double dlg = 0.5 * log(a) + ia * (-0.125 + ia2 * (1.0 / 192.0 + ia2 * (-1.0 / 640.0 + ia2 * (17.0 / 14336.0))));
    lbeta = 0.57236494292470008707 - dlg;   // lgamma(1/2) = log(pi) / 2
  } else {
    lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
  }
  const double front = exp(a * log(x) + b * log(x1) - lbeta);
  if (x < (a + 1.0) / (a + b + 2.0)) return front * incbeta_cf(a, b, x) / a;
  return 1.0 - front * incbeta_cf(b, a, x1) / b;
}

// two-sided tail probability: Student-t with df degrees of freedom (df > 0) or normal (df <= 0)
__device__ double two_sided_p(double t, double df) {
  if (isnan(t) || isnan(df)) return NAN;
  const double at = fabs(t);
  if (!(df > 0.0)) return erfc(at * 0.70710678118654752440);
  if (isinf(at)) return 0.0;
  if (isinf(df)) return erfc(at * 0.70710678118654752440);
  const double t2 = at * at, den = df + t2;
  // P(|T| > t) = I_x(df / 2, 1 / 2),  x = df / (df + t^2)
  return incbeta(0.5 * df, 0.5, df / den, t2 / den);
}

struct TailRows {
  const double* t;     // a x p statistics (NaN = not tested)
  const double* df;    // a x p or nullptr (normal)
  int a, p, m;         // m = p rounded up to a power of two
  double* key;         // a x m scratch
  int* idx;            // a x m scratch
  double* pv;          // a x p
  double* qv;          // a x p
  double* pva;         // a x p
  double* qva;         // a x p
  double* med_mad;     // a x 2 (median, MAD / Phi^{-1}(3/4))
  int sort_in_smem;
};

// bitonic sort of (key, idx) ascending, one CTA, data in global memory
__device__ void cta_bitonic(double* key, int* idx, int m) {
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = key[i], b = key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[l] = a;
            const int ia = idx[i];
            idx[i] = idx[l];
            idx[l] = ia;
          }
        }
      }
      __syncthreads();
    }
}

__device__ int cta_fill_keys(const double* src, int p, int m, double* key, int* idx, int* s_cnt) {
  if (threadIdx.x == 0) *s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    double v = i < p ? src[i] : NAN;
    const bool ok = !isnan(v);
    key[i] = ok ? v : INFINITY;      // NaN (and padding) sort last
    idx[i] = i;
    c += ok;
  }
  atomicAdd(s_cnt, c);
  __syncthreads();
  return *s_cnt;
}

// median of the cnt smallest keys of a sorted array (numpy: mean of the two middle values)
__device__ double sorted_median(const double* key, int cnt) {
  if (cnt == 0) return NAN;
  const int h = cnt >> 1;
  return (cnt & 1) ? key[h] : 0.5 * (key[h - 1] + key[h]);
}

// Benjamini-Hochberg of the p-values `pv` (row of length p, NaN = left out) into `q`.
__device__ void cta_bh(const double* pv, int p, int m, double* key, int* idx, int* s_cnt, double* s_scan, double* q) {
  const int cnt = cta_fill_keys(pv, p, m, key, idx, s_cnt);
  cta_bitonic(key, idx, m);
  // raw(r) = p_(r) cnt / (r + 1); q_(r) = min(1, min_{r' >= r} raw(r')): reverse running minimum.
  // Each thread owns a contiguous chunk; chunk minima are combined by a serial pass of thread 0
  // (at most blockDim.x values).
  const int nt = blockDim.x;
  const int chunk = (cnt + nt - 1) / nt;
  const int lo = threadIdx.x * chunk, hi = min(cnt, lo + chunk);
  double mn = INFINITY;
  for (int r = hi - 1; r >= lo; --r) {
    const double raw = key[r] * (double)cnt / (double)(r + 1);
    mn = raw < mn ? raw : mn;
  }
  s_scan[threadIdx.x] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = INFINITY;
    for (int w = nt - 1; w >= 0; --w) {
      const double own = s_scan[w];
      s_scan[w] = run;          // minimum over the chunks to the right
      run = own < run ? own : run;
    }
  }
  __syncthreads();
  double run = s_scan[threadIdx.x];
  for (int r = hi - 1; r >= lo; --r) {
    const double raw = key[r] * (double)cnt / (double)(r + 1);
    run = raw < run ? raw : run;
    q[idx[r]] = run < 1.0 ? run : 1.0;
  }
  for (int i = threadIdx.x; i < p; i += nt)
    if (isnan(pv[i])) q[i] = NAN;
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_tail_rows(TailRows R) {
  __shared__ int s_cnt;
  __shared__ double s_scan[1024];
  __shared__ double s_med, s_mad;
  const int row = blockIdx.x;
  const double* t = R.t + (size_t)row * R.p;
  const double* df = R.df ? R.df + (size_t)row * R.p : nullptr;
  // sort buffers: shared memory when the padded row fits (96 KB at p = 8000; the 4 x 91 compare-exchange
  // stages of a row then run at shared-memory instead of L2 latency), global scratch otherwise
  extern __shared__ __align__(16) unsigned char s_sort[];
  double* key = R.sort_in_smem ? reinterpret_cast<double*>(s_sort) : R.key + (size_t)row * R.m;
  int* idx = R.sort_in_smem ? reinterpret_cast<int*>(s_sort + sizeof(double) * R.m) : R.idx + (size_t)row * R.m;
  double* pv = R.pv + (size_t)row * R.p;
  double* qv = R.qv + (size_t)row * R.p;
  double* pva = R.pva + (size_t)row * R.p;
  double* qva = R.qva + (size_t)row * R.p;
  const int p = R.p, m = R.m;
  // ---- p-values and BH ----
  for (int i = threadIdx.x; i < p; i += blockDim.x) pv[i] = two_sided_p(t[i], df ? df[i] : -1.0);
  __syncthreads();
  cta_bh(pv, p, m, key, idx, &s_cnt, s_scan, qv);
  // ---- empirical null: median and MAD of the tested statistics ----
  int cnt = cta_fill_keys(t, p, m, key, idx, &s_cnt);
  cta_bitonic(key, idx, m);
  if (threadIdx.x == 0) s_med = sorted_median(key, cnt);
  __syncthreads();
  const double med = s_med;
  // |t - med| of the tested entries: pva serves as a temporary row
  for (int i = threadIdx.x; i < p; i += blockDim.x) pva[i] = isnan(t[i]) ? NAN : fabs(t[i] - med);
  __syncthreads();
  cnt = cta_fill_keys(pva, p, m, key, idx, &s_cnt);
  cta_bitonic(key, idx, m);
  if (threadIdx.x == 0) s_mad = sorted_median(key, cnt) / 0.6744897501960817;  // scipy scale="normal"
  __syncthreads();
  const double mad = s_mad;
  if (threadIdx.x == 0) {
    R.med_mad[2 * row] = med;
    R.med_mad[2 * row + 1] = mad;
  }
  if (!(mad > 0.0) || cnt == 0) {  // the reference leaves the adjusted columns NaN
    for (int i = threadIdx.x; i < p; i += blockDim.x) pva[i] = qva[i] = NAN;
    return;
  }
  for (int i = threadIdx.x; i < p; i += blockDim.x)
    pva[i] = isnan(t[i]) ? NAN : two_sided_p((t[i] - med) / mad, df ? df[i] : -1.0);
  __syncthreads();
  cta_bh(pva, p, m, key, idx, &s_cnt, s_scan, qva);
}

}  // namespace ca

using namespace ca;

// t, df: a x p host arrays (df may be null: normal tails); outputs a x p host arrays; med_mad a x 2 or null.
extern "C" int ca_inference_tail(const double* t, const double* df, int a, int p, double* pv, double* qv, double* pva,
                                 double* qva, double* med_mad) {
  CA_REQUIRE(t && pv && qv && pva && qva && a > 0 && p > 0, "bad arguments");
  int m = 1;
  while (m < p) m <<= 1;
  const size_t np_ = (size_t)a * p;
  DevBuf<double> dt, ddf, key, out, mm;
  DevBuf<int> idx;
  if (dt.alloc(np_) || (df && ddf.alloc(np_)) || key.alloc((size_t)a * m) || idx.alloc((size_t)a * m) ||
      out.alloc(4 * np_) || mm.alloc(2 * (size_t)a))
    return -1;
  cudaStream_t st = 0;
  CA_CUDA(cudaMemcpyAsync(dt.ptr, t, sizeof(double) * np_, cudaMemcpyHostToDevice, st));
  if (df) CA_CUDA(cudaMemcpyAsync(ddf.ptr, df, sizeof(double) * np_, cudaMemcpyHostToDevice, st));
  TailRows R;
  R.t = dt.ptr;
  R.df = df ? ddf.ptr : nullptr;
  R.a = a;
  R.p = p;
  R.m = m;
  R.key = key.ptr;
  R.idx = idx.ptr;
  R.pv = out.ptr;
  R.qv = out.ptr + np_;
  R.pva = out.ptr + 2 * np_;
  R.qva = out.ptr + 3 * np_;
  R.med_mad = mm.ptr;
  {
    ProfScope prof(PROF_AIPW, st);
    size_t smem = (size_t)m * (sizeof(double) + sizeof(int));
    R.sort_in_smem = smem <= 200 * 1024;
    if (!R.sort_in_smem) smem = 0;
    CA_CUDA(cudaFuncSetAttribute(k_tail_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tail_rows<<<a, 1024, smem, st>>>(R);
    CA_LAUNCH_CHECK();
  }
  CA_CUDA(cudaMemcpyAsync(pv, R.pv, sizeof(double) * np_, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(qv, R.qv, sizeof(double) * np_, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(pva, R.pva, sizeof(double) * np_, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(qva, R.qva, sizeof(double) * np_, cudaMemcpyDeviceToHost, st));
  if (med_mad) CA_CUDA(cudaMemcpyAsync(med_mad, mm.ptr, sizeof(double) * 2 * a, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  return 0;
}
