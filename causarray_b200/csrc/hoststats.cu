// Host-side (CPU) helpers of the inference tail, exported next to the CUDA entry points so that the
// Python layer does not have to run them as dozens of small numpy calls: Benjamini-Hochberg within
// each treatment and the median / MAD of the empirical-null standardisation
// (causarray/DR_inference.py:153-201; statsmodels multipletests(method='fdr_bh')).  They operate on
// (a, p) host arrays -- a few hundred KB -- one std::thread per row; no device work.
#include <math.h>
#include <algorithm>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/causarray_b200.h"
#include "common.cuh"

namespace {

void bh_row(const double* p, int m, double* q) {
  std::vector<std::pair<double, int>> v;
  v.reserve(m);
  for (int j = 0; j < m; ++j) {
    q[j] = NAN;
    if (!isnan(p[j])) v.emplace_back(p[j], j);
  }
  const int cnt = (int)v.size();
  if (cnt == 0) return;
  std::sort(v.begin(), v.end());
  double run = INFINITY;
  for (int r = cnt - 1; r >= 0; --r) {  // step-up: reverse cumulative minimum of p * cnt / rank, capped at 1
    const double raw = v[r].first * (double)cnt / (double)(r + 1);
    run = raw < run ? raw : run;
    q[v[r].second] = run < 1.0 ? run : 1.0;
  }
}

double median_of(std::vector<double>& x) {  // numpy's convention: mean of the two middle values
  const size_t n = x.size();
  if (n == 0) return NAN;
  const size_t h = n / 2;
  std::nth_element(x.begin(), x.begin() + h, x.end());
  const double hi = x[h];
  if (n & 1) return hi;
  const double lo = *std::max_element(x.begin(), x.begin() + h);
  return 0.5 * (lo + hi);
}

void median_mad_row(const double* t, int m, double* med, double* mad) {
  std::vector<double> x;
  x.reserve(m);
  for (int j = 0; j < m; ++j)
    if (!isnan(t[j])) x.push_back(t[j]);
  const double c = median_of(x);
  *med = c;
  for (double& v : x) v = fabs(v - c);
  *mad = median_of(x);
}

template <typename F>
void for_rows(int a, F f) {
  const int nt = std::max(1, std::min(a, (int)std::thread::hardware_concurrency()));
  if (nt <= 1) {
    for (int i = 0; i < a; ++i) f(i);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int w = 0; w < nt; ++w)
    th.emplace_back([=]() {
      for (int i = w; i < a; i += nt) f(i);
    });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" int ca_host_bh_rows(const double* p, int a, int m, double* q) {
  CA_REQUIRE(p && q && a >= 0 && m >= 0, "null argument");
  for_rows(a, [=](int i) { bh_row(p + (size_t)i * m, m, q + (size_t)i * m); });
  return 0;
}

extern "C" int ca_host_median_mad_rows(const double* t, int a, int m, double* med, double* mad) {
  CA_REQUIRE(t && med && mad && a >= 0 && m >= 0, "null argument");
  for_rows(a, [=](int i) { median_mad_row(t + (size_t)i * m, m, med + i, mad + i); });
  return 0;
}
