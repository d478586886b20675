// Device-resident state behind ca_glm_t.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "devbuf.h"

namespace ca {
// Split of a design into dense columns + one-hot groups for the persistent IRLS kernel
// (glm_irls.cu: plan_irls; glm_irls_impl.cuh explains the formulation).
struct IrlsPlan {
  bool ok = false;
  int kx = 0, kde = 0, ngroups = 1, npad = 0, nchunks = 0, nseg_tot = 0;
  signed char dcol[16], gcol[64], cdense[64];
  unsigned char cgroup[64];
  std::vector<int> perm;    // sorted padded position -> cell, -1 for a padding row
  std::vector<int> gpair;   // group of every 8-cell pair-step
  int chunk_lo[9], seg_ptr[9];
  unsigned char gseg_ptr[66], gseg_list[96];
};
bool plan_irls(const double* X, int n, int kx, IrlsPlan& L);
}  // namespace ca

struct ca_glm {
  int n = 0, p = 0;
  long ldn = 0;
  cudaStream_t st = 0;
  ca::DevBuf<double> Yraw_own, Yt_own;
  const double* Yraw = nullptr;  // n x p cell-major raw counts
  const double* Yt = nullptr;    // p x ldn gene-major raw counts
  ca::DevBuf<double> X, beta, ybar, cst, dev, dev_prev, step, part, mu_t, offset, nu, ridge, resid, resid_t;
  ca::DevBuf<int> status, n_iter, counter, genelist, not_counts;
  // persistent IRLS path: counts in group-sorted padded cell order (cached), dense design rows
  ca::DevBuf<double> Yp, Xd, irls_scratch;
  ca::DevBuf<int> perm_dev, gpair;
  std::vector<int> perm_host;
  const double* Yp_src = nullptr;
  bool mu_valid = false;       // mu_t holds the fitted means of the last fit
  // layout of the last fit's design in X (pass kernel's internal column order), for the lazy
  // fitted-means pass
  int lay_kxi = 0, lay_t0 = 0;
  signed char lay_cmap[64];
  bool has_offset = false, has_nu = false;
  int family = 0;
  double thres = 100.0;
  int iters_done = 0;
  int fit_kx = 0;
};
