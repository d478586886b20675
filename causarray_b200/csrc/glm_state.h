// Device-resident state behind ca_glm_t.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "devbuf.h"

struct ca_glm {
  int n = 0, p = 0;
  long ldn = 0;
  cudaStream_t st = 0;
  ca::DevBuf<double> Yraw_own, Yt_own;
  const double* Yraw = nullptr;  // n x p cell-major raw counts
  const double* Yt = nullptr;    // p x ldn gene-major raw counts
  ca::DevBuf<double> X, beta, ybar, cst, dev, dev_prev, step, part, mu_t, offset, nu, ridge, resid;
  ca::DevBuf<int> status, n_iter, counter, genelist;
  bool has_offset = false, has_nu = false;
  int family = 0;
  double thres = 100.0;
  int iters_done = 0;
  int fit_kx = 0;
};
