// K8 of SURVEY.md section 2b: AIPW pseudo-outcomes + LFC estimand as fused
// per-gene segmented reductions.
//
// Replaces AIPW_mean (causarray/DR_estimation.py:624-659), the size-factor
// normalisation (causarray/DR_learner.py:208), the per-treatment cell
// selection (:212-219) and the LFC `estimand` closure (:403-482):
//   psi0 = (1-A)/(1-pi) (Y - Yhat0) + Yhat0 ,  psi1 = A/pi (Y - Yhat1) + Yhat1   (both / size_factor)
//   m_a = mean(psi_a) ; estimable = finite & > 0 ; tau_a = max(m_a, thres_diff) (1 if not estimable)
//   tau = log(tau_1 / tau_0) ; IF = psi1/tau_1 - psi0/tau_0
//   Welch: v_a = (var_{A=a}(IF, ddof=1) + eps)/n_a ; var = v0 + v1 ; df = Satterthwaite
//   pooled: var = (var(IF, ddof=1) + eps)/n
//   filter (~estimable | max(m) < thres_min | |m1 - m0| < thres_diff) -> tau 0, var inf, df nan
// The (n, p, a, 2) tensors of the reference are never formed: Yhat0 =
// min(exp(W Bcov^T + offset), cap) is recomputed per (cell, gene) and
// Yhat1_k = min(exp(.) * exp(Btrt_k), cap).
//
// One warp owns one gene; lanes stride over cells.  Per-cell records
// (W row, offset, 1/size_factor, and per treatment the two AIPW weights and
// the inclusion flag) are packed on the host into rows of odd length so that
// lane-strided shared-memory reads are conflict free, and stream through
// shared memory in chunks (cp.async).  Treatments are processed in groups of
// JT so that the 4*JT running sums stay in registers.
#include <vector>
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "devbuf.h"
#include "kernels.h"
#include "glm_state.h"
#include "prof.h"
#include "fastmath.cuh"

namespace ca {

constexpr int DR_WARPS = 8;
constexpr int DR_THREADS = DR_WARPS * 32;
constexpr int DR_CHUNK = 64;
constexpr int JT = 8;

struct DrArgs {
  const double* Yt;   // p x ldn gene-major raw counts
  long ldn;
  int n, p, kw, a;
  const double* rec;  // n x ldr packed records for this treatment group, cells ordered controls first, then by treatment
  int ldr;            // kw + 3 + 3*JT (the last field is the cell's index into Yt), forced odd
  int j0, nj;         // treatment group [j0, j0 + nj)
  const double* Bcov; // p x kw
  const double* Btrt; // p x a
  const double* cnt;  // a x 3 : N included, n0, n1
  double cap;
  int usevar;
  double thres_min, thres_diff, eps_var;
  double* S;          // a x p x 2 sums of psi0, psi1 (pass A out / pass B in)
  double* mean0; double* mean1; double* tau; double* var; double* df; uint8_t* estimable;  // a x p
  // explicit-Yhat variant
  const double* Yhat;  // n x p x a x 2 or nullptr
  const double* fm_tables;
};

// PASS 0: sums of psi0 / psi1.  PASS 1: centred second moments -> outputs.
template <int PASS>
__global__ void __launch_bounds__(DR_THREADS, 2) k_aipw(DrArgs D) {
  extern __shared__ __align__(16) double smem[];
  double* Rs = smem;  // 2 * DR_CHUNK * ldr, then the exp / log tables
  // pass 1: per-treatment constants of the warp's gene (1 / tau_0, 1 / tau_1, shift) live in shared memory --
  // in registers they kept the kernel at 170 registers and one CTA per SM
  __shared__ double s_c[DR_WARPS][3 * JT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const FmTables FT = fm_load_tables(smem + 2 * DR_CHUNK * D.ldr, D.fm_tables);
  const int gene = blockIdx.x * DR_WARPS + warp;
  const bool ok_gene = gene < D.p;
  const int ldr = D.ldr, kw = D.kw;
  const double* bcov = D.Bcov ? D.Bcov + (size_t)(ok_gene ? gene : 0) * kw : nullptr;
  double ej[JT];
  double* t0 = s_c[warp];            // in pass 1, t0 / t1 hold the RECIPROCALS of tau_0 / tau_1
  double* t1 = s_c[warp] + JT;
  double* shift = s_c[warp] + 2 * JT;
#pragma unroll
  for (int j = 0; j < JT; ++j) {
    ej[j] = 1.0;
    if (lane == 0) {
      t0[j] = t1[j] = 1.0;
      shift[j] = 0.0;
    }
    if (ok_gene && j < D.nj) {
      if (D.Btrt) ej[j] = exp(D.Btrt[(size_t)gene * D.a + D.j0 + j]);
      if (PASS == 1) {
        const int jj = D.j0 + j;
        const double N = D.cnt[jj * 3 + 0];
        const double m0 = D.S[((size_t)jj * D.p + gene) * 2 + 0] / N;
        const double m1 = D.S[((size_t)jj * D.p + gene) * 2 + 1] / N;
        const bool est = isfinite(m0) && isfinite(m1) && m0 > 0.0 && m1 > 0.0;
        const double tt0 = est ? fmax(m0, D.thres_diff) : 1.0;
        const double tt1 = est ? fmax(m1, D.thres_diff) : 1.0;
        if (lane == 0) {
          shift[j] = m1 / tt1 - m0 / tt0;
          t0[j] = 1.0 / tt0;   // the per-cell influence values are scaled by multiplication
          t1[j] = 1.0 / tt1;
        }
      }
    }
  }
  __syncwarp();
  double acc0[JT], acc1[JT], acc2[JT], acc3[JT];
#pragma unroll
  for (int j = 0; j < JT; ++j) acc0[j] = acc1[j] = acc2[j] = acc3[j] = 0.0;
  const double* yrow = D.Yt + (size_t)(ok_gene ? gene : 0) * D.ldn;

  const int nchunks = (D.n + DR_CHUNK - 1) / DR_CHUNK;
  auto stage = [&](int c, int buf) {
    const int c0 = c * DR_CHUNK;
    const int nvalid = min(DR_CHUNK, D.n - c0);
    const int ndbl = nvalid * ldr;
    const double* src = D.rec + (size_t)c0 * ldr;
    double* dst = Rs + buf * DR_CHUNK * ldr;
    // records are only 8-byte aligned (odd ldr): plain loads
    for (int i = tid; i < ndbl; i += DR_THREADS) dst[i] = src[i];
  };
  stage(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    __syncthreads();
    if (c + 1 < nchunks) stage(c + 1, (c + 1) & 1);
    if (ok_gene) {
      const double* Rc = Rs + (c & 1) * DR_CHUNK * ldr;
      for (int blk = 0; blk < DR_CHUNK; blk += 32) {
        const int cell = c * DR_CHUNK + blk + lane;
        if (cell >= D.n) continue;
        const double* rec = Rc + (blk + lane) * ldr;
        const int ci = (int)rec[kw + 2 + 3 * JT];   // the cell's position in Yt (records are sorted by arm)
        const double y = yrow[ci];
        const double isf = rec[kw + 1];
        double yh0_raw = 0.0;
        if (!D.Yhat) {
          double eta = rec[kw];
          for (int q = 0; q < kw; ++q) eta = fma(rec[q], bcov[q], eta);
          // table-driven exp (fastmath.cuh); anything beyond the clamp is cut by the cap below
          yh0_raw = fm_exp(eta < 700.0 ? (eta > -700.0 ? eta : -700.0) : 700.0, FT.exp_t);
        }
#pragma unroll
        for (int j = 0; j < JT; ++j) {
          if (j >= D.nj) break;
          const double w0 = rec[kw + 2 + 3 * j], w1 = rec[kw + 3 + 3 * j], inc = rec[kw + 4 + 3 * j];
          if (inc == 0.0) continue;
          double yh0, yh1;
          if (D.Yhat) {
            const size_t base = (((size_t)ci * D.p + gene) * D.a + (D.j0 + j)) * 2;
            yh0 = fmin(D.Yhat[base], D.cap);
            yh1 = fmin(D.Yhat[base + 1], D.cap);
          } else {
            const double r1 = yh0_raw * ej[j];
            yh0 = yh0_raw < D.cap ? yh0_raw : D.cap;
            yh1 = r1 < D.cap ? r1 : D.cap;
          }
          // w0 = (1-A)/(1-pi)/sf, w1 = A/pi/sf, isf = 1/sf
          const double psi0 = w0 * (y - yh0) + yh0 * isf;
          const double psi1 = w1 * (y - yh1) + yh1 * isf;
          if (PASS == 0) {
            acc0[j] += psi0;
            acc1[j] += psi1;
          } else {
            const double dlt = fma(psi1, t1[j], -fma(psi0, t0[j], shift[j]));
            if (w1 != 0.0) {  // A == 1 arm
              acc2[j] += dlt;
              acc3[j] += dlt * dlt;
            } else {
              acc0[j] += dlt;
              acc1[j] += dlt * dlt;
            }
          }
        }
      }
    }
  }
  if (!ok_gene) return;
#pragma unroll
  for (int j = 0; j < JT; ++j) {
    if (j >= D.nj) break;
    const int jj = D.j0 + j;
    const double s0 = warp_sum(acc0[j]), s1 = warp_sum(acc1[j]);
    if (PASS == 0) {
      if (lane == 0) {
        D.S[((size_t)jj * D.p + gene) * 2 + 0] = s0;
        D.S[((size_t)jj * D.p + gene) * 2 + 1] = s1;
      }
    } else {
      const double s2 = warp_sum(acc2[j]), s3 = warp_sum(acc3[j]);
      if (lane == 0) {
        const double N = D.cnt[jj * 3 + 0], n0 = D.cnt[jj * 3 + 1], n1 = D.cnt[jj * 3 + 2];
        const double m0 = D.S[((size_t)jj * D.p + gene) * 2 + 0] / N;
        const double m1 = D.S[((size_t)jj * D.p + gene) * 2 + 1] / N;
        const bool est = isfinite(m0) && isfinite(m1) && m0 > 0.0 && m1 > 0.0;
        double tau = log(t0[j] / t1[j]);   // (reciprocals: tau_1 / tau_0)
        double var, dfe = NAN;
        if (D.usevar == 1) {
          const double var0 = (s1 - s0 * s0 / n0) / (n0 - 1.0);
          const double var1 = (s3 - s2 * s2 / n1) / (n1 - 1.0);
          const double v0 = (var0 + D.eps_var) / n0, v1 = (var1 + D.eps_var) / n1;
          var = v0 + v1;
          dfe = (v0 + v1) * (v0 + v1) / (v0 * v0 / (n0 - 1.0) + v1 * v1 / (n1 - 1.0));
        } else {
          const double sd = s0 + s2, sq = s1 + s3;
          var = ((sq - sd * sd / N) / (N - 1.0) + D.eps_var) / N;
        }
        const bool filt = !est || (fmax(m0, m1) < D.thres_min) || (fabs(m1 - m0) < D.thres_diff);
        if (filt) {
          tau = 0.0;
          var = INFINITY;
          dfe = NAN;
        }
        const size_t o = (size_t)jj * D.p + gene;
        D.mean0[o] = m0;
        D.mean1[o] = m1;
        D.tau[o] = tau;
        D.var[o] = var;
        D.df[o] = dfe;
        D.estimable[o] = est ? 1 : 0;
      }
    }
  }
}

// Host driver shared by the three entry points.
static int run_aipw(const double* Yt_dev, long ldn, int n, int p, int kw, int a, const double* W, const double* Bcov,
                    const double* Btrt, const double* offset, const double* size_factor, const double* A,
                    const double* pi, const uint8_t* cell_mask, const double* Yhat_dev, const ca_lfc_params* prm,
                    double* mean0, double* mean1, double* tau, double* var, double* df, uint8_t* estimable,
                    cudaStream_t st) {
  CA_REQUIRE(prm && A && pi && size_factor, "null argument");
  CA_REQUIRE(prm->usevar == 0 || prm->usevar == 1, "usevar must be either \"pooled\" or \"unequal\"");
  const bool dbg = getenv("CA_TIMING") != nullptr;
  auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    cudaStreamSynchronize(st);
    auto T1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[run_aipw] %-12s %.3f ms\n", what, std::chrono::duration<double, std::milli>(T1 - T0).count());
    T0 = T1;
  };
  int ldr = kw + 3 + 3 * JT;
  if (ldr % 2 == 0) ldr += 1;
  // inclusion, arm counts (causarray/DR_learner.py:212-219, 426-427)
  std::vector<uint8_t> ctrl(n);
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < a; ++j) s += A[(size_t)i * a + j];
    ctrl[i] = (s == 0.0);
  }
  std::vector<double> cnt((size_t)a * 3, 0.0);
  const int ngroups = (a + JT - 1) / JT;
  DevBuf<double> d_rec, d_Bcov, d_Btrt, d_cnt, d_S, d_out;
  DevBuf<uint8_t> d_est;
  if (d_rec.alloc((size_t)n * ldr) || d_cnt.alloc((size_t)a * 3) || d_S.alloc((size_t)a * p * 2) ||
      d_out.alloc((size_t)a * p * 5) || d_est.alloc((size_t)a * p))
    return -1;
  if (Bcov) {
    if (d_Bcov.alloc((size_t)p * kw) || d_Btrt.alloc((size_t)p * a)) return -1;
    CA_CUDA(cudaMemcpyAsync(d_Bcov.ptr, Bcov, sizeof(double) * (size_t)p * kw, cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaMemcpyAsync(d_Btrt.ptr, Btrt, sizeof(double) * (size_t)p * a, cudaMemcpyHostToDevice, st));
  }
  lap("alloc+h2d");
  // Cells are visited controls first, then by treatment: a warp's 32 cells then share their inclusion pattern
  // (a control cell is part of every treatment's contrast, a treated cell of one) and the per-treatment loop of
  // the kernel does not diverge; in input order half of the lanes idled through it.  The sums only change in
  // their rounding.
  std::vector<int> order(n);
  {
    std::vector<int> key(n);
    for (int i = 0; i < n; ++i) {
      int k = 0;
      if (!ctrl[i])
        for (int j = 0; j < a; ++j)
          if (A[(size_t)i * a + j] != 0.0) {
            k = 1 + j;
            break;
          }
      key[i] = k;
      order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
  }
  std::vector<std::vector<double>> recs(ngroups, std::vector<double>((size_t)n * ldr, 0.0));
  for (int g = 0; g < ngroups; ++g) {
    std::vector<double>& R = recs[g];
    const int j0 = g * JT, nj = std::min(JT, a - j0);
    for (int c = 0; c < n; ++c) {
      const int i = order[c];
      double* r = R.data() + (size_t)c * ldr;
      r[kw + 2 + 3 * JT] = (double)i;
      for (int q = 0; q < kw; ++q) r[q] = W ? W[(size_t)i * kw + q] : 0.0;
      r[kw] = offset ? offset[i] : 0.0;
      const double isf = 1.0 / size_factor[i];
      r[kw + 1] = isf;
      for (int j = 0; j < nj; ++j) {
        const int jj = j0 + j;
        const double Aij = A[(size_t)i * a + jj], pij = pi[(size_t)i * a + jj];
        const bool inc = cell_mask ? (cell_mask[(size_t)i * a + jj] != 0) : (ctrl[i] || Aij == 1.0);
        if (!inc) continue;
        r[kw + 2 + 3 * j] = ((1.0 - Aij) / (1.0 - pij)) / size_factor[i];
        r[kw + 3 + 3 * j] = (Aij / pij) / size_factor[i];
        r[kw + 4 + 3 * j] = 1.0;
        cnt[(size_t)jj * 3 + 0] += 1.0;
        if (Aij == 0.0) cnt[(size_t)jj * 3 + 1] += 1.0;
        if (Aij == 1.0) cnt[(size_t)jj * 3 + 2] += 1.0;
      }
    }
  }
  lap("pack");
  CA_CUDA(cudaMemcpyAsync(d_cnt.ptr, cnt.data(), sizeof(double) * a * 3, cudaMemcpyHostToDevice, st));
  DrArgs D;
  D.Yt = Yt_dev;
  D.ldn = ldn;
  D.n = n;
  D.p = p;
  D.kw = kw;
  D.a = a;
  D.rec = d_rec.ptr;
  D.ldr = ldr;
  D.Bcov = Bcov ? d_Bcov.ptr : nullptr;
  D.Btrt = Bcov ? d_Btrt.ptr : nullptr;
  D.cnt = d_cnt.ptr;
  D.cap = prm->yhat_cap;
  D.usevar = prm->usevar;
  D.thres_min = prm->thres_min;
  D.thres_diff = prm->thres_diff;
  D.eps_var = prm->eps_var;
  D.S = d_S.ptr;
  D.mean0 = d_out.ptr;
  D.mean1 = d_out.ptr + (size_t)a * p;
  D.tau = d_out.ptr + (size_t)2 * a * p;
  D.var = d_out.ptr + (size_t)3 * a * p;
  D.df = d_out.ptr + (size_t)4 * a * p;
  D.estimable = d_est.ptr;
  D.Yhat = Yhat_dev;
  D.fm_tables = fastmath_tables();
  CA_REQUIRE(D.fm_tables, "fastmath tables unavailable");
  const size_t sm = sizeof(double) * (2 * DR_CHUNK * ldr + FM_TABLE_DOUBLES);
  CA_CUDA(cudaFuncSetAttribute(k_aipw<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  CA_CUDA(cudaFuncSetAttribute(k_aipw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int grid = (p + DR_WARPS - 1) / DR_WARPS;
  for (int g = 0; g < ngroups; ++g) {
    D.j0 = g * JT;
    D.nj = std::min(JT, a - D.j0);
    CA_CUDA(cudaMemcpyAsync(d_rec.ptr, recs[g].data(), sizeof(double) * (size_t)n * ldr, cudaMemcpyHostToDevice, st));
    {
      ProfScope prof(PROF_AIPW, st);
      k_aipw<0><<<grid, DR_THREADS, sm, st>>>(D);
      CA_LAUNCH_CHECK();
    }
    {
      ProfScope prof(PROF_AIPW, st);
      k_aipw<1><<<grid, DR_THREADS, sm, st>>>(D);
      CA_LAUNCH_CHECK();
    }
    CA_CUDA(cudaStreamSynchronize(st));  // recs[g] is reused as the staging source
  }
  lap("kernels");
  const size_t ap = (size_t)a * p;
  CA_CUDA(cudaMemcpyAsync(mean0, D.mean0, sizeof(double) * ap, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(mean1, D.mean1, sizeof(double) * ap, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(tau, D.tau, sizeof(double) * ap, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(var, D.var, sizeof(double) * ap, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(df, D.df, sizeof(double) * ap, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(estimable, d_est.ptr, ap, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  lap("d2h");
  return 0;
}

}  // namespace ca

using namespace ca;

static int upload_transposed(const double* Y, int n, int p, DevBuf<double>& raw, DevBuf<double>& yt, long ldn,
                             cudaStream_t st) {
  if (raw.alloc((size_t)n * p) || yt.alloc((size_t)p * ldn)) return -1;
  CA_CUDA(cudaMemcpyAsync(raw.ptr, Y, sizeof(double) * (size_t)n * p, cudaMemcpyHostToDevice, st));
  CA_CUDA(cudaMemsetAsync(yt.ptr, 0, yt.bytes(), st));
  return launch_transpose(raw.ptr, n, p, p, yt.ptr, ldn, st);
}

extern "C" int ca_aipw_lfc(int n, int p, int kw, int a, const double* Y, const double* W, const double* Bcov,
                           const double* Btrt, const double* offset, const double* size_factor, const double* A,
                           const double* pi, const uint8_t* cell_mask, const ca_lfc_params* prm, double* mean0,
                           double* mean1, double* tau, double* var, double* df, uint8_t* estimable) {
  CA_REQUIRE(Y && W && Bcov && Btrt && n > 0 && p > 0 && a > 0 && kw > 0, "bad argument");
  DevBuf<double> raw, yt;
  const long ldn = round_up(n, 8);
  if (upload_transposed(Y, n, p, raw, yt, ldn, 0)) return -1;
  return run_aipw(yt.ptr, ldn, n, p, kw, a, W, Bcov, Btrt, offset, size_factor, A, pi, cell_mask, nullptr, prm, mean0,
                  mean1, tau, var, df, estimable, 0);
}

extern "C" int ca_aipw_lfc_shared(ca_glm_t h, int kw, int a, const double* W, const double* Bcov, const double* Btrt,
                                  const double* offset, const double* size_factor, const double* A, const double* pi,
                                  const uint8_t* cell_mask, const ca_lfc_params* prm, double* mean0, double* mean1,
                                  double* tau, double* var, double* df, uint8_t* estimable) {
  CA_REQUIRE(h && h->Yt && W && Bcov && Btrt && a > 0 && kw > 0, "bad argument");
  return run_aipw(h->Yt, h->ldn, h->n, h->p, kw, a, W, Bcov, Btrt, offset, size_factor, A, pi, cell_mask, nullptr, prm,
                  mean0, mean1, tau, var, df, estimable, h->st);
}

extern "C" int ca_aipw_lfc_yhat(int n, int p, int a, const double* Y, const double* Yhat, const double* size_factor,
                                const double* A, const double* pi, const uint8_t* cell_mask, const ca_lfc_params* prm,
                                double* mean0, double* mean1, double* tau, double* var, double* df,
                                uint8_t* estimable) {
  CA_REQUIRE(Y && Yhat && n > 0 && p > 0 && a > 0, "bad argument");
  DevBuf<double> raw, yt, yh;
  const long ldn = round_up(n, 8);
  if (upload_transposed(Y, n, p, raw, yt, ldn, 0)) return -1;
  const size_t nyh = (size_t)n * p * a * 2;
  if (yh.alloc(nyh)) return -1;
  CA_CUDA(cudaMemcpyAsync(yh.ptr, Yhat, sizeof(double) * nyh, cudaMemcpyHostToDevice, 0));
  return run_aipw(yt.ptr, ldn, n, p, 0, a, nullptr, nullptr, nullptr, nullptr, size_factor, A, pi, cell_mask, yh.ptr,
                  prm, mean0, mean1, tau, var, df, estimable, 0);
}
