// K4 / K5 / K6 of SURVEY.md section 2b: batched per-gene GLM fits by IRLS.
//
// Replaces the per-gene statsmodels fit driven from fit_glm
// (causarray/gcate_glm.py:187-236) -- algorithm restated in oracle/glm.py:
//   mu0 = (y + mean y)/2, eta0 = log mu0, dev0 = deviance(mu0)
//   repeat: w = mu (Poisson) | mu/(1 + mu/nu) (NB);  z = eta + (y - mu)/mu - offset
//           beta = argmin || sqrt(w) (z - X beta) ||   (normal equations, Cholesky)
//           eta = X beta + offset; mu = exp(eta); dev = deviance(mu)
//           stop when |dev - dev_prev| <= tol
// plus the method-of-moments dispersion (gcate_glm.py:301-306) and signed
// deviance residuals (statsmodels resid_deviance / nb_glm_fast.py:729-777).
//
// One warp owns one gene.  The shared design X streams through shared memory in
// chunks of 128 cells (bulk async copies on mbarriers, double buffered; glm_pass_impl.cuh)
// and is reused by the 8 genes of the CTA.  Per 32 cells a lane evaluates eta/mu/w/z for ONE cell, then the
// weighted Gram  X^T diag(w) X  is accumulated on the FP64 tensor cores: for a
// 4-cell step the fragment X[cell = lane%4][col = lane/4 + 8t] is the B
// operand as is and, multiplied by w[cell], the A operand (DMMA m8n8k4); only
// the upper-triangular 8x8 tiles are kept.  A second kernel solves the
// (Jacobi-equilibrated) normal equations with a warp-level Cholesky.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <vector>
#include <new>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "devbuf.h"
#include "kernels.h"
#include "glm_state.h"
#include "gcate_state.h"
#include "prof.h"
#include "fastmath.cuh"
#include "glm_pass_impl.cuh"
#include "glm_solve.cuh"

namespace ca {

// Per running gene (one warp): combine the cell-range partials in fixed order, evaluate the
// deviance at the current beta and the convergence test of the previous update
// (|dev - dev_prev| <= tol), and -- if the gene keeps running -- solve the Jacobi-equilibrated
// normal equations of the next update with a warp-level Cholesky.
struct GlmSolve {
  int p, kx, nslots, nsplit, part_ld;
  int kxi;               // width of the pass kernel's internal column layout
  signed char emap[64];  // caller's column -> internal column
  const int* genelist;   // nullptr = identity
  const double* part;
  const double* ridge;   // kx or nullptr
  const double* l1;      // kx or nullptr: L1 weights n * alpha_k of the lasso refit (caller's column order)
  const double* nu;
  int family;
  double thres;
  const double* cst;
  double* dev;
  double* dev_prev;
  int* status;
  int* n_iter;
  double* beta;
  int iter;
  double tol;
  double* step;          // p: relative size of the gene's last coefficient update (lasso refit)
  double steptol;        // lasso refit: also require step <= steptol (see the convergence test)
  int finalize;          // only the deviance
  int* n_active;
};

__global__ void k_glm_solve(GlmSolve Q) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  const int slot = blockIdx.x * nw + warp;
  if (slot >= Q.nslots) return;
  const int gene = Q.genelist ? Q.genelist[slot] : slot;
  const int kx = Q.kx, kxi = Q.kxi;
  if (!Q.finalize && Q.status[gene] != 0) return;
  const double* part0 = Q.part + (size_t)slot * Q.nsplit * Q.part_ld;
  // ---- deviance ----
  double s_yeta = 0.0, s_mu = 0.0, s_log = 0.0;
  for (int sp = 0; sp < Q.nsplit; ++sp) {
    const double* ps = part0 + (size_t)sp * Q.part_ld + kxi * kxi + kxi;
    s_yeta += ps[0];
    s_mu += ps[1];
    s_log += ps[2];
  }
  bool pois = true;
  if (Q.family == CA_FAMILY_NB && Q.nu) pois = Q.nu[gene] > Q.thres;
  const double sy = Q.cst[(size_t)gene * 3 + 0], sylogy = Q.cst[(size_t)gene * 3 + 1], synu = Q.cst[(size_t)gene * 3 + 2];
  double dev = pois ? 2.0 * ((sylogy - s_yeta) - (sy - s_mu)) : 2.0 * ((sylogy - s_yeta) - (synu - s_log));
  if (Q.l1 && !Q.finalize) {
    // lasso refit: the convergence test runs on the penalised objective dev + 2 sum_k l1_k |beta_k|
    // (= 2 n times statsmodels' -loglike / n + sum_k alpha_k |beta_k|, up to a constant)
    double pen = 0.0;
    for (int i = 0; i < kx; ++i) pen += Q.l1[i] * fabs(Q.beta[(size_t)gene * kx + i]);
    dev += 2.0 * pen;
  }
  if (Q.finalize) {
    if (lane == 0) Q.dev[gene] = dev;
    return;
  }
  bool run = true;
  {
    const double prev = Q.dev_prev[gene];
    if (!(dev == dev) || isinf(dev)) {
      run = false;
      if (lane == 0) {
        Q.status[gene] = 2;
        Q.n_iter[gene] = Q.iter - 1;
      }
    } else if (Q.iter > 1 && fabs(dev - prev) <= Q.tol && (!Q.l1 || Q.step[gene] <= Q.steptol)) {
      // (lasso refit: the penalised objective is flat along the weakly identified columns -- a change
      // of 1e-8 in it still leaves their coefficients 1e-4 from the minimiser -- so the refit also
      // waits for the update itself to become small)
      run = false;
      if (lane == 0) {
        Q.status[gene] = 1;
        Q.n_iter[gene] = Q.iter - 1;
      }
    }
    if (lane == 0) {
      Q.dev[gene] = dev;
      if (run) {
        Q.dev_prev[gene] = dev;
        atomicAdd(Q.n_active, 1);
      }
    }
  }
  if (!run) return;
  // ---- normal equations ----
  const int ld = kx + 1;
  // (the lasso refit keeps a second copy of the system; plain fits do not pay its shared memory)
  const GlmSolveWork W = glm_solve_carve(sm + (size_t)warp * glm_solve_work_doubles(kx, Q.l1 != nullptr), kx);
  double* L = W.L;
  double* sc = W.sc;
  double* b = W.b;
  {
    // fixed summation order over the splits; a lane walks the flat index e = i * kx + j in steps of
    // 32 (row / column tracked incrementally: no division) and keeps 2 x nsplit loads in flight
    int i0 = lane / kx, j0 = lane - i0 * kx;
    const int di = 32 / kx, dj = 32 - di * kx;
    for (int e = lane; e < kx * kx; e += 64) {
      int i1 = i0 + di, j1 = j0 + dj;
      if (j1 >= kx) {
        j1 -= kx;
        ++i1;
      }
      const bool two = e + 32 < kx * kx;
      const int ea = (int)Q.emap[i0] * kxi + (int)Q.emap[j0];
      const int eb = two ? (int)Q.emap[i1] * kxi + (int)Q.emap[j1] : ea;
      double ga = 0.0, gb = 0.0;
      int sp = 0;
      for (; sp + 8 <= Q.nsplit; sp += 8) {
        double va[8], vb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          va[u] = part0[(size_t)(sp + u) * Q.part_ld + ea];
          vb[u] = part0[(size_t)(sp + u) * Q.part_ld + eb];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          ga += va[u];
          gb += vb[u];
        }
      }
      for (; sp < Q.nsplit; ++sp) {
        ga += part0[(size_t)sp * Q.part_ld + ea];
        gb += part0[(size_t)sp * Q.part_ld + eb];
      }
      if (Q.ridge && i0 == j0) ga += Q.ridge[i0];
      L[i0 * ld + j0] = ga;
      if (two) {
        if (Q.ridge && i1 == j1) gb += Q.ridge[i1];
        L[i1 * ld + j1] = gb;
      }
      // advance both cursors by 64
      i0 = i1 + di;
      j0 = j1 + dj;
      if (j0 >= kx) {
        j0 -= kx;
        ++i0;
      }
    }
  }
  for (int i = lane; i < kx; i += 32) {
    double r = 0.0;
    for (int sp = 0; sp < Q.nsplit; ++sp) r += part0[(size_t)sp * Q.part_ld + kxi * kxi + (int)Q.emap[i]];
    b[i] = r;
  }
  __syncwarp();
  if (glm_solve_system(W, kx, Q.l1, lane)) {
    if (lane == 0) {
      Q.status[gene] = 2;
      Q.n_iter[gene] = Q.iter;
      atomicAdd(Q.n_active, -1);
    }
    return;
  }
  bool fin = true;
  double stepmax = 0.0;
  for (int i = lane; i < kx; i += 32) {
    const double v = b[i] * sc[i];
    const double old = Q.beta[(size_t)gene * kx + i];
    stepmax = fmax(stepmax, fabs(v - old) / (1.0 + fabs(v)));
    Q.beta[(size_t)gene * kx + i] = v;
    fin = fin && (v == v) && !isinf(v);
  }
  if (Q.l1) {
    for (int o = 16; o > 0; o >>= 1) stepmax = fmax(stepmax, __shfl_xor_sync(0xffffffffu, stepmax, o));
    if (lane == 0) Q.step[gene] = stepmax;
  }
  fin = __all_sync(0xffffffffu, fin);
  if (!fin && lane == 0) {
    Q.status[gene] = 2;
    Q.n_iter[gene] = Q.iter;
    atomicAdd(Q.n_active, -1);
  }
}

// per-gene constants: mean y, sum y, sum y log y, sum (y+nu) log(y+nu)
__global__ void k_glm_consts(const double* __restrict__ Yt, long ldn, int n, const double* __restrict__ nu,
                             int family, double thres, double* __restrict__ ybar, double* __restrict__ cst) {
  __shared__ double scratch[32];
  const int gene = blockIdx.x;
  const double* y = Yt + (size_t)gene * ldn;
  const bool nb = family == CA_FAMILY_NB && nu && !(nu[gene] > thres);
  const double nuj = nb ? nu[gene] : 1.0;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = y[i];
    s0 += v;
    if (v > 0.0) s1 += v * log(v);
    if (nb) s2 += (v + nuj) * log(v + nuj);
  }
  s0 = block_sum(s0, scratch);
  s1 = block_sum(s1, scratch);
  s2 = block_sum(s2, scratch);
  if (threadIdx.x == 0) {
    ybar[gene] = s0 / (double)n;
    cst[(size_t)gene * 3 + 0] = s0;
    cst[(size_t)gene * 3 + 1] = s1;
    cst[(size_t)gene * 3 + 2] = s2;
  }
}

// list of genes still running (status == 0), in gene order (single CTA, deterministic)
__global__ void k_glm_active_list(const int* __restrict__ status, int p, int* __restrict__ list, int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < p; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const int flag = (i < p && status[i] == 0) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int pre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (flag) list[off + pre] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

// guards of causarray/gcate_glm.py:205
__global__ void k_glm_guard(int p, int kx, int d_guard, const double* __restrict__ beta, int* __restrict__ status) {
  const int gene = blockIdx.x * blockDim.x + threadIdx.x;
  if (gene >= p) return;
  if (status[gene] == 2) return;
  bool bad = false;
  for (int c = 0; c < kx; ++c) {
    const double v = beta[(size_t)gene * kx + c];
    if (!(v == v) || isinf(v)) bad = true;
    if (c < d_guard ? fabs(v) > 50.0 : fabs(v) > 10.0) bad = true;
  }
  status[gene] = bad ? 3 : status[gene];
}

// signed deviance residuals, written cell-major (n x p) through a tiled transpose
__global__ void k_glm_resid(const double* __restrict__ Yt, const double* __restrict__ mu_t, long ldn, int n, int p,
                            const double* __restrict__ nu, int family, double thres, const int* __restrict__ status,
                            double* __restrict__ out, long ldo, double* __restrict__ out_t) {
  __shared__ double tile[32][33];
  const int g0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int gene = g0 + rr, cell = i0 + threadIdx.x;
    double v = 0.0;
    if (gene < p && cell < n && (status == nullptr || status[gene] <= 1)) {
      const double y = Yt[(size_t)gene * ldn + cell], mu = mu_t[(size_t)gene * ldn + cell];
      const bool nb = family == CA_FAMILY_NB && nu && !(nu[gene] > thres);
      const double r = fmax(y / mu, 2.220446049250313e-16);
      double dd = y * log(r);
      if (nb) {
        const double ia = nu[gene];
        dd -= (y + ia) * log((y + ia) / (mu + ia));
      } else {
        dd -= (y - mu);
      }
      dd *= 2.0;
      const double sg = (y > mu) ? 1.0 : ((y < mu) ? -1.0 : 0.0);
      v = sg * sqrt(fmax(dd, 0.0));
    }
    tile[rr][threadIdx.x] = v;
    if (out_t && gene < p && cell < n) out_t[(size_t)gene * ldn + cell] = v;   // gene-major copy (for R^T products)
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int cell = i0 + rr, gene = g0 + threadIdx.x;
    if (cell < n && gene < p) out[(size_t)cell * ldo + gene] = tile[threadIdx.x][rr];
  }
}

// method-of-moments NB size (gcate_glm.py:301-306), one CTA per gene
__global__ void k_disp_moments(const double* __restrict__ Yt, const double* __restrict__ mu_t, long ldn, int n,
                               const double* __restrict__ offset, double* __restrict__ disp) {
  __shared__ double scratch[32];
  const int gene = blockIdx.x;
  const double* y = Yt + (size_t)gene * ldn;
  const double* m = mu_t + (size_t)gene * ldn;
  double mx = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double sf = offset ? exp(offset[i]) : 1.0;
    mx = fmax(mx, y[i] / sf);
  }
  // block max through the sum helper: use shuffles
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, scratch[w]);
  __syncthreads();
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double sf = offset ? exp(offset[i]) : 1.0;
    const double yn = y[i] / sf;
    const double yh = fmin(fmax(m[i] / sf, 0.0), mx);
    const double dlt = yn - yh;
    s1 += dlt * dlt - yh;
    s2 += yh * yh;
  }
  s1 = block_sum(s1, scratch);
  s2 = block_sum(s2, scratch);
  if (threadIdx.x == 0) {
    const double phi = (s1 / (double)n) / (s2 / (double)n);
    double v = 1.0 / fmin(fmax(phi, 0.01), 100.0);
    if (!(phi == phi)) v = 1.0;
    disp[gene] = v;
  }
}

static int launch_pass(const GlmPass& P, cudaStream_t st) {
  const int nt = (P.kxi + 7) / 8;
  switch (nt) {
    case 1: return launch_glm_pass_nt<1>(P, st);
    case 2: return launch_glm_pass_nt<2>(P, st);
    case 3: return launch_glm_pass_nt<3>(P, st);
    case 4: return launch_glm_pass_nt<4>(P, st);
    case 5: return launch_glm_pass_nt<5>(P, st);
    case 6: return launch_glm_pass_nt<6>(P, st);
    case 7: return launch_glm_pass_nt<7>(P, st);
    case 8: return launch_glm_pass_nt<8>(P, st);
  }
  set_error("design wider than 64 columns");
  return -2;
}

int kpad_for(int k);

int glm_fit_irls(ca_glm* h, const IrlsPlan& L, const double* X, const double* offset, int family, double thres,
                 bool has_nu, bool has_ridge, bool lasso, double tol, int maxiter, bool warm, bool freeze_dense,
                 const int* genelist_dev, int n_list);

// Internal column layout of the pass kernel.  A contiguous block of design columns of which every
// cell has at most one non-zero entry (one-hot treatment indicators, causarray/DR_learner.py builds
// the design as [W, A]) has a diagonal weighted Gram block whatever the IRLS weights are.  Moving
// that block behind the other ("dense") columns, aligned to an 8-column tile, lets the kernel skip
// the DMMA tiles that lie inside it (3 of 10 for [1, U(10) | A(15)]).  The layout is used only when
// it lowers the number of tiles.
struct GlmLayout {
  int kxi;
  signed char cmap[64];  // internal -> caller (-1: padding)
  signed char emap[64];  // caller -> internal
  int t0;                // first tile of the orthogonal block (= number of tiles: none)
};

static int tri_tiles(int nt) { return nt * (nt + 1) / 2; }

static GlmLayout plan_layout(const double* X, int n, int kx) {
  GlmLayout L;
  L.kxi = kx;
  L.t0 = (kx + 7) / 8;
  for (int c = 0; c < 64; ++c) L.cmap[c] = L.emap[c] = (signed char)(c < kx ? c : -1);
  if (kx < 9 || getenv("CA_GLM_NO_ORTHO")) return L;
  // conflict[i] bit j: some cell has columns i and j both non-zero
  unsigned long long conflict[64] = {0};
  for (int i = 0; i < n; ++i) {
    const double* x = X + (size_t)i * kx;
    unsigned long long m = 0;
    for (int c = 0; c < kx; ++c)
      if (x[c] != 0.0) m |= 1ull << c;
    for (unsigned long long r = m; r; r &= r - 1) conflict[__builtin_ctzll(r)] |= m;
  }
  // longest contiguous range [lo, hi) without a conflict between two distinct columns
  int best_lo = 0, best_n = 0;
  for (int lo = 0; lo < kx; ++lo) {
    unsigned long long in = 0;
    int hi = lo;
    while (hi < kx && !(conflict[hi] & in & ~(1ull << hi))) {
      in |= 1ull << hi;
      ++hi;
    }
    if (hi - lo > best_n) {
      best_n = hi - lo;
      best_lo = lo;
    }
  }
  if (best_n < 9) return L;  // at least one full tile pair inside the block is needed to skip anything
  const int n_dense = kx - best_n;
  const int o0 = round_up(n_dense, 8);
  const int kxi = o0 + best_n;
  if (kxi > 64) return L;
  const int nt_i = (kxi + 7) / 8, nt_e = (kx + 7) / 8, t0 = o0 / 8;
  const int tiles_i = tri_tiles(nt_i) - tri_tiles(nt_i - t0);
  if (tiles_i >= tri_tiles(nt_e)) return L;
  if (t0 < 1 || t0 > 2 || nt_i - t0 < 2) return L;  // variants built in glm_pass_inst.cu
  L.kxi = kxi;
  for (int c = 0; c < 64; ++c) L.cmap[c] = -1;
  int di = 0;
  for (int c = 0; c < kx; ++c) {
    const bool ortho = c >= best_lo && c < best_lo + best_n;
    const int ci = ortho ? o0 + (c - best_lo) : di++;
    L.cmap[ci] = (signed char)c;
    L.emap[c] = (signed char)ci;
  }
  L.t0 = t0;
  return L;
}

}  // namespace ca

using namespace ca;

extern "C" int ca_glm_create(ca_glm_t* out, int n, int p) {
  CA_REQUIRE(out && n > 0 && p > 0, "bad dimensions");
  ca_glm* h = new (std::nothrow) ca_glm();
  CA_REQUIRE(h, "out of host memory");
  h->n = n;
  h->p = p;
  h->ldn = round_up(n, 8);
  h->st = 0;
  *out = h;
  return 0;
}

extern "C" int ca_glm_destroy(ca_glm_t h) {
  if (h) {
    cudaStreamSynchronize(h->st);
    delete h;
  }
  return 0;
}

extern "C" int ca_glm_load_counts(ca_glm_t h, const double* Y) {
  CA_REQUIRE(h && Y, "null argument");
  const int n = h->n, p = h->p;
  if (h->Yraw_own.alloc((size_t)n * p) || h->Yt_own.alloc((size_t)p * h->ldn)) return -1;
  if (h2d_staged(h->Yraw_own.ptr, Y, sizeof(double) * (size_t)n * p, h->st)) return -1;
  CA_CUDA(cudaMemsetAsync(h->Yt_own.ptr, 0, h->Yt_own.bytes(), h->st));
  if (launch_transpose(h->Yraw_own.ptr, n, p, p, h->Yt_own.ptr, h->ldn, h->st)) return -1;
  CA_CUDA(cudaStreamSynchronize(h->st));
  h->Yraw = h->Yraw_own.ptr;
  h->Yt = h->Yt_own.ptr;
  h->Yp_src = nullptr;   // the group-sorted copy of the IRLS kernel belongs to the previous counts (same buffer!)
  h->mu_valid = false;
  return 0;
}

extern "C" int ca_glm_share_counts(ca_glm_t h, ca_gcate_t g) {
  CA_REQUIRE(h && g && g->raw_loaded, "GCATE handle holds no raw counts");
  CA_REQUIRE(h->n == g->n && h->p == g->p, "shape mismatch");
  h->Yraw = g->Yraw.ptr;
  h->Yt = g->Yraw_t.ptr;
  h->ldn = g->ldn;
  h->Yp_src = nullptr;
  h->mu_valid = false;
  return 0;
}

extern "C" int ca_glm_fit(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                          double thres_disp, int maxiter, double tol, int d_guard, double* B, double* dev,
                          int32_t* n_iter, int32_t* status) {
  return ca_glm_fit_ex(h, family, X, kx, offset, nu, thres_disp, maxiter, tol, d_guard, nullptr, nullptr, B, dev, n_iter,
                       status);
}

static int glm_fit_impl(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                        double thres_disp, int maxiter, double tol, int d_guard, const double* ridge, const double* l1,
                        const uint8_t* gene_mask, bool warm, double* B, double* dev, int32_t* n_iter, int32_t* status,
                        bool freeze_dense = false);

extern "C" int ca_glm_fit_ex(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                             double thres_disp, int maxiter, double tol, int d_guard, const double* ridge,
                             const uint8_t* gene_mask, double* B, double* dev, int32_t* n_iter, int32_t* status) {
  return glm_fit_impl(h, family, X, kx, offset, nu, thres_disp, maxiter, tol, d_guard, ridge, nullptr, gene_mask, false, B,
                      dev, n_iter, status);
}

extern "C" int ca_glm_refit_l1(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                               double thres_disp, int maxiter, double tol, const double* l1, const uint8_t* gene_mask,
                               double* B, double* dev, int32_t* n_iter, int32_t* status) {
  CA_REQUIRE(l1 && gene_mask, "the lasso refit needs its weights and the gene subset");
  return glm_fit_impl(h, family, X, kx, offset, nu, thres_disp, maxiter, tol, -1, nullptr, l1, gene_mask, true, B, dev,
                      n_iter, status);
}

extern "C" int ca_glm_fit_treatment(ca_glm_t h, int family, const double* X, int kx, const double* offset,
                                    const double* nu, double thres_disp, int maxiter, double tol, double* B, double* dev,
                                    int32_t* n_iter, int32_t* status) {
  return glm_fit_impl(h, family, X, kx, offset, nu, thres_disp, maxiter, tol, -1, nullptr, nullptr, nullptr, true, B, dev,
                      n_iter, status, true);
}

static int glm_fit_impl(ca_glm_t h, int family, const double* X, int kx, const double* offset, const double* nu,
                        double thres_disp, int maxiter, double tol, int d_guard, const double* ridge, const double* l1,
                        const uint8_t* gene_mask, bool warm, double* B, double* dev, int32_t* n_iter, int32_t* status,
                        bool freeze_dense) {
  CA_REQUIRE(h && h->Yt && X && B, "load counts first");
  CA_REQUIRE(kx > 0 && kx <= 64, "design width must be in [1, 64]");
  CA_REQUIRE(family == CA_FAMILY_POISSON || family == CA_FAMILY_NB, "Family not recognized");
  const int n = h->n, p = h->p;
  cudaStream_t st = h->st;
  const GlmLayout lay = plan_layout(X, n, kx);
  const int kxi = lay.kxi;
  const int ldx = 8 * ((kxi + 7) / 8) + 4;
  if (h->X.alloc((size_t)n * ldx) || h->beta.alloc((size_t)p * kx) || h->ybar.alloc(p) || h->cst.alloc((size_t)p * 3) ||
      h->dev.alloc(p) || h->dev_prev.alloc(p) || h->step.alloc(p) || h->status.alloc(p) || h->n_iter.alloc(p) ||
      h->part.alloc((size_t)p * (kxi * kxi + kxi + 4)) || h->mu_t.alloc((size_t)p * h->ldn) ||
      h->counter.alloc(2) || h->offset.alloc((size_t)n + 8) || h->nu.alloc(p) || h->ridge.alloc(2 * kx))
    return -1;
  CA_REQUIRE(!gene_mask || h->fit_kx == kx, "gene_mask refit needs a previous fit of the same width");
  h->fit_kx = kx;
  {
    // design in the internal column layout, zero padded to ldx
    std::vector<double> xi((size_t)n * ldx, 0.0);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < kx; ++c) xi[(size_t)i * ldx + lay.emap[c]] = X[(size_t)i * kx + c];
    CA_CUDA(cudaMemcpyAsync(h->X.ptr, xi.data(), sizeof(double) * xi.size(), cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaStreamSynchronize(st));
  }
  CA_CUDA(cudaMemsetAsync(h->offset.ptr, 0, h->offset.bytes(), st));
  if (offset) CA_CUDA(cudaMemcpyAsync(h->offset.ptr, offset, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  if (nu) CA_CUDA(cudaMemcpyAsync(h->nu.ptr, nu, sizeof(double) * p, cudaMemcpyHostToDevice, st));
  if (ridge) CA_CUDA(cudaMemcpyAsync(h->ridge.ptr, ridge, sizeof(double) * kx, cudaMemcpyHostToDevice, st));
  if (l1) CA_CUDA(cudaMemcpyAsync(h->ridge.ptr + kx, l1, sizeof(double) * kx, cudaMemcpyHostToDevice, st));
  if (gene_mask) {
    // refit only the flagged genes; the others keep their coefficients and are frozen
    std::vector<int> st0(p);
    std::vector<double> hb((size_t)p * kx);
    CA_CUDA(cudaMemcpy(hb.data(), h->beta.ptr, sizeof(double) * (size_t)p * kx, cudaMemcpyDeviceToHost));
    for (int j = 0; j < p; ++j) {
      st0[j] = gene_mask[j] ? 0 : 1;
      if (!gene_mask[j]) continue;
      // a warm refit starts from the coefficients of the previous fit -- unless that fit broke down
      // numerically, in which case the gene starts from zero like a cold fit
      bool reset = !warm;
      for (int c = 0; c < kx && !reset; ++c) reset = !std::isfinite(hb[(size_t)j * kx + c]);
      if (reset)
        for (int c = 0; c < kx; ++c) hb[(size_t)j * kx + c] = 0.0;
    }
    CA_CUDA(cudaMemcpy(h->beta.ptr, hb.data(), sizeof(double) * (size_t)p * kx, cudaMemcpyHostToDevice));
    CA_CUDA(cudaMemcpy(h->status.ptr, st0.data(), sizeof(int) * p, cudaMemcpyHostToDevice));
  } else {
    if (freeze_dense)   // the caller's coefficients: the dense ones are kept, the group ones are the starting point
      CA_CUDA(cudaMemcpyAsync(h->beta.ptr, B, sizeof(double) * (size_t)p * kx, cudaMemcpyHostToDevice, st));
    else
      CA_CUDA(cudaMemsetAsync(h->beta.ptr, 0, h->beta.bytes(), st));
    CA_CUDA(cudaMemsetAsync(h->status.ptr, 0, h->status.bytes(), st));
    CA_CUDA(cudaMemsetAsync(h->n_iter.ptr, 0, h->n_iter.bytes(), st));
  }
  CA_CUDA(cudaMemsetAsync(h->dev_prev.ptr, 0, h->dev_prev.bytes(), st));
  CA_CUDA(cudaMemsetAsync(h->step.ptr, 0x7f, h->step.bytes(), st));  // "large": expected-information weights first
  h->has_offset = offset != nullptr;
  h->family = family;
  h->thres = thres_disp;
  h->has_nu = (nu != nullptr) && family == CA_FAMILY_NB;
  h->lay_kxi = kxi;
  h->lay_t0 = lay.t0;
  memcpy(h->lay_cmap, lay.cmap, sizeof h->lay_cmap);
  h->mu_valid = false;
  if (h->genelist.alloc(p)) return -1;
  // ---- persistent per-octet IRLS (glm_irls_impl.cuh): one launch for the whole fit ----
  bool done = false;
  {
    IrlsPlan ip;
    if (plan_irls(X, n, kx, ip)) {
      int nlist = 0;
      if (gene_mask) {
        k_glm_active_list<<<1, 1024, 0, st>>>(h->status.ptr, p, h->genelist.ptr, h->counter.ptr + 1);
        CA_LAUNCH_CHECK();
        CA_CUDA(cudaMemcpyAsync(&nlist, h->counter.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CA_CUDA(cudaStreamSynchronize(st));
      }
      const int rc = (gene_mask && nlist == 0)
                         ? 0
                         : glm_fit_irls(h, ip, X, offset, family, thres_disp, h->has_nu, ridge != nullptr, l1 != nullptr,
                                        tol, maxiter, warm, freeze_dense, gene_mask ? h->genelist.ptr : nullptr, nlist);
      if (rc < 0) return rc;
      done = rc == 0;
    }
  }
  CA_REQUIRE(done || !freeze_dense,
             "the treatment-only fit needs a design with a one-hot block of at least two columns and at most 16 others");
  if (!done) {
  k_glm_consts<<<p, 256, 0, st>>>(h->Yt, h->ldn, n, h->has_nu ? h->nu.ptr : nullptr, family, thres_disp, h->ybar.ptr,
                                  h->cst.ptr);
  CA_LAUNCH_CHECK();
  }
  GlmPass P;
  P.Yt = h->Yt;
  P.ldn = h->ldn;
  P.n = n;
  P.p = p;
  P.X = h->X.ptr;
  P.kx = kx;
  P.ldx = ldx;
  P.kxi = kxi;
  memcpy(P.cmap, lay.cmap, sizeof P.cmap);
  P.t0 = lay.t0;
  P.offset = h->offset.ptr;
  P.nu = h->has_nu ? h->nu.ptr : nullptr;
  P.family = family;
  P.thres = thres_disp;
  P.beta = h->beta.ptr;
  P.ybar = h->ybar.ptr;
  P.status = h->status.ptr;
  P.part = h->part.ptr;
  P.part_ld = kxi * kxi + kxi + 4;
  P.nsplit = 1;
  P.mu_t = nullptr;
  P.finalize = 0;
  P.genelist = nullptr;
  P.n_list = 0;
  P.fm_tables = fastmath_tables();
  P.newton_step = l1 ? h->step.ptr : nullptr;
  CA_REQUIRE(P.fm_tables, "fastmath tables unavailable");
  const int solve_warps = 4;
  const size_t solve_sm =
      sizeof(double) * solve_warps * ((size_t)kx * (kx + 1) + 3 * kx + (l1 ? (size_t)kx * (kx + 1) + 3 * kx : 0));
  CA_CUDA(cudaFuncSetAttribute(k_glm_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_sm));
  GlmSolve Q;
  Q.p = p;
  Q.kx = kx;
  Q.kxi = kxi;
  memcpy(Q.emap, lay.emap, sizeof Q.emap);
  Q.part_ld = P.part_ld;
  Q.part = h->part.ptr;
  Q.ridge = ridge ? h->ridge.ptr : nullptr;
  Q.l1 = l1 ? h->ridge.ptr + kx : nullptr;
  Q.nu = P.nu;
  Q.family = family;
  Q.thres = thres_disp;
  Q.cst = h->cst.ptr;
  Q.dev = h->dev.ptr;
  Q.dev_prev = h->dev_prev.ptr;
  Q.step = h->step.ptr;
  Q.steptol = 1e-8;
  Q.status = h->status.ptr;
  Q.n_iter = h->n_iter.ptr;
  Q.beta = h->beta.ptr;
  Q.tol = tol;
  Q.finalize = 0;
  Q.n_active = h->counter.ptr;
  int sm_count = 148;
  {
    int devid = 0;
    cudaGetDevice(&devid);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, devid);
  }
  const int nchunks = (n + GLM_CHUNK - 1) / GLM_CHUNK;
  auto launch_solve = [&](int nslots) -> int {
    if (nslots <= 0) return 0;
    ProfScope prof(PROF_GLM_SOLVE, st);
    k_glm_solve<<<(nslots + solve_warps - 1) / solve_warps, solve_warps * 32, solve_sm, st>>>(Q);
    CA_LAUNCH_CHECK();
    return 0;
  };
  int iters_done = 0;
  int nrun = p;  // genes still running (gene_mask refits start from the flagged subset)
  for (int it = 1; it <= maxiter && !done; ++it) {
    // pass `it` accumulates, per running gene, the deviance terms at the current beta and the
    // normal equations of update `it`; the solve kernel then applies the convergence test of
    // update it-1 and, for genes that keep running, performs update `it`
    if (it > 1 || gene_mask) {
      k_glm_active_list<<<1, 1024, 0, st>>>(h->status.ptr, p, h->genelist.ptr, h->counter.ptr + 1);
      CA_LAUNCH_CHECK();
      if (gene_mask && it == 1) {
        CA_CUDA(cudaMemcpyAsync(&nrun, h->counter.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CA_CUDA(cudaStreamSynchronize(st));
      }
      P.genelist = h->genelist.ptr;
      P.n_list = nrun;
    }
    if (nrun <= 0) break;
    // few running genes: split the cells of each gene over several CTAs so that the tail
    // iterations are not bound by the latency of one CTA walking all n cells
    const int ctas = (nrun + GLM_WARPS - 1) / GLM_WARPS;
    int nsplit = 1;
    if (ctas < 2 * sm_count) nsplit = std::min(std::min(nchunks, 8), std::max(1, (2 * sm_count) / ctas));
    while (nsplit > 1 && (long)nrun * nsplit > (long)p) --nsplit;
    P.nsplit = nsplit;
    Q.genelist = P.genelist;
    Q.nslots = nrun;
    Q.nsplit = nsplit;
    // tail (few running genes): enqueue several iterations per host synchronisation; genes that
    // converge in between are skipped on the device through their status flag
    const int burst = (nrun <= 64) ? std::min(8, maxiter - it + 1) : 1;
    int nact = 0;
    for (int b = 0; b < burst; ++b) {
      P.iter = it + b + (warm ? 1 : 0);  // a warm refit has no "first" pass from mu0 = (y + ybar) / 2
      Q.iter = it + b + (warm ? 1 : 0);
      CA_CUDA(cudaMemsetAsync(h->counter.ptr, 0, sizeof(int), st));
      if (launch_pass(P, st)) return -1;
      if (launch_solve(nrun)) return -1;
    }
    CA_CUDA(cudaMemcpyAsync(&nact, h->counter.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    it += burst - 1;
    if (nact == 0) break;  // every gene has converged (its iteration count was recorded on the device)
    nrun = nact;
    iters_done = it;
  }
  if (!done) {
    // finalize: fitted means + deviance at the final beta for every gene
    P.genelist = nullptr;
    P.n_list = 0;
    P.nsplit = 1;
    P.finalize = 1;
    P.iter = 2;  // never "first" in finalize
    P.mu_t = h->mu_t.ptr;
    if (launch_pass(P, st)) return -1;
    Q.genelist = nullptr;
    Q.nslots = p;
    Q.nsplit = 1;
    Q.finalize = 1;
    if (launch_solve(p)) return -1;
    h->mu_valid = true;
  }
  h->iters_done = iters_done;
  // genes still "running" after maxiter updates: n_iter = maxiter
  if (d_guard >= 0) {
    k_glm_guard<<<(p + 127) / 128, 128, 0, st>>>(p, kx, d_guard, h->beta.ptr, h->status.ptr);
    CA_LAUNCH_CHECK();
  }
  std::vector<int> hs(p), hi(p);
  CA_CUDA(cudaMemcpyAsync(B, h->beta.ptr, sizeof(double) * (size_t)p * kx, cudaMemcpyDeviceToHost, st));
  if (dev) CA_CUDA(cudaMemcpyAsync(dev, h->dev.ptr, sizeof(double) * p, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(hs.data(), h->status.ptr, sizeof(int) * p, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(hi.data(), h->n_iter.ptr, sizeof(int) * p, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  for (int j = 0; j < p; ++j) {
    int s = hs[j];
    int it = hi[j];
    if (s == 0) it = maxiter;  // hit maxiter without meeting the tolerance
    // report: 0 ok (converged or maxiter), 1 guard tripped, 2 numeric failure
    int rep = (s == 3) ? 1 : (s == 2 ? 2 : 0);
    if (status) status[j] = rep;
    if (n_iter) n_iter[j] = it;
  }
  return 0;
}

// Fitted means of the last fit, computed on demand (the persistent IRLS kernel does not write them:
// only the residuals of the first initialisation fit and the dispersion estimate need them).
static int glm_ensure_mu(ca_glm_t h) {
  if (h->mu_valid) return 0;
  CA_REQUIRE(h->fit_kx > 0 && h->X.ptr && h->beta.ptr, "fit first");
  const int kxi = h->lay_kxi;
  GlmPass P;
  memset(&P, 0, sizeof P);
  P.Yt = h->Yt;
  P.ldn = h->ldn;
  P.n = h->n;
  P.p = h->p;
  P.X = h->X.ptr;
  P.kx = h->fit_kx;
  P.ldx = 8 * ((kxi + 7) / 8) + 4;
  P.kxi = kxi;
  memcpy(P.cmap, h->lay_cmap, sizeof P.cmap);
  P.t0 = h->lay_t0;
  P.offset = h->offset.ptr;
  P.nu = h->has_nu ? h->nu.ptr : nullptr;
  P.family = h->family;
  P.thres = h->thres;
  P.beta = h->beta.ptr;
  P.ybar = h->ybar.ptr;
  P.status = h->status.ptr;
  P.part = h->part.ptr;
  P.part_ld = kxi * kxi + kxi + 4;
  P.nsplit = 1;
  P.mu_t = h->mu_t.ptr;
  P.iter = 2;
  P.finalize = 1;
  P.genelist = nullptr;
  P.n_list = 0;
  P.fm_tables = fastmath_tables();
  P.newton_step = nullptr;
  CA_REQUIRE(P.fm_tables, "fastmath tables unavailable");
  if (launch_pass(P, h->st)) return -1;
  h->mu_valid = true;
  return 0;
}

extern "C" int ca_glm_resid(ca_glm_t h, double* resid) {
  CA_REQUIRE(h && h->mu_t.ptr && resid, "fit first");
  if (glm_ensure_mu(h)) return -1;
  const int n = h->n, p = h->p;
  DevBuf<double> out;
  if (out.alloc((size_t)n * p)) return -1;
  dim3 grid((n + 31) / 32, (p + 31) / 32), block(32, 8);
  k_glm_resid<<<grid, block, 0, h->st>>>(h->Yt, h->mu_t.ptr, h->ldn, n, p, h->has_nu ? h->nu.ptr : nullptr, h->family,
                                         h->thres, h->status.ptr, out.ptr, p, nullptr);
  CA_LAUNCH_CHECK();
  CA_CUDA(cudaMemcpyAsync(resid, out.ptr, sizeof(double) * (size_t)n * p, cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_glm_resid_device(ca_glm_t h, void** dev_ptr) {
  CA_REQUIRE(h && h->mu_t.ptr && dev_ptr, "fit first");
  if (glm_ensure_mu(h)) return -1;
  const int n = h->n, p = h->p;
  if (h->resid.alloc((size_t)n * p)) return -1;
  dim3 grid((n + 31) / 32, (p + 31) / 32), block(32, 8);
  k_glm_resid<<<grid, block, 0, h->st>>>(h->Yt, h->mu_t.ptr, h->ldn, n, p, h->has_nu ? h->nu.ptr : nullptr, h->family,
                                         h->thres, h->status.ptr, h->resid.ptr, p, nullptr);
  CA_LAUNCH_CHECK();
  CA_CUDA(cudaStreamSynchronize(h->st));
  *dev_ptr = h->resid.ptr;
  return 0;
}

extern "C" int ca_glm_resid_device2(ca_glm_t h, void** dev_ptr, void** dev_ptr_t, int64_t* ldt) {
  CA_REQUIRE(h && h->mu_t.ptr && dev_ptr && dev_ptr_t && ldt, "fit first");
  if (glm_ensure_mu(h)) return -1;
  const int n = h->n, p = h->p;
  if (h->resid.alloc((size_t)n * p) || h->resid_t.alloc((size_t)p * h->ldn)) return -1;
  CA_CUDA(cudaMemsetAsync(h->resid_t.ptr, 0, h->resid_t.bytes(), h->st));
  dim3 grid((n + 31) / 32, (p + 31) / 32), block(32, 8);
  k_glm_resid<<<grid, block, 0, h->st>>>(h->Yt, h->mu_t.ptr, h->ldn, n, p, h->has_nu ? h->nu.ptr : nullptr, h->family,
                                         h->thres, h->status.ptr, h->resid.ptr, p, h->resid_t.ptr);
  CA_LAUNCH_CHECK();
  CA_CUDA(cudaStreamSynchronize(h->st));
  *dev_ptr = h->resid.ptr;
  *dev_ptr_t = h->resid_t.ptr;
  *ldt = h->ldn;
  return 0;
}

extern "C" int ca_glm_fitted(ca_glm_t h, double* mu) {
  CA_REQUIRE(h && h->mu_t.ptr && mu, "fit first");
  if (glm_ensure_mu(h)) return -1;
  const int n = h->n, p = h->p;
  DevBuf<double> out;
  if (out.alloc((size_t)n * p)) return -1;
  if (launch_transpose(h->mu_t.ptr, p, n, h->ldn, out.ptr, p, h->st)) return -1;
  CA_CUDA(cudaMemcpyAsync(mu, out.ptr, sizeof(double) * (size_t)n * p, cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_glm_disp_moments(ca_glm_t h, const double* offset, double* disp) {
  CA_REQUIRE(h && h->mu_t.ptr && disp, "fit first");
  if (glm_ensure_mu(h)) return -1;
  const int n = h->n, p = h->p;
  DevBuf<double> off, out;
  if (out.alloc(p)) return -1;
  if (offset) {
    if (off.alloc(n)) return -1;
    CA_CUDA(cudaMemcpyAsync(off.ptr, offset, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
  }
  k_disp_moments<<<p, 256, 0, h->st>>>(h->Yt, h->mu_t.ptr, h->ldn, n, offset ? off.ptr : nullptr, out.ptr);
  CA_LAUNCH_CHECK();
  CA_CUDA(cudaMemcpyAsync(disp, out.ptr, sizeof(double) * p, cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}
