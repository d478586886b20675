// K4, second form: the WHOLE per-gene IRLS of eight genes inside one CTA ("octet"), no launch and
// no host round trip per iteration (replaces the statsmodels fit of causarray/gcate_glm.py:204;
// algorithm restated in oracle/glm.py).
//
// Why: the launch-per-iteration form (glm_pass_impl.cuh) spends 7-10 DMMA tiles per 4 cells on ONE
// gene (448-640 multiply-adds for 351 + 26 useful products at kx = 26), pays a launch, a solve
// launch and a host synchronisation per iteration, and its tail iterations (few running genes)
// run on a nearly empty machine: 0.32 of the FP64 peak over a benchmark step.
//
// Formulation.  The design is split by the host (glm_irls.cu: plan_groups) into KDE "dense" columns
// (the last one is the unit column: the intercept, or an appended column of ones) and a one-hot
// block (treatment indicators of [W, A]): every cell belongs to at most one "group" g >= 1, cells
// are sorted by group and every group is padded to a multiple of 8 cells with all-zero design rows.
// Inside a group the Gram matrix only needs the dense pair products:
//     G[i][j]     = sum_cells      w x_i x_j            (dense x dense, all cells)
//     G[i][a_g]   = sum_cells in g w x_i * 1            (the (i, unit) products of group g)
//     G[a_g][a_g] = sum_cells in g w                    (the (unit, unit) product of group g)
// i.e. 66 + 11 instead of 351 + 26 products per cell and gene for [1, U(10) | A(15)].
// The contraction over cells runs on the FP64 tensor cores with the GENES as the n-dimension:
//     C[pair row, gene] += P[pair row, cell] * w[cell, gene]            (DMMA m8n8k4)
// A lane (lr = lane / 4, lc = lane % 4) evaluates eta / mu / w / z of cell lc of a 4-cell step for
// gene lr -- exactly the B fragment -- and the A fragment of "strip" tile (b, j) is
// x[cell lc][8 b + lr] * x[cell lc][j] with j a compile-time index: one LDS per row block and step,
// no table lookups.  Strips (b, j >= 8 b) cover every unordered pair; the tiles of the unit column
// j = u and the last right-hand-side tile are flushed to scratch at every group boundary, which
// yields the per-group sums above.
//
// A CTA of 8 warps owns one octet: the warps split the (sorted, padded) cells into contiguous
// chunk ranges, each with its own double-buffered bulk-async staging of the design rows (no CTA
// barrier inside a pass), the partial tiles are added in warp order (deterministic), warp g then
// assembles and solves gene g's normal equations (glm_solve.cuh) and tests convergence on the sum
// of the per-cell unit deviances (the form statsmodels sums, oracle/glm.py:62-75).  Octets are
// handed out through an atomic counter; an octet runs until its slowest gene has converged
// (+15 % work at the benchmark shape, measured from the iteration histogram).
#pragma once
#include "common.cuh"
#include "fastmath.cuh"
#include "glm_solve.cuh"
#include <type_traits>

namespace ca {

constexpr int IRLS_WARPS = 8;
constexpr int IRLS_THREADS = IRLS_WARPS * 32;
constexpr int IRLS_CHUNK = 32;    // cells per staged chunk (per warp): four pair-steps of 8 cells
constexpr int IRLS_MAX_KDE = 16;
constexpr int IRLS_MAX_SEG = 96;

struct GlmIrls {
  const double* Yp;        // p x ldp raw counts, gene-major, cells in group-sorted padded order
  long ldp;
  int p;
  const double* Xd;        // (32 nchunks) x ldxd: dense columns (unit last), then the offset; zero rows pad
  int ldxd;
  int nchunks;
  int chunk_lo[IRLS_WARPS + 1];        // warp w walks chunks [chunk_lo[w], chunk_lo[w + 1])
  const int* gpair;                    // group of every 8-cell pair-step (4 nchunks entries)
  int nseg_tot;                        // flush segments of one pass, in (warp, time) order
  int seg_ptr[IRLS_WARPS + 1];         // first segment of warp w
  unsigned char gseg_ptr[66];          // segments of group g: gseg_list[gseg_ptr[g] .. gseg_ptr[g + 1])
  unsigned char gseg_list[IRLS_MAX_SEG];
  int ngroups;                         // including group 0 (no one-hot column set)
  int kx, kde;
  signed char dcol[IRLS_MAX_KDE];      // dense internal column -> caller's column (-1: appended unit column)
  signed char gcol[64];                // group -> caller's column (gcol[0] = -1)
  signed char cdense[64];              // caller's column -> dense internal column or -1
  unsigned char cgroup[64];            // caller's column -> group (0: a dense column)
  const double* nu;                    // p or nullptr
  int family;
  double thres;
  const double* ybar;                  // p
  const double* ridge;                 // kx or nullptr
  const double* l1;                    // kx or nullptr (lasso refit)
  double tol, steptol;
  int maxiter;
  int freeze_dense;                    // 1: the dense coefficients stay at their input values; only the one-hot
                                       // group coefficients are updated (two-stage fit: per-perturbation 1-D
                                       // fits against the covariate offset, causarray/nb_glm_fast.py:500-567)
  int iter0;                           // 0: cold start (first pass from mu0 = (y + ybar) / 2); 1: warm start from beta
  const int* genelist;                 // genes to fit (nullptr: all p genes)
  int n_list;
  int* counter;                        // octet dispenser (zeroed by the host)
  double* scratch;                     // per CTA: nseg_tot x (NB + 1) x 64 doubles
  double* beta;                        // p x kx (in: warm start; out)
  double* dev;                         // p
  double* step;                        // p (lasso refit: relative size of the last update)
  int* status;                         // p : 0 running (hit maxiter), 1 converged, 2 numeric failure
  int* n_iter;                         // p
  const double* fm_tables;
};

template <int KDE>
struct IrlsShape {
  static constexpr int NB = (KDE + 7) / 8;             // row blocks
  static constexpr int U = KDE - 1;                    // unit column
  static constexpr int NTW = NB * KDE - 8 * (NB * (NB - 1) / 2);  // strip tiles: sum_b (KDE - 8 b)
  static constexpr int NT = NTW + NB;                  // + right-hand-side tiles
  static constexpr int FLT = NB + 1;                   // tiles flushed per group: (b, U) for every b, rhs(NB - 1)
  static constexpr int LDXD = (KDE + 1) | 1;           // dense columns + offset; odd: a half-warp reading one
                                                       // row per lane then touches 16 distinct bank pairs
  __host__ __device__ static constexpr int tw(int b, int j) { return b * KDE - 8 * (b * (b - 1) / 2) + (j - 8 * b); }
};

inline size_t glm_irls_smem_bytes(int kde, int kx, bool lasso) {
  const int nb = (kde + 7) / 8;
  const int ntw = nb * kde - 8 * (nb * (nb - 1) / 2);
  const int nt = ntw + nb;
  const int ldxd = (kde + 1) | 1;
  size_t d = 0;
  d += (size_t)IRLS_WARPS * 2 * (((IRLS_CHUNK * ldxd + 8) + 1) & ~1);  // staging (+ slack for the masked row-block reads)
  d += (size_t)IRLS_WARPS * 2 * IRLS_CHUNK * 9;               // weights / weighted responses handed from the evaluation to the DMMA phase
  d += (size_t)8 * (kde + 4);                                 // dense coefficients, nu, ybar, flags of the octet
  d += (size_t)nt * 64;                                       // dense totals
  d += (size_t)IRLS_WARPS * 8;                                // deviance partials
  d += (size_t)8 * kx;                                        // coefficients of the octet
  d += (size_t)IRLS_WARPS * glm_solve_work_doubles(kx, lasso);
  d += FM_TABLE_DOUBLES;
  return d * sizeof(double);
}

template <int KDE, bool NBF>
__global__ void __launch_bounds__(IRLS_THREADS, 1) k_glm_irls(const __grid_constant__ GlmIrls P) {
  using S = IrlsShape<KDE>;
  constexpr int NB = S::NB, U = S::U, NTW = S::NTW, NT = S::NT, FLT = S::FLT, LDXD = S::LDXD;
  constexpr int STG = ((IRLS_CHUNK * LDXD + 8) + 1) & ~1;   // even: every stage starts 16-byte aligned (bulk copies)
  constexpr int WLD = 9;                                    // row stride of the weight hand-off buffers (odd)
  extern __shared__ __align__(16) double smem[];
  const int kx = P.kx;
  const bool lasso = P.l1 != nullptr;
  double* xs_all = smem;                                   // IRLS_WARPS * 2 * STG
  double* Tt = xs_all + IRLS_WARPS * 2 * STG;              // NT * 64
  double* Dp = Tt + NT * 64;                               // IRLS_WARPS * 8
  double* betas = Dp + IRLS_WARPS * 8;                     // 8 * kx
  double* work_all = betas + 8 * kx;
  const size_t work_d = glm_solve_work_doubles(kx, lasso);
  double* wbuf_all = work_all + IRLS_WARPS * work_d;       // IRLS_WARPS * 2 * IRLS_CHUNK * WLD
  double* bdense = wbuf_all + IRLS_WARPS * 2 * IRLS_CHUNK * WLD;  // 8 * KDE
  double* s_nu = bdense + 8 * KDE;                         // 8
  double* s_ybar = s_nu + 8;                               // 8
  double* s_flag = s_ybar + 8;                             // 8: bit 0 Poisson branch, bit 1 observed-information weights
  double* s_tab = s_flag + 16;
  __shared__ __align__(8) unsigned long long bars[IRLS_WARPS * 2];
  __shared__ int s_octet, s_nrun, s_anynewton;
  __shared__ int s_gene[8], s_status[8], s_niter[8];
  __shared__ double s_devprev[8], s_dev[8], s_step[8];
  __shared__ signed char s_gcol[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const FmTables FT = fm_load_tables(s_tab, P.fm_tables);
  double* xs = xs_all + warp * 2 * STG;
  unsigned long long* bar = bars + warp * 2;
  if (lane == 0) {
    mbar_init(bar + 0, 1);
    mbar_init(bar + 1, 1);
  }
  for (int i = lane; i < 2 * STG; i += 32) xs[i] = 0.0;
  if (tid < 64) s_gcol[tid] = P.gcol[tid];
  if (tid == 0) mbar_fence_init();
  __syncthreads();
  fence_proxy_async();
  unsigned fphase = 0;

  const int c_lo = P.chunk_lo[warp], c_hi = P.chunk_lo[warp + 1];
  const int n_oct = ((P.genelist ? P.n_list : P.p) + 7) >> 3;
  double* scratch = P.scratch + (size_t)blockIdx.x * P.nseg_tot * (FLT * 64);
  constexpr double LOG_EPS = -36.04365338911715;  // log(2^-52): statsmodels clips y / mu at FLOAT_EPS

  for (;;) {
    if (tid == 0) s_octet = atomicAdd(P.counter, 1);
    __syncthreads();
    const int oct = s_octet;
    if (oct >= n_oct) break;
    // ---- octet set-up ----
    if (tid < 8) {
      const int slot = oct * 8 + tid;
      const int nl = P.genelist ? P.n_list : P.p;
      const int g = slot < nl ? (P.genelist ? P.genelist[slot] : slot) : -1;
      s_gene[tid] = g;
      s_status[tid] = g >= 0 ? 0 : 1;
      s_niter[tid] = 0;
      s_devprev[tid] = 0.0;
      s_dev[tid] = 0.0;
      s_step[tid] = 1e300;   // "large": expected-information weights first
    }
    __syncthreads();
    for (int e = tid; e < 8 * kx; e += IRLS_THREADS) {
      const int g = s_gene[e / kx];
      betas[e] = (g >= 0 && P.iter0) ? P.beta[(size_t)g * kx + (e % kx)] : 0.0;
    }
    __syncthreads();
    // per-gene constants of the octet (a padding slot of the last octet borrows gene 0's row and is never stored)
    if (tid < 8) {
      const int g = s_gene[tid] >= 0 ? s_gene[tid] : (s_gene[0] >= 0 ? s_gene[0] : 0);
      double nu = 1.0;
      bool pois = true;
      if (NBF && P.nu) {
        nu = P.nu[g];
        pois = nu > P.thres;
      }
      s_nu[tid] = nu;
      s_ybar[tid] = P.ybar[g];
      s_flag[tid] = pois ? 1.0 : 0.0;
    }
    __syncthreads();
    const double* ybase = P.Yp + (size_t)(c_lo < c_hi ? c_lo : 0) * IRLS_CHUNK + lane;
    long yoff[8];
#pragma unroll
    for (int g = 0; g < 8; ++g)
      yoff[g] = (long)(s_gene[g] >= 0 ? s_gene[g] : (s_gene[0] >= 0 ? s_gene[0] : 0)) * P.ldp;
    double* wbuf = wbuf_all + warp * 2 * IRLS_CHUNK * WLD;   // w, then w z: [cell of the chunk][gene], row stride WLD
    double* wzbuf = wbuf + IRLS_CHUNK * WLD;

    for (int it = 1;; ++it) {
      const int iter = it + P.iter0;            // 1: the pass from mu0 = (y + ybar) / 2
      const bool first = iter == 1;
      // dense coefficients of the 8 genes in the internal column order; observed-information switch
      for (int e = tid; e < 8 * KDE; e += IRLS_THREADS) {
        const int g = e / KDE, i = e - g * KDE;
        bdense[e] = P.dcol[i] >= 0 ? betas[g * kx + P.dcol[i]] : 0.0;
      }
      if (tid == 0) s_anynewton = 0;
      __syncthreads();
      if (tid < 8) {
        const bool pois = ((int)s_flag[tid] & 1) != 0;
        const bool nwt = NBF && lasso && !first && !pois && s_step[tid] <= 1e-2;
        s_flag[tid] = (pois ? 1.0 : 0.0) + (nwt ? 2.0 : 0.0);
        if (nwt) s_anynewton = 1;
      }
      __syncthreads();
      const bool any_newton = s_anynewton != 0;
      // ---------------- pass over this warp's chunks ----------------
      double acc[NT][2];
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
      double dacc[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) dacc[g] = 0.0;
      int gcur = -1, seg = P.seg_ptr[warp];
      auto issue = [&](int c, int stage) {
        const unsigned bytes = (unsigned)(IRLS_CHUNK * LDXD * sizeof(double));
        mbar_expect_tx(bar + stage, bytes);
        bulk_g2s(xs + stage * STG, P.Xd + (size_t)c * IRLS_CHUNK * LDXD, bytes, bar + stage);
      };
      auto flush = [&]() {
        double* dst = scratch + (size_t)seg * (FLT * 64) + lr * 8 + 2 * lc;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          *reinterpret_cast<double2*>(dst + b * 64) = make_double2(acc[S::tw(b, U)][0], acc[S::tw(b, U)][1]);
          acc[S::tw(b, U)][0] = acc[S::tw(b, U)][1] = 0.0;
        }
        *reinterpret_cast<double2*>(dst + NB * 64) = make_double2(acc[NTW + NB - 1][0], acc[NTW + NB - 1][1]);
        acc[NTW + NB - 1][0] = acc[NTW + NB - 1][1] = 0.0;
        ++seg;
      };
      if (lane == 0) {
        if (c_lo < c_hi) issue(c_lo, 0);
        if (c_lo + 1 < c_hi) issue(c_lo + 1, 1);
      }
      // counts of this lane's cell of the chunk, one per gene (256 contiguous bytes per gene and load)
      double ycur[8], ynext[8];
      if (c_lo < c_hi) {
#pragma unroll
        for (int g = 0; g < 8; ++g) ycur[g] = ybase[yoff[g]];
      }
      for (int c = c_lo; c < c_hi; ++c) {
        const int stage = (c - c_lo) & 1;
        if (c + 1 < c_hi) {
#pragma unroll
          for (int g = 0; g < 8; ++g) ynext[g] = ybase[yoff[g] + (size_t)(c + 1 - c_lo) * IRLS_CHUNK];
        }
        mbar_wait(bar + stage, (fphase >> stage) & 1u);
        fphase ^= 1u << stage;
        const double* xc = xs + stage * STG;
        // ===== evaluation: lane = cell of the chunk, the 8 genes are 8 independent chains =====
        // (one straight-line body per mode -- first pass / regular / observed-information weights --
        // so that ptxas interleaves the eight chains: with the mode tests inside the gene loop every
        // gene was its own basic block and the chains ran one after the other)
        {
          const double* xr = xc + lane * LDXD;
          double x[KDE];
#pragma unroll
          for (int i = 0; i < KDE; ++i) x[i] = xr[i];
          const double off = xr[KDE];
          const double vld = x[U];                 // 1 for a cell, 0 for a padding row
          const int grp = __ldg(P.gpair + c * 4 + (lane >> 3));
          const int gc = grp > 0 ? (int)s_gcol[grp] : 0;
          const double gsel = grp > 0 ? 1.0 : 0.0;
          auto eval8 = [&](auto first_c, auto newton_c) {
            constexpr bool FIRST = decltype(first_c)::value;
            constexpr bool NEWTON = decltype(newton_c)::value;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const double y = ycur[g];
              const double nu = s_nu[g];
              const int fl = (int)s_flag[g];
              const bool pois = (fl & 1) != 0;
              double eta, ec, mu;
              if (FIRST) {
                mu = 0.5 * (y + s_ybar[g]);
                eta = fm_log(mu, FT.log_t);
                ec = eta;
              } else {
                const double* bg = bdense + g * KDE;
                double ea = fma(gsel, betas[g * kx + gc], off), eb = 0.0;
#pragma unroll
                for (int i = 0; i + 1 < KDE; i += 2) {
                  ea = fma(x[i], bg[i], ea);
                  eb = fma(x[i + 1], bg[i + 1], eb);
                }
                if (KDE & 1) ea = fma(x[KDE - 1], bg[KDE - 1], ea);
                eta = ea + eb;
                ec = eta < 700.0 ? eta : 700.0;
                ec = ec > -700.0 ? ec : -700.0;
                mu = fm_exp(ec, FT.exp_t);
              }
              // unit deviance / 2 : y log(max(y / mu, eps)) - (y - mu)                          (Poisson)
              //                     y log(max(y / mu, eps)) - (y + nu) log((y + nu) / (mu + nu)) (NB)
              const double ly = fm_log(y > 0.0 ? y : 1.0, FT.log_t);
              double lrat = ly - ec;
              lrat = lrat > LOG_EPS ? lrat : LOG_EPS;
              double d2 = y - mu;
              double wgt = mu;
              double rc = 0.0;
              if (NBF) {
                rc = fm_rcp(nu + mu);
                const double lt = fm_log((y + nu) * rc, FT.log_t);
                d2 = pois ? d2 : (y + nu) * lt;
                wgt = pois ? mu : mu * nu * rc;
              }
              dacc[g] = fma(vld, fma(y, lrat, -d2), dacc[g]);
              const double z = (eta - off) + (y - mu) * fm_rcp(mu);
              double w = wgt, wz = wgt * z;
              if (NBF && NEWTON) {
                // observed information of the NB2 log-likelihood: nu mu (nu + y) / (nu + mu)^2;
                // rhs = w (eta - off) + score (same minimiser, quadratic tail; lasso refit only)
                const double wn = mu * nu * rc * ((nu + y) * rc);
                const double wzn = fma(wn, eta - off, (y - mu) * (nu * rc));
                const bool nw = (fl & 2) != 0;
                w = nw ? wn : w;
                wz = nw ? wzn : wz;
              }
              wbuf[lane * WLD + g] = w;
              wzbuf[lane * WLD + g] = wz;
            }
          };
          if (first)
            eval8(std::true_type{}, std::false_type{});
          else if (any_newton)
            eval8(std::false_type{}, std::true_type{});
          else
            eval8(std::false_type{}, std::false_type{});
        }
        __syncwarp();
        // ===== Gram strips and right-hand side on the FP64 tensor cores: lane = (cell lc of a step, gene lr) =====
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const int g = __ldg(P.gpair + c * 4 + q);
          if (g != gcur) {
            if (gcur >= 0) flush();
            gcur = g;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cell = (2 * q + h) * 4 + lc;
            const double* xr = xc + cell * LDXD;
            const double w = wbuf[cell * WLD + lr], wz = wzbuf[cell * WLD + lr];
            double xa[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const double t = xr[8 * b + lr];
              xa[b] = (8 * b + lr < KDE) ? t : 0.0;
            }
#pragma unroll
            for (int j = 0; j < KDE; ++j) {
              const double bw = w * xr[j];
#pragma unroll
              for (int b = 0; b < NB; ++b)
                if (j >= 8 * b) dmma884(acc[S::tw(b, j)][0], acc[S::tw(b, j)][1], xa[b], bw);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) dmma884(acc[NTW + b][0], acc[NTW + b][1], xa[b], wz);
          }
        }
        __syncwarp();
        if (lane == 0 && c + 2 < c_hi) issue(c + 2, stage);
#pragma unroll
        for (int g = 0; g < 8; ++g) ycur[g] = ynext[g];
      }
      if (gcur >= 0) flush();
      // ---------------- combine the warps (fixed order) ----------------
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const double v = warp_sum(dacc[g]);
        if (lane == g) Dp[warp * 8 + g] = v;
      }
      for (int w = 0; w < IRLS_WARPS; ++w) {
        if (warp == w) {
          double* dst = Tt + lr * 8 + 2 * lc;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            double2* pp = reinterpret_cast<double2*>(dst + t * 64);
            if (w == 0) {
              *pp = make_double2(acc[t][0], acc[t][1]);
            } else {
              double2 v = *pp;
              v.x += acc[t][0];
              v.y += acc[t][1];
              *pp = v;
            }
          }
        }
        __syncthreads();
      }
      if (tid == 0) s_nrun = 0;
      __syncthreads();
      // ---------------- warp g: convergence test and update of gene g ----------------
      {
        const int g = warp;
        const int gene = s_gene[g];
        bool run = gene >= 0 && s_status[g] == 0;
        if (run) {
          double dev = 0.0;
          for (int w = 0; w < IRLS_WARPS; ++w) dev += Dp[w * 8 + g];
          dev *= 2.0;
          if (lasso) {
            double pen = 0.0;
            for (int i = 0; i < kx; ++i) pen += P.l1[i] * fabs(betas[g * kx + i]);
            dev += 2.0 * pen;
          }
          const double prev = s_devprev[g];
          int newst = 0;
          if (!(dev == dev) || isinf(dev)) {
            newst = 2;
          } else if (iter > 1 && fabs(dev - prev) <= P.tol && (!lasso || s_step[g] <= P.steptol)) {
            newst = 1;
          }
          __syncwarp();
          if (lane == 0) {
            s_dev[g] = dev;
            if (newst) {
              s_status[g] = newst;
              s_niter[g] = iter - 1;
            } else {
              s_devprev[g] = dev;
            }
          }
          if (newst) run = false;
          if (run && it > P.maxiter) {
            // maxiter updates done and the last test failed: keep status 0 (the host reports n_iter = maxiter)
            run = false;
            if (lane == 0) s_status[g] = 3;
          }
        }
        if (run) {
          const double* sc0 = scratch + g;   // column g of every flushed tile
          int newst = 0;
          if (!lasso) {
            // Arrow structure of the normal equations: dense block D (KDE columns) + one diagonal entry
            // d_g per one-hot group, coupled through c_g = (the (i, unit) products of group g):
            //   G_DD b_D + sum_g c_g b_g = r_D,   c_g^T b_D + d_g b_g = r_g
            // => (G_DD - sum_g c_g c_g^T / d_g) b_D = r_D - sum_g c_g r_g / d_g,  b_g = (r_g - c_g^T b_D) / d_g:
            // an 11 x 11 Cholesky instead of a 26 x 26 one for [1, U(10) | A(15)] (the full solve and the
            // kx^2 assembly were 32 % of the kernel's time, all of it with the FP64 pipe idle).
            const int nd = P.dcol[U] >= 0 ? KDE : KDE - 1;     // an appended unit column is not a coefficient
            const int ngr = P.ngroups - 1;
            double* wk = work_all + (size_t)g * work_d;
            const GlmSolveWork W = glm_solve_carve(wk, nd);
            double* cg = wk + glm_solve_work_doubles(nd, false);  // ngr x (KDE + 1): c_g[0..KDE-1] (entry U = d_g), r_g
            const int ld = nd + 1;
            const int CL = KDE + 1;
            for (int e = lane; e < ngr * CL; e += 32) {
              const int gg = e / CL + 1, i = e - (gg - 1) * CL;
              double v = 0.0;
              const int off = i < KDE ? (i >> 3) * 64 + (i & 7) * 8 : NB * 64 + (U & 7) * 8;
              for (int q = P.gseg_ptr[gg]; q < P.gseg_ptr[gg + 1]; ++q) v += sc0[(size_t)P.gseg_list[q] * (FLT * 64) + off];
              if (i == U && P.ridge) v += P.ridge[P.gcol[gg]];
              cg[e] = v;
            }
            for (int e = lane; e < nd * nd; e += 32) {
              const int i = e / nd, j = e - i * nd;
              const int a = i < j ? i : j, bb = i < j ? j : i;
              double v = 0.0;
              if (bb == U) {
                for (int s2 = 0; s2 < P.nseg_tot; ++s2) v += sc0[(size_t)s2 * (FLT * 64) + (a >> 3) * 64 + (a & 7) * 8];
              } else {
                v = Tt[S::tw(a >> 3, bb) * 64 + (a & 7) * 8 + g];
              }
              if (P.ridge && i == j) v += P.ridge[P.dcol[i]];
              W.L[i * ld + j] = v;
            }
            for (int i = lane; i < nd; i += 32) {
              double v = 0.0;
              if ((i >> 3) == NB - 1) {
                for (int s2 = 0; s2 < P.nseg_tot; ++s2) v += sc0[(size_t)s2 * (FLT * 64) + NB * 64 + (i & 7) * 8];
              } else {
                v = Tt[(NTW + (i >> 3)) * 64 + (i & 7) * 8 + g];
              }
              W.b[i] = v;
            }
            __syncwarp();
            bool bad = false;
            for (int gg = lane; gg < ngr; gg += 32)
              if (!(cg[gg * CL + U] > 0.0)) bad = true;
            bad = __any_sync(0xffffffffu, bad);
            if (P.freeze_dense) {
              // the dense block keeps its coefficients: W.b carries them (unscaled) into the group update
              for (int i = lane; i < nd; i += 32) {
                W.b[i] = betas[g * kx + P.dcol[i]];
                W.sc[i] = 1.0;
              }
              __syncwarp();
            } else if (!bad) {
              for (int e = lane; e < nd * nd; e += 32) {
                const int i = e / nd, j = e - i * nd;
                double v = W.L[i * ld + j];
                for (int gg = 0; gg < ngr; ++gg) v = fma(-cg[gg * CL + i] / cg[gg * CL + U], cg[gg * CL + j], v);
                W.L[i * ld + j] = v;
              }
              for (int i = lane; i < nd; i += 32) {
                double v = W.b[i];
                for (int gg = 0; gg < ngr; ++gg) v = fma(-cg[gg * CL + i] / cg[gg * CL + U], cg[gg * CL + KDE], v);
                W.b[i] = v;
              }
              __syncwarp();
              bad = glm_solve_system(W, nd, nullptr, lane);
            }
            if (bad) {
              newst = 2;
            } else {
              bool fin = true;
              for (int i = lane; i < nd; i += 32) {
                const double v = W.b[i] * W.sc[i];
                W.b[i] = v;
                betas[g * kx + P.dcol[i]] = v;
                fin = fin && (v == v) && !isinf(v);
              }
              __syncwarp();
              for (int gg = lane; gg < ngr; gg += 32) {
                double v = cg[gg * CL + KDE];
                for (int i = 0; i < nd; ++i) v = fma(-cg[gg * CL + i], W.b[i], v);
                v /= cg[gg * CL + U];
                betas[g * kx + P.gcol[gg + 1]] = v;
                fin = fin && (v == v) && !isinf(v);
              }
              fin = __all_sync(0xffffffffu, fin);
              if (!fin) newst = 2;
            }
          } else {
          const GlmSolveWork W = glm_solve_carve(work_all + (size_t)g * work_d, kx);
          const int ld = kx + 1;
          for (int e = lane; e < kx * kx; e += 32) {
            const int ci = e / kx, cj = e - ci * kx;
            const int di = P.cdense[ci], dj = P.cdense[cj];
            const int gi = P.cgroup[ci], gj = P.cgroup[cj];
            double v = 0.0;
            if (di >= 0 && dj >= 0) {
              const int a = di < dj ? di : dj, bb = di < dj ? dj : di;
              if (bb == U) {
                for (int s = 0; s < P.nseg_tot; ++s) v += sc0[(size_t)s * (FLT * 64) + (a >> 3) * 64 + (a & 7) * 8];
              } else {
                v = Tt[S::tw(a >> 3, bb) * 64 + (a & 7) * 8 + g];
              }
            } else if (di >= 0 || dj >= 0) {
              const int a = di >= 0 ? di : dj, gg = di >= 0 ? gj : gi;
              for (int q = P.gseg_ptr[gg]; q < P.gseg_ptr[gg + 1]; ++q)
                v += sc0[(size_t)P.gseg_list[q] * (FLT * 64) + (a >> 3) * 64 + (a & 7) * 8];
            } else if (gi == gj) {
              for (int q = P.gseg_ptr[gi]; q < P.gseg_ptr[gi + 1]; ++q)
                v += sc0[(size_t)P.gseg_list[q] * (FLT * 64) + (U >> 3) * 64 + (U & 7) * 8];
            }
            if (P.ridge && ci == cj) v += P.ridge[ci];
            W.L[ci * ld + cj] = v;
          }
          for (int c = lane; c < kx; c += 32) {
            const int di = P.cdense[c], gg = P.cgroup[c];
            double v = 0.0;
            if (di >= 0) {
              if ((di >> 3) == NB - 1) {
                for (int s = 0; s < P.nseg_tot; ++s) v += sc0[(size_t)s * (FLT * 64) + NB * 64 + (di & 7) * 8];
              } else {
                v = Tt[(NTW + (di >> 3)) * 64 + (di & 7) * 8 + g];
              }
            } else {
              for (int q = P.gseg_ptr[gg]; q < P.gseg_ptr[gg + 1]; ++q)
                v += sc0[(size_t)P.gseg_list[q] * (FLT * 64) + NB * 64 + (U & 7) * 8];
            }
            W.b[c] = v;
          }
          __syncwarp();
          if (glm_solve_system(W, kx, P.l1, lane)) {
            newst = 2;
          } else {
            bool fin = true;
            double stepmax = 0.0;
            for (int i = lane; i < kx; i += 32) {
              const double v = W.b[i] * W.sc[i];
              const double old = betas[g * kx + i];
              stepmax = fmax(stepmax, fabs(v - old) / (1.0 + fabs(v)));
              betas[g * kx + i] = v;
              fin = fin && (v == v) && !isinf(v);
            }
            for (int o = 16; o > 0; o >>= 1) stepmax = fmax(stepmax, __shfl_xor_sync(0xffffffffu, stepmax, o));
            if (lane == 0) s_step[g] = stepmax;
            fin = __all_sync(0xffffffffu, fin);
            if (!fin) newst = 2;
          }
          }
          if (lane == 0) {
            if (newst) {
              s_status[g] = newst;
              s_niter[g] = iter;
            } else {
              atomicAdd(&s_nrun, 1);
            }
          }
        }
      }
      __syncthreads();
      if (s_nrun == 0) break;
    }
    // ---- results of the octet ----
    for (int e = tid; e < 8 * kx; e += IRLS_THREADS) {
      const int g = s_gene[e / kx];
      if (g >= 0) P.beta[(size_t)g * kx + (e % kx)] = betas[e];
    }
    if (tid < 8 && s_gene[tid] >= 0) {
      const int g = s_gene[tid];
      const int st = s_status[tid];
      P.dev[g] = s_dev[tid];
      P.status[g] = st == 3 ? 0 : st;
      P.n_iter[g] = s_niter[tid];
      if (P.step) P.step[g] = s_step[tid];
    }
    __syncthreads();
  }
}

template <int KDE>
int launch_glm_irls_kde(const GlmIrls& P, int grid, size_t smem, cudaStream_t st);

}  // namespace ca
