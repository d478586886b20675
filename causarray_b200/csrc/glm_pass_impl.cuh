// K4 pass kernel of the batched per-gene IRLS (see glm.cu for the algorithm): for every running
// gene accumulate, over a range of cells, the deviance terms at the current beta and the normal
// equations  X^T diag(w) X,  X^T (w z)  of the next update.
//
// One warp owns one gene.  The shared design X (n x LDX, LDX = 8 NT + 4) streams through shared
// memory in chunks of 128 cells with bulk async copies (cp.async.bulk / UBLKCP) completed on
// "full" mbarriers; a stage is refilled once all 8 warps have arrived on its "empty" mbarrier, so
// warps drift apart by up to a chunk instead of meeting at a CTA barrier (ncu on the first version:
// the per-chunk __syncthreads kept all warps in the same phase -- all evaluating, then all issuing
// DMMAs -- and the FP64/tensor pipe idle half of the time).  Per 32 cells a lane evaluates
// eta / mu / w / z for ONE cell (straight-line code, table-driven exp / log), software-pipelined
// one block ahead of the Gram accumulation: for a 4-cell step the fragment
// X[cell = lane%4][col = lane/4 + 8t] is the DMMA B operand as is and, times w[cell], the A operand;
// only the upper-triangular 8x8 tiles are kept.  T0 < NT: the tiles from T0 on hold mutually
// orthogonal columns (GlmLayout in glm.cu) and the DMMAs between them are skipped.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace ca {

constexpr int GLM_WARPS = 8;
constexpr int GLM_THREADS = GLM_WARPS * 32;
constexpr int GLM_CHUNK = 128;  // cells per staged chunk

struct GlmPass {
  const double* Yt;      // p x ldn raw counts, gene-major
  long ldn;
  int n, p;
  const double* X;       // n x ldx in the INTERNAL column layout, ldx = 8 * ceil(kxi / 8) + 4, zero padded
  int kx, ldx;           // kx: width of the caller's design (layout of beta)
  int kxi;               // width of the internal layout (see GlmLayout in glm.cu)
  signed char cmap[64];  // internal column -> caller's column, -1 for a padding column
  int t0;                // tiles >= t0 hold mutually orthogonal columns: a tile (mt, nt) with both
                         // indices >= t0 has only its diagonal (mt == nt) non-zero (t0 = NT: none)
  const double* offset;  // n (zeros when the model has no offset)
  const double* nu;      // p or nullptr
  int family;
  double thres;
  const double* beta;    // p x kx
  const double* ybar;    // p
  int* status;           // p : 0 running, 1 converged, 2 numeric failure
  double* part;          // per (slot, split): kx*kx Gram | kx rhs | 3 deviance sums  (stride part_ld)
  int part_ld;
  int nsplit;            // cell-range splits per gene (gridDim.y); > 1 when few genes are still running
  double* mu_t;          // p x ldn or nullptr (finalize)
  int iter;              // 1-based index of the update this pass prepares
  int finalize;          // 1: only deviance (+ mu_t), no Gram
  const int* genelist;   // compacted list of running genes (nullptr = all genes)
  int n_list;            // entries of genelist
  const double* fm_tables;
  const double* newton_step;  // lasso refit only (else nullptr): relative size of each gene's last update;
                              // once it is below 1e-2 the NB weights switch from expected to observed
                              // information (same minimiser, quadratic instead of linear convergence)
};

inline size_t glm_pass_smem_bytes(int nt) {
  const int nt8 = nt * 8, ldx = nt8 + 4;
  return sizeof(double) * ((size_t)2 * GLM_CHUNK * ldx + 2 * GLM_CHUNK + GLM_WARPS * nt8 + FM_TABLE_DOUBLES);
}

template <int NT, int T0, bool FIRST, bool FIN, bool NB>
__global__ void __launch_bounds__(GLM_THREADS, (NT <= 5 ? 2 : 1)) k_glm_pass(GlmPass P) {
  constexpr int NT8 = NT * 8, LDX = NT8 + 4;
  constexpr int NTRI = NT * (NT + 1) / 2;
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;                            // 2 * GLM_CHUNK * LDX
  double* offs = Xs + 2 * GLM_CHUNK * LDX;      // 2 * GLM_CHUNK
  double* betas = offs + 2 * GLM_CHUNK;         // GLM_WARPS * NT8
  double* s_tab = betas + GLM_WARPS * NT8;      // FM_TABLE_DOUBLES
  __shared__ __align__(8) unsigned long long bars[4];  // full[0,1], empty[0,1]
  __shared__ int s_any;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int kx = P.kx, kxi = P.kxi;
  const int slot = blockIdx.x * GLM_WARPS + warp;
  const int gene = P.genelist ? (slot < P.n_list ? P.genelist[slot] : P.p) : slot;
  const bool gene_ok = gene < P.p;
  const bool active = gene_ok && (FIN ? true : (P.status[gene] == 0));
  if (tid == 0) s_any = 0;
  __syncthreads();
  if (active && lane == 0) atomicOr(&s_any, 1);
  __syncthreads();
  if (!s_any) return;  // CTA-wide early out

  const int nchunks_all = (P.n + GLM_CHUNK - 1) / GLM_CHUNK;
  const int c_lo = (int)((long)blockIdx.y * nchunks_all / P.nsplit);
  const int c_hi = (int)((long)(blockIdx.y + 1) * nchunks_all / P.nsplit);

  const FmTables FT = fm_load_tables(s_tab, P.fm_tables);
  for (int i = tid; i < 2 * GLM_CHUNK * LDX; i += GLM_THREADS) Xs[i] = 0.0;  // rows past n stay finite
  for (int i = tid; i < 2 * GLM_CHUNK; i += GLM_THREADS) offs[i] = 0.0;
  for (int c = lane; c < NT8; c += 32) {
    const int ce = c < kxi ? (int)P.cmap[c] : -1;
    betas[warp * NT8 + c] = (gene_ok && ce >= 0 && !FIRST) ? P.beta[(size_t)gene * kx + ce] : 0.0;
  }
  if (tid == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_init(bars + 2, GLM_WARPS);
    mbar_init(bars + 3, GLM_WARPS);
    mbar_fence_init();
  }
  __syncthreads();
  fence_proxy_async();  // order the generic zero fill before the async-proxy writes of the bulk copies

  auto issue = [&](int c, int stage) {
    const int c0 = c * GLM_CHUNK;
    const int nvalid = min(GLM_CHUNK, P.n - c0);
    const unsigned bX = (unsigned)(nvalid * LDX * sizeof(double));
    const unsigned bO = (unsigned)(round_up(nvalid, 2) * sizeof(double));
    mbar_expect_tx(bars + stage, bX + bO);
    bulk_g2s(Xs + stage * GLM_CHUNK * LDX, P.X + (size_t)c0 * LDX, bX, bars + stage);
    bulk_g2s(offs + stage * GLM_CHUNK, P.offset + c0, bO, bars + stage);
  };
  if (tid == 0) {
    if (c_lo < c_hi) issue(c_lo, 0);
    if (c_lo + 1 < c_hi) issue(c_lo + 1, 1);
  }

  double nu = 1.0;
  bool pois = true;
  if (NB && gene_ok && P.nu) {
    nu = P.nu[gene];
    pois = nu > P.thres;
  }
  const bool newton = NB && !FIRST && !FIN && gene_ok && !pois && P.newton_step && P.newton_step[gene] <= 1e-2;
  const double ybar = gene_ok ? P.ybar[gene] : 1.0;
  const int rot0 = lane % NT8;
  const double* yrow = P.Yt + (size_t)(gene_ok ? gene : 0) * P.ldn;
  double* murow = (FIN && P.mu_t && gene_ok) ? P.mu_t + (size_t)gene * P.ldn : nullptr;
  const double* bw = betas + warp * NT8;

  double acc[NTRI][2];
#pragma unroll
  for (int i = 0; i < NTRI; ++i) acc[i][0] = acc[i][1] = 0.0;
  double racc[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) racc[i] = 0.0;
  double s_yeta = 0.0, s_mu = 0.0, s_log = 0.0;
  unsigned fphase = 0, ephase = 0;

  for (int c = c_lo; c < c_hi; ++c) {
    const int cb = (c - c_lo) & 1;
    const int cell0 = c * GLM_CHUNK;
    // counts of this lane's four cells of the chunk (issued before the wait: hides the global latency)
    double y4[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cell = cell0 + b * 32 + lane;
      y4[b] = (active && cell < P.n) ? yrow[cell] : 0.0;
    }
    mbar_wait(bars + cb, (fphase >> cb) & 1u);
    fphase ^= 1u << cb;
    if (active) {
      const double* Xc = Xs + cb * GLM_CHUNK * LDX;
      const double* oc = offs + cb * GLM_CHUNK;
      // ---- per-cell evaluation of block b (straight-line) ----
      auto eval = [&](int b, double& w, double& wz) {
        const int cell = cell0 + b * 32 + lane;
        const bool ok = cell < P.n;
        const double y = b == 0 ? y4[0] : (b == 1 ? y4[1] : (b == 2 ? y4[2] : y4[3]));
        const double off = oc[b * 32 + lane];
        double eta, mu;
        if (FIRST) {
          mu = 0.5 * (y + ybar);
          eta = fm_log(mu, FT.log_t);
        } else {
          // lane-rotated column order: lane l reads column (q + l) mod NT8 at step q, which makes the
          // row-strided shared-memory reads conflict free; two partial sums for ILP
          const double* xr = Xc + (b * 32 + lane) * LDX;
          double ea = 0.0, eb = 0.0;
          int idx = rot0;
#pragma unroll 4
          for (int q = 0; q < NT8; q += 2) {
            ea = fma(xr[idx], bw[idx], ea);
            idx = (idx + 1 == NT8) ? 0 : idx + 1;
            eb = fma(xr[idx], bw[idx], eb);
            idx = (idx + 1 == NT8) ? 0 : idx + 1;
          }
          eta = (ea + eb) + off;
          const double ec = eta < 700.0 ? eta : 700.0;
          mu = fm_exp(ec > -700.0 ? ec : -700.0, FT.exp_t);
        }
        const double okf = ok ? 1.0 : 0.0;
        s_yeta = fma(y, eta, s_yeta);          // y = 0 for cells past n; y * log(mu): eta IS log(mu)
        s_mu = fma(okf, mu, s_mu);
        if (NB) {
          const double lg = fm_log(mu + nu, FT.log_t);
          s_log = fma((pois || !ok) ? 0.0 : (y + nu), lg, s_log);
        }
        if (FIN) {
          if (murow && ok) murow[cell] = mu;
        } else {
          double wv = mu;
          if (NB) wv = pois ? mu : mu * nu * fm_rcp(nu + mu);
          const double z = eta + (y - mu) * fm_rcp(mu) - off;
          w = okf * wv;
          wz = w * z;
          if (NB && newton) {
            // -d score / d eta of the NB2 log-likelihood: nu mu (nu + y) / (nu + mu)^2; rhs = w (eta - off) + score
            const double rc = fm_rcp(nu + mu);
            const double wn = mu * nu * rc * ((nu + y) * rc);
            w = okf * wn;
            wz = okf * fma(wn, eta - off, (y - mu) * (nu * rc));
          }
        }
      };
      // ---- Gram / rhs accumulation of block b on the FP64 tensor cores ----
      auto gram = [&](int b, double w, double wz) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int src = s * 4 + lc;
          const double ws = __shfl_sync(0xffffffffu, w, src);
          const double wzs = __shfl_sync(0xffffffffu, wz, src);
          const double* xp = Xc + (b * 32 + src) * LDX + lr;
          double xv[NT], av[NT];
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            xv[t] = xp[8 * t];
            av[t] = ws * xv[t];
            racc[t] = fma(wzs, xv[t], racc[t]);
          }
          int idx = 0;
#pragma unroll
          for (int mt = 0; mt < NT; ++mt)
#pragma unroll
            for (int nt = mt; nt < NT; ++nt) {
              // tiles between mutually orthogonal columns (one-hot treatment indicators): every
              // cross product vanishes whatever the weights, so only the diagonal is accumulated
              if (!(mt >= T0 && nt >= T0))
                dmma884(acc[idx][0], acc[idx][1], av[mt], xv[nt]);
              else if (mt == nt)
                acc[idx][0] = fma(av[mt], xv[mt], acc[idx][0]);
              ++idx;
            }
        }
      };
      const int nblk = min(4, (P.n - cell0 + 31) >> 5);
      // (overlap of the evaluation with the DMMA stream comes from the other warps of the SM: with
      // the mbarrier pipeline the warps are not phase-aligned; an intra-warp software pipeline
      // spilled ~1 KB of registers at the 128-register budget of two CTAs per SM)
#pragma unroll 1
      for (int b = 0; b < nblk; ++b) {
        double w0 = 0.0, wz0 = 0.0;
        eval(b, w0, wz0);
        if (!FIN) gram(b, w0, wz0);
      }
    }
    // release the stage; thread 0 refills it once all warps have released it
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 2 + cb);
    if (tid == 0 && c + 2 < c_hi) {
      mbar_wait(bars + 2 + cb, (ephase >> cb) & 1u);
      issue(c + 2, cb);
    }
    if (c + 2 < c_hi) ephase ^= 1u << cb;  // (kept in step on every thread; only thread 0 uses it)
  }
  if (!active) return;
  // ---- partial sums of this (gene, cell range): Gram (upper tiles mirrored), rhs, deviance terms ----
  s_yeta = warp_sum(s_yeta);
  s_mu = warp_sum(s_mu);
  s_log = warp_sum(s_log);
  double* part = P.part + ((size_t)slot * P.nsplit + blockIdx.y) * P.part_ld;
  if (lane == 0) {
    part[kxi * kxi + kxi + 0] = s_yeta;
    part[kxi * kxi + kxi + 1] = s_mu;
    part[kxi * kxi + kxi + 2] = s_log;
  }
  if (FIN) return;
  double* G = part;
  int idx = 0;
#pragma unroll
  for (int mt = 0; mt < NT; ++mt)
#pragma unroll
    for (int nt = mt; nt < NT; ++nt) {
      const int m = 8 * mt + lr;
      const bool sk = mt >= T0 && nt >= T0;
      double dg = 0.0;
      if (sk && mt == nt) {  // diagonal of a skipped tile: column m, partial sums over the 4 cell lanes
        dg = acc[idx][0];
        dg += __shfl_xor_sync(0xffffffffu, dg, 1);
        dg += __shfl_xor_sync(0xffffffffu, dg, 2);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int nn = 8 * nt + 2 * lc + i;
        if (m < kxi && nn < kxi) {
          const double v = sk ? ((mt == nt && nn == m) ? dg : 0.0) : acc[idx][i];
          G[m * kxi + nn] = v;
          if (nt != mt) G[nn * kxi + m] = v;
        }
      }
      ++idx;
    }
  // rhs: lane holds partial over its cells (lc) for column lr + 8t -> reduce over lc
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    double v = racc[t];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    const int col = lr + 8 * t;
    if (lc == 0 && col < kxi) part[kxi * kxi + col] = v;
  }
}

template <int NT>
int launch_glm_pass_nt(const GlmPass& P, cudaStream_t st);

}  // namespace ca
