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

constexpr int IRLS_WARPS = 4;      // cell ranges of an octet (each with its own staging); NH warps work on a range
constexpr int IRLS_NH = 2;         // warps per cell range: 8 warps per CTA, two CTAs (octets) per SM
constexpr int IRLS_LOGY = 256;     // counts below this read log y from shared memory (log y is the same in every
                                   // iteration; the table holds the values fm_log would return, bit for bit)
constexpr int IRLS_SOLVERS = 8;    // (the lasso refit, whose work space is 4.5 x larger, uses IRLS_SOLVERS_L1 warps
constexpr int IRLS_SOLVERS_L1 = 4; //  in two rounds, and keeps its copy of the Gram matrix in global scratch, so that two
                                   //  CTAs fit on an SM here too: one octet's solves then overlap the other's arithmetic)
static_assert(IRLS_WARPS <= 4, "group_sync() names four barriers");    // warps with solve work space (one gene each)
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
  const int* not_counts;               // 0: every entry of Yp is a non-negative integer (log y from a table)
  double* scratch;                     // per CTA: nseg_tot x (NB + 1) x 64 doubles
  double* gsum;                        // per CTA: (ngroups + 1) x (NB + 1) x 64 doubles: group sums and totals of the flushed tiles
  double* gc_scratch;                  // lasso refit, per CTA: IRLS_SOLVERS_L1 x kx x (kx + 1) doubles
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

// Work space of one solving warp: the lasso refit factorises the full kx x kx system (+ a copy), the
// unpenalised fit only the dense block (arrow structure) next to the per-group couplings.
__host__ __device__ inline size_t glm_irls_work_doubles(int kx, int kde, int ngroups, bool lasso) {
  return lasso ? glm_solve_work_doubles(kx, true) - (size_t)kx * (kx + 1)   // (the Gram copy is in global scratch)
               : glm_solve_work_doubles(kde, false) + (size_t)(ngroups > 1 ? ngroups - 1 : 0) * (kde + 1);
}

inline size_t glm_irls_smem_bytes(int kde, int kx, int ngroups, bool lasso) {
  const int nb = (kde + 7) / 8;
  const int ntw = nb * kde - 8 * (nb * (nb - 1) / 2);
  const int nt = ntw + nb;
  const int ldxd = (kde + 1) | 1;
  size_t d = 0;
  d += (size_t)IRLS_WARPS * 2 * ((((IRLS_CHUNK * ldxd + 8) + 1) & ~1) + 8 * IRLS_CHUNK);  // staging: design rows + counts
  d += (size_t)IRLS_WARPS * 2 * IRLS_CHUNK * 9;               // weights / weighted responses handed from the evaluation to the DMMA phase
  d += (size_t)8 * (kde + 4);                                 // dense coefficients, nu, ybar, flags of the octet
  d += (size_t)nt * 64;                                       // dense totals
  d += (size_t)IRLS_WARPS * 8;                                // deviance partials
  d += (size_t)8 * kx;                                        // coefficients of the octet
  d += (size_t)(lasso ? IRLS_SOLVERS_L1 : IRLS_SOLVERS) * glm_irls_work_doubles(kx, kde, ngroups, lasso);
  d += FM_TABLE_DOUBLES;
  d += IRLS_LOGY;                                             // log 1, log 1, log 2 .. log(IRLS_LOGY - 1)
  return d * sizeof(double);
}

// NH warps share a cell range: each evaluates 8 / NH genes of the chunk's 32 cells and owns
// ceil(NT / NH) of the accumulator tiles (NH = 1: 32 accumulator doubles per lane and 228 registers, one
// CTA of 8 warps per SM; NH = 2: 16 warps per SM at <= 128 registers -- twice the warps to hide the
// fixed-latency stalls that left the FP64 pipe 52 % idle).
template <int KDE, bool NBF, int NH>
__global__ void __launch_bounds__(IRLS_WARPS * 32 * NH, 2) k_glm_irls(const __grid_constant__ GlmIrls P) {
  using S = IrlsShape<KDE>;
  constexpr int NB = S::NB, U = S::U, NTW = S::NTW, NT = S::NT, FLT = S::FLT, LDXD = S::LDXD;
  constexpr int IRLS_THREADS = IRLS_WARPS * 32 * NH;
  constexpr int GH = 8 / NH;                // genes evaluated per warp
  constexpr int NTH = (NT + NH - 1) / NH;   // accumulator tiles per warp: tile t belongs to warp t / NTH of the range
  constexpr int XSTG = ((IRLS_CHUNK * LDXD + 8) + 1) & ~1;  // design rows of a chunk (+ slack for the masked row-block reads); even
  constexpr int STG = XSTG + 8 * IRLS_CHUNK;                // + the chunk's counts of the 8 genes; every stage starts 16-byte aligned
  constexpr int WLD = 9;                                    // row stride of the weight hand-off buffers (odd)
  extern __shared__ __align__(16) double smem[];
  const int kx = P.kx;
  const bool lasso = P.l1 != nullptr;
  double* xs_all = smem;                                   // IRLS_WARPS * 2 * STG
  double* Tt = xs_all + IRLS_WARPS * 2 * STG;              // NT * 64
  double* Dp = Tt + NT * 64;                               // IRLS_WARPS * 8
  double* betas = Dp + IRLS_WARPS * 8;                     // 8 * kx
  double* work_all = betas + 8 * kx;
  const size_t work_d = glm_irls_work_doubles(kx, KDE, P.ngroups, lasso);
  const int nsolv = lasso ? IRLS_SOLVERS_L1 : IRLS_SOLVERS;
  double* wbuf_all = work_all + nsolv * work_d;       // IRLS_WARPS * 2 * IRLS_CHUNK * WLD
  double* bdense = wbuf_all + IRLS_WARPS * 2 * IRLS_CHUNK * WLD;  // 8 * KDE
  double* s_nu = bdense + 8 * KDE;                         // 8
  double* s_ybar = s_nu + 8;                               // 8
  double* s_flag = s_ybar + 8;                             // 8: bit 0 Poisson branch, bit 1 observed-information weights
  double* s_logy = s_flag + 16;                            // IRLS_LOGY
  double* s_tab = s_logy + IRLS_LOGY;
  __shared__ __align__(8) unsigned long long bars[IRLS_WARPS * 2];
  __shared__ int s_octet, s_nrun, s_anynewton;
  __shared__ int s_gene[8], s_status[8], s_niter[8];
  __shared__ double s_devprev[8], s_dev[8], s_step[8];
  __shared__ signed char s_gcol[64];
  __shared__ const double* s_yrow[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cr = warp / NH, half = warp - cr * NH;   // cell range, position inside the range's warp group
  const int lr = lane >> 2, lc = lane & 3;
  const FmTables FT = fm_load_tables(s_tab, P.fm_tables);
  double* xs = xs_all + cr * 2 * STG;
  unsigned long long* bar = bars + cr * 2;
  auto group_sync = [&]() {
    if (NH == 1)
      __syncwarp();
    else
      switch (cr) {   // immediate barrier numbers: ptxas then reserves IRLS_WARPS + 1 barriers, not all 16
        case 0: asm volatile("bar.sync 1, %0;" ::"n"(32 * NH) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"n"(32 * NH) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"n"(32 * NH) : "memory"); break;
        default: asm volatile("bar.sync 4, %0;" ::"n"(32 * NH) : "memory"); break;
      }
  };
  if (lane == 0 && half == 0) {
    mbar_init(bar + 0, 1);
    mbar_init(bar + 1, 1);
  }
  for (int i = lane; i < 2 * STG; i += 32) xs[i] = 0.0;
  if (tid < 64) s_gcol[tid] = P.gcol[tid];
  if (tid == 0) mbar_fence_init();
  __syncthreads();   // (the exp / log tables are in place)
  for (int i = tid; i < IRLS_LOGY; i += IRLS_THREADS) s_logy[i] = fm_log(i > 0 ? (double)i : 1.0, FT.log_t);
  const bool ytab = *P.not_counts == 0;
  __syncthreads();
  fence_proxy_async();
  unsigned fphase = 0;

  const int c_lo = P.chunk_lo[cr], c_hi = P.chunk_lo[cr + 1];
  const int n_oct = ((P.genelist ? P.n_list : P.p) + 7) >> 3;
  double* scratch = P.scratch + (size_t)blockIdx.x * P.nseg_tot * (FLT * 64);
  double* gsum = P.gsum + (size_t)blockIdx.x * (P.ngroups + 1) * (FLT * 64);
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
    if (tid < 8) {
      const int sg = s_gene[tid];
      s_yrow[tid] = P.Yp + (size_t)(sg >= 0 ? sg : (s_gene[0] >= 0 ? s_gene[0] : 0)) * P.ldp;
    }
    __syncthreads();
    double* wbuf = wbuf_all + cr * 2 * IRLS_CHUNK * WLD;   // w, then w z: [cell of the chunk][gene], row stride WLD
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
      double acc[NTH][2];
#pragma unroll
      for (int t = 0; t < NTH; ++t) acc[t][0] = acc[t][1] = 0.0;
      double dacc[GH];
#pragma unroll
      for (int g = 0; g < GH; ++g) dacc[g] = 0.0;
      int gcur = -1, seg = P.seg_ptr[cr];
      auto issue = [&](int c, int stage) {
        const unsigned bytes = (unsigned)(IRLS_CHUNK * LDXD * sizeof(double));
        mbar_expect_tx(bar + stage, bytes + (unsigned)(8 * IRLS_CHUNK * sizeof(double)));
        bulk_g2s(xs + stage * STG, P.Xd + (size_t)c * IRLS_CHUNK * LDXD, bytes, bar + stage);
        // the counts of the chunk: 256 contiguous bytes per gene (registers held them before; the prefetched
        // set was spilled and the spill store waited for the loads, 4 % of the kernel)
#pragma unroll
        for (int g = 0; g < 8; ++g)
          bulk_g2s(xs + stage * STG + XSTG + g * IRLS_CHUNK, s_yrow[g] + (size_t)c * IRLS_CHUNK,
                   (unsigned)(IRLS_CHUNK * sizeof(double)), bar + stage);
      };
      auto flush = [&](auto half_c) {
        constexpr int H = decltype(half_c)::value;
        double* dst = scratch + (size_t)seg * (FLT * 64) + lr * 8 + 2 * lc;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (S::tw(b, U) / NTH == H) {
            *reinterpret_cast<double2*>(dst + b * 64) =
                make_double2(acc[S::tw(b, U) % NTH][0], acc[S::tw(b, U) % NTH][1]);
            acc[S::tw(b, U) % NTH][0] = acc[S::tw(b, U) % NTH][1] = 0.0;
          }
        }
        if ((NTW + NB - 1) / NTH == H) {
          *reinterpret_cast<double2*>(dst + NB * 64) =
              make_double2(acc[(NTW + NB - 1) % NTH][0], acc[(NTW + NB - 1) % NTH][1]);
          acc[(NTW + NB - 1) % NTH][0] = acc[(NTW + NB - 1) % NTH][1] = 0.0;
        }
        ++seg;
      };
      if (lane == 0 && half == 0) {
        if (c_lo < c_hi) issue(c_lo, 0);
        if (c_lo + 1 < c_hi) issue(c_lo + 1, 1);
      }
      // counts of this lane's cell of the chunk, one per gene (256 contiguous bytes per gene and load)
      for (int c = c_lo; c < c_hi; ++c) {
        const int stage = (c - c_lo) & 1;
        mbar_wait(bar + stage, (fphase >> stage) & 1u);
        fphase ^= 1u << stage;
        const double* xc = xs + stage * STG;
        double ycur[GH];
#pragma unroll
        for (int g = 0; g < GH; ++g) ycur[g] = xc[XSTG + (half * GH + g) * IRLS_CHUNK + lane];
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
          // log y: the same in every iteration -- from the table for counts below IRLS_LOGY, computed (whole
          // warp, rare) otherwise; kept out of the main body, which must stay ONE basic block
          double lyv[GH];
          {
            bool slow = !ytab;
#pragma unroll
            for (int gl = 0; gl < GH; ++gl) {
              const double y = ycur[gl];
              slow = slow || !(y < (double)IRLS_LOGY);
              int yi = __double2int_rz(y);
              yi = yi < IRLS_LOGY - 1 ? yi : IRLS_LOGY - 1;
              lyv[gl] = s_logy[yi > 0 ? yi : 0];
            }
            if (__any_sync(0xffffffffu, slow)) {
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) lyv[gl] = fm_log(ycur[gl] > 0.0 ? ycur[gl] : 1.0, FT.log_t);
            }
          }
          // The stores of w / w z are deferred to the end of the body: a shared-memory store between two genes
          // keeps ptxas from hoisting the next gene's table / coefficient loads above it, and the GH chains then
          // run one after the other (profile of the first 16-warp version: every dependent DFMA stalled).
          auto eval8 = [&](auto first_c, auto newton_c) {
            constexpr bool FIRST = decltype(first_c)::value;
            constexpr bool NEWTON = decltype(newton_c)::value;
            // every statement runs across the GH genes (fm_*_n): the GH dependent chains are interleaved in
            // source order, which is the order ptxas keeps
            double eta[GH], ec[GH], mu[GH], nu[GH];
            bool pois[GH], nwt[GH];
#pragma unroll
            for (int gl = 0; gl < GH; ++gl) {
              const int g = half * GH + gl;
              nu[gl] = s_nu[g];
              const int fl = (int)s_flag[g];
              pois[gl] = (fl & 1) != 0;
              nwt[gl] = (fl & 2) != 0;
            }
            if (FIRST) {
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) mu[gl] = 0.5 * (ycur[gl] + s_ybar[half * GH + gl]);
              fm_log_n<GH>(mu, eta, FT.log_t);
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) ec[gl] = eta[gl];
            } else {
              double ea[GH], eb[GH];
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) {
                ea[gl] = fma(gsel, betas[(half * GH + gl) * kx + gc], off);
                eb[gl] = 0.0;
              }
#pragma unroll
              for (int i = 0; i + 1 < KDE; i += 2) {
#pragma unroll
                for (int gl = 0; gl < GH; ++gl) {
                  const double* bg = bdense + (half * GH + gl) * KDE;
                  ea[gl] = fma(x[i], bg[i], ea[gl]);
                  eb[gl] = fma(x[i + 1], bg[i + 1], eb[gl]);
                }
              }
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) {
                if (KDE & 1) ea[gl] = fma(x[KDE - 1], bdense[(half * GH + gl) * KDE + KDE - 1], ea[gl]);
                eta[gl] = ea[gl] + eb[gl];
                double e = eta[gl] < 700.0 ? eta[gl] : 700.0;
                ec[gl] = e > -700.0 ? e : -700.0;
              }
              fm_exp_n<GH>(ec, mu, FT.exp_t);
            }
            // unit deviance / 2 : y log(max(y / mu, eps)) - (y - mu)                          (Poisson)
            //                     y log(max(y / mu, eps)) - (y + nu) log((y + nu) / (mu + nu)) (NB)
            double rc[GH], lt[GH];
            if (NBF) {
              double den[GH], arg[GH];
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) den[gl] = nu[gl] + mu[gl];
              fm_rcp_n<GH>(den, rc);
#pragma unroll
              for (int gl = 0; gl < GH; ++gl) arg[gl] = (ycur[gl] + nu[gl]) * rc[gl];
              fm_log_n<GH>(arg, lt, FT.log_t);
            }
#pragma unroll
            for (int gl = 0; gl < GH; ++gl) {
              const double y = ycur[gl];
              double lrat = lyv[gl] - ec[gl];
              lrat = lrat > LOG_EPS ? lrat : LOG_EPS;
              double d2 = y - mu[gl];
              double wgt = mu[gl];
              // w z = w (eta - off) + (y - mu) w / mu with w / mu = nu / (nu + mu) resp. 1: no reciprocal of mu
              double sfac = 1.0;
              if (NBF) {
                d2 = pois[gl] ? d2 : (y + nu[gl]) * lt[gl];
                wgt = pois[gl] ? mu[gl] : mu[gl] * nu[gl] * rc[gl];
                sfac = pois[gl] ? 1.0 : nu[gl] * rc[gl];
              }
              dacc[gl] = fma(vld, fma(y, lrat, -d2), dacc[gl]);
              double w = wgt, wz = fma(wgt, eta[gl] - off, (y - mu[gl]) * sfac);
              if (NBF && NEWTON) {
                // observed information of the NB2 log-likelihood: nu mu (nu + y) / (nu + mu)^2;
                // rhs = w (eta - off) + score (same minimiser, quadratic tail; lasso refit only)
                const double wn = mu[gl] * nu[gl] * rc[gl] * ((nu[gl] + y) * rc[gl]);
                const double wzn = fma(wn, eta[gl] - off, (y - mu[gl]) * (nu[gl] * rc[gl]));
                w = nwt[gl] ? wn : w;
                wz = nwt[gl] ? wzn : wz;
              }
              wbuf[lane * WLD + half * GH + gl] = w;
              wzbuf[lane * WLD + half * GH + gl] = wz;
            }
          };
          if (first)
            eval8(std::true_type{}, std::false_type{});
          else if (any_newton)
            eval8(std::false_type{}, std::true_type{});
          else
            eval8(std::false_type{}, std::false_type{});
        }
        group_sync();
        // ===== Gram strips and right-hand side on the FP64 tensor cores: lane = (cell lc of a step, gene lr) =====
        auto gram = [&](auto half_c) {
          constexpr int H = decltype(half_c)::value;
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            const int g = __ldg(P.gpair + c * 4 + q);
            if (g != gcur) {
              if (gcur >= 0) flush(half_c);
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
                  if (j >= 8 * b && S::tw(b, j) / NTH == H)
                    dmma884(acc[S::tw(b, j) % NTH][0], acc[S::tw(b, j) % NTH][1], xa[b], bw);
              }
#pragma unroll
              for (int b = 0; b < NB; ++b)
                if ((NTW + b) / NTH == H) dmma884(acc[(NTW + b) % NTH][0], acc[(NTW + b) % NTH][1], xa[b], wz);
            }
          }
        };
        if (NH == 1 || half == 0)
          gram(std::integral_constant<int, 0>{});
        else
          gram(std::integral_constant<int, NH - 1>{});
        group_sync();
        if (lane == 0 && half == 0 && c + 2 < c_hi) issue(c + 2, stage);
      }
      if (gcur >= 0) {
        if (NH == 1 || half == 0)
          flush(std::integral_constant<int, 0>{});
        else
          flush(std::integral_constant<int, NH - 1>{});
      }
      // ---------------- combine the cell ranges (fixed order) ----------------
#pragma unroll
      for (int g = 0; g < GH; ++g) {
        const double v = warp_sum(dacc[g]);
        if (lane == g) Dp[cr * 8 + half * GH + g] = v;
      }
      for (int w = 0; w < IRLS_WARPS; ++w) {
        if (cr == w) {
          double* dst = Tt + lr * 8 + 2 * lc;
#pragma unroll
          for (int t = 0; t < NTH; ++t) {
            const int tg = half * NTH + t;
            if (tg < NT) {
              double2* pp = reinterpret_cast<double2*>(dst + tg * 64);
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
        }
        __syncthreads();
      }
      // ---------------- per-group and total sums of the flushed tiles ----------------
      // (all threads, independent outputs; the solving warps then read ONE value per matrix entry -- summing
      // up to nseg_tot L2-resident values per entry inside the assembly was a chain of dependent L2
      // latencies, ~10 % of an unpenalised octet iteration and a third of a lasso one.  Same summation order.)
      {
        constexpr int GS = FLT * 64;
        for (int e = tid; e < (P.ngroups + 1) * GS; e += IRLS_THREADS) {
          const int gg = e / GS, o = e - gg * GS;
          double v = 0.0;
          if (gg < P.ngroups) {
            for (int q = P.gseg_ptr[gg]; q < P.gseg_ptr[gg + 1]; ++q) v += scratch[(size_t)P.gseg_list[q] * GS + o];
          } else {
            for (int s2 = 0; s2 < P.nseg_tot; ++s2) v += scratch[(size_t)s2 * GS + o];
          }
          gsum[e] = v;
        }
      }
      if (tid == 0) s_nrun = 0;
      __syncthreads();
      // ---------------- warp g: convergence test and update of gene g ----------------
      for (int g = warp; g < 8 && warp < nsolv; g += nsolv) {   // (work space slot = warp)
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
          const double* gs0 = gsum + g;                                  // column g of the group sums ...
          const double* gt0 = gsum + (size_t)P.ngroups * (FLT * 64) + g; // ... and of the totals over all segments
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
            double* wk = work_all + (size_t)warp * work_d;
            const GlmSolveWork W = glm_solve_carve(wk, nd);
            double* cg = wk + glm_solve_work_doubles(nd, false);  // ngr x (KDE + 1): c_g[0..KDE-1] (entry U = d_g), r_g
            const int ld = nd + 1;
            const int CL = KDE + 1;
            for (int e = lane; e < ngr * CL; e += 32) {
              const int gg = e / CL + 1, i = e - (gg - 1) * CL;
              double v = 0.0;
              const int off = i < KDE ? (i >> 3) * 64 + (i & 7) * 8 : NB * 64 + (U & 7) * 8;
              v = gs0[(size_t)gg * (FLT * 64) + off];
              if (i == U && P.ridge) v += P.ridge[P.gcol[gg]];
              cg[e] = v;
            }
            for (int e = lane; e < nd * nd; e += 32) {
              const int i = e / nd, j = e - i * nd;
              const int a = i < j ? i : j, bb = i < j ? j : i;
              double v = 0.0;
              if (bb == U) {
                v = gt0[(a >> 3) * 64 + (a & 7) * 8];
              } else {
                v = Tt[S::tw(a >> 3, bb) * 64 + (a & 7) * 8 + g];
              }
              if (P.ridge && i == j) v += P.ridge[P.dcol[i]];
              W.L[i * ld + j] = v;
            }
            for (int i = lane; i < nd; i += 32) {
              double v = 0.0;
              if ((i >> 3) == NB - 1) {
                v = gt0[NB * 64 + (i & 7) * 8];
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
          const GlmSolveWork W = glm_solve_carve(work_all + (size_t)warp * work_d, kx,
                                                 P.gc_scratch + ((size_t)blockIdx.x * IRLS_SOLVERS_L1 + warp) * kx * (kx + 1));
          const int ld = kx + 1;
          for (int e = lane; e < kx * kx; e += 32) {
            const int ci = e / kx, cj = e - ci * kx;
            const int di = P.cdense[ci], dj = P.cdense[cj];
            const int gi = P.cgroup[ci], gj = P.cgroup[cj];
            double v = 0.0;
            if (di >= 0 && dj >= 0) {
              const int a = di < dj ? di : dj, bb = di < dj ? dj : di;
              if (bb == U) {
                v = gt0[(a >> 3) * 64 + (a & 7) * 8];
              } else {
                v = Tt[S::tw(a >> 3, bb) * 64 + (a & 7) * 8 + g];
              }
            } else if (di >= 0 || dj >= 0) {
              const int a = di >= 0 ? di : dj, gg = di >= 0 ? gj : gi;
              v = gs0[(size_t)gg * (FLT * 64) + (a >> 3) * 64 + (a & 7) * 8];
            } else if (gi == gj) {
              v = gs0[(size_t)gi * (FLT * 64) + (U >> 3) * 64 + (U & 7) * 8];
            }
            if (P.ridge && ci == cj) v += P.ridge[ci];
            W.L[ci * ld + cj] = v;
          }
          for (int c = lane; c < kx; c += 32) {
            const int di = P.cdense[c], gg = P.cgroup[c];
            double v = 0.0;
            if (di >= 0) {
              if ((di >> 3) == NB - 1) {
                v = gt0[NB * 64 + (di & 7) * 8];
              } else {
                v = Tt[(NTW + (di >> 3)) * 64 + (di & 7) * 8 + g];
              }
            } else {
              v = gs0[(size_t)gg * (FLT * 64) + NB * 64 + (U & 7) * 8];
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
