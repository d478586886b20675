// K1 / K2 of the GCATE alternating minimisation, templated on
// KS = kpad / 4 (number of k-steps of the DMMA chain) so that the hot loop is fully
// unrolled with immediate shared-memory offsets.
//
//   K1  row pass  : for a tile of 8 "rows" (cells in the A-step, genes in the B-step)
//                   Theta = X_rows * M^T on the FP64 tensor cores (DMMA m8n8k4); the
//                   epilogue of every 8x8 accumulator tile evaluates the Poisson / NB
//                   log-likelihood terms (table-driven exp / log, fastmath.cuh),
//                   accumulates the row sums and -- in the gradient variant -- feeds
//                   d ell/d theta straight back into a second DMMA that contracts it with
//                   M (the gradient block).  Theta is never written to memory.
//   K2  line search: two phases (k_search: first candidate t = alpha for every row of a tile;
//                   k_search_multi: the remaining candidates t = alpha * beta^j of the rows that
//                   failed it, four per pass over the data); the accept / stop / fallback rule
//                   of causarray/gcate_opt.py:30-82 is replayed over the objectives in order.
//
// Replaces: nll / grad (causarray/gcate_likelihood.py:46-142) and line_search as driven
// from update_with_mask (causarray/gcate_opt.py:168-175, 182-200).
//
// Layout: Y is (rows x ldY) row-major for the entity being updated (Y for cells, Y^T for
// genes), M is (ncols x kpad) row-major with kpad % 8 == 4 so that the B-fragment loads
// of a half-warp hit 16 distinct bank pairs; per-column NB parameters are packed as
// (nu, 1/nu, log nu, poisson flag) and staged with the M chunk.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"
#include "kernels.h"
#include <type_traits>

namespace ca {

constexpr int TILE_ROWS = 8;
constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;
#ifndef CA_MINB
#define CA_MINB 2
#endif
#ifndef CA_TC
#define CA_TC 128
#endif
#ifndef CA_NACC
#define CA_NACC 2
#endif

struct PassCfg {
  const double* Y;
  long ldY;
  int ncols;
  const double* M;
  const double* nu;       // per row (nu_per_row) : nu, inv_nu, log_nu indexed by row
  const double* inv_nu;
  const double* log_nu;
  const double4* colparm; // per column (!nu_per_row): (nu, 1/nu, log nu, pois)
  int nu_per_row;
  int any_pois;  // some nu exceeds thres (NB entries of those rows / columns take the Poisson branch)
  int family;
  double thres;
  int tc;  // columns per staged chunk (multiple of 8)
  const double* fm_tables;
  NbClip clip;
  double* theta_out;      // optional, gradient passes: see RowProblem
  const int* theta_pos;
  long theta_ld;
};

struct SearchArgs {
  double* X;
  const double* g;
  const double* f0;
  const double* normg;
  const double* tys;
  const double* lam;
  int ldlam;
  int prox_lo, prox_hi;
  int mv_lo, mv_hi;         // columns the step direction can be non-zero in ([0, kpad) when unknown)
  const double* alpha_row;  // nullptr -> alpha
  double alpha, beta, tol;
  int max_iters;
  const int* rowlist;
  const int* count;
  int nrows;
  double* ell_fin;
  int* neval;
  unsigned long long* eval_total;
  // two-phase search: phase A (k_search with pend != nullptr) evaluates the first candidate only and
  // hands the rows that neither accepted nor terminated to phase B (k_search_multi)
  unsigned char* pend;   // per row: 1 = continue in phase B
  double* f01s;          // per row: objective / log-likelihood sum of the first candidate
  double* e01s;
};

template <int KS>
struct Lay {
  static constexpr int KPAD = 4 * KS;
};

// shared-memory carve-up (doubles)
struct Smem {
  double* tab;     // FM_TABLE_DOUBLES
  double* M;       // 2 * tc * kpad + 64 slack
  double4* cp;     // 2 * tc   (column parameters)
  double* Ys;      // 2 * 8 * (tc + 8)  (rows of Y for the chunk)
  double* X;       // 8 * kpad
  double* X0;      // 8 * kpad (search)
  double* Gd;      // 8 * kpad (search)
  double* red;     // 64
  double* rowsum;  // 8
  double* Gred;    // 64 * NGT*8
  double* G;       // 8 * NGT*8
};

// nx: number of (8 x kpad) row blocks: 1 row pass, 3 line search (X, X0, g), NC + 3 multi-candidate search
__host__ __device__ inline size_t smem_doubles(int tc, int kpad, int ngt, int nx) {
  return FM_TABLE_DOUBLES + 2 * (size_t)tc * kpad + 64 + 2 * (size_t)tc * 4 + 2 * 8 * ((size_t)tc + 8) + (size_t)nx * 8 * kpad + 64 +
         8 + 64 * (size_t)ngt * 8 + 8 * (size_t)ngt * 8;
}

__device__ __forceinline__ Smem carve(double* base, int tc, int kpad, int ngt8, int nx) {
  Smem s;
  s.tab = base;
  base += FM_TABLE_DOUBLES;
  s.M = base;
  base += 2 * tc * kpad + 64;
  s.cp = reinterpret_cast<double4*>(base);
  base += 2 * tc * 4;
  s.Ys = base;
  base += 2 * 8 * (tc + 8);
  s.X = base;
  s.X0 = base + 8 * kpad;
  s.Gd = base + 16 * kpad;
  base += nx * 8 * kpad;
  s.red = base;
  base += 64;
  s.rowsum = base;
  base += 8;
  s.Gred = base;
  base += 64 * ngt8;
  s.G = base;
  return s;
}

// Stage chunk c (M rows, column parameters, the 8 rows of Y) into shared memory with bulk async
// copies (cp.async.bulk / UBLKCP) issued by ONE thread and completed on an mbarrier: the first
// version staged with per-thread 16-byte cp.async and spent ~40 % of the non-math instructions
// of the kernel on address arithmetic and LDGSTS issue.
struct Pipe {
  unsigned long long* bar;  // two mbarriers (one per stage) in shared memory
  unsigned phase;           // bit s = parity of the next wait on stage s
};

template <int KS>
__device__ __forceinline__ void stage_chunk(const PassCfg& P, int c, double* dstM, double4* dstC, double* dstY,
                                            const int* s_rows, unsigned long long* bar) {
  constexpr int KPAD = 4 * KS;
  const int col0 = c * P.tc;
  const int nvalid = min(P.tc, P.ncols - col0);
  const int npad = min(P.tc, round_up(nvalid, 8));
  const int ldy = P.tc + 8;
  // The ten copies of a chunk are issued by ten different lanes: lane 0 of warp w requests row w of
  // Y, warp 0 also arms the barrier and requests M, warp 1 the column parameters.  (With one thread
  // issuing everything its warp was ~200 cycles late at every per-chunk CTA barrier.)  A copy that
  // completes before the expect_tx is harmless: the phase cannot complete before warp 0's arrive.
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    const unsigned bM = (unsigned)(nvalid * KPAD * sizeof(double));
    const unsigned bC = P.colparm ? (unsigned)(nvalid * sizeof(double4)) : 0u;
    const unsigned bY = (unsigned)(npad * sizeof(double));
    if (w == 0) {
      mbar_expect_tx(bar, bM + bC + 8u * bY);
      bulk_g2s(dstM, P.M + (size_t)col0 * KPAD, bM, bar);
    }
    if (w == 1 && P.colparm) bulk_g2s(dstC, P.colparm + col0, bC, bar);
    const int row = s_rows[w];
    bulk_g2s(dstY + w * ldy, P.Y + (size_t)(row >= 0 ? row : 0) * P.ldY + col0, bY, bar);
  }
  // ragged tail of the last chunk: zero rows of M and neutral column parameters (generic stores
  // to addresses the bulk copies do not touch)
  if (npad > nvalid) {
    for (int i = nvalid * KPAD + threadIdx.x; i < npad * KPAD; i += NTHREADS) dstM[i] = 0.0;
    if (P.colparm)
      for (int i = nvalid + threadIdx.x; i < npad; i += NTHREADS) dstC[i] = make_double4(1.0, 1.0, 0.0, 0.0);
  }
}

// *_F: NB without the Poisson switch -- no nu exceeds thres_disp (always the case for
// method-of-moments dispersions, which are clipped to [0.01, 100]); the per-entry Poisson selects
// drop out and the log-likelihood-only pass needs a single clamp of log u.
enum PassMode { MODE_POIS = 0, MODE_NB_ROW = 1, MODE_NB_COL = 2, MODE_NB_ROW_F = 3, MODE_NB_COL_F = 4 };
__host__ __device__ constexpr bool mode_row(int m) { return m == MODE_NB_ROW || m == MODE_NB_ROW_F; }
__host__ __device__ constexpr bool mode_col(int m) { return m == MODE_NB_COL || m == MODE_NB_COL_F; }
__host__ __device__ constexpr bool mode_fast(int m) { return m == MODE_NB_ROW_F || m == MODE_NB_COL_F; }

// Likelihood terms of one matrix entry (branch free), accumulated straight into `acc`:
//   acc += ell(y, theta);   d = -d ell / d theta (GRAD).
// The last operations of ell are fused multiply-adds into the running sum (y*a - c*b with
// (a, b, c) = (log u, log1p u, y + nu) for NB and (xi, m, 1) on the Poisson branch), which saves
// the separate product, the Poisson alternative and the add into the row sum of the first version.
// VALID = false (ragged last tile only) removes the entry from the sums.
template <int MODE, bool GRAD, bool RAGGED>
__device__ __forceinline__ void eval_entry(const FmTables& FT, const NbClip& clip, double theta, double y, double nu,
                                           double inu, double lnu, bool pois, bool valid, double& acc, double& d) {
  if (MODE == MODE_POIS) {
    const double x = fm_lt_pos(theta, 100.0) ? theta : 100.0;
    double m = fm_exp(fm_gt_neg(x, -700.0) ? x : -700.0, FT.exp_t);
    if (RAGGED) {
      y = valid ? y : 0.0;
      m = valid ? m : 0.0;
    }
    acc = fma(y, x, acc) - m;
    if (GRAD) d = m - y;
  } else if (mode_fast(MODE) && !GRAD) {
    // ell = y log u_c - (y + nu) log1p(u_c) with log u_c = clamp(xi - log nu) and u_c = exp(log u_c):
    // one clamp instead of the paired clamps of u and log u, and no range guard on the exponential
    const double x = fm_lt_pos(theta, 100.0) ? theta : 100.0;
    double lu = x - lnu;
    lu = fm_lt_neg(lu, clip.lu_lo) ? clip.lu_lo : lu;
    lu = fm_gt_pos(lu, clip.lu_hi) ? clip.lu_hi : lu;
    const double u = fm_exp(lu, FT.exp_t);
    const double L = fm_log(1.0 + u, FT.log_t);
    double c = y + nu;
    if (RAGGED) {
      y = valid ? y : 0.0;
      c = valid ? c : 0.0;
    }
    acc = fma(-c, L, fma(y, lu, acc));
  } else if (mode_fast(MODE)) {
    double x, m, lu, L, w;
    fm_terms_nb(theta, inu, lnu, FT, clip, x, m, lu, L, w);
    d = (m - y) * fm_rcp(w);
    double c = y + nu;
    if (RAGGED) {
      d = valid ? d : 0.0;
      y = valid ? y : 0.0;
      c = valid ? c : 0.0;
    }
    acc = fma(-c, L, fma(y, lu, acc));
  } else {
    double x, m, lu, L, w;
    fm_terms_nb(theta, inu, lnu, FT, clip, x, m, lu, L, w);
    const double a = pois ? x : lu;
    const double b = pois ? m : L;
    double c = pois ? 1.0 : y + nu;
    if (GRAD) {
      const double q = pois ? 1.0 : fm_rcp(w);
      d = (m - y) * q;
      if (RAGGED) d = valid ? d : 0.0;
    }
    if (RAGGED) {
      y = valid ? y : 0.0;
      c = valid ? c : 0.0;
    }
    acc = fma(-c, b, fma(y, a, acc));
  }
}

// The four entries a lane has in flight (A0, B0, A1, B1 of the tile pair), for the switch-free NB modes:
// the same operations as four eval_entry calls, written statement by statement ACROSS the entries
// (fm_exp_n / fm_log_n / fm_rcp_n).  ptxas keeps source order: with four consecutive calls it overlapped
// the chains only in pairs.  accA takes entries 0 and 2 (in that order), accB entries 1 and 3.
template <int MODE, bool GRAD, bool RAGGED>
__device__ __forceinline__ void eval_entries4(const FmTables& FT, const NbClip& clip, const double (&theta)[4],
                                              const double (&yin)[4], const double (&nu)[4], const double (&inu)[4],
                                              const double (&lnu)[4], const bool (&valid)[4], double& accA,
                                              double& accB, double (&d)[4]) {
  static_assert(mode_fast(MODE), "switch-free NB modes only");
  double x[4], lu[4], L[4], y[4], c[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    x[e] = fm_lt_pos(theta[e], 100.0) ? theta[e] : 100.0;
    y[e] = yin[e];
  }
  if (!GRAD) {
    double u[4], w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) lu[e] = x[e] - lnu[e];
#pragma unroll
    for (int e = 0; e < 4; ++e) lu[e] = fm_lt_neg(lu[e], clip.lu_lo) ? clip.lu_lo : lu[e];
#pragma unroll
    for (int e = 0; e < 4; ++e) lu[e] = fm_gt_pos(lu[e], clip.lu_hi) ? clip.lu_hi : lu[e];
    fm_exp_n<4>(lu, u, FT.exp_t);
#pragma unroll
    for (int e = 0; e < 4; ++e) w[e] = 1.0 + u[e];
    fm_log_n<4>(w, L, FT.log_t);
  } else {
    double xe[4], m[4], uc[4], w[4], r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) xe[e] = fm_gt_neg(x[e], -700.0) ? x[e] : -700.0;
    fm_exp_n<4>(xe, m, FT.exp_t);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uc[e] = m[e] * inu[e];
      lu[e] = x[e] - lnu[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool lo = fm_lt_pos(uc[e], clip.u_lo), hi = fm_gt_pos(uc[e], clip.u_hi);
      uc[e] = lo ? clip.u_lo : uc[e];
      lu[e] = lo ? clip.lu_lo : lu[e];
      uc[e] = hi ? clip.u_hi : uc[e];
      lu[e] = hi ? clip.lu_hi : lu[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) w[e] = 1.0 + uc[e];
    fm_log_n<4>(w, L, FT.log_t);
    fm_rcp_n<4>(w, r);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      d[e] = (m[e] - y[e]) * r[e];
      if (RAGGED) d[e] = valid[e] ? d[e] : 0.0;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    c[e] = y[e] + nu[e];
    if (RAGGED) {
      y[e] = valid[e] ? y[e] : 0.0;
      c[e] = valid[e] ? c[e] : 0.0;
    }
  }
  accA = fma(-c[0], L[0], fma(y[0], lu[0], accA));
  accB = fma(-c[1], L[1], fma(y[1], lu[1], accB));
  accA = fma(-c[2], L[2], fma(y[2], lu[2], accA));
  accB = fma(-c[3], L[3], fma(y[3], lu[3], accB));
}

// One pass over all columns for the 8 rows listed in s_rows (-1 = empty slot; its sums are
// garbage and ignored by the callers).  s_X: 8 x kpad candidate rows.  On return (after a
// __syncthreads) s_rowsum[r] holds sum_col ell(y, theta) and, if GRAD, s_G[r*NGT*8 + c] holds
// sum_col dl * M[col, gc_lo + c].
//
// Per iteration a warp handles TWO 8x8 tiles (t and t + 8): four independent per-entry
// dependency chains per lane and 2 x NACC independent DMMA accumulators (measured latencies:
// DFMA 8 cycles, DMMA 26 cycles, one DMMA per 16 cycles per SM sub-partition).
template <int KS, bool GRAD, int NGT, int MODE, bool TH = false>
__device__ __forceinline__ void tile_pass_mode(const PassCfg& P, const FmTables& FT, const int* s_rows,
                                               const double* s_X, const Smem& S, int gc_lo, Pipe& pipe) {
  constexpr int KPAD = 4 * KS;
  constexpr int NACC = KS < CA_NACC ? KS : CA_NACC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int tc = P.tc;
  const int row = s_rows[lr];
  double nu_r = 1.0, inu_r = 1.0, lnu_r = 0.0;
  bool pois_r = false;
  if (mode_row(MODE) && row >= 0) {
    nu_r = P.nu[row];
    inu_r = P.inv_nu[row];
    lnu_r = P.log_nu[row];
    pois_r = nu_r > P.thres;
  }
  const NbClip clip = P.clip;
  double afrag[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) afrag[ks] = s_X[lr * KPAD + ks * 4 + lc];
  double ellA = 0.0, ellB = 0.0;  // running row sums of the two tiles of an iteration
  double gacc[NGT][2], gacb[NGT][2];  // two accumulator sets (the two tiles of an iteration)
#pragma unroll
  for (int j = 0; j < NGT; ++j) gacc[j][0] = gacc[j][1] = gacb[j][0] = gacb[j][1] = 0.0;

  const int nchunks = (P.ncols + tc - 1) / tc;
  const int chunk_elems = tc * KPAD;
  const int ldy = tc + 8;
  stage_chunk<KS>(P, 0, S.M, S.cp, S.Ys, s_rows, pipe.bar + 0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      const int sn = (c + 1) & 1;
      stage_chunk<KS>(P, c + 1, S.M + sn * chunk_elems, S.cp + sn * tc, S.Ys + sn * 8 * ldy, s_rows, pipe.bar + sn);
    }
    mbar_wait(pipe.bar + (c & 1), (pipe.phase >> (c & 1)) & 1u);
    pipe.phase ^= 1u << (c & 1);
    if (c == nchunks - 1) __syncthreads();  // the ragged-tail zero fill (generic stores) of the last chunk
    const double* Mb = S.M + (c & 1) * chunk_elems;
    const double4* Cb = S.cp + (c & 1) * tc;
    const double* Yb = S.Ys + (c & 1) * 8 * ldy + lr * ldy + 2 * lc;
    const int col0 = c * tc;
    const int nrem = P.ncols - col0;
    const int ntile = min(tc, round_up(nrem, 8)) >> 3;
    auto pair_body = [&](auto ragged_tag, int tA, int tB, bool hasB) {
      constexpr bool RAGGED = decltype(ragged_tag)::value;
      const int clA = tA * 8 + 2 * lc, clB = tB * 8 + 2 * lc;  // columns within the chunk
      const double2 yA = *reinterpret_cast<const double2*>(Yb + tA * 8);
      const double2 yB = *reinterpret_cast<const double2*>(Yb + tB * 8);
      // theta output: the positions of the four columns are requested here, ahead of the products and the
      // likelihood terms, so that their (L1 / L2) latency is hidden; loaded next to the stores they were
      // exposed at the end of every tile pair (+30 % on the pass)
      int qA0 = -1, qA1 = -1, qB0 = -1, qB1 = -1;
      if (TH && GRAD) {
        const int cA = col0 + clA, cB = col0 + clB;
        if (!RAGGED || clA < nrem) qA0 = __ldg(P.theta_pos + cA);
        if (!RAGGED || clA + 1 < nrem) qA1 = __ldg(P.theta_pos + cA + 1);
        if (hasB && (!RAGGED || clB < nrem)) qB0 = __ldg(P.theta_pos + cB);
        if (hasB && (!RAGGED || clB + 1 < nrem)) qB1 = __ldg(P.theta_pos + cB + 1);
      }
      double pa[NACC][2], pb[NACC][2];
      const double* bpA = Mb + (tA * 8 + lr) * KPAD + lc;
      const double* bpB = Mb + (tB * 8 + lr) * KPAD + lc;
#pragma unroll
      for (int ks = 0; ks < NACC; ++ks) {
        pa[ks][0] = pa[ks][1] = pb[ks][0] = pb[ks][1] = 0.0;
        dmma884(pa[ks][0], pa[ks][1], afrag[ks], bpA[ks * 4]);
        dmma884(pb[ks][0], pb[ks][1], afrag[ks], bpB[ks * 4]);
      }
#pragma unroll
      for (int ks = NACC; ks < KS; ++ks) {
        dmma884(pa[ks % NACC][0], pa[ks % NACC][1], afrag[ks], bpA[ks * 4]);
        dmma884(pb[ks % NACC][0], pb[ks % NACC][1], afrag[ks], bpB[ks * 4]);
      }
#pragma unroll
      for (int st = 1; st < NACC; st <<= 1)
#pragma unroll
        for (int i = 0; i + st < NACC; i += 2 * st) {
          pa[i][0] += pa[i + st][0];
          pa[i][1] += pa[i + st][1];
          pb[i][0] += pb[i + st][0];
          pb[i][1] += pb[i + st][1];
        }
      double4 cA0 = make_double4(nu_r, inu_r, lnu_r, pois_r ? 1.0 : 0.0), cA1 = cA0, cB0 = cA0, cB1 = cA0;
      if (mode_col(MODE)) {
        cA0 = Cb[clA];
        cA1 = Cb[clA + 1];
        cB0 = Cb[clB];
        cB1 = Cb[clB + 1];
      }
      bool vA0 = true, vA1 = true, vB0 = true, vB1 = true;
      if (RAGGED) {  // missing second tile / columns past the end
        vA0 = clA < nrem;
        vA1 = clA + 1 < nrem;
        vB0 = hasB && clB < nrem;
        vB1 = hasB && clB + 1 < nrem;
      }
      double dA0 = 0.0, dA1 = 0.0, dB0 = 0.0, dB1 = 0.0;
      if constexpr (mode_fast(MODE)) {
        const double th4[4] = {pa[0][0], pb[0][0], pa[0][1], pb[0][1]};
        const double y4[4] = {yA.x, yB.x, yA.y, yB.y};
        const double nu4[4] = {cA0.x, cB0.x, cA1.x, cB1.x};
        const double in4[4] = {cA0.y, cB0.y, cA1.y, cB1.y};
        const double ln4[4] = {cA0.z, cB0.z, cA1.z, cB1.z};
        const bool v4[4] = {vA0, vB0, vA1, vB1};
        double d4[4] = {0.0, 0.0, 0.0, 0.0};
        eval_entries4<MODE, GRAD, RAGGED>(FT, clip, th4, y4, nu4, in4, ln4, v4, ellA, ellB, d4);
        dA0 = d4[0];
        dB0 = d4[1];
        dA1 = d4[2];
        dB1 = d4[3];
      } else {
        eval_entry<MODE, GRAD, RAGGED>(FT, clip, pa[0][0], yA.x, cA0.x, cA0.y, cA0.z, cA0.w != 0.0, vA0, ellA, dA0);
        eval_entry<MODE, GRAD, RAGGED>(FT, clip, pb[0][0], yB.x, cB0.x, cB0.y, cB0.z, cB0.w != 0.0, vB0, ellB, dB0);
        eval_entry<MODE, GRAD, RAGGED>(FT, clip, pa[0][1], yA.y, cA1.x, cA1.y, cA1.z, cA1.w != 0.0, vA1, ellA, dA1);
        eval_entry<MODE, GRAD, RAGGED>(FT, clip, pb[0][1], yB.y, cB1.x, cB1.y, cB1.z, cB1.w != 0.0, vB1, ellB, dB1);
      }
      if (GRAD) {
        const double* gA = Mb + (tA * 8 + 2 * lc) * KPAD + gc_lo + lr;
        const double* gB = Mb + (tB * 8 + 2 * lc) * KPAD + gc_lo + lr;
#pragma unroll
        for (int j = 0; j < NGT; ++j) {
          dmma884(gacc[j][0], gacc[j][1], dA0, gA[8 * j]);
          dmma884(gacb[j][0], gacb[j][1], dB0, gB[8 * j]);
          dmma884(gacc[j][0], gacc[j][1], dA1, gA[KPAD + 8 * j]);
          dmma884(gacb[j][0], gacb[j][1], dB1, gB[KPAD + 8 * j]);
        }
        if (TH && row >= 0) {
          // linear predictor of the selected columns (the treated cells of a stage-2 gene pass) for the
          // one-hot line search that follows; placed after the likelihood terms so that their
          // straight-line block stays intact
          double* tr = P.theta_out + (size_t)row * P.theta_ld;
          if (qA0 >= 0) tr[qA0] = pa[0][0];
          if (qA1 >= 0) tr[qA1] = pa[0][1];
          if (qB0 >= 0) tr[qB0] = pb[0][0];
          if (qB1 >= 0) tr[qB1] = pb[0][1];
        }
      }
    };
#pragma unroll 1
    for (int t = warp; t < ntile; t += 2 * NWARPS) {
      const bool hasB = t + NWARPS < ntile;            // warp uniform
      const int tB = hasB ? t + NWARPS : t;
      if (!hasB || tB * 8 + 8 > nrem)
        pair_body(std::true_type{}, t, tB, hasB);
      else
        pair_body(std::false_type{}, t, tB, hasB);
    }
    __syncthreads();
  }
  // fixed-order reduction: 4 lanes of a row, then the 8 warps
  double ellsum = ellA + ellB;
  ellsum += __shfl_xor_sync(0xffffffffu, ellsum, 1);
  ellsum += __shfl_xor_sync(0xffffffffu, ellsum, 2);
  if (lc == 0) S.red[warp * 8 + lr] = ellsum;
  if (GRAD) {
#pragma unroll
    for (int j = 0; j < NGT; ++j) {
      S.Gred[(warp * 8 + lr) * (NGT * 8) + 8 * j + 2 * lc] = gacc[j][0] + gacb[j][0];
      S.Gred[(warp * 8 + lr) * (NGT * 8) + 8 * j + 2 * lc + 1] = gacc[j][1] + gacb[j][1];
    }
  }
  __syncthreads();
  if (tid < 8) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) s += S.red[w * 8 + tid];
    S.rowsum[tid] = s;
  }
  if (GRAD) {
    for (int e = tid; e < 8 * NGT * 8; e += NTHREADS) {
      const int r = e / (NGT * 8), cc = e % (NGT * 8);
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) s += S.Gred[(w * 8 + r) * (NGT * 8) + cc];
      S.G[e] = s;
    }
  }
  __syncthreads();
}

template <int KS, bool GRAD, int NGT, bool TH = false>
__device__ __forceinline__ void tile_pass(const PassCfg& P, const FmTables& FT, const int* s_rows, const double* s_X,
                                          const Smem& S, int gc_lo, Pipe& pipe) {
  if (TH) {   // (theta output: gene passes only, i.e. per-row dispersions)
    if (P.family != CA_FAMILY_NB)
      tile_pass_mode<KS, GRAD, NGT, MODE_POIS, TH>(P, FT, s_rows, s_X, S, gc_lo, pipe);
    else if (P.any_pois)
      tile_pass_mode<KS, GRAD, NGT, MODE_NB_ROW, TH>(P, FT, s_rows, s_X, S, gc_lo, pipe);
    else
      tile_pass_mode<KS, GRAD, NGT, MODE_NB_ROW_F, TH>(P, FT, s_rows, s_X, S, gc_lo, pipe);
    return;
  }
  if (P.family != CA_FAMILY_NB)
    tile_pass_mode<KS, GRAD, NGT, MODE_POIS>(P, FT, s_rows, s_X, S, gc_lo, pipe);
  else if (P.nu_per_row && P.any_pois)
    tile_pass_mode<KS, GRAD, NGT, MODE_NB_ROW>(P, FT, s_rows, s_X, S, gc_lo, pipe);
  else if (P.nu_per_row)
    tile_pass_mode<KS, GRAD, NGT, MODE_NB_ROW_F>(P, FT, s_rows, s_X, S, gc_lo, pipe);
  else if (P.any_pois)
    tile_pass_mode<KS, GRAD, NGT, MODE_NB_COL>(P, FT, s_rows, s_X, S, gc_lo, pipe);
  else
    tile_pass_mode<KS, GRAD, NGT, MODE_NB_COL_F>(P, FT, s_rows, s_X, S, gc_lo, pipe);
}

__device__ __forceinline__ void pipe_init(Pipe& pipe, unsigned long long* bars) {
  pipe.bar = bars;
  pipe.phase = 0;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_fence_init();
  }
}

// ---------------------------------------------------------------------------
// Gradient / log-likelihood pass over a row list.
//   rowlist == nullptr -> rows 0..nrows-1, else rowlist[0..*count)
//   ell_out[row]                      = sum_col ell
//   G_out[row * ldG + c], c < ngt*8   = sum_col dl * M[col, gc_lo + c]   (GRAD)
// ---------------------------------------------------------------------------
template <int KS, bool GRAD, int NGT, bool TH = false>
__global__ void __launch_bounds__(NTHREADS, CA_MINB) k_rowpass(PassCfg P, const double* X, const int* rowlist,
                                                       const int* count, int nrows, int gc_lo,
                                                       double* ell_out, double* G_out, int ldG) {
  constexpr int KPAD = 4 * KS;
  extern __shared__ __align__(16) double smem[];
  const Smem S = carve(smem, P.tc, KPAD, NGT * 8, 1);
  __shared__ int s_rows[8];
  __shared__ __align__(8) unsigned long long s_bars[2];
  const int total = rowlist ? *count : nrows;
  const int tile0 = blockIdx.x * TILE_ROWS;
  if (tile0 >= total) return;
  Pipe pipe;
  pipe_init(pipe, s_bars);
  const FmTables FT = fm_load_tables(S.tab, P.fm_tables);
  if (threadIdx.x < 8) {
    const int i = tile0 + threadIdx.x;
    s_rows[threadIdx.x] = (i < total) ? (rowlist ? rowlist[i] : i) : -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD;
    const int row = s_rows[r];
    S.X[e] = (row >= 0) ? X[(size_t)row * KPAD + c] : 0.0;
  }
  __syncthreads();
  tile_pass<KS, GRAD, NGT, TH>(P, FT, s_rows, S.X, S, gc_lo, pipe);
  if (threadIdx.x < 8 && s_rows[threadIdx.x] >= 0) ell_out[s_rows[threadIdx.x]] = S.rowsum[threadIdx.x];
  if (GRAD) {
    for (int e = threadIdx.x; e < 8 * NGT * 8; e += NTHREADS) {
      const int r = e / (NGT * 8), cc = e % (NGT * 8);
      if (s_rows[r] >= 0 && cc < ldG) G_out[(size_t)s_rows[r] * ldG + cc] = S.G[e];
    }
  }
}

__device__ __forceinline__ double soft_thr(double x, double lam) {
  const double a = fabs(x) - lam;
  const double m = a > 0.0 ? a : 0.0;
  return (x > 0.0) ? m : ((x < 0.0) ? -m : 0.0);
}

// ---------------------------------------------------------------------------
// Backtracking line search for a row list (causarray/gcate_opt.py:30-82).
//   g[row, 0..kpad)   projected gradient (zeros outside the moving block)
//   f0[row], normg[row] from k_prep_search
//   prox on columns [prox_lo, prox_hi) with thresholds lam[row*ldlam + c-prox_lo]
//   tys[row] = sum_col Tys (constant term of the row objective)
// On exit X[row] is overwritten with the accepted point, ell_fin[row] holds the raw
// log-likelihood sum at that point.
// ---------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(NTHREADS, CA_MINB) k_search(PassCfg P, SearchArgs A) {
  constexpr int KPAD = 4 * KS;
  extern __shared__ __align__(16) double smem[];
  const Smem S = carve(smem, P.tc, KPAD, 8, 3);
  __shared__ int s_rows[8];
  __shared__ __align__(8) unsigned long long s_bars[2];
  __shared__ double s_t[8], s_f0[8], s_ng[8], s_f01[8], s_e01[8], s_alpha[8], s_efin[8];
  __shared__ unsigned char s_done[8], s_pend[8];
  __shared__ int s_mode[8], s_ne[8], s_alldone;
  const int tid = threadIdx.x;
  const int total = A.rowlist ? *A.count : A.nrows;
  const int tile0 = blockIdx.x * TILE_ROWS;
  if (tile0 >= total) return;
  Pipe pipe;
  pipe_init(pipe, s_bars);
  const FmTables FT = fm_load_tables(S.tab, P.fm_tables);
  if (tid < 8) {
    const int i = tile0 + tid;
    const int row = (i < total) ? (A.rowlist ? A.rowlist[i] : i) : -1;
    s_rows[tid] = row;
    s_done[tid] = row < 0;
    s_pend[tid] = 0;
    s_mode[tid] = 0;
    s_ne[tid] = 0;
    if (row >= 0) {
      s_alpha[tid] = A.alpha_row ? A.alpha_row[row] : A.alpha;
      s_t[tid] = s_alpha[tid];
      s_f0[tid] = A.f0[row];
      s_ng[tid] = A.normg[row];
    }
  }
  __syncthreads();
  for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD, row = s_rows[r];
    S.X0[e] = (row >= 0) ? A.X[(size_t)row * KPAD + c] : 0.0;
    S.Gd[e] = (row >= 0) ? A.g[(size_t)row * KPAD + c] : 0.0;
    S.X[e] = 0.0;
  }
  __syncthreads();
  const double ncols_d = (double)P.ncols;
  const int nprox = A.prox_hi - A.prox_lo;
  for (int it = 0; it < A.max_iters; ++it) {
    for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
      const int r = e / KPAD, c = e % KPAD;
      if (!s_done[r]) {
        double v = S.X0[e] - s_t[r] * S.Gd[e];
        if (A.lam && c >= A.prox_lo && c < A.prox_hi)
          v = soft_thr(v, A.lam[(size_t)s_rows[r] * A.ldlam + (c - A.prox_lo)]);
        S.X[e] = v;
      }
    }
    __syncthreads();
    tile_pass<KS, false, 1>(P, FT, s_rows, S.X, S, 0, pipe);
    if (tid < 8 && !s_done[tid]) {
      const int row = s_rows[tid];
      double pen = 0.0;
      if (A.lam && nprox > 0) {
        for (int c = 0; c < nprox; ++c)
          pen += A.lam[(size_t)row * A.ldlam + c] * fabs(S.X[tid * KPAD + A.prox_lo + c]);
        pen /= (double)nprox;
      }
      const double ell = S.rowsum[tid];
      const double f1 = -(ell + A.tys[row]) / ncols_d + pen;
      s_ne[tid] += 1;
      if (it == 0) {
        s_f01[tid] = f1;
        s_e01[tid] = ell;
      }
      const double t = s_t[tid];
      if (f1 < s_f0[tid] - A.tol * t * s_ng[tid]) {
        s_done[tid] = 1;
        s_efin[tid] = ell;
      } else if (t < 1e-4 || it == A.max_iters - 1) {
        s_done[tid] = 1;
        if (s_f01[tid] < 1.1 * f1) {
          s_mode[tid] = 1;  // fall back to the full initial step
          s_efin[tid] = s_e01[tid];
        } else {
          s_efin[tid] = ell;
        }
      } else if (A.pend) {  // phase A: the remaining candidates of this row are evaluated together in phase B
        s_done[tid] = 1;
        s_pend[tid] = 1;
        A.f01s[row] = f1;
        A.e01s[row] = ell;
      } else {
        s_t[tid] = t * A.beta;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int all = 1;
      for (int r = 0; r < 8; ++r) all &= (int)s_done[r];
      s_alldone = all;
    }
    __syncthreads();
    if (s_alldone) break;
  }
  for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD, row = s_rows[r];
    if (row < 0 || s_pend[r]) continue;
    double v = S.X[e];
    if (s_mode[r] == 1) {
      v = S.X0[e] - s_alpha[r] * S.Gd[e];
      if (A.lam && c >= A.prox_lo && c < A.prox_hi) v = soft_thr(v, A.lam[(size_t)row * A.ldlam + (c - A.prox_lo)]);
    }
    A.X[(size_t)row * KPAD + c] = v;
  }
  if (tid < 8 && s_rows[tid] >= 0) {
    if (A.pend) A.pend[s_rows[tid]] = s_pend[tid];
    if (!s_pend[tid]) {
      A.ell_fin[s_rows[tid]] = s_efin[tid];
      if (A.neval) A.neval[s_rows[tid]] = s_ne[tid];
    }
    if (A.eval_total) atomicAdd(A.eval_total, (unsigned long long)s_ne[tid]);
  }
  if (tid == 0 && A.eval_total) {  // lockstep cost of the tile: rows x the slowest row's candidates
    int mx = 0;
    for (int r = 0; r < 8; ++r) mx = max(mx, s_ne[r]);
    atomicAdd(A.eval_total + 2, (unsigned long long)(8 * mx));
  }
}

// ---------------------------------------------------------------------------
// Phase B of the line search: NC candidates per pass over the data.
//
// In stage 2 most genes fail the Armijo test for every candidate (the histogram of candidates per
// gene is 1 | 11-13, nothing in between), so walking the candidates one pass at a time re-stages
// M and Y a dozen times and gives a lane only four dependency chains.  Here a warp keeps the two
// tiles of Y and the B fragments of a k-step in registers and runs NC candidate rows (A fragments
// from shared memory) against them: 2 NC independent DMMA chains and 4 NC independent epilogues per
// lane.  The accept / stop / fallback rule is then applied to the NC objectives in candidate order,
// exactly as the sequential walk does.
// ---------------------------------------------------------------------------
template <int KS, int MODE, int NC>
__device__ __forceinline__ void tile_pass_multi_mode(const PassCfg& P, const FmTables& FT, const int* s_rows,
                                                     const double* s_Xc, const Smem& S, Pipe& pipe, unsigned mv) {
  constexpr int KPAD = 4 * KS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int tc = P.tc;
  const int row = s_rows[lr];
  double nu_r = 1.0, inu_r = 1.0, lnu_r = 0.0;
  bool pois_r = false;
  if (mode_row(MODE) && row >= 0) {
    nu_r = P.nu[row];
    inu_r = P.inv_nu[row];
    lnu_r = P.log_nu[row];
    pois_r = nu_r > P.thres;
  }
  const NbClip clip = P.clip;
  const double* ap = s_Xc + lr * KPAD + lc;  // A fragment of candidate c, k-step ks: ap[c * 8 * KPAD + 4 * ks]
  double accA[NC], accB[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) accA[c] = accB[c] = 0.0;

  const int nchunks = (P.ncols + tc - 1) / tc;
  const int chunk_elems = tc * KPAD;
  const int ldy = tc + 8;
  stage_chunk<KS>(P, 0, S.M, S.cp, S.Ys, s_rows, pipe.bar + 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 < nchunks) {
      const int sn = (ch + 1) & 1;
      stage_chunk<KS>(P, ch + 1, S.M + sn * chunk_elems, S.cp + sn * tc, S.Ys + sn * 8 * ldy, s_rows, pipe.bar + sn);
    }
    mbar_wait(pipe.bar + (ch & 1), (pipe.phase >> (ch & 1)) & 1u);
    pipe.phase ^= 1u << (ch & 1);
    if (ch == nchunks - 1) __syncthreads();  // the ragged-tail zero fill (generic stores) of the last chunk
    const double* Mb = S.M + (ch & 1) * chunk_elems;
    const double4* Cb = S.cp + (ch & 1) * tc;
    const double* Yb = S.Ys + (ch & 1) * 8 * ldy + lr * ldy + 2 * lc;
    const int nrem = P.ncols - ch * tc;
    const int ntile = min(tc, round_up(nrem, 8)) >> 3;
    auto pair_body = [&](auto ragged_tag, int tA, int tB, bool hasB) {
      constexpr bool RAGGED = decltype(ragged_tag)::value;
      const int clA = tA * 8 + 2 * lc, clB = tB * 8 + 2 * lc;
      const double2 yA = *reinterpret_cast<const double2*>(Yb + tA * 8);
      const double2 yB = *reinterpret_cast<const double2*>(Yb + tB * 8);
      const double* bpA = Mb + (tA * 8 + lr) * KPAD + lc;
      const double* bpB = Mb + (tB * 8 + lr) * KPAD + lc;
      double pa[NC][2], pb[NC][2];
#pragma unroll
      for (int c = 0; c < NC; ++c) pa[c][0] = pa[c][1] = pb[c][0] = pb[c][1] = 0.0;
      // k-steps whose four columns the step direction never touches (bit clear in mv) contribute
      // the same product to every candidate: they are accumulated once (stage 2 moves only the
      // treatment columns: 4 of the 7 k-steps at k = 26)
      double fA0 = 0.0, fA1 = 0.0, fB0 = 0.0, fB1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double bA = bpA[ks * 4], bB = bpB[ks * 4];
        if (!((mv >> ks) & 1u)) {  // block uniform
          const double a = ap[ks * 4];
          dmma884(fA0, fA1, a, bA);
          dmma884(fB0, fB1, a, bB);
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const double a = ap[c * 8 * KPAD + ks * 4];
            dmma884(pa[c][0], pa[c][1], a, bA);
            dmma884(pb[c][0], pb[c][1], a, bB);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        pa[c][0] += fA0;
        pa[c][1] += fA1;
        pb[c][0] += fB0;
        pb[c][1] += fB1;
      }
      double4 cA0 = make_double4(nu_r, inu_r, lnu_r, pois_r ? 1.0 : 0.0), cA1 = cA0, cB0 = cA0, cB1 = cA0;
      if (mode_col(MODE)) {
        cA0 = Cb[clA];
        cA1 = Cb[clA + 1];
        cB0 = Cb[clB];
        cB1 = Cb[clB + 1];
      }
      bool vA0 = true, vA1 = true, vB0 = true, vB1 = true;
      if (RAGGED) {
        vA0 = clA < nrem;
        vA1 = clA + 1 < nrem;
        vB0 = hasB && clB < nrem;
        vB1 = hasB && clB + 1 < nrem;
      }
      double dummy = 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if constexpr (mode_fast(MODE)) {
          const double th4[4] = {pa[c][0], pb[c][0], pa[c][1], pb[c][1]};
          const double y4[4] = {yA.x, yB.x, yA.y, yB.y};
          const double nu4[4] = {cA0.x, cB0.x, cA1.x, cB1.x};
          const double in4[4] = {cA0.y, cB0.y, cA1.y, cB1.y};
          const double ln4[4] = {cA0.z, cB0.z, cA1.z, cB1.z};
          const bool v4[4] = {vA0, vB0, vA1, vB1};
          double d4[4];
          eval_entries4<MODE, false, RAGGED>(FT, clip, th4, y4, nu4, in4, ln4, v4, accA[c], accB[c], d4);
        } else {
          eval_entry<MODE, false, RAGGED>(FT, clip, pa[c][0], yA.x, cA0.x, cA0.y, cA0.z, cA0.w != 0.0, vA0, accA[c], dummy);
          eval_entry<MODE, false, RAGGED>(FT, clip, pb[c][0], yB.x, cB0.x, cB0.y, cB0.z, cB0.w != 0.0, vB0, accB[c], dummy);
          eval_entry<MODE, false, RAGGED>(FT, clip, pa[c][1], yA.y, cA1.x, cA1.y, cA1.z, cA1.w != 0.0, vA1, accA[c], dummy);
          eval_entry<MODE, false, RAGGED>(FT, clip, pb[c][1], yB.y, cB1.x, cB1.y, cB1.z, cB1.w != 0.0, vB1, accB[c], dummy);
        }
      }
    };
#pragma unroll 1
    for (int t = warp; t < ntile; t += 2 * NWARPS) {
      const bool hasB = t + NWARPS < ntile;
      const int tB = hasB ? t + NWARPS : t;
      if (!hasB || tB * 8 + 8 > nrem)
        pair_body(std::true_type{}, t, tB, hasB);
      else
        pair_body(std::false_type{}, t, tB, hasB);
    }
    __syncthreads();
  }
  // fixed-order reduction per candidate: 4 lanes of a row, then the 8 warps (scratch: S.Gred / S.G)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double e = accA[c] + accB[c];
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    if (lc == 0) S.Gred[(c * NWARPS + warp) * 8 + lr] = e;
  }
  __syncthreads();
  if (tid < 8 * NC) {
    const int c = tid >> 3, r = tid & 7;
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) sum += S.Gred[(c * NWARPS + w) * 8 + r];
    S.G[c * 8 + r] = sum;
  }
  __syncthreads();
}

template <int KS, int NC>
__device__ __forceinline__ void tile_pass_multi(const PassCfg& P, const FmTables& FT, const int* s_rows,
                                                const double* s_Xc, const Smem& S, Pipe& pipe, unsigned mv) {
  if (P.family != CA_FAMILY_NB)
    tile_pass_multi_mode<KS, MODE_POIS, NC>(P, FT, s_rows, s_Xc, S, pipe, mv);
  else if (P.nu_per_row && P.any_pois)
    tile_pass_multi_mode<KS, MODE_NB_ROW, NC>(P, FT, s_rows, s_Xc, S, pipe, mv);
  else if (P.nu_per_row)
    tile_pass_multi_mode<KS, MODE_NB_ROW_F, NC>(P, FT, s_rows, s_Xc, S, pipe, mv);
  else if (P.any_pois)
    tile_pass_multi_mode<KS, MODE_NB_COL, NC>(P, FT, s_rows, s_Xc, S, pipe, mv);
  else
    tile_pass_multi_mode<KS, MODE_NB_COL_F, NC>(P, FT, s_rows, s_Xc, S, pipe, mv);
}

#ifndef CA_SEARCH_NC
#define CA_SEARCH_NC 4
#endif
constexpr int SEARCH_NC = CA_SEARCH_NC;  // candidates per pass in phase B

// Rows listed in A.rowlist failed their first candidate in phase A (A.f01s / A.e01s hold its
// objective and log-likelihood sum; X is untouched).  Candidates j = 1, 2, ... are evaluated NC at a
// time; per row the walk of causarray/gcate_opt.py:60-82 is replayed over the objectives in order.
template <int KS, int NC>
__global__ void __launch_bounds__(NTHREADS, CA_MINB) k_search_multi(PassCfg P, SearchArgs A) {
  constexpr int KPAD = 4 * KS;
  extern __shared__ __align__(16) double smem[];
  const Smem S = carve(smem, P.tc, KPAD, 8, NC + 3);
  double* Xc = S.X;                        // NC x 8 x KPAD candidate rows
  double* X0 = S.X + NC * 8 * KPAD;
  double* Gd = X0 + 8 * KPAD;
  double* Xfin = Gd + 8 * KPAD;
  __shared__ int s_rows[8];
  __shared__ __align__(8) unsigned long long s_bars[2];
  __shared__ double s_tlast[8], s_tc[NC][8], s_f0[8], s_ng[8], s_f01[8], s_e01[8], s_alpha[8], s_efin[8];
  __shared__ unsigned char s_done[8];
  __shared__ int s_mode[8], s_ne[8], s_sel[8], s_alldone;
  const int tid = threadIdx.x;
  const int total = A.rowlist ? *A.count : A.nrows;
  const int tile0 = blockIdx.x * TILE_ROWS;
  if (tile0 >= total) return;
  Pipe pipe;
  pipe_init(pipe, s_bars);
  const FmTables FT = fm_load_tables(S.tab, P.fm_tables);
  if (tid < 8) {
    const int i = tile0 + tid;
    const int row = (i < total) ? (A.rowlist ? A.rowlist[i] : i) : -1;
    s_rows[tid] = row;
    s_done[tid] = row < 0;
    s_mode[tid] = 0;
    s_ne[tid] = 1;  // the first candidate was evaluated in phase A
    s_sel[tid] = -1;
    if (row >= 0) {
      s_alpha[tid] = A.alpha_row ? A.alpha_row[row] : A.alpha;
      s_tlast[tid] = s_alpha[tid];
      s_f0[tid] = A.f0[row];
      s_ng[tid] = A.normg[row];
      s_f01[tid] = A.f01s[row];
      s_e01[tid] = A.e01s[row];
    } else {
      s_tlast[tid] = 0.0;
    }
  }
  __syncthreads();
  for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD, row = s_rows[r];
    X0[e] = (row >= 0) ? A.X[(size_t)row * KPAD + c] : 0.0;
    Gd[e] = (row >= 0) ? A.g[(size_t)row * KPAD + c] : 0.0;
    Xfin[e] = X0[e];
  }
  const double ncols_d = (double)P.ncols;
  const int nprox = A.prox_hi - A.prox_lo;
  unsigned mv = 0;  // k-steps that overlap the moving columns
  for (int ks = 0; ks < KS; ++ks)
    if (4 * ks < A.mv_hi && 4 * ks + 4 > A.mv_lo) mv |= 1u << ks;
  int groups = 0;
  for (int j0 = 1; j0 < A.max_iters; j0 += NC) {
    ++groups;
    if (tid < 8) {
      double t = s_tlast[tid];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        t = t * A.beta;
        s_tc[c][tid] = t;
      }
      s_tlast[tid] = t;
    }
    __syncthreads();
    for (int e = tid; e < NC * 8 * KPAD; e += NTHREADS) {
      const int c = e / (8 * KPAD), rc = e % (8 * KPAD), r = rc / KPAD, col = rc % KPAD;
      double v = X0[rc] - s_tc[c][r] * Gd[rc];
      if (A.lam && col >= A.prox_lo && col < A.prox_hi && s_rows[r] >= 0)
        v = soft_thr(v, A.lam[(size_t)s_rows[r] * A.ldlam + (col - A.prox_lo)]);
      Xc[e] = v;
    }
    __syncthreads();
    tile_pass_multi<KS, NC>(P, FT, s_rows, Xc, S, pipe, mv);
    if (tid < 8 && !s_done[tid]) {
      const int row = s_rows[tid];
      for (int c = 0; c < NC; ++c) {
        const int it = j0 + c;  // iteration index of the reference's loop
        if (it >= A.max_iters) break;
        double pen = 0.0;
        if (A.lam && nprox > 0) {
          for (int q = 0; q < nprox; ++q)
            pen += A.lam[(size_t)row * A.ldlam + q] * fabs(Xc[(c * 8 + tid) * KPAD + A.prox_lo + q]);
          pen /= (double)nprox;
        }
        const double ell = S.G[c * 8 + tid];
        const double f1 = -(ell + A.tys[row]) / ncols_d + pen;
        const double t = s_tc[c][tid];
        s_ne[tid] += 1;
        if (f1 < s_f0[tid] - A.tol * t * s_ng[tid]) {
          s_done[tid] = 1;
          s_sel[tid] = c;
          s_efin[tid] = ell;
          break;
        } else if (t < 1e-4 || it == A.max_iters - 1) {
          s_done[tid] = 1;
          if (s_f01[tid] < 1.1 * f1) {
            s_mode[tid] = 1;  // fall back to the full initial step
            s_sel[tid] = NC;
            s_efin[tid] = s_e01[tid];
          } else {
            s_sel[tid] = c;
            s_efin[tid] = ell;
          }
          break;
        }
      }
    }
    __syncthreads();
    // keep the chosen point of rows that finished in this group (Xc is overwritten by the next group)
    for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
      const int r = e / KPAD, col = e % KPAD, sel = s_sel[r];
      if (sel < 0) continue;
      double v;
      if (sel == NC) {
        v = X0[e] - s_alpha[r] * Gd[e];
        if (A.lam && col >= A.prox_lo && col < A.prox_hi)
          v = soft_thr(v, A.lam[(size_t)s_rows[r] * A.ldlam + (col - A.prox_lo)]);
      } else {
        v = Xc[(sel * 8 + r) * KPAD + col];
      }
      Xfin[e] = v;
    }
    __syncthreads();
    if (tid == 0) {
      int all = 1;
      for (int r = 0; r < 8; ++r) {
        all &= (int)s_done[r];
        s_sel[r] = -1;
      }
      s_alldone = all;
    }
    __syncthreads();
    if (s_alldone) break;
  }
  for (int e = tid; e < 8 * KPAD; e += NTHREADS) {
    const int r = e / KPAD, c = e % KPAD, row = s_rows[r];
    if (row >= 0) A.X[(size_t)row * KPAD + c] = Xfin[e];
  }
  if (tid < 8 && s_rows[tid] >= 0) {
    A.ell_fin[s_rows[tid]] = s_efin[tid];
    if (A.neval) A.neval[s_rows[tid]] = s_ne[tid];
    if (A.eval_total) atomicAdd(A.eval_total, (unsigned long long)(s_ne[tid] - 1));
  }
  if (tid == 0 && A.eval_total) atomicAdd(A.eval_total + 2, (unsigned long long)(8 * NC * groups));
}

// per-KS launchers (explicitly instantiated in rowpass_inst.cu, one object file per KS)
template <int KS>
int launch_rowpass_ks(const PassCfg& P, const double* X, const int* rowlist, const int* count, int nrows,
                      bool want_grad, int gc_lo, int gc_n, double* ell_out, double* G_out, int ldG, cudaStream_t st);
template <int KS>
int launch_search_ks(const PassCfg& P, const SearchArgs& A, cudaStream_t st);
template <int KS>
int launch_search_multi_ks(const PassCfg& P, const SearchArgs& A, cudaStream_t st);

}  // namespace ca
