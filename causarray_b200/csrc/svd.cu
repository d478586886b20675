// K7 of SURVEY.md section 2b: top-r singular triplets of the (n x p) deviance-residual matrix for the
// GCATE initialisation (replaces scipy.sparse.linalg.svds / ARPACK, causarray/gcate_opt.py:312).
//
// Block subspace iteration with Rayleigh-Ritz extraction, all on the device and without a library GEMM
// or factorisation (round 1 drove cuBLAS / cuSOLVER through torch: ~150 small launches and a dozen host
// synchronisations, 8 ms per batch for ~1 ms of kernel time):
//   V <- orth(R^T orth(R V))            two tall-skinny FP64 DMMA products per iteration (k_ts_gemm), each
//                                       one streaming pass over R (resp. its gene-major copy R^T, which the
//                                       residual kernel writes anyway), block width b = r + 20 <= 64
//   orth = CholeskyQR2                  b x b Gram by a fixed-order two-stage reduction, Cholesky + triangular
//                                       solve of every row inside one kernel (each CTA refactors the tiny
//                                       Gram matrix itself)
//   Ritz step every 4th iteration       T = (R V)^T (R V) on the device, cyclic Jacobi eigen-solver of the
//                                       b x b matrix on the host, rotation applied on the device; stop when
//                                       max_k || R^T u_k - s_k v_k || <= tol s_1
// Singular subspaces are deterministic; signs and the ascending order of svds are fixed by the caller.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "devbuf.h"
#include "prof.h"

namespace ca {

constexpr int TS_ROWS = 16;   // output rows per CTA (small tiles: 250-500 CTAs, 4 per SM -- the product is bound by
                              // the latency of the chunk loads, so bytes in flight matter more than operand reuse)
constexpr int TS_KC = 64;     // contraction chunk

// Y (m x BW) = M (m x k, row-major, leading dimension ldm) * V (k x BW, leading dimension BW)
template <int BW>
__global__ void __launch_bounds__(256, 4) k_ts_gemm(const double* __restrict__ M, long ldm, int m, int k,
                                                 const double* __restrict__ V, double* __restrict__ Y) {
  constexpr int LDM = TS_KC + 4, LDV = BW + 4;
  constexpr int NT = BW / 8;              // n-tiles
  constexpr int TPW = 2 * NT / 8;         // output tiles per warp (2 m-tiles x NT n-tiles over 8 warps)
  extern __shared__ __align__(16) double sm[];
  double* Ms = sm;                          // 2 x TS_ROWS x LDM
  double* Vs = sm + 2 * TS_ROWS * LDM;      // 2 x TS_KC x LDV
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int row0 = blockIdx.x * TS_ROWS;
  const int mt = warp >> 2;                 // m-tile of this warp (0..1)
  const int nt0 = (warp & 3) * TPW;         // first n-tile
  double acc[TPW][2];
#pragma unroll
  for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = 0.0;
  const int nchunks = (k + TS_KC - 1) / TS_KC;
  auto load = [&](int c, int stage) {
    const int k0 = c * TS_KC;
    double* ms = Ms + stage * TS_ROWS * LDM;
    double* vs = Vs + stage * TS_KC * LDV;
    for (int e = tid; e < TS_ROWS * (TS_KC / 2); e += 256) {      // 16-byte pieces of the M tile
      const int r = e / (TS_KC / 2), q = (e % (TS_KC / 2)) * 2;
      const int gr = row0 + r, gk = k0 + q;
      double* dst = ms + r * LDM + q;
      if (gr < m && gk + 1 < k && ((ldm & 1) == 0)) {
        cp_async16(dst, M + (size_t)gr * ldm + gk);
      } else {
        dst[0] = (gr < m && gk < k) ? M[(size_t)gr * ldm + gk] : 0.0;
        dst[1] = (gr < m && gk + 1 < k) ? M[(size_t)gr * ldm + gk + 1] : 0.0;
      }
    }
    for (int e = tid; e < TS_KC * (BW / 2); e += 256) {
      const int r = e / (BW / 2), q = (e % (BW / 2)) * 2;
      const int gk = k0 + r;
      double* dst = vs + r * LDV + q;
      if (gk < k) {
        cp_async16(dst, V + (size_t)gk * BW + q);
      } else {
        dst[0] = dst[1] = 0.0;
      }
    }
    cp_async_commit();
  };
  load(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load(c + 1, (c + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* ms = Ms + (c & 1) * TS_ROWS * LDM + (mt * 8 + lr) * LDM + lc;
    const double* vs = Vs + (c & 1) * TS_KC * LDV + lc * LDV + lr;
#pragma unroll 4
    for (int ks = 0; ks < TS_KC / 4; ++ks) {
      const double a = ms[ks * 4];
#pragma unroll
      for (int t = 0; t < TPW; ++t) dmma884(acc[t][0], acc[t][1], a, vs[ks * 4 * LDV + (nt0 + t) * 8]);
    }
    __syncthreads();
  }
  const int gr = row0 + mt * 8 + lr;
  if (gr < m) {
#pragma unroll
    for (int t = 0; t < TPW; ++t)
      *reinterpret_cast<double2*>(Y + (size_t)gr * BW + (nt0 + t) * 8 + 2 * lc) = make_double2(acc[t][0], acc[t][1]);
  }
}

// partial Gram matrices: part[blk][i][j] = sum over the CTA's rows of X[row][i] * Z[row][j]
template <int BW>
__global__ void __launch_bounds__(256) k_ts_gram(const double* __restrict__ X, const double* __restrict__ Z, int m,
                                                 double* __restrict__ part) {
  __shared__ double xs[32][BW + 1], zs[32][BW + 1];
  const int tid = threadIdx.x;
  const int rows_per = (m + gridDim.x - 1) / gridDim.x;
  const int r_lo = blockIdx.x * rows_per, r_hi = min(m, r_lo + rows_per);
  constexpr int EPT = (BW * BW + 255) / 256;
  double acc[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) acc[e] = 0.0;
  for (int r0 = r_lo; r0 < r_hi; r0 += 32) {
    for (int e = tid; e < 32 * BW; e += 256) {
      const int r = e / BW, c = e % BW;
      const bool ok = r0 + r < r_hi;
      xs[r][c] = ok ? X[(size_t)(r0 + r) * BW + c] : 0.0;
      zs[r][c] = ok ? Z[(size_t)(r0 + r) * BW + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = tid + e * 256;
      if (idx < BW * BW) {
        const int i = idx / BW, j = idx % BW;
        double s = acc[e];
#pragma unroll 8
        for (int r = 0; r < 32; ++r) s = fma(xs[r][i], zs[r][j], s);
        acc[e] = s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int idx = tid + e * 256;
    if (idx < BW * BW) part[(size_t)blockIdx.x * BW * BW + idx] = acc[e];
  }
}

// G = sum of the partials in block order (deterministic)
__global__ void k_sum_parts(const double* __restrict__ part, int nparts, int len, double* __restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int b = 0; b < nparts; ++b) s += part[(size_t)b * len + i];
  G[i] = s;
}

// One round of Cholesky-QR: X <- X L^-T with L L^T = G (bw x bw leading block of a BW-wide buffer).
// Every CTA factors G itself (a few microseconds) and transforms its rows.  flag != 0: G not positive definite.
template <int BW>
__global__ void __launch_bounds__(256) k_cholqr_apply(double* __restrict__ X, int m, int bw, const double* __restrict__ G,
                                                      int* __restrict__ flag) {
  __shared__ double L[BW][BW + 1];
  __shared__ double invd[BW];
  __shared__ int bad;
  const int tid = threadIdx.x;
  for (int e = tid; e < BW * BW; e += 256) L[e / BW][e % BW] = G[e];
  if (tid == 0) bad = 0;
  __syncthreads();
  if (tid < 32) {
    for (int j = 0; j < bw; ++j) {
      const double d = L[j][j];
      if (!(d > 0.0) && tid == 0) bad = 1;
      const double inv = rsqrt(d > 1e-300 ? d : 1e-300);
      __syncwarp();
      for (int i = j + tid; i < bw; i += 32) L[i][j] *= inv;        // column j (diagonal becomes sqrt(d))
      __syncwarp();
      for (int i = j + 1 + tid; i < bw; i += 32) {
        const double lij = L[i][j];
        for (int c = j + 1; c <= i; ++c) L[i][c] = fma(-lij, L[c][j], L[i][c]);
      }
      __syncwarp();
    }
    for (int j = tid; j < bw; j += 32) invd[j] = 1.0 / L[j][j];
  }
  __syncthreads();
  if (bad) {
    if (tid == 0 && blockIdx.x == 0) *flag = 1;
    return;
  }
  for (int row = blockIdx.x * 256 + tid; row < m; row += gridDim.x * 256) {
    double* x = X + (size_t)row * BW;
    double y[BW];
#pragma unroll
    for (int j = 0; j < BW; ++j) y[j] = x[j];
    // solve y L^T = x  <=>  L y^T = x^T  (forward substitution)
#pragma unroll
    for (int j = 0; j < BW; ++j) {
      if (j < bw) {
        double v = y[j];
#pragma unroll
        for (int c = 0; c < j; ++c) v = fma(-L[j][c], y[c], v);
        y[j] = v * invd[j];
      }
    }
#pragma unroll
    for (int j = 0; j < BW; ++j) x[j] = j < bw ? y[j] : 0.0;
  }
}

// X <- X W (row-wise product with a small BW x BW matrix held in shared memory)
template <int BW>
__global__ void __launch_bounds__(256) k_mul_small(const double* __restrict__ X, int m, const double* __restrict__ W,
                                                   double* __restrict__ out) {
  __shared__ double Ws[BW][BW + 1];
  for (int e = threadIdx.x; e < BW * BW; e += 256) Ws[e / BW][e % BW] = W[e];
  __syncthreads();
  for (int row = blockIdx.x * 256 + threadIdx.x; row < m; row += gridDim.x * 256) {
    double x[BW], y[BW];
#pragma unroll
    for (int j = 0; j < BW; ++j) x[j] = X[(size_t)row * BW + j];
#pragma unroll
    for (int j = 0; j < BW; ++j) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < BW; ++c) s = fma(x[c], Ws[c][j], s);
      y[j] = s;
    }
#pragma unroll
    for (int j = 0; j < BW; ++j) out[(size_t)row * BW + j] = y[j];
  }
}

// squared column norms of (Z - Vr diag(s)) restricted to the first r columns: partial sums per CTA
template <int BW>
__global__ void __launch_bounds__(256) k_ritz_resid(const double* __restrict__ Z, const double* __restrict__ Vr,
                                                    const double* __restrict__ s, int m, int r, double* __restrict__ part) {
  __shared__ double red[256];
  const int rows_per = (m + gridDim.x - 1) / gridDim.x;
  const int r_lo = blockIdx.x * rows_per, r_hi = min(m, r_lo + rows_per);
  for (int c = 0; c < r; ++c) {
    double a = 0.0;
    for (int row = r_lo + threadIdx.x; row < r_hi; row += 256) {
      const double d = Z[(size_t)row * BW + c] - Vr[(size_t)row * BW + c] * s[c];
      a = fma(d, d, a);
    }
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * BW + c] = red[0];
    __syncthreads();
  }
}

// standard-normal start block: counter-based (splitmix64 of the element index and the seed) + Box-Muller
__global__ void k_randn_block(double* V, long count, int BWc, int b, unsigned long long seed) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % BWc);
    unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    unsigned long long w = z * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull;
    w = (w ^ (w >> 30)) * 0xBF58476D1CE4E5B9ull;
    w = (w ^ (w >> 27)) * 0x94D049BB133111EBull;
    w ^= w >> 31;
    const double u1 = ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(w >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    V[i] = col < b ? sqrt(-2.0 * log(u1)) * cospi(2.0 * u2) : 0.0;
  }
}

// zero the listed columns of R (n x p) and rows of Rt (p x ldt)
__global__ void k_zero_genes(double* R, int n, int p, double* Rt, long ldt, const unsigned char* __restrict__ mask) {
  const int g = blockIdx.x;
  if (!mask[g]) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    R[(size_t)i * p + g] = 0.0;
    if (Rt) Rt[(size_t)g * ldt + i] = 0.0;
  }
}

// symmetric eigen-decomposition by cyclic Jacobi rotations (host, b <= 64): A = W diag(lam) W^T
static void jacobi_eigh(std::vector<double>& A, int b, std::vector<double>& W, std::vector<double>& lam) {
  W.assign((size_t)b * b, 0.0);
  for (int i = 0; i < b; ++i) W[(size_t)i * b + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < b; ++i)
      for (int j = 0; j < b; ++j) (i == j ? diag : off) += A[(size_t)i * b + j] * A[(size_t)i * b + j];
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < b - 1; ++p)
      for (int q = p + 1; q < b; ++q) {
        const double apq = A[(size_t)p * b + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * b + p], aqq = A[(size_t)q * b + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < b; ++k) {
          const double akp = A[(size_t)k * b + p], akq = A[(size_t)k * b + q];
          A[(size_t)k * b + p] = c * akp - s * akq;
          A[(size_t)k * b + q] = s * akp + c * akq;
        }
        for (int k = 0; k < b; ++k) {
          const double apk = A[(size_t)p * b + k], aqk = A[(size_t)q * b + k];
          A[(size_t)p * b + k] = c * apk - s * aqk;
          A[(size_t)q * b + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < b; ++k) {
          const double wkp = W[(size_t)k * b + p], wkq = W[(size_t)k * b + q];
          W[(size_t)k * b + p] = c * wkp - s * wkq;
          W[(size_t)k * b + q] = s * wkp + c * wkq;
        }
      }
  }
  lam.resize(b);
  for (int i = 0; i < b; ++i) lam[i] = A[(size_t)i * b + i];
}

template <int BW>
static int svd_topr_impl(double* R, double* Rt, long ldt, int n, int p, int r, int b, int max_iter, double tol,
                         unsigned seed, double* U, double* S, double* Vt, int* iters_out, cudaStream_t st) {
  const int NPART = 32;
  DevBuf<double> V, Q, Z, Bm, part, G, Wd, sd;
  DevBuf<int> flag;
  if (V.alloc((size_t)p * BW) || Q.alloc((size_t)n * BW) || Z.alloc((size_t)p * BW) || Bm.alloc((size_t)n * BW) ||
      part.alloc((size_t)NPART * BW * BW) || G.alloc((size_t)BW * BW) || Wd.alloc((size_t)BW * BW) || sd.alloc(BW) ||
      flag.alloc(1))
    return -1;
  CA_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int), st));
  k_randn_block<<<296, 256, 0, st>>>(V.ptr, (long)p * BW, BW, b, (unsigned long long)seed);
  CA_LAUNCH_CHECK();
  const size_t sm_gemm = sizeof(double) * (2 * TS_ROWS * (TS_KC + 4) + 2 * TS_KC * (BW + 4));
  CA_CUDA(cudaFuncSetAttribute(k_ts_gemm<BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gemm));
  auto gemm = [&](const double* M, long ldm, int m, int k, const double* Vin, double* Yout) -> int {
    k_ts_gemm<BW><<<(m + TS_ROWS - 1) / TS_ROWS, 256, sm_gemm, st>>>(M, ldm, m, k, Vin, Yout);
    CA_LAUNCH_CHECK();
    return 0;
  };
  auto gram = [&](const double* X, const double* Zz, int m) -> int {
    k_ts_gram<BW><<<NPART, 256, 0, st>>>(X, Zz, m, part.ptr);
    CA_LAUNCH_CHECK();
    k_sum_parts<<<(BW * BW + 255) / 256, 256, 0, st>>>(part.ptr, NPART, BW * BW, G.ptr);
    CA_LAUNCH_CHECK();
    return 0;
  };
  // Cholesky-QR: one round keeps the block well conditioned between iterations (its loss of
  // orthogonality, ~eps kappa^2 of a block that was orthonormal one product ago, is removed by the next
  // round); two rounds (CholeskyQR2: orthonormal to machine precision) before a Rayleigh-Ritz step
  auto orth = [&](double* X, int m, int rounds) -> int {
    for (int round = 0; round < rounds; ++round) {
      if (gram(X, X, m)) return -1;
      k_cholqr_apply<BW><<<std::min(148, (m + 255) / 256), 256, 0, st>>>(X, m, b, G.ptr, flag.ptr);
      CA_LAUNCH_CHECK();
    }
    return 0;
  };
  std::vector<double> T((size_t)BW * BW), Tb, W, lam, Wfull((size_t)BW * BW), sv(BW), hpart((size_t)NPART * BW);
  bool have = false;
  int it = 0;
  for (; it < max_iter; ++it) {
    const bool ritz = it % 4 == 3 || it == max_iter - 1;
    if (gemm(R, p, n, p, V.ptr, Q.ptr)) return -1;          // Q = R V
    if (orth(Q.ptr, n, 1)) return -1;
    if (gemm(Rt, ldt, p, n, Q.ptr, V.ptr)) return -1;       // V = R^T Q
    if (orth(V.ptr, p, ritz ? 2 : 1)) return -1;
    if (ritz) {
      // Rayleigh-Ritz on span(V): Bm = R V, T = Bm^T Bm = W diag(s^2) W^T
      if (gemm(R, p, n, p, V.ptr, Bm.ptr)) return -1;
      if (gram(Bm.ptr, Bm.ptr, n)) return -1;
      int hflag = 0;
      CA_CUDA(cudaMemcpyAsync(T.data(), G.ptr, sizeof(double) * BW * BW, cudaMemcpyDeviceToHost, st));
      CA_CUDA(cudaMemcpyAsync(&hflag, flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
      CA_CUDA(cudaStreamSynchronize(st));
      if (hflag) {
        set_error("Cholesky-QR broke down (rank-deficient block)");
        return -3;
      }
      Tb.assign((size_t)b * b, 0.0);
      for (int i = 0; i < b; ++i)
        for (int j = 0; j < b; ++j) Tb[(size_t)i * b + j] = 0.5 * (T[(size_t)i * BW + j] + T[(size_t)j * BW + i]);
      jacobi_eigh(Tb, b, W, lam);
      std::vector<int> ord(b);
      for (int i = 0; i < b; ++i) ord[i] = i;
      std::sort(ord.begin(), ord.end(), [&](int x, int y) { return lam[x] > lam[y]; });
      std::fill(Wfull.begin(), Wfull.end(), 0.0);
      for (int c = 0; c < b; ++c) {
        sv[c] = sqrt(std::max(lam[ord[c]], 0.0));
        for (int i = 0; i < b; ++i) Wfull[(size_t)i * BW + c] = W[(size_t)i * b + ord[c]];
      }
      CA_CUDA(cudaMemcpyAsync(Wd.ptr, Wfull.data(), sizeof(double) * BW * BW, cudaMemcpyHostToDevice, st));
      CA_CUDA(cudaMemcpyAsync(sd.ptr, sv.data(), sizeof(double) * BW, cudaMemcpyHostToDevice, st));
      // Ub S = Bm W (into Q), Vr = V W (into Z... kept in V after the test): residual || R^T (Ub) - Vr S ||
      k_mul_small<BW><<<std::min(148, (n + 255) / 256), 256, 0, st>>>(Bm.ptr, n, Wd.ptr, Q.ptr);   // Q = Ub diag(s)
      CA_LAUNCH_CHECK();
      k_mul_small<BW><<<std::min(148, (p + 255) / 256), 256, 0, st>>>(V.ptr, p, Wd.ptr, Z.ptr);    // Z = Vr
      CA_LAUNCH_CHECK();
      // R^T (Ub s) = Vr s^2 at convergence: compare R^T Q with Vr diag(s^2), scaled back by s afterwards
      if (gemm(Rt, ldt, p, n, Q.ptr, V.ptr)) return -1;   // V <- R^T (Ub s)  (V is rebuilt from Vr below)
      std::vector<double> s2(BW, 0.0);
      for (int c = 0; c < b; ++c) s2[c] = sv[c] * sv[c];
      CA_CUDA(cudaMemcpyAsync(sd.ptr, s2.data(), sizeof(double) * BW, cudaMemcpyHostToDevice, st));
      k_ritz_resid<BW><<<NPART, 256, 0, st>>>(V.ptr, Z.ptr, sd.ptr, p, r, part.ptr);
      CA_LAUNCH_CHECK();
      CA_CUDA(cudaMemcpyAsync(hpart.data(), part.ptr, sizeof(double) * NPART * BW, cudaMemcpyDeviceToHost, st));
      CA_CUDA(cudaStreamSynchronize(st));
      double res = 0.0;
      for (int c = 0; c < r; ++c) {
        double a = 0.0;
        for (int bl = 0; bl < NPART; ++bl) a += hpart[(size_t)bl * BW + c];
        const double rc = sv[c] > 0.0 ? sqrt(a) / sv[c] : 0.0;     // || R^T u_c - s_c v_c ||
        res = std::max(res, rc);
      }
      // continue the iteration from the rotated basis Vr (orthonormal: V orthonormal, W orthogonal)
      CA_CUDA(cudaMemcpyAsync(V.ptr, Z.ptr, sizeof(double) * (size_t)p * BW, cudaMemcpyDeviceToDevice, st));
      have = true;
      if (!(res == res)) {
        set_error("non-finite Ritz residual");
        return -3;
      }
      if (res <= tol * sv[0]) {
        ++it;
        break;
      }
    }
  }
  if (!have) {
    set_error("subspace iteration ended without a Ritz step");
    return -2;
  }
  // outputs: U = (Ub s) / s (n x r), S (r), Vt (r x p) from Vr; all in descending order
  std::vector<double> hq((size_t)n * BW), hv((size_t)p * BW);
  CA_CUDA(cudaMemcpyAsync(hq.data(), Q.ptr, sizeof(double) * hq.size(), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(hv.data(), Z.ptr, sizeof(double) * hv.size(), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < r; ++c) {
    S[c] = sv[c];
    const double inv = sv[c] > 0.0 ? 1.0 / sv[c] : 0.0;
    for (int i = 0; i < n; ++i) U[(size_t)i * r + c] = hq[(size_t)i * BW + c] * inv;
    for (int j = 0; j < p; ++j) Vt[(size_t)c * p + j] = hv[(size_t)j * BW + c];
  }
  if (iters_out) *iters_out = it;
  if (getenv("CA_SVD_DEBUG")) fprintf(stderr, "ca_svd_topr: n=%d p=%d r=%d b=%d iterations=%d\n", n, p, r, b, it);
  return 0;
}

}  // namespace ca

using namespace ca;

// R: n x p device matrix (row-major, leading dimension p); Rt: its transpose p x ldt on the device.
// zero_genes (p, host, may be NULL): columns to zero first (in both copies).  U (n x r), S (r), Vt (r x p)
// are host outputs in DESCENDING order of the singular values.  Returns 0, -1 (CUDA), -2 (argument),
// -3 (numerical breakdown: the caller may retry with another method).
extern "C" int ca_svd_topr(void* R, void* Rt, int64_t ldt, int n, int p, int r, int oversample, int max_iter, double tol,
                           uint32_t seed, const uint8_t* zero_genes, double* U, double* S, double* Vt, int32_t* n_iter) {
  CA_REQUIRE(R && Rt && U && S && Vt && n > 0 && p > 0 && r > 0, "bad arguments");
  const int b = std::min(std::min(n, p), r + oversample);
  CA_REQUIRE(b >= r && b <= 64, "block width r + oversample must be at most 64");
  CA_REQUIRE(ldt >= n, "leading dimension of the transposed copy too small");
  cudaStream_t st = 0;
  ProfScope prof(PROF_PREP, st);
  if (zero_genes) {
    bool any = false;
    for (int g = 0; g < p && !any; ++g) any = zero_genes[g] != 0;
    if (any) {
      DevBuf<unsigned char> m;
      if (m.alloc(p)) return -1;
      CA_CUDA(cudaMemcpyAsync(m.ptr, zero_genes, p, cudaMemcpyHostToDevice, st));
      k_zero_genes<<<p, 128, 0, st>>>((double*)R, n, p, (double*)Rt, ldt, m.ptr);
      CA_LAUNCH_CHECK();
      CA_CUDA(cudaStreamSynchronize(st));
    }
  }
  int it = 0;
  int rc;
  if (b <= 32)
    rc = svd_topr_impl<32>((double*)R, (double*)Rt, ldt, n, p, r, b, max_iter, tol, seed, U, S, Vt, &it, st);
  else
    rc = svd_topr_impl<64>((double*)R, (double*)Rt, ldt, n, p, r, b, max_iter, tol, seed, U, S, Vt, &it, st);
  if (n_iter) *n_iter = it;
  return rc;
}
