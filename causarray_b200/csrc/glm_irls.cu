// Host side of the persistent per-octet IRLS kernel (glm_irls_impl.cuh): split of the design into
// dense columns + one-hot groups, group-sorted padded cell order, the permuted copy of the counts,
// and the launch.  Replaces the iteration loop of the per-gene statsmodels fit
// (causarray/gcate_glm.py:193-236) for designs with at most IRLS_MAX_KDE dense columns.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "devbuf.h"
#include "glm_state.h"
#include "prof.h"
#include "fastmath.cuh"
#include "glm_irls_impl.cuh"

namespace ca {

// Yp[g][c] = perm[c] >= 0 ? Yt[g][perm[c]] : 0   (one CTA row per gene, coalesced writes)
// *not_counts is raised when an entry is not a non-negative integer (the IRLS kernel then computes every
// log y instead of reading its table of log 0 .. log 255).
__global__ void k_glm_permute_counts(const double* __restrict__ Yt, long ldn, int p, const int* __restrict__ perm,
                                     int npad, double* __restrict__ Yp, long ldp, int* __restrict__ not_counts) {
  const int gene = blockIdx.y;
  const double* src = Yt + (size_t)gene * ldn;
  double* dst = Yp + (size_t)gene * ldp;
  bool odd = false;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < npad; c += gridDim.x * blockDim.x) {
    const int s = perm[c];
    const double v = s >= 0 ? src[s] : 0.0;
    odd = odd || !(v >= 0.0) || v != floor(v);
    dst[c] = v;
  }
  if (__any_sync(0xffffffffu, odd) && (threadIdx.x & 31) == 0) *not_counts = 1;
}

// mean count per gene (the IRLS starting value mu0 = (y + ybar) / 2)
__global__ void k_glm_ybar(const double* __restrict__ Yt, long ldn, int n, double* __restrict__ ybar) {
  __shared__ double scratch[32];
  const int gene = blockIdx.x;
  const double* y = Yt + (size_t)gene * ldn;
  double s0 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s0 += y[i];
  s0 = block_sum(s0, scratch);
  if (threadIdx.x == 0) ybar[gene] = s0 / (double)n;
}

bool plan_irls(const double* X, int n, int kx, IrlsPlan& L) {
  L.ok = false;
  if (kx > 64 || getenv("CA_GLM_NO_IRLS")) return false;
  // ---- the one-hot block: longest contiguous run of {0, 1}-valued columns with at most one
  //      non-zero entry per cell (the treatment indicators of [W, A]) ----
  unsigned long long conflict[64] = {0};
  unsigned long long binary = kx >= 64 ? ~0ull : ((1ull << kx) - 1), ones = binary;
  for (int i = 0; i < n; ++i) {
    const double* x = X + (size_t)i * kx;
    unsigned long long m = 0;
    for (int c = 0; c < kx; ++c) {
      const double v = x[c];
      if (v != 0.0) {
        m |= 1ull << c;
        if (v != 1.0) {
          binary &= ~(1ull << c);
          ones &= ~(1ull << c);
        }
      } else {
        ones &= ~(1ull << c);
      }
    }
    for (unsigned long long r = m; r; r &= r - 1) conflict[__builtin_ctzll(r)] |= m;
  }
  int best_lo = 0, best_n = 0;
  for (int lo = 0; lo < kx; ++lo) {
    unsigned long long in = 0;
    int hi = lo;
    while (hi < kx && ((binary >> hi) & 1) && !((ones >> hi) & 1) && !(conflict[hi] & in & ~(1ull << hi))) {
      in |= 1ull << hi;
      ++hi;
    }
    if (hi - lo > best_n) {
      best_n = hi - lo;
      best_lo = lo;
    }
  }
  if (best_n < 2) best_n = 0;
  // ---- dense columns, the unit column last ----
  memset(L.dcol, -1, sizeof L.dcol);
  memset(L.gcol, -1, sizeof L.gcol);
  memset(L.cdense, -1, sizeof L.cdense);
  memset(L.cgroup, 0, sizeof L.cgroup);
  int unit = -1;
  for (int c = 0; c < kx && unit < 0; ++c) {
    const bool in_block = c >= best_lo && c < best_lo + best_n;
    if (!in_block && ((ones >> c) & 1)) unit = c;
  }
  const int n_dense_caller = kx - best_n;
  const int kde = n_dense_caller + (unit < 0 ? 1 : 0);
  if (kde > IRLS_MAX_KDE || kde < 1) return false;
  {
    int di = 0;
    for (int c = 0; c < kx; ++c) {
      const bool in_block = c >= best_lo && c < best_lo + best_n;
      if (in_block) {
        L.cgroup[c] = (unsigned char)(1 + c - best_lo);
        L.gcol[1 + c - best_lo] = (signed char)c;
      } else if (c != unit) {
        L.dcol[di] = (signed char)c;
        L.cdense[c] = (signed char)di;
        ++di;
      }
    }
    if (unit >= 0) {
      L.dcol[kde - 1] = (signed char)unit;
      L.cdense[unit] = (signed char)(kde - 1);
    }
  }
  L.kde = kde;
  L.kx = kx;
  L.ngroups = best_n + 1;
  // ---- cells sorted by group, every group padded to a multiple of 8 ----
  std::vector<int> grp(n, 0), cnt(L.ngroups + 1, 0);
  for (int i = 0; i < n; ++i) {
    const double* x = X + (size_t)i * kx + best_lo;
    int g = 0;
    for (int c = 0; c < best_n; ++c)
      if (x[c] != 0.0) g = 1 + c;
    grp[i] = g;
    ++cnt[g];
  }
  std::vector<int> start(L.ngroups + 1, 0);
  for (int g = 0; g < L.ngroups; ++g) start[g + 1] = start[g] + (cnt[g] + 7) / 8 * 8;
  const int npad = (start[L.ngroups] + IRLS_CHUNK - 1) / IRLS_CHUNK * IRLS_CHUNK;
  if (npad == 0) return false;
  L.npad = npad;
  L.nchunks = npad / IRLS_CHUNK;
  L.perm.assign(npad, -1);
  {
    std::vector<int> cur(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) L.perm[cur[grp[i]]++] = i;
  }
  L.gpair.assign(npad / 8, 0);
  {
    int last = 0;
    for (int g = 0; g < L.ngroups; ++g)
      for (int s = start[g] / 8; s < start[g + 1] / 8; ++s) L.gpair[s] = last = g;
    for (int s = start[L.ngroups] / 8; s < npad / 8; ++s) L.gpair[s] = last;  // trailing pad rows are all zero
  }
  // ---- chunk ranges of the warps and the flush segments they produce ----
  for (int w = 0; w <= IRLS_WARPS; ++w) L.chunk_lo[w] = (int)((long)w * L.nchunks / IRLS_WARPS);
  std::vector<int> seg_group;
  for (int w = 0; w < IRLS_WARPS; ++w) {
    L.seg_ptr[w] = (int)seg_group.size();
    int gcur = -1;
    for (int s = L.chunk_lo[w] * 4; s < L.chunk_lo[w + 1] * 4; ++s)
      if (L.gpair[s] != gcur) {
        gcur = L.gpair[s];
        seg_group.push_back(gcur);
      }
  }
  L.seg_ptr[IRLS_WARPS] = (int)seg_group.size();
  L.nseg_tot = (int)seg_group.size();
  if (L.nseg_tot > IRLS_MAX_SEG || L.nseg_tot > 255) return false;
  {
    int q = 0;
    for (int g = 0; g < L.ngroups; ++g) {
      L.gseg_ptr[g] = (unsigned char)q;
      for (int s = 0; s < L.nseg_tot; ++s)
        if (seg_group[s] == g) L.gseg_list[q++] = (unsigned char)s;
    }
    L.gseg_ptr[L.ngroups] = (unsigned char)q;
  }
  L.ok = true;
  return true;
}

static int launch_irls(const GlmIrls& P, int grid, size_t smem, cudaStream_t st) {
  switch (P.kde) {
    case 1: return launch_glm_irls_kde<1>(P, grid, smem, st);
    case 2: return launch_glm_irls_kde<2>(P, grid, smem, st);
    case 3: return launch_glm_irls_kde<3>(P, grid, smem, st);
    case 4: return launch_glm_irls_kde<4>(P, grid, smem, st);
    case 5: return launch_glm_irls_kde<5>(P, grid, smem, st);
    case 6: return launch_glm_irls_kde<6>(P, grid, smem, st);
    case 7: return launch_glm_irls_kde<7>(P, grid, smem, st);
    case 8: return launch_glm_irls_kde<8>(P, grid, smem, st);
    case 9: return launch_glm_irls_kde<9>(P, grid, smem, st);
    case 10: return launch_glm_irls_kde<10>(P, grid, smem, st);
    case 11: return launch_glm_irls_kde<11>(P, grid, smem, st);
    case 12: return launch_glm_irls_kde<12>(P, grid, smem, st);
    case 13: return launch_glm_irls_kde<13>(P, grid, smem, st);
    case 14: return launch_glm_irls_kde<14>(P, grid, smem, st);
    case 15: return launch_glm_irls_kde<15>(P, grid, smem, st);
    case 16: return launch_glm_irls_kde<16>(P, grid, smem, st);
  }
  set_error("unsupported number of dense design columns");
  return -2;
}

// Runs the whole IRLS of the genes in `genelist` (nullptr: all) on the device.  Expects in the
// handle: beta (warm start or zeros), nu, ridge (ridge | l1), status / n_iter prepared by the caller.
// Returns 1 when the design does not fit this path (the caller then runs the per-iteration loop).
int glm_fit_irls(ca_glm* h, const IrlsPlan& L, const double* X, const double* offset, int family, double thres,
                 bool has_nu, bool has_ridge, bool lasso, double tol, int maxiter, bool warm, bool freeze_dense,
                 const int* genelist_dev, int n_list) {
  if (freeze_dense && L.ngroups < 2) return 1;
  const int n = h->n, p = h->p, kx = L.kx, kde = L.kde;
  cudaStream_t st = h->st;
  const size_t smem = glm_irls_smem_bytes(kde, kx, L.ngroups, lasso);
  if (smem > 226 * 1024) return 1;   // (227 KB of dynamic shared memory per CTA on sm_100a)
  const int ldxd = (kde + 1) | 1;
  // ---- counts in the sorted padded cell order (cached across the fits of a batch) ----
  const long ldp = L.npad;
  if (h->perm_host != L.perm || !h->Yp.ptr || h->Yp_src != h->Yt) {
    if (h->Yp.alloc((size_t)p * ldp) || h->perm_dev.alloc(L.npad) || h->not_counts.alloc(1)) return -1;
    CA_CUDA(cudaMemcpyAsync(h->perm_dev.ptr, L.perm.data(), sizeof(int) * L.npad, cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaMemsetAsync(h->not_counts.ptr, 0, sizeof(int), st));
    ProfScope prof(PROF_PREP, st);
    k_glm_permute_counts<<<dim3((L.npad + 1023) / 1024, p), 256, 0, st>>>(h->Yt, h->ldn, p, h->perm_dev.ptr, L.npad,
                                                                           h->Yp.ptr, ldp, h->not_counts.ptr);
    CA_LAUNCH_CHECK();
    CA_CUDA(cudaStreamSynchronize(st));  // L.perm (host) is read by the copy above
    h->perm_host = L.perm;
    h->Yp_src = h->Yt;
  }
  // ---- dense design rows + offset in that order ----
  {
    std::vector<double> xd((size_t)L.npad * ldxd, 0.0);
    for (int c = 0; c < L.npad; ++c) {
      const int i = L.perm[c];
      if (i < 0) continue;
      double* r = xd.data() + (size_t)c * ldxd;
      for (int d = 0; d < kde; ++d) r[d] = L.dcol[d] >= 0 ? X[(size_t)i * kx + L.dcol[d]] : 1.0;
      r[kde] = offset ? offset[i] : 0.0;
    }
    if (h->Xd.alloc(xd.size()) || h->gpair.alloc(L.gpair.size())) return -1;
    CA_CUDA(cudaMemcpyAsync(h->Xd.ptr, xd.data(), sizeof(double) * xd.size(), cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaMemcpyAsync(h->gpair.ptr, L.gpair.data(), sizeof(int) * L.gpair.size(), cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaStreamSynchronize(st));
  }
  {
    ProfScope prof(PROF_PREP, st);
    k_glm_ybar<<<p, 256, 0, st>>>(h->Yt, h->ldn, n, h->ybar.ptr);
    CA_LAUNCH_CHECK();
  }
  int sm_count = 148;
  {
    int devid = 0;
    cudaGetDevice(&devid);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, devid);
  }
  const int n_oct = ((genelist_dev ? n_list : p) + 7) / 8;
  if (n_oct <= 0) return 0;
  const int grid = std::min(n_oct, 2 * sm_count);   // two CTAs per SM when the shared memory allows (the kernel is persistent either way)
  const int nb = (kde + 7) / 8;
  const size_t flush_d = (size_t)grid * L.nseg_tot * (nb + 1) * 64;
  const size_t gc_d = lasso ? (size_t)grid * IRLS_SOLVERS_L1 * kx * (kx + 1) : 0;
  const size_t gsum_d = (size_t)grid * (L.ngroups + 1) * (nb + 1) * 64;
  if (h->irls_scratch.alloc(flush_d + gc_d + gsum_d)) return -1;
  CA_CUDA(cudaMemsetAsync(h->counter.ptr, 0, sizeof(int) * 2, st));
  GlmIrls P;
  memset(&P, 0, sizeof P);
  P.Yp = h->Yp.ptr;
  P.ldp = ldp;
  P.p = p;
  P.Xd = h->Xd.ptr;
  P.ldxd = ldxd;
  P.nchunks = L.nchunks;
  memcpy(P.chunk_lo, L.chunk_lo, sizeof P.chunk_lo);
  P.gpair = h->gpair.ptr;
  P.nseg_tot = L.nseg_tot;
  memcpy(P.seg_ptr, L.seg_ptr, sizeof P.seg_ptr);
  memcpy(P.gseg_ptr, L.gseg_ptr, sizeof P.gseg_ptr);
  memcpy(P.gseg_list, L.gseg_list, sizeof P.gseg_list);
  P.ngroups = L.ngroups;
  P.kx = kx;
  P.kde = kde;
  memcpy(P.dcol, L.dcol, sizeof P.dcol);
  memcpy(P.gcol, L.gcol, sizeof P.gcol);
  memcpy(P.cdense, L.cdense, sizeof P.cdense);
  memcpy(P.cgroup, L.cgroup, sizeof P.cgroup);
  P.nu = has_nu ? h->nu.ptr : nullptr;
  P.family = family;
  P.thres = thres;
  P.ybar = h->ybar.ptr;
  P.ridge = has_ridge ? h->ridge.ptr : nullptr;
  P.l1 = lasso ? h->ridge.ptr + kx : nullptr;
  P.tol = tol;
  P.steptol = 1e-8;
  P.maxiter = maxiter;
  P.iter0 = warm ? 1 : 0;
  P.freeze_dense = freeze_dense ? 1 : 0;
  P.genelist = genelist_dev;
  P.n_list = n_list;
  P.counter = h->counter.ptr;
  P.not_counts = h->not_counts.ptr;
  P.scratch = h->irls_scratch.ptr;
  P.gc_scratch = h->irls_scratch.ptr + flush_d;
  P.gsum = h->irls_scratch.ptr + flush_d + gc_d;
  P.beta = h->beta.ptr;
  P.dev = h->dev.ptr;
  P.step = h->step.ptr;
  P.status = h->status.ptr;
  P.n_iter = h->n_iter.ptr;
  P.fm_tables = fastmath_tables();
  CA_REQUIRE(P.fm_tables, "fastmath tables unavailable");
  return launch_irls(P, grid, smem, st);
}

}  // namespace ca
