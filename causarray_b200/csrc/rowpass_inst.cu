// One object file per KS (compiled with -DCA_KS=<n>): explicit instantiation of the row-pass
// and line-search kernels for kpad = 4 * KS.
#include "rowpass_impl.cuh"
#include "prof.h"

#ifndef CA_KS
#error "compile with -DCA_KS=<kpad/4>"
#endif

namespace ca {

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  CA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int KS>
int launch_rowpass_ks(const PassCfg& P, const double* X, const int* rowlist, const int* count, int nrows,
                      bool want_grad, int gc_lo, int gc_n, double* ell_out, double* G_out, int ldG, cudaStream_t st) {
  constexpr int KPAD = 4 * KS;
  const int grid = (nrows + TILE_ROWS - 1) / TILE_ROWS;
  ProfScope prof(want_grad ? PROF_ROWPASS_GRAD : PROF_ROWPASS_ELL, st);
  if (!want_grad) {
    const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 1, 1);
    if (set_smem(k_rowpass<KS, false, 1>, sm)) return -1;
    k_rowpass<KS, false, 1><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, 0, ell_out, nullptr, 0);
  } else if (P.theta_out && P.nu_per_row && gc_n <= 16) {   // gene pass that also leaves theta of the treated cells
    const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 2, 1);
    if (set_smem(k_rowpass<KS, true, 2, true>, sm)) return -1;
    k_rowpass<KS, true, 2, true><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
  } else if (P.theta_out && P.nu_per_row && gc_n <= 32) {
    const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 4, 1);
    if (set_smem(k_rowpass<KS, true, 4, true>, sm)) return -1;
    k_rowpass<KS, true, 4, true><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
  } else if (gc_n <= 16) {
    const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 2, 1);
    if (set_smem(k_rowpass<KS, true, 2>, sm)) return -1;
    k_rowpass<KS, true, 2><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
  } else if (gc_n <= 32) {
    const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 4, 1);
    if (set_smem(k_rowpass<KS, true, 4>, sm)) return -1;
    k_rowpass<KS, true, 4><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
  } else {
    const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 8, 1);
    if (set_smem(k_rowpass<KS, true, 8>, sm)) return -1;
    k_rowpass<KS, true, 8><<<grid, NTHREADS, sm, st>>>(P, X, rowlist, count, nrows, gc_lo, ell_out, G_out, ldG);
  }
  CA_LAUNCH_CHECK();
  return 0;
}

template <int KS>
int launch_search_ks(const PassCfg& P, const SearchArgs& A, cudaStream_t st) {
  constexpr int KPAD = 4 * KS;
  const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 1, 3);
  if (set_smem(k_search<KS>, sm)) return -1;
  const int grid = (A.nrows + TILE_ROWS - 1) / TILE_ROWS;
  ProfScope prof(PROF_SEARCH, st);
  k_search<KS><<<grid, NTHREADS, sm, st>>>(P, A);
  CA_LAUNCH_CHECK();
  return 0;
}

template <int KS>
int launch_search_multi_ks(const PassCfg& P, const SearchArgs& A, cudaStream_t st) {
  constexpr int KPAD = 4 * KS;
  const size_t sm = sizeof(double) * smem_doubles(P.tc, KPAD, 1, SEARCH_NC + 3);
  if (set_smem(k_search_multi<KS, SEARCH_NC>, sm)) return -1;
  const int grid = (A.nrows + TILE_ROWS - 1) / TILE_ROWS;
  ProfScope prof(PROF_SEARCH, st);
  k_search_multi<KS, SEARCH_NC><<<grid, NTHREADS, sm, st>>>(P, A);
  CA_LAUNCH_CHECK();
  return 0;
}

template int launch_rowpass_ks<CA_KS>(const PassCfg&, const double*, const int*, const int*, int, bool, int, int,
                                      double*, double*, int, cudaStream_t);
template int launch_search_ks<CA_KS>(const PassCfg&, const SearchArgs&, cudaStream_t);
template int launch_search_multi_ks<CA_KS>(const PassCfg&, const SearchArgs&, cudaStream_t);

}  // namespace ca
