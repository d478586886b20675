// One object file per number of dense design columns KDE (compiled with -DCA_KDE=<n>): explicit
// instantiation of the persistent per-octet IRLS kernel, Poisson and NB.
#include "glm_irls_impl.cuh"
#include "prof.h"

#ifndef CA_KDE
#error "compile with -DCA_KDE=<dense columns>"
#endif

namespace ca {

template <int KDE>
int launch_glm_irls_kde(const GlmIrls& P, int grid, size_t smem, cudaStream_t st) {
  const bool nb = P.family == CA_FAMILY_NB && P.nu != nullptr;
  ProfScope prof(PROF_GLM_PASS, st);
  constexpr int NTHR = IRLS_WARPS * 32 * IRLS_NH;
  if (nb) {
    CA_CUDA(cudaFuncSetAttribute(k_glm_irls<KDE, true, IRLS_NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_glm_irls<KDE, true, IRLS_NH><<<grid, NTHR, smem, st>>>(P);
  } else {
    CA_CUDA(cudaFuncSetAttribute(k_glm_irls<KDE, false, IRLS_NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_glm_irls<KDE, false, IRLS_NH><<<grid, NTHR, smem, st>>>(P);
  }
  CA_LAUNCH_CHECK();
  return 0;
}

template int launch_glm_irls_kde<CA_KDE>(const GlmIrls&, int, size_t, cudaStream_t);

}  // namespace ca
