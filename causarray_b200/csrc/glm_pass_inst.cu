// One object file per NT = ceil(kx / 8) (compiled with -DCA_NT=<n>): explicit instantiation of
// the IRLS pass kernel variants (first iteration / regular / finalize) x (Poisson / NB).
#include "glm_pass_impl.cuh"
#include "prof.h"

#ifndef CA_NT
#error "compile with -DCA_NT=<ceil(kx/8)>"
#endif

namespace ca {

template <int NT, bool FIRST, bool FIN, bool NB>
static int launch_variant(const GlmPass& P, dim3 grid, size_t sm, cudaStream_t st) {
  CA_CUDA(cudaFuncSetAttribute(k_glm_pass<NT, FIRST, FIN, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_glm_pass<NT, FIRST, FIN, NB><<<grid, GLM_THREADS, sm, st>>>(P);
  CA_LAUNCH_CHECK();
  return 0;
}

template <int NT>
int launch_glm_pass_nt(const GlmPass& P, cudaStream_t st) {
  const int nslots = P.genelist ? P.n_list : P.p;
  if (nslots <= 0) return 0;
  const dim3 grid((nslots + GLM_WARPS - 1) / GLM_WARPS, P.nsplit);
  const size_t sm = glm_pass_smem_bytes(NT);
  const bool nb = P.family == CA_FAMILY_NB && P.nu != nullptr;
  const bool first = P.iter == 1 && !P.finalize;
  ProfScope prof(PROF_GLM_PASS, st);
  if (P.finalize) return nb ? launch_variant<NT, false, true, true>(P, grid, sm, st) : launch_variant<NT, false, true, false>(P, grid, sm, st);
  if (first) return nb ? launch_variant<NT, true, false, true>(P, grid, sm, st) : launch_variant<NT, true, false, false>(P, grid, sm, st);
  return nb ? launch_variant<NT, false, false, true>(P, grid, sm, st) : launch_variant<NT, false, false, false>(P, grid, sm, st);
}

template int launch_glm_pass_nt<CA_NT>(const GlmPass&, cudaStream_t);

}  // namespace ca
