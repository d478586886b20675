// One object file per NT = ceil(kx / 8) (compiled with -DCA_NT=<n>): explicit instantiation of
// the IRLS pass kernel variants (first iteration / regular / finalize) x (Poisson / NB).
#include "glm_pass_impl.cuh"
#include "prof.h"

#ifndef CA_NT
#error "compile with -DCA_NT=<ceil(kx/8)>"
#endif

namespace ca {

template <int NT, int T0, bool FIRST, bool FIN, bool NB>
static int launch_variant(const GlmPass& P, dim3 grid, size_t sm, cudaStream_t st) {
  CA_CUDA(cudaFuncSetAttribute(k_glm_pass<NT, T0, FIRST, FIN, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_glm_pass<NT, T0, FIRST, FIN, NB><<<grid, GLM_THREADS, sm, st>>>(P);
  CA_LAUNCH_CHECK();
  return 0;
}

template <int NT, int T0>
static int launch_t0(const GlmPass& P, dim3 grid, size_t sm, cudaStream_t st) {
  const bool nb = P.family == CA_FAMILY_NB && P.nu != nullptr;
  const bool first = P.iter == 1 && !P.finalize;
  if (P.finalize) return nb ? launch_variant<NT, T0, false, true, true>(P, grid, sm, st) : launch_variant<NT, T0, false, true, false>(P, grid, sm, st);
  if (first) return nb ? launch_variant<NT, T0, true, false, true>(P, grid, sm, st) : launch_variant<NT, T0, true, false, false>(P, grid, sm, st);
  return nb ? launch_variant<NT, T0, false, false, true>(P, grid, sm, st) : launch_variant<NT, T0, false, false, false>(P, grid, sm, st);
}

// Orthogonal-block variants are built for a dense part of one or two tiles (glm_pass_t0_supported).
template <int NT>
int launch_glm_pass_nt(const GlmPass& P, cudaStream_t st) {
  const int nslots = P.genelist ? P.n_list : P.p;
  if (nslots <= 0) return 0;
  const dim3 grid((nslots + GLM_WARPS - 1) / GLM_WARPS, P.nsplit);
  const size_t sm = glm_pass_smem_bytes(NT);
  ProfScope prof(PROF_GLM_PASS, st);
  if constexpr (NT >= 3) {
    if (P.t0 == 1) return launch_t0<NT, 1>(P, grid, sm, st);
  }
  if constexpr (NT >= 4) {
    if (P.t0 == 2) return launch_t0<NT, 2>(P, grid, sm, st);
  }
  CA_REQUIRE(P.t0 >= NT, "unsupported orthogonal-block layout");
  return launch_t0<NT, NT>(P, grid, sm, st);
}

template int launch_glm_pass_nt<CA_NT>(const GlmPass&, cudaStream_t);

}  // namespace ca
