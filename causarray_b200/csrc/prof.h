// Optional per-kernel-class device timing (CUDA events on the launching stream)
// used by bench.py for the roofline numbers, and a global launch counter.
#pragma once
#include <cuda_runtime.h>

namespace ca {

enum ProfClass {
  PROF_ROWPASS_GRAD = 0,
  PROF_ROWPASS_ELL,
  PROF_SEARCH,
  PROF_GLM_PASS,
  PROF_GLM_SOLVE,
  PROF_AIPW,
  PROF_PREP,   // upload-side kernels: transpose / normalise / constant sums / size factors
  PROF_SEARCH_ONEHOT,   // stage-2 gene line search over the treated cells only (search_onehot.cu)
  PROF_NCLASS
};

extern long long g_launches;
void prof_begin(int cls, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);

struct ProfScope {
  int cls;
  cudaStream_t st;
  ProfScope(int c, cudaStream_t s) : cls(c), st(s) { prof_begin(c, s); }
  ~ProfScope() { prof_end(cls, st); }
};

}  // namespace ca
