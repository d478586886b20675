// Host side of fastmath.cuh: the exp / log tables, rounded from long double.
#include <math.h>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "fastmath.cuh"

namespace ca {

static double* g_tables[16] = {nullptr};
static std::mutex g_tab_mu;

const double* fastmath_tables() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  std::lock_guard<std::mutex> lock(g_tab_mu);
  if (g_tables[dev]) return g_tables[dev];
  std::vector<double> h(FM_TABLE_DOUBLES);
  for (int j = 0; j < FM_EXP_N; ++j) h[j] = (double)exp2l((long double)j / FM_EXP_N);
  for (int i = 0; i < FM_LOG_N; ++i) {
    const long double c = 1.0L + ((long double)i + 0.5L) / FM_LOG_N;
    const double inv = (double)(1.0L / c);
    h[FM_EXP_N + 2 * i] = inv;
    h[FM_EXP_N + 2 * i + 1] = (double)(-logl((long double)inv));  // log c for the ROUNDED 1/c
  }
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * FM_TABLE_DOUBLES) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), sizeof(double) * FM_TABLE_DOUBLES, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  g_tables[dev] = d;
  return d;
}

NbClip nb_clip_constants() {
  NbClip c;
  c.u_lo = 1.0 / (1.0 - 1e-6) - 1.0;
  c.u_hi = 1e6 - 1.0;
  c.lu_lo = log(c.u_lo);
  c.lu_hi = log(c.u_hi);
  return c;
}

}  // namespace ca

// ---- test hook: evaluate the device functions on an array -------------------
namespace ca {
__global__ void k_fm_test(const double* __restrict__ x, int n, int op, const double* __restrict__ g_tab,
                          double* __restrict__ out) {
  __shared__ double s_tab[FM_TABLE_DOUBLES];
  const FmTables T = fm_load_tables(s_tab, g_tab);
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = op == 0 ? fm_exp(x[i], T.exp_t) : fm_log(x[i], T.log_t);
}
}  // namespace ca

extern "C" int ca_test_fastmath(const double* x, int n, int op, double* out) {
  using namespace ca;
  CA_REQUIRE(x && out && n > 0 && (op == 0 || op == 1), "bad argument");
  const double* tab = fastmath_tables();
  CA_REQUIRE(tab, "fastmath tables unavailable");
  double *dx = nullptr, *dout = nullptr;
  CA_CUDA(cudaMalloc((void**)&dx, sizeof(double) * n));
  CA_CUDA(cudaMalloc((void**)&dout, sizeof(double) * n));
  CA_CUDA(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  k_fm_test<<<148, 256>>>(dx, n, op, tab, dout);
  CA_LAUNCH_CHECK();
  CA_CUDA(cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(dx);
  cudaFree(dout);
  return 0;
}
