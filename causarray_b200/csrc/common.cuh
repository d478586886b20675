// Shared device/host helpers for libcausarray_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/causarray_b200.h"

namespace ca {

// ---------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative code and
// leaves a message retrievable through ca_last_error().
// ---------------------------------------------------------------------------
extern long long g_launches;
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, int line);

#define CA_CUDA(x)                                                       \
  do {                                                                   \
    cudaError_t _e = (x);                                                \
    if (_e != cudaSuccess) return ca::cuda_fail(_e, #x, __LINE__);       \
  } while (0)

#define CA_LAUNCH_CHECK()            \
  do {                               \
    ++ca::g_launches;                \
    CA_CUDA(cudaGetLastError());     \
  } while (0)

#define CA_REQUIRE(cond, msg)                                            \
  do {                                                                   \
    if (!(cond)) {                                                       \
      ca::set_error("%s (%s:%d)", msg, __FILE__, __LINE__);              \
      return -2;                                                         \
    }                                                                    \
  } while (0)


__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------
// Per-entry likelihood terms.  Restates the elementwise arithmetic of the
// reference (causarray/gcate_likelihood.py:74-88 for nll, :126-138 for grad):
//   xi = min(theta, 100); m = exp(xi)
//   Poisson branch (family poisson, or NB with nu > thres_disp):
//        ell = y*xi - m                       d ell/d theta * (-1) = m - y
//   NB:  q = clip(1/(1 + m/nu), 1e-6, 1-1e-6)
//        ell = y*log1p(-q) + nu*log(q)        -(y - m) * q
// ---------------------------------------------------------------------------
template <bool WANT_DL>
__device__ __forceinline__ void entry_terms(double theta, double y, double nu, bool pois,
                                            double& ell, double& dl) {
  const double xi = fmin(theta, 100.0);
  const double m = exp(xi);
  if (pois) {
    ell = y * xi - m;
    if (WANT_DL) dl = m - y;
  } else {
    double q = 1.0 / (1.0 + m / nu);
    q = fmin(fmax(q, 1e-6), 1.0 - 1e-6);
    ell = y * log1p(-q) + nu * log(q);
    if (WANT_DL) dl = (m - y) * q;
  }
}

// m8n8k4 FP64 tensor-core MMA (DMMA.8x8x4): D = A(8x4,row) * B(4x8,col) + C.
// Fragment ownership (lane = 4*g + t): A[g][t], B[t][g], C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ---- TMA-style bulk async copies (cp.async.bulk, SASS UBLKCP) completed on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
#ifdef CA_WAIT_SLEEP
    __nanosleep(CA_WAIT_SLEEP);
#endif
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Deterministic block-wide sum (fixed tree): every thread passes a value, the
// result is returned to all threads.  `scratch` needs blockDim.x/32 doubles.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < nw; ++w) s += scratch[w];
  return s;
}

}  // namespace ca
