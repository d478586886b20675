// DMMA / DFMA latency and per-warp issue-rate probe for sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048;

template <int NCHAIN>
__global__ void k_dmma(double* out, double a, double b, long long* cyc) {
  double c[NCHAIN][2];
  double fa[NCHAIN], fb[NCHAIN];
  for (int i = 0; i < NCHAIN; ++i) { c[i][0] = c[i][1] = 0; fa[i] = a + threadIdx.x + i; fb[i] = b + threadIdx.x * 2 + i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa[i]), "d"(fb[i]));
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NCHAIN; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NCHAIN>
__global__ void k_dfma(double* out, double a, double b, long long* cyc) {
  double x[NCHAIN];
  for (int i = 0; i < NCHAIN; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NCHAIN; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
long long run(F f) {
  long long* d; cudaMalloc(&d, 8); f(d); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); cudaFree(d); return h;
}

int main() {
  double* out; CK(cudaMalloc(&out, 8 * 1024 * 1024));
  // one warp on one SM: dependent-chain latency and independent-issue rate
#define RUN_DMMA(N, W) { long long c = run([&](long long* d) { k_dmma<N><<<1, 32 * W>>>(out, 1.0000001, 1e-9, d); }); \
    printf("dmma chains/warp=%d warps=%d : %.1f cycles per DMMA per warp, %.2f cycles per DMMA per SM\n", N, W, double(c) / ITERS / N, double(c) / ITERS / N / W); }
  RUN_DMMA(1, 1) RUN_DMMA(2, 1) RUN_DMMA(4, 1) RUN_DMMA(8, 1) RUN_DMMA(16, 1)
  RUN_DMMA(1, 4) RUN_DMMA(4, 4) RUN_DMMA(8, 4) RUN_DMMA(1, 16) RUN_DMMA(4, 16) RUN_DMMA(8, 8) RUN_DMMA(2, 32)
#define RUN_DFMA(N, W) { long long c = run([&](long long* d) { k_dfma<N><<<1, 32 * W>>>(out, 1.0000001, 1e-9, d); }); \
    printf("dfma chains/warp=%d warps=%d : %.2f cycles per DFMA per warp, %.2f per SM\n", N, W, double(c) / ITERS / N, double(c) / ITERS / N / W); }
  RUN_DFMA(1, 1) RUN_DFMA(2, 1) RUN_DFMA(4, 1) RUN_DFMA(8, 1) RUN_DFMA(1, 4) RUN_DFMA(4, 4) RUN_DFMA(8, 4) RUN_DFMA(4, 16) RUN_DFMA(8, 16)
  return 0;
}
