// FP64 pipe micro-benchmark for B200 (sm_100a): DFMA, DMMA (mma.sync f64),
// exp/log/log1p/div throughput.  These are the roofline denominators for the
// FP64 kernels (MEASURED_PEAKS.json only holds HBM and bf16 tensor numbers).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// NOTE: the four chains must not share operands AND start values: the first version passed the same
// (fa, fb) to four zero-initialised accumulators, ptxas merged them into ONE chain (the SASS held
// ITERS DMMAs, not 4 ITERS) and the printed rate was 4x too high (146 instead of 36.5 TFLOP/s).
__global__ void k_dmma884(double* out, double a, double b) {
  double c0[2] = {0, 1}, c1[2] = {2, 3}, c2[2] = {4, 5}, c3[2] = {6, 7};
  double fa = a + threadIdx.x, fb = b + threadIdx.x;
  double fa1 = fa + 1, fa2 = fa + 2, fa3 = fa + 3;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa1), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa2), "d"(fb));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa3), "d"(fb));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

// m16n8k8: A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs)
__global__ void k_dmma1688(double* out, double a, double b) {
  double c0[4] = {0, 1, 2, 3}, c1[4] = {4, 5, 6, 7}, c2[4] = {8, 9, 10, 11}, c3[4] = {12, 13, 14, 15};
  double fa0 = a + threadIdx.x, fa1 = fa0 + 1, fa2 = fa0 + 2, fa3 = fa0 + 3, fb0 = b + threadIdx.x, fb1 = fb0 + 1;
  const double fq1 = fa0 + 1.5, fq2 = fa0 + 2.5, fq3 = fa0 + 3.5;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
#define MMA(c, q) asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(q), "d"(fa1), "d"(fa2), "d"(fa3), "d"(fb0), "d"(fb1));
    MMA(c0, fa0) MMA(c1, fq1) MMA(c2, fq2) MMA(c3, fq3)
#undef MMA
  }
  double s = 0;
  for (int j = 0; j < 4; ++j) s += c0[j] + c1[j] + c2[j] + c3[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4
__global__ void k_dmma1684(double* out, double a, double b) {
  double c0[4] = {0, 1, 2, 3}, c1[4] = {4, 5, 6, 7}, c2[4] = {8, 9, 10, 11}, c3[4] = {12, 13, 14, 15};
  double fa0 = a + threadIdx.x, fa1 = fa0 + 1, fb0 = b + threadIdx.x;
  const double fq1 = fa0 + 1.5, fq2 = fa0 + 2.5, fq3 = fa0 + 3.5;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
#define MMA(c, q) asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" \
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(q), "d"(fa1), "d"(fb0));
    MMA(c0, fa0) MMA(c1, fq1) MMA(c2, fq2) MMA(c3, fq3)
#undef MMA
  }
  double s = 0;
  for (int j = 0; j < 4; ++j) s += c0[j] + c1[j] + c2[j] + c3[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__global__ void k_func(double* out, double seed, double step) {
  double x0 = seed + threadIdx.x * 1e-3, x1 = x0 + 0.1, x2 = x0 + 0.2, x3 = x0 + 0.3;
  double acc = 0;
  for (int i = 0; i < ITERS / 8; ++i) {
    double y0, y1, y2, y3;
    if (OP == 0) { y0 = exp(x0); y1 = exp(x1); y2 = exp(x2); y3 = exp(x3); }
    if (OP == 1) { y0 = log(x0); y1 = log(x1); y2 = log(x2); y3 = log(x3); }
    if (OP == 2) { y0 = log1p(x0); y1 = log1p(x1); y2 = log1p(x2); y3 = log1p(x3); }
    if (OP == 3) { y0 = 1.0 / x0; y1 = 1.0 / x1; y2 = 1.0 / x2; y3 = 1.0 / x3; }
    if (OP == 4) { y0 = sqrt(x0); y1 = sqrt(x1); y2 = sqrt(x2); y3 = sqrt(x3); }
    if (OP == 5) { y0 = __drcp_rn(x0) ; y1 = __drcp_rn(x1); y2 = __drcp_rn(x2); y3 = __drcp_rn(x3); }
    acc += y0 + y1 + y2 + y3;
    x0 += step; x1 += step; x2 += step; x3 += step;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_it(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, prop.clockRate);
  const int threads = 256, blocks = sms * 8;
  double* out; CK(cudaMalloc(&out, sizeof(double) * threads * blocks));
  double nthreads = double(threads) * blocks;
  float ms;
  ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); }, 5);
  printf(" \"dfma_tflops\": %.3f,\n", nthreads * ITERS * 8 * 2 / (ms * 1e-3) / 1e12);
  ms = time_it([&] { k_dmma884<<<blocks, threads>>>(out, 1.0000001, 1e-9); }, 5);
  printf(" \"dmma_m8n8k4_tflops\": %.3f,\n", (nthreads / 32) * ITERS * 4 * (8 * 8 * 4 * 2.0) / (ms * 1e-3) / 1e12);
  ms = time_it([&] { k_dmma1684<<<blocks, threads>>>(out, 1.0000001, 1e-9); }, 5);
  printf(" \"dmma_m16n8k4_tflops\": %.3f,\n", (nthreads / 32) * ITERS * 4 * (16 * 8 * 4 * 2.0) / (ms * 1e-3) / 1e12);
  ms = time_it([&] { k_dmma1688<<<blocks, threads>>>(out, 1.0000001, 1e-9); }, 5);
  printf(" \"dmma_m16n8k8_tflops\": %.3f,\n", (nthreads / 32) * ITERS * 4 * (16 * 8 * 8 * 2.0) / (ms * 1e-3) / 1e12);
  const char* names[6] = {"exp", "log", "log1p", "div", "sqrt", "rcp"};
  double nops = nthreads * (ITERS / 8) * 4;
  ms = time_it([&] { k_func<0><<<blocks, threads>>>(out, -3.0, 1e-3); }, 5);
  printf(" \"%s_gops\": %.2f,\n", names[0], nops / (ms * 1e-3) / 1e9);
  ms = time_it([&] { k_func<1><<<blocks, threads>>>(out, 0.5, 1e-3); }, 5);
  printf(" \"%s_gops\": %.2f,\n", names[1], nops / (ms * 1e-3) / 1e9);
  ms = time_it([&] { k_func<2><<<blocks, threads>>>(out, 0.01, 1e-3); }, 5);
  printf(" \"%s_gops\": %.2f,\n", names[2], nops / (ms * 1e-3) / 1e9);
  ms = time_it([&] { k_func<3><<<blocks, threads>>>(out, 0.5, 1e-3); }, 5);
  printf(" \"%s_gops\": %.2f,\n", names[3], nops / (ms * 1e-3) / 1e9);
  ms = time_it([&] { k_func<4><<<blocks, threads>>>(out, 0.5, 1e-3); }, 5);
  printf(" \"%s_gops\": %.2f,\n", names[4], nops / (ms * 1e-3) / 1e9);
  ms = time_it([&] { k_func<5><<<blocks, threads>>>(out, 0.5, 1e-3); }, 5);
  printf(" \"%s_gops\": %.2f,\n", names[5], nops / (ms * 1e-3) / 1e9);
  // HBM copy
  size_t nbytes = size_t(2) << 30;
  double *a, *b; CK(cudaMalloc(&a, nbytes)); CK(cudaMalloc(&b, nbytes));
  CK(cudaMemset(a, 1, nbytes));
  ms = time_it([&] { CK(cudaMemcpyAsync(b, a, nbytes, cudaMemcpyDeviceToDevice)); }, 5);
  printf(" \"hbm_copy_gbs\": %.1f}\n", 2.0 * nbytes / (ms * 1e-3) / 1e9);
  return 0;
}
