// Minimal owning device buffer (cudaMalloc / cudaFree), no exceptions.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "common.cuh"

namespace ca {

template <typename T>
struct DevBuf {
  T* ptr = nullptr;
  size_t count = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  // (re)allocate to hold n elements; returns 0 or -1 (error recorded)
  int alloc(size_t n) {
    if (n == count && ptr) return 0;
    release();
    if (n == 0) return 0;
    cudaError_t e = cudaMalloc((void**)&ptr, n * sizeof(T));
    if (e != cudaSuccess) {
      ptr = nullptr;
      return cuda_fail(e, "cudaMalloc", __LINE__);
    }
    count = n;
    return 0;
  }
  size_t bytes() const { return count * sizeof(T); }
};

}  // namespace ca
