// Owning device buffer on top of a small caching allocator, no exceptions.
//
// cudaFree of the ~1.6 GB a perturbation batch needs costs ~230 ms (measured through the
// Python profiler on the B200 box), more than the whole device-side LFC stage, and
// gcate_lfc_batch creates and destroys one workspace per batch.  Freed blocks therefore go to a
// size-keyed free list (batches have identical shapes, so reuse is exact); the cache is bounded
// (CA_POOL_GB, default 24 GB) and ca_release_cached() returns everything to the driver.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "common.cuh"

namespace ca {

void* pool_alloc(size_t bytes);          // nullptr on failure (error recorded)
void pool_free(void* p, size_t bytes);
void pool_release_all();

// Host -> device copy of a large PAGEABLE array: threaded memcpy into a pinned double buffer with the
// chunk uploads pipelined behind it (a plain cudaMemcpy of pageable memory ran at ~12 GB/s: 21 ms
// for the 256 MB of one batch).  Asynchronous on `st` like cudaMemcpyAsync; returns 0 or -1.
int h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t st);
// rows[i] selects the source row (leading dimension ld_bytes) of destination row i; see devbuf.cu
int h2d_staged_rows(void* dst, const void* src, size_t ld_bytes, const int64_t* rows, size_t nrows, size_t row_bytes,
                    cudaStream_t st);

template <typename T>
struct DevBuf {
  T* ptr = nullptr;
  size_t count = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (ptr) pool_free(ptr, count * sizeof(T));
    ptr = nullptr;
    count = 0;
  }
  // (re)allocate to hold n elements; returns 0 or -1 (error recorded)
  int alloc(size_t n) {
    if (n == count && ptr) return 0;
    release();
    if (n == 0) return 0;
    ptr = static_cast<T*>(pool_alloc(n * sizeof(T)));
    if (!ptr) return -1;
    count = n;
    return 0;
  }
  size_t bytes() const { return count * sizeof(T); }
};

}  // namespace ca
