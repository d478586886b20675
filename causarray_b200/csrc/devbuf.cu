#include <cstdlib>
#include <map>
#include <mutex>
#include "../../include/causarray_b200.h"
#include "devbuf.h"

namespace ca {

struct Pool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_list;  // (device, bytes) -> block
  size_t cached = 0;
  size_t cap = 0;
};
static Pool& pool() {
  static Pool p;
  if (p.cap == 0) {
    const char* e = getenv("CA_POOL_GB");
    const double gb = e ? atof(e) : 24.0;
    p.cap = (size_t)(gb * (double)(1ull << 30)) + 1;
  }
  return p;
}

void* pool_alloc(size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  Pool& P = pool();
  {
    std::lock_guard<std::mutex> lock(P.mu);
    auto it = P.free_list.find({dev, bytes});
    if (it != P.free_list.end()) {
      void* p = it->second;
      P.free_list.erase(it);
      P.cached -= bytes;
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    // out of memory with blocks cached: give them back and retry once
    pool_release_all();
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cuda_fail(e, "cudaMalloc", __LINE__);
    return nullptr;
  }
  return p;
}

void pool_free(void* p, size_t bytes) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  Pool& P = pool();
  {
    std::lock_guard<std::mutex> lock(P.mu);
    if (P.cached + bytes <= P.cap) {
      P.free_list.insert({{dev, bytes}, p});
      P.cached += bytes;
      return;
    }
  }
  cudaFree(p);
}

void pool_release_all() {
  Pool& P = pool();
  std::lock_guard<std::mutex> lock(P.mu);
  for (auto& kv : P.free_list) cudaFree(kv.second);
  P.free_list.clear();
  P.cached = 0;
}

}  // namespace ca

extern "C" int ca_release_cached(void) {
  ca::pool_release_all();
  return 0;
}
