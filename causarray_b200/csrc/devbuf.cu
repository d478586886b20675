#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
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

namespace ca {

namespace {
constexpr size_t STAGE_CHUNK = 32u << 20;
struct Staging {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2];
  bool used[2] = {false, false};
  bool ok = false;
  std::mutex mu;
};
Staging g_stage;
}  // namespace

int h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes < (16u << 20) || getenv("CA_NO_STAGED_H2D")) {
    CA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  std::lock_guard<std::mutex> lock(g_stage.mu);
  if (!g_stage.ok) {
    for (int b = 0; b < 2; ++b) {
      CA_CUDA(cudaHostAlloc(&g_stage.buf[b], STAGE_CHUNK, cudaHostAllocDefault));
      CA_CUDA(cudaEventCreateWithFlags(&g_stage.ev[b], cudaEventDisableTiming));
    }
    g_stage.ok = true;
  }
  const int nt = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / 2));
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  int i = 0;
  for (size_t off = 0; off < bytes; off += STAGE_CHUNK, ++i) {
    const int b = i & 1;
    const size_t len = std::min(STAGE_CHUNK, bytes - off);
    if (g_stage.used[b]) CA_CUDA(cudaEventSynchronize(g_stage.ev[b]));  // the previous upload from this buffer
    char* pin = static_cast<char*>(g_stage.buf[b]);
    const size_t piece = (len / nt + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (int w = 1; w < nt; ++w) {
      const size_t lo = (size_t)w * piece;
      if (lo >= len) break;
      const size_t n = std::min(piece, len - lo);
      th.emplace_back([=]() { memcpy(pin + lo, s + off + lo, n); });
    }
    memcpy(pin, s + off, std::min(piece, len));
    for (auto& t : th) t.join();
    CA_CUDA(cudaMemcpyAsync(d + off, pin, len, cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaEventRecord(g_stage.ev[b], st));
    g_stage.used[b] = true;
  }
  return 0;
}

// Same pipeline, but the source is `nrows` rows (row_bytes each) of a larger host matrix picked by
// `rows` (row i of the destination = row rows[i] of the source, leading dimension ld_bytes): the
// gather happens inside the threaded copy into the pinned buffer, so a perturbation batch goes from
// the full count matrix to the device without an intermediate host copy (numpy's fancy-index copy of
// one 4 000 x 8 000 batch cost 30 ms, as much as the upload itself).
int h2d_staged_rows(void* dst, const void* src, size_t ld_bytes, const int64_t* rows, size_t nrows, size_t row_bytes,
                    cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_stage.mu);
  if (!g_stage.ok) {
    for (int b = 0; b < 2; ++b) {
      CA_CUDA(cudaHostAlloc(&g_stage.buf[b], STAGE_CHUNK, cudaHostAllocDefault));
      CA_CUDA(cudaEventCreateWithFlags(&g_stage.ev[b], cudaEventDisableTiming));
    }
    g_stage.ok = true;
  }
  if (row_bytes == 0 || nrows == 0) return 0;
  if (row_bytes > STAGE_CHUNK) {
    set_error("row longer than the staging buffer");
    return -2;
  }
  const int nt = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / 2));
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  const size_t rows_per_chunk = STAGE_CHUNK / row_bytes;
  int i = 0;
  for (size_t r0 = 0; r0 < nrows; r0 += rows_per_chunk, ++i) {
    const int b = i & 1;
    const size_t nr = std::min(rows_per_chunk, nrows - r0);
    if (g_stage.used[b]) CA_CUDA(cudaEventSynchronize(g_stage.ev[b]));
    char* pin = static_cast<char*>(g_stage.buf[b]);
    auto work = [=](size_t lo, size_t hi) {
      for (size_t r = lo; r < hi; ++r) memcpy(pin + r * row_bytes, s + (size_t)rows[r0 + r] * ld_bytes, row_bytes);
    };
    const size_t piece = (nr + nt - 1) / nt;
    std::vector<std::thread> th;
    for (int w = 1; w < nt; ++w) {
      const size_t lo = (size_t)w * piece;
      if (lo >= nr) break;
      th.emplace_back(work, lo, std::min(nr, lo + piece));
    }
    work(0, std::min(nr, piece));
    for (auto& t : th) t.join();
    CA_CUDA(cudaMemcpyAsync(d + r0 * row_bytes, pin, nr * row_bytes, cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaEventRecord(g_stage.ev[b], st));
    g_stage.used[b] = true;
  }
  return 0;
}

}  // namespace ca
