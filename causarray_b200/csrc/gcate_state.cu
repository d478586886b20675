// S1 seam: device-resident GCATE problem + the host-side driver of one
// alternating iteration (update_with_mask, causarray/gcate_opt.py:148-206) and
// of the optimiser loop (alter_min, causarray/gcate_opt.py:337-458).
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "../../include/causarray_b200.h"
#include "common.cuh"
#include "devbuf.h"
#include "comm.h"
#include "kernels.h"
#include "gcate_state.h"

namespace ca {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}
int cuda_fail(cudaError_t e, const char* what, int line) {
  set_error("CUDA error %d (%s) in %s at line %d", (int)e, cudaGetErrorString(e), what, line);
  return -1;
}
const char* last_error() { return g_err.c_str(); }

int kpad_for(int k) {
  int kp = round_up(k, 4);
  while (kp % 8 != 4) kp += 4;
  return kp;
}

}  // namespace ca

using namespace ca;

// gene sharding: sums over genes are completed over the ranks (no-op for a single rank)
static inline bool sharded(const ca_gcate* h) { return h->comm != nullptr && h->world > 1; }
static inline int allreduce(ca_gcate* h, double* buf, size_t count) {
  if (!sharded(h)) return 0;
  if (h->peer && count <= PEER_MAX_DOUBLES) return peer_allreduce_sum((PeerArena*)h->peer, buf, count, h->st);
  return comm_allreduce_sum(h->comm, buf, count, h->st);
}
static inline int allreduce2(ca_gcate* h, double* b0, size_t c0, double* b1, size_t c1) {
  if (!sharded(h)) return 0;
  if (h->peer && c0 + c1 <= PEER_MAX_DOUBLES) return peer_allreduce_sum2((PeerArena*)h->peer, b0, c0, b1, c1, h->st);
  return allreduce(h, b0, c0) || allreduce(h, b1, c1);
}
// a peer that never reached a reduction makes the kernel give up after a few seconds instead of hanging
static inline int peer_check(ca_gcate* h) {
  if (h->peer && peer_error((PeerArena*)h->peer, h->st)) {
    set_error("a rank did not arrive at a peer-memory reduction (timed out); results are invalid");
    return -1;
  }
  return 0;
}
static inline double genes_total(const ca_gcate* h) { return h->p_total > 0 ? (double)h->p_total : (double)h->p; }


static int gcate_alloc(ca_gcate* h) {
  const int n = h->n, p = h->p, kp = h->kpad;
  if (h->Yn.alloc((size_t)n * h->ldp) || h->Ynt.alloc((size_t)p * h->ldn)) return -1;
  if (h->nu.alloc(p) || h->inv_nu.alloc(p) || h->log_nu.alloc(p) || h->colparm.alloc((size_t)4 * p) || h->tys_row.alloc(n) || h->tys_col.alloc(p) || h->tys_total.alloc(1) || h->sparsity.alloc(p))
    return -1;
  if (h->A.alloc((size_t)n * kp) || h->B.alloc((size_t)p * kp) || h->Aprev.alloc((size_t)n * kp) ||
      h->Bprev.alloc((size_t)p * kp))
    return -1;
  if (h->gene_active.alloc(p) || h->cell_active.alloc(n) || h->list_c.alloc(n) || h->list_g.alloc(p) ||
      h->list_clip.alloc(p) || h->counts.alloc(8))
    return -1;
  if (h->ell_c.alloc(n) || h->ell_g.alloc(p) || h->ellfin_c.alloc(n) || h->ellfin_g.alloc(p) ||
      h->Graw_c.alloc((size_t)n * h->ldG) || h->Graw_g.alloc((size_t)p * h->ldG) || h->g_c.alloc((size_t)n * kp) ||
      h->g_g.alloc((size_t)p * kp) || h->f0_c.alloc(n) || h->f0_g.alloc(p) || h->ng_c.alloc(n) || h->ng_g.alloc(p) ||
      h->W.alloc(64 * 64) || h->scal.alloc(16) || h->evals.alloc(8) || h->neval_g.alloc(p) || h->alpha_gene.alloc(p) || h->ls_pend.alloc(std::max(n, p)) ||
      h->ls_f01.alloc(std::max(n, p)) || h->ls_e01.alloc(std::max(n, p)) || h->ls_list2.alloc(std::max(n, p)))
    return -1;
  CA_CUDA(cudaMemsetAsync(h->Yn.ptr, 0, h->Yn.bytes(), h->st));
  CA_CUDA(cudaMemsetAsync(h->Ynt.ptr, 0, h->Ynt.bytes(), h->st));
  CA_CUDA(cudaMemsetAsync(h->A.ptr, 0, h->A.bytes(), h->st));
  CA_CUDA(cudaMemsetAsync(h->B.ptr, 0, h->B.bytes(), h->st));
  CA_CUDA(cudaMemsetAsync(h->evals.ptr, 0, h->evals.bytes(), h->st));
  CA_CUDA(cudaMemsetAsync(h->neval_g.ptr, 0, h->neval_g.bytes(), h->st));
  if (launch_fill_u8(h->gene_active.ptr, p, 1, h->st) || launch_fill_u8(h->cell_active.ptr, n, 1, h->st)) return -1;
  return 0;
}

extern "C" int ca_version(void) { return 100; }
extern "C" const char* ca_last_error(void) { return ca::last_error(); }

extern "C" int ca_set_device(int device) {
  CA_CUDA(cudaSetDevice(device));
  return 0;
}

extern "C" int ca_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem) {
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CA_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = (uint64_t)prop.totalGlobalMem;
  return 0;
}

extern "C" int ca_gcate_create(ca_gcate_t* out, int n, int p, int d, int r, int family, double thres_disp) {
  CA_REQUIRE(out && n > 0 && p > 0 && d >= 0 && r >= 0 && d + r > 0, "bad dimensions");
  CA_REQUIRE(family == CA_FAMILY_POISSON || family == CA_FAMILY_NB, "Family not recognized");
  CA_REQUIRE(d + r <= 64, "d + r > 64 is not supported");
  ca_gcate* h = new (std::nothrow) ca_gcate();
  CA_REQUIRE(h, "out of host memory");
  h->n = n;
  h->p = p;
  h->d = d;
  h->r = r;
  h->k = d + r;
  h->kpad = kpad_for(h->k);
  h->family = family;
  h->thres = thres_disp;
  h->ldp = round_up(p, 8);
  h->ldn = round_up(n, 8);
  h->st = 0;
  if (gcate_alloc(h)) {
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

extern "C" int ca_gcate_destroy(ca_gcate_t h) {
  if (h) {
    cudaStreamSynchronize(h->st);
    if (h->peer) peer_destroy(h->comm, (PeerArena*)h->peer, h->st);
    if (h->comm) comm_destroy(h->comm);
    delete h;
  }
  return 0;
}

extern "C" int ca_comm_unique_id(uint8_t* out128) {
  CA_REQUIRE(out128, "null argument");
  return comm_unique_id(out128);
}

extern "C" int ca_gcate_set_shard(ca_gcate_t h, int rank, int world, int64_t p_total, const uint8_t* id128) {
  CA_REQUIRE(h && id128 && world >= 1 && rank >= 0 && rank < world && p_total >= h->p, "bad shard description");
  CA_REQUIRE(!h->counts_loaded, "set the shard before the counts are loaded (the constant row sums are reduced there)");
  if (h->peer) {
    peer_destroy(h->comm, (PeerArena*)h->peer, h->st);
    h->peer = nullptr;
  }
  if (h->comm) {
    comm_destroy(h->comm);
    h->comm = nullptr;
  }
  h->rank = rank;
  h->world = world;
  h->p_total = (long)p_total;
  if (world > 1) {
    if (comm_init(&h->comm, rank, world, id128)) return -1;
    PeerArena* pa = nullptr;
    if (peer_setup(h->comm, rank, world, h->st, &pa)) return -1;
    h->peer = pa;
    const size_t n = (size_t)h->n;
    if (h->ls_Xc.alloc(n * h->kpad) || h->ls_t.alloc(n) || h->ls_ell.alloc(n) || h->ls_ne.alloc(n)) return -1;
  }
  return 0;
}

// 0: one rank (no exchange), 1: NCCL all-reduce, 2: one-shot all-reduce over NVLink peer memory (comm.cu)
extern "C" int ca_gcate_transport(ca_gcate_t h) {
  if (!h || h->world <= 1) return 0;
  return h->peer ? 2 : 1;
}

static int finish_load(ca_gcate* h, const double* nu_host, const double* Ys_dev_or_null) {
  const int n = h->n, p = h->p;
  // new counts: the gathered treated-cell counts of the one-hot stage-2 search belong to the previous ones
  h->s2_cells_host.clear();
  h->s2_checked = false;
  h->s2_ok = false;
  std::vector<double> ones;
  if (!nu_host) {
    ones.assign(p, 1.0);
    nu_host = ones.data();
  }
  std::vector<double> hin(p), hln(p), hcp((size_t)4 * p);
  h->any_pois = 0;
  for (int j = 0; j < p; ++j) {
    hin[j] = 1.0 / nu_host[j];
    hln[j] = std::log(nu_host[j]);
    hcp[4 * j + 0] = nu_host[j];
    hcp[4 * j + 1] = hin[j];
    hcp[4 * j + 2] = hln[j];
    hcp[4 * j + 3] = (nu_host[j] > h->thres) ? 1.0 : 0.0;
    if (nu_host[j] > h->thres) h->any_pois = 1;
  }
  CA_CUDA(cudaMemcpyAsync(h->colparm.ptr, hcp.data(), sizeof(double) * 4 * p, cudaMemcpyHostToDevice, h->st));
  CA_CUDA(cudaMemcpyAsync(h->nu.ptr, nu_host, sizeof(double) * p, cudaMemcpyHostToDevice, h->st));
  CA_CUDA(cudaMemcpyAsync(h->inv_nu.ptr, hin.data(), sizeof(double) * p, cudaMemcpyHostToDevice, h->st));
  CA_CUDA(cudaMemcpyAsync(h->log_nu.ptr, hln.data(), sizeof(double) * p, cudaMemcpyHostToDevice, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  if (Ys_dev_or_null) {
    // row sums directly; column sums through a transposed copy staged in Graw buffers is too small,
    // so reuse Ynt's shape with a temporary
    DevBuf<double> tmp;
    if (tmp.alloc((size_t)p * h->ldn)) return -1;
    CA_CUDA(cudaMemsetAsync(tmp.ptr, 0, tmp.bytes(), h->st));
    if (launch_rowsum(Ys_dev_or_null, p, n, p, h->tys_row.ptr, h->st)) return -1;
    if (launch_transpose(Ys_dev_or_null, n, p, p, tmp.ptr, h->ldn, h->st)) return -1;
    if (launch_rowsum(tmp.ptr, h->ldn, p, n, h->tys_col.ptr, h->st)) return -1;
    CA_CUDA(cudaStreamSynchronize(h->st));
  } else {
    if (launch_tys_rowsum(h->Yn.ptr, h->ldp, n, p, h->nu.ptr, 0, h->family, h->thres, h->tys_row.ptr, h->st)) return -1;
    if (launch_tys_rowsum(h->Ynt.ptr, h->ldn, p, n, h->nu.ptr, 1, h->family, h->thres, h->tys_col.ptr, h->st)) return -1;
  }
  if (launch_sum(h->tys_col.ptr, p, h->tys_total.ptr, h->st)) return -1;
  if (allreduce(h, h->tys_row.ptr, (size_t)n) || allreduce(h, h->tys_total.ptr, 1)) return -1;
  if (launch_zero_fraction(h->Ynt.ptr, h->ldn, p, n, h->sparsity.ptr, h->st)) return -1;
  CA_CUDA(cudaStreamSynchronize(h->st));
  h->counts_loaded = true;
  return 0;
}

// common tail of the two upload entry points: gene-major copy, input checks
static int finish_upload(ca_gcate* h, int64_t* n_nonint, int64_t* n_nonfinite, int64_t* n_empty_genes) {
  const int n = h->n, p = h->p;
  CA_CUDA(cudaMemsetAsync(h->Yraw_t.ptr, 0, h->Yraw_t.bytes(), h->st));
  if (launch_transpose(h->Yraw.ptr, n, p, p, h->Yraw_t.ptr, h->ldn, h->st)) return -1;
  h->raw_loaded = true;
  h->counts_loaded = false;
  if (n_nonint || n_nonfinite || n_empty_genes) {
    DevBuf<unsigned long long> flags;
    if (flags.alloc(2)) return -1;
    if (launch_check_counts(h->Yraw.ptr, (long)n * p, flags.ptr, h->st)) return -1;
    if (launch_zero_fraction(h->Yraw_t.ptr, h->ldn, p, n, h->sparsity.ptr, h->st)) return -1;
    unsigned long long hf[2];
    std::vector<double> sp(p);
    CA_CUDA(cudaMemcpyAsync(hf, flags.ptr, sizeof hf, cudaMemcpyDeviceToHost, h->st));
    CA_CUDA(cudaMemcpyAsync(sp.data(), h->sparsity.ptr, sizeof(double) * p, cudaMemcpyDeviceToHost, h->st));
    CA_CUDA(cudaStreamSynchronize(h->st));
    int64_t empty = 0;
    for (int j = 0; j < p; ++j) empty += (sp[j] >= 1.0);
    if (n_nonint) *n_nonint = (int64_t)hf[0];
    if (n_nonfinite) *n_nonfinite = (int64_t)hf[1];
    if (n_empty_genes) *n_empty_genes = empty;
  }
  return 0;
}

extern "C" int ca_gcate_upload_counts(ca_gcate_t h, const double* Y, int64_t* n_nonint, int64_t* n_nonfinite,
                                      int64_t* n_empty_genes) {
  CA_REQUIRE(h && Y, "null argument");
  const int n = h->n, p = h->p;
  if (h->Yraw.alloc((size_t)n * p) || h->Yraw_t.alloc((size_t)p * h->ldn)) return -1;
  if (h2d_staged(h->Yraw.ptr, Y, sizeof(double) * (size_t)n * p, h->st)) return -1;
  return finish_upload(h, n_nonint, n_nonfinite, n_empty_genes);
}

extern "C" int ca_gcate_upload_counts_typed(ca_gcate_t h, const void* Y, int dtype, int64_t* n_nonint,
                                            int64_t* n_nonfinite, int64_t* n_empty_genes) {
  CA_REQUIRE(h && Y, "null argument");
  if (dtype == CA_F64) return ca_gcate_upload_counts(h, (const double*)Y, n_nonint, n_nonfinite, n_empty_genes);
  static const size_t width[6] = {8, 4, 4, 8, 2, 1};
  CA_REQUIRE(dtype > 0 && dtype <= CA_U8, "unknown element type");
  const int n = h->n, p = h->p;
  const size_t count = (size_t)n * p;
  if (h->Yraw.alloc(count) || h->Yraw_t.alloc((size_t)p * h->ldn)) return -1;
  DevBuf<unsigned char> narrow;
  if (narrow.alloc(count * width[dtype])) return -1;
  if (h2d_staged(narrow.ptr, Y, count * width[dtype], h->st)) return -1;
  if (launch_widen_counts(narrow.ptr, dtype, (long)count, h->Yraw.ptr, h->st)) return -1;
  if (int rc = finish_upload(h, n_nonint, n_nonfinite, n_empty_genes)) return rc;
  CA_CUDA(cudaStreamSynchronize(h->st));  // the narrow staging copy goes out of scope
  return 0;
}

extern "C" int ca_gcate_upload_counts_rows(ca_gcate_t h, const void* Y, int64_t ld, const int64_t* rows, int dtype,
                                           int64_t* n_nonint, int64_t* n_nonfinite, int64_t* n_empty_genes) {
  CA_REQUIRE(h && Y && rows, "null argument");
  static const size_t width[6] = {8, 4, 4, 8, 2, 1};
  CA_REQUIRE(dtype >= 0 && dtype <= CA_U8, "unknown element type");
  const int n = h->n, p = h->p;
  CA_REQUIRE(ld >= p, "leading dimension smaller than the number of genes");
  const size_t count = (size_t)n * p, wd = width[dtype];
  if (h->Yraw.alloc(count) || h->Yraw_t.alloc((size_t)p * h->ldn)) return -1;
  if (dtype == CA_F64) {
    if (h2d_staged_rows(h->Yraw.ptr, Y, (size_t)ld * wd, rows, (size_t)n, (size_t)p * wd, h->st)) return -1;
    if (int rc = finish_upload(h, n_nonint, n_nonfinite, n_empty_genes)) return rc;
    CA_CUDA(cudaStreamSynchronize(h->st));
    return 0;
  }
  DevBuf<unsigned char> narrow;
  if (narrow.alloc(count * wd)) return -1;
  if (h2d_staged_rows(narrow.ptr, Y, (size_t)ld * wd, rows, (size_t)n, (size_t)p * wd, h->st)) return -1;
  if (launch_widen_counts(narrow.ptr, dtype, (long)count, h->Yraw.ptr, h->st)) return -1;
  if (int rc = finish_upload(h, n_nonint, n_nonfinite, n_empty_genes)) return rc;
  CA_CUDA(cudaStreamSynchronize(h->st));  // the narrow staging copy goes out of scope
  return 0;
}

extern "C" int ca_gcate_upload_counts_csr(ca_gcate_t h, const int64_t* indptr, const int32_t* indices, const double* data,
                                          int64_t nnz, int64_t* n_nonint, int64_t* n_nonfinite,
                                          int64_t* n_empty_genes) {
  CA_REQUIRE(h && indptr && (nnz == 0 || (indices && data)) && nnz >= 0, "null argument");
  const int n = h->n, p = h->p;
  CA_REQUIRE(indptr[0] == 0 && indptr[n] == nnz, "indptr does not describe n rows with nnz entries");
  if (h->Yraw.alloc((size_t)n * p) || h->Yraw_t.alloc((size_t)p * h->ldn)) return -1;
  DevBuf<long long> d_ptr;
  DevBuf<int> d_idx;
  DevBuf<double> d_val;
  if (d_ptr.alloc((size_t)n + 1) || d_idx.alloc((size_t)std::max<int64_t>(nnz, 1)) ||
      d_val.alloc((size_t)std::max<int64_t>(nnz, 1)))
    return -1;
  CA_CUDA(cudaMemcpyAsync(d_ptr.ptr, indptr, sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, h->st));
  if (nnz > 0) {
    if (h2d_staged(d_idx.ptr, indices, sizeof(int) * (size_t)nnz, h->st)) return -1;
    if (h2d_staged(d_val.ptr, data, sizeof(double) * (size_t)nnz, h->st)) return -1;
  }
  CA_CUDA(cudaMemsetAsync(h->Yraw.ptr, 0, h->Yraw.bytes(), h->st));
  if (launch_csr_expand(d_ptr.ptr, d_idx.ptr, d_val.ptr, n, p, h->Yraw.ptr, h->st)) return -1;
  if (int rc = finish_upload(h, n_nonint, n_nonfinite, n_empty_genes)) return rc;
  CA_CUDA(cudaStreamSynchronize(h->st));  // the CSR staging buffers go out of scope
  return 0;
}

extern "C" int ca_gcate_size_factor_medians(ca_gcate_t h, double* log_sf) {
  CA_REQUIRE(h && h->raw_loaded && log_sf, "upload counts first");
  const int n = h->n, p = h->p;
  DevBuf<double> lg, out;
  DevBuf<unsigned long long> scratch;
  if (lg.alloc(p) || out.alloc(n) || scratch.alloc((size_t)n * p)) return -1;
  if (launch_size_factor_medians(h->Yraw.ptr, p, h->Yraw_t.ptr, h->ldn, n, p, lg.ptr, scratch.ptr, out.ptr, h->st))
    return -1;
  CA_CUDA(cudaMemcpyAsync(log_sf, out.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_gcate_prepare(ca_gcate_t h, const double* size_factor, const double* nu) {
  CA_REQUIRE(h && h->raw_loaded, "upload counts first");
  const int n = h->n, p = h->p;
  DevBuf<double> sf;
  const double* sfp = nullptr;
  if (size_factor) {
    if (sf.alloc(n)) return -1;
    CA_CUDA(cudaMemcpyAsync(sf.ptr, size_factor, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
    sfp = sf.ptr;
  }
  if (launch_normalize_transpose(h->Yraw.ptr, n, p, p, sfp, h->Yn.ptr, h->ldp, h->Ynt.ptr, h->ldn, h->st)) return -1;
  return finish_load(h, nu, nullptr);
}

extern "C" int ca_gcate_load_counts(ca_gcate_t h, const double* Y, const double* size_factor, const double* nu) {
  if (int rc = ca_gcate_upload_counts(h, Y, nullptr, nullptr, nullptr)) return rc;
  return ca_gcate_prepare(h, size_factor, nu);
}

extern "C" int ca_gcate_load_normalized(ca_gcate_t h, const double* Yn, const double* Ys, const double* nu) {
  CA_REQUIRE(h && Yn, "null argument");
  const int n = h->n, p = h->p;
  DevBuf<double> tmp, ys;
  if (tmp.alloc((size_t)n * p)) return -1;
  if (h2d_staged(tmp.ptr, Yn, sizeof(double) * (size_t)n * p, h->st)) return -1;
  if (launch_normalize_transpose(tmp.ptr, n, p, p, nullptr, h->Yn.ptr, h->ldp, h->Ynt.ptr, h->ldn, h->st)) return -1;
  if (Ys) {
    if (ys.alloc((size_t)n * p)) return -1;
    if (h2d_staged(ys.ptr, Ys, sizeof(double) * (size_t)n * p, h->st)) return -1;
  }
  return finish_load(h, nu, Ys ? ys.ptr : nullptr);
}

static int upload_padded(double* dst, const double* src, int rows, int k, int kpad, cudaStream_t st) {
  CA_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * (size_t)rows * kpad, st));
  CA_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * kpad, src, sizeof(double) * k, sizeof(double) * k, rows,
                            cudaMemcpyHostToDevice, st));
  return 0;
}

extern "C" int ca_gcate_set_factors(ca_gcate_t h, const double* A, const double* B) {
  CA_REQUIRE(h, "null handle");
  if (A && upload_padded(h->A.ptr, A, h->n, h->k, h->kpad, h->st)) return -1;
  if (A) {
    h->hostX.resize((size_t)h->n * h->d);
    for (int i = 0; i < h->n; ++i)
      for (int c = 0; c < h->d; ++c) h->hostX[(size_t)i * h->d + c] = A[(size_t)i * h->k + c];
    h->s2_checked = false;
  }
  if (B && upload_padded(h->B.ptr, B, h->p, h->k, h->kpad, h->st)) return -1;
  CA_CUDA(cudaStreamSynchronize(h->st));
  h->factors_set = true;
  return 0;
}

extern "C" int ca_gcate_get_factors(ca_gcate_t h, double* A, double* B) {
  CA_REQUIRE(h, "null handle");
  if (A)
    CA_CUDA(cudaMemcpy2DAsync(A, sizeof(double) * h->k, h->A.ptr, sizeof(double) * h->kpad, sizeof(double) * h->k,
                              h->n, cudaMemcpyDeviceToHost, h->st));
  if (B)
    CA_CUDA(cudaMemcpy2DAsync(B, sizeof(double) * h->k, h->B.ptr, sizeof(double) * h->kpad, sizeof(double) * h->k,
                              h->p, cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_gcate_set_stage(ca_gcate_t h, const double* P1, int d1, const double* P2, int r2, const double* lam,
                                  int a) {
  CA_REQUIRE(h, "null handle");
  CA_REQUIRE(a >= 0 && a <= h->d, "a must be in [0, d]");
  CA_REQUIRE(!(P1 && P2), "P1 and P2 are mutually exclusive");
  h->has_P1 = P1 != nullptr;
  h->has_P2 = P2 != nullptr;
  h->has_lam = lam != nullptr && a > 0;
  h->a = a;
  if (P1) {
    CA_REQUIRE(d1 > 0 && d1 <= 64, "bad P1 width");
    h->d1 = d1;
    if (h->P1.alloc((size_t)h->n * d1)) return -1;
    CA_CUDA(cudaMemcpyAsync(h->P1.ptr, P1, sizeof(double) * (size_t)h->n * d1, cudaMemcpyHostToDevice, h->st));
  }
  if (P2) {
    CA_REQUIRE(r2 > 0 && r2 <= 64, "bad P2 width");
    h->r2 = r2;
    if (h->P2.alloc((size_t)h->p * r2)) return -1;
    CA_CUDA(cudaMemcpyAsync(h->P2.ptr, P2, sizeof(double) * (size_t)h->p * r2, cudaMemcpyHostToDevice, h->st));
  }
  if (a > 0) {
    if (h->lam.alloc((size_t)h->p * a)) return -1;
    if (lam)
      CA_CUDA(cudaMemcpyAsync(h->lam.ptr, lam, sizeof(double) * (size_t)h->p * a, cudaMemcpyHostToDevice, h->st));
    else
      CA_CUDA(cudaMemsetAsync(h->lam.ptr, 0, h->lam.bytes(), h->st));
  }
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_gcate_set_masks(ca_gcate_t h, const uint8_t* gene_active, const uint8_t* cell_active,
                                  const double* alpha_gene) {
  CA_REQUIRE(h, "null handle");
  if (gene_active)
    CA_CUDA(cudaMemcpyAsync(h->gene_active.ptr, gene_active, h->p, cudaMemcpyHostToDevice, h->st));
  else if (launch_fill_u8(h->gene_active.ptr, h->p, 1, h->st))
    return -1;
  if (cell_active)
    CA_CUDA(cudaMemcpyAsync(h->cell_active.ptr, cell_active, h->n, cudaMemcpyHostToDevice, h->st));
  else if (launch_fill_u8(h->cell_active.ptr, h->n, 1, h->st))
    return -1;
  h->has_alpha_gene = alpha_gene != nullptr;
  if (alpha_gene)
    CA_CUDA(cudaMemcpyAsync(h->alpha_gene.ptr, alpha_gene, sizeof(double) * h->p, cudaMemcpyHostToDevice, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_gcate_get_masks(ca_gcate_t h, uint8_t* gene_active, uint8_t* cell_active) {
  CA_REQUIRE(h, "null handle");
  if (gene_active) CA_CUDA(cudaMemcpyAsync(gene_active, h->gene_active.ptr, h->p, cudaMemcpyDeviceToHost, h->st));
  if (cell_active) CA_CUDA(cudaMemcpyAsync(cell_active, h->cell_active.ptr, h->n, cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

static RowProblem cell_problem(ca_gcate* h) {
  RowProblem rp;
  rp.Y = h->Yn.ptr;
  rp.ldY = h->ldp;
  rp.ncols = h->p;
  rp.M = h->B.ptr;
  rp.kpad = h->kpad;
  rp.nu = h->nu.ptr;
  rp.inv_nu = h->inv_nu.ptr;
  rp.log_nu = h->log_nu.ptr;
  rp.colparm = h->colparm.ptr;
  rp.nu_per_row = 0;
  rp.any_pois = h->any_pois;
  rp.family = h->family;
  rp.thres_disp = h->thres;
  return rp;
}
static RowProblem gene_problem(ca_gcate* h) {
  RowProblem rp;
  rp.Y = h->Ynt.ptr;
  rp.ldY = h->ldn;
  rp.ncols = h->n;
  rp.M = h->A.ptr;
  rp.kpad = h->kpad;
  rp.nu = h->nu.ptr;
  rp.inv_nu = h->inv_nu.ptr;
  rp.log_nu = h->log_nu.ptr;
  rp.colparm = nullptr;
  rp.nu_per_row = 1;
  rp.any_pois = h->any_pois;
  rp.family = h->family;
  rp.thres_disp = h->thres;
  return rp;
}

// Cell line search with the genes split over ranks (see k_ls_candidates / k_ls_decide): per round
// the candidate rows are built, their partial log-likelihoods come from the ordinary row pass over
// this rank's genes, one all-reduce completes them and every rank takes the same decisions.  The
// number of rounds is data dependent but identical on all ranks (the reduced sums are), so the
// pending count is read back once per round.
static int cell_search_sharded(ca_gcate* h, const RowProblem& rp, double alpha, double beta, int max_iters, double tol) {
  const int n = h->n, kp = h->kpad;
  cudaStream_t st = h->st;
  int* cnt = h->counts.ptr;
  LsShard a;
  a.X = h->A.ptr;
  a.g = h->g_c.ptr;
  a.Xc = h->ls_Xc.ptr;
  a.t = h->ls_t.ptr;
  a.ne = h->ls_ne.ptr;
  a.ell = h->ls_ell.ptr;
  a.tys = h->tys_row.ptr;
  a.f0 = h->f0_c.ptr;
  a.normg = h->ng_c.ptr;
  a.f01 = h->ls_f01.ptr;
  a.e01 = h->ls_e01.ptr;
  a.ell_fin = h->ellfin_c.ptr;
  a.pend = h->ls_pend.ptr;
  a.n_pending = cnt + 7;
  a.eval_total = h->evals.ptr + 0;
  a.alpha = alpha;
  a.beta = beta;
  a.tol = tol;
  a.ncols_total = genes_total(h);
  a.kpad = kp;
  a.max_iters = max_iters;
  a.rowlist = h->list_c.ptr;
  a.count = cnt + 0;
  CA_CUDA(cudaMemsetAsync(h->ls_pend.ptr, 0, (size_t)n, st));
  for (int round = 0; round < max_iters; ++round) {
    a.round = round;
    if (launch_ls_candidates(a, n, st)) return -1;
    CA_CUDA(cudaMemsetAsync(h->ls_ell.ptr, 0, sizeof(double) * n, st));
    if (launch_rowpass(rp, h->ls_Xc.ptr, a.rowlist, a.count, n, false, 0, 0, h->ls_ell.ptr, nullptr, 0, st)) return -1;
    if (allreduce(h, h->ls_ell.ptr, (size_t)n)) return -1;
    CA_CUDA(cudaMemsetAsync(cnt + 7, 0, sizeof(int), st));
    if (launch_ls_decide(a, n, st)) return -1;
    int npend = 0;
    CA_CUDA(cudaMemcpyAsync(&npend, cnt + 7, sizeof(int), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    if (npend == 0) break;
    if (launch_compact(h->ls_pend.ptr, n, h->ls_list2.ptr, cnt + 5, st)) return -1;
    a.rowlist = h->ls_list2.ptr;
    a.count = cnt + 5;
  }
  return 0;
}

// Stage 2 with one-hot treatment indicators: every treated cell belongs to exactly one of the `a`
// treatment columns of A (values 0 / 1).  Builds the group-sorted treated-cell list, the column map
// for the linear predictor written by the gene gradient pass and the gathered counts.
static int s2_prepare(ca_gcate* h) {
  h->s2_checked = true;
  h->s2_ok = false;
  const int n = h->n, d = h->d, a = h->a;
  if (getenv("CA_NO_S2_ONEHOT") || !h->has_P2 || a < 1 || a > 32 || (int)h->hostX.size() != n * d) return 0;
  std::vector<int> grp(n, -1), cnt(a, 0);
  for (int i = 0; i < n; ++i) {
    const double* x = h->hostX.data() + (size_t)i * d + (d - a);
    for (int c = 0; c < a; ++c) {
      if (x[c] == 0.0) continue;
      if (x[c] != 1.0 || grp[i] >= 0) return 0;     // not an indicator / two treatments in one cell
      grp[i] = c;
    }
    if (grp[i] >= 0) ++cnt[grp[i]];
  }
  int ntr = 0;
  std::vector<int> start(a + 1, 0);
  for (int c = 0; c < a; ++c) start[c + 1] = start[c] + cnt[c];
  ntr = start[a];
  if (ntr == 0) return 0;
  std::vector<int> cells(ntr), pos(n, -1), cur(start.begin(), start.end() - 1);
  std::vector<signed char> g8(ntr);
  for (int i = 0; i < n; ++i)
    if (grp[i] >= 0) {
      const int q = cur[grp[i]]++;
      cells[q] = i;
      pos[i] = q;
      g8[q] = (signed char)grp[i];
    }
  const long ld = (ntr + 7) / 8 * 8;
  const bool same = h->s2_cells_host == cells && h->s2_a == a && h->s2_Ytr.ptr;
  if (h->s2_cells.alloc(ntr) || h->s2_pos.alloc(n) || h->s2_grp.alloc(ntr) || h->s2_Ytr.alloc((size_t)h->p * ld) ||
      h->s2_Th.alloc((size_t)h->p * ld))
    return -1;
  if (!same) {
    CA_CUDA(cudaMemcpyAsync(h->s2_cells.ptr, cells.data(), sizeof(int) * ntr, cudaMemcpyHostToDevice, h->st));
    CA_CUDA(cudaMemcpyAsync(h->s2_pos.ptr, pos.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->st));
    CA_CUDA(cudaMemcpyAsync(h->s2_grp.ptr, g8.data(), ntr, cudaMemcpyHostToDevice, h->st));
    if (launch_s2_gather(h->Ynt.ptr, h->ldn, h->p, h->s2_cells.ptr, ntr, h->s2_Ytr.ptr, ld, h->st)) return -1;
    CA_CUDA(cudaStreamSynchronize(h->st));
    h->s2_cells_host = cells;
  }
  h->s2_ntr = ntr;
  h->s2_a = a;
  h->s2_ld = ld;
  h->s2_ok = true;
  return 0;
}

// One alternating iteration, asynchronous on h->st.  On completion
// scal[0] = sum_j ell_fin_g[j] (log-likelihood sum at the new A, B).
static int enqueue_update(ca_gcate* h, double C, double alpha, double beta, int max_iters, double tol) {
  const int n = h->n, p = h->p, d = h->d, k = h->k, kp = h->kpad, r = h->r;
  cudaStream_t st = h->st;
  int* cnt = h->counts.ptr;
  // ---------------- A-step (cells) ----------------
  if (r > 0) {
    RowProblem rp = cell_problem(h);
    if (launch_compact(h->cell_active.ptr, n, h->list_c.ptr, cnt + 0, st)) return -1;
    if (launch_rowpass(rp, h->A.ptr, h->list_c.ptr, cnt + 0, n, true, d, r, h->ell_c.ptr, h->Graw_c.ptr, h->ldG, st))
      return -1;
    // gene shards: the per-cell log-likelihoods and gradient blocks are sums over all genes
    if (allreduce2(h, h->Graw_c.ptr, (size_t)n * h->ldG, h->ell_c.ptr, (size_t)n)) return -1;
    if (launch_div_scalar(h->Graw_c.ptr, (long)n * h->ldG, genes_total(h), st)) return -1;
    PrepArgs pa;
    pa.nrows = n;
    pa.rowlist = h->list_c.ptr;
    pa.count = cnt + 0;
    pa.kpad = kp;
    pa.ncols = (int)genes_total(h);
    pa.G = h->Graw_c.ptr;
    pa.ldG = h->ldG;
    pa.gc_lo = d;
    pa.gc_n = r;
    pa.ball_lo = d;
    pa.ball_hi = k;
    pa.radius = 2.0 * C;
    pa.X = h->A.ptr;
    pa.lam = nullptr;
    pa.ldlam = 0;
    pa.prox_lo = pa.prox_hi = 0;
    pa.ell0 = h->ell_c.ptr;
    pa.tys = h->tys_row.ptr;
    pa.g = h->g_c.ptr;
    pa.f0 = h->f0_c.ptr;
    pa.normg = h->ng_c.ptr;
    if (launch_prep_search(pa, st)) return -1;
    SearchLaunch s;
    s.X = h->A.ptr;
    s.g = h->g_c.ptr;
    s.f0 = h->f0_c.ptr;
    s.normg = h->ng_c.ptr;
    s.tys = h->tys_row.ptr;
    s.lam = nullptr;
    s.ldlam = 0;
    s.prox_lo = s.prox_hi = 0;
    s.mv_lo = d;
    s.mv_hi = k;
    s.alpha_row = nullptr;
    s.alpha = alpha;
    s.beta = beta;
    s.tol = tol;
    s.max_iters = max_iters;
    s.rowlist = h->list_c.ptr;
    s.count = cnt + 0;
    s.nrows = n;
    s.ell_fin = h->ellfin_c.ptr;
    s.neval = nullptr;
    s.eval_total = h->evals.ptr + 0;
    s.pend = h->ls_pend.ptr;
    s.f01s = h->ls_f01.ptr;
    s.e01s = h->ls_e01.ptr;
    s.list2 = h->ls_list2.ptr;
    s.count2 = cnt + 5;
    if (sharded(h)) {
      if (cell_search_sharded(h, rp, alpha, beta, max_iters, tol)) return -1;
    } else if (launch_search(rp, s, st)) {
      return -1;
    }
    if (h->has_P1) {
      if (launch_pt_v(h->P1.ptr, h->d1, h->d1, h->A.ptr + d, kp, r, n, h->W.ptr, st)) return -1;
      if (launch_sub_p_w(h->P1.ptr, h->d1, h->d1, h->A.ptr + d, kp, r, n, h->W.ptr, st)) return -1;
    }
  }
  // ---------------- B-step (genes) ----------------
  {
    RowProblem rp = gene_problem(h);
    const int a = h->a;
    int gc_lo, gc_n, ball_lo, ball_hi;
    if (h->has_P1) {
      gc_lo = d;
      gc_n = r;
      ball_lo = d;
      ball_hi = k;
    } else if (h->has_P2) {
      gc_lo = d - a;
      gc_n = a;
      ball_lo = 0;
      ball_hi = d;
    } else {
      gc_lo = 0;
      gc_n = k;
      ball_lo = d;
      ball_hi = k;
    }
    const bool any_grad = gc_n > 0;
    if (h->has_P2 && !h->s2_checked && s2_prepare(h)) return -1;
    const bool s2 = h->has_P2 && h->s2_ok && h->s2_a == a && a > 0 && gc_lo == d - a && gc_n == a;
    if (s2) {   // the gradient pass also leaves the linear predictor of the treated cells behind
      rp.theta_out = h->s2_Th.ptr;
      rp.theta_pos = h->s2_pos.ptr;
      rp.theta_ld = h->s2_ld;
    }
    // all genes: the column log-likelihood of inactive genes changes with A
    if (launch_rowpass(rp, h->B.ptr, nullptr, nullptr, p, any_grad, gc_lo, gc_n, h->ell_g.ptr, h->Graw_g.ptr, h->ldG, st))
      return -1;
    CA_CUDA(cudaMemcpyAsync(h->ellfin_g.ptr, h->ell_g.ptr, sizeof(double) * p, cudaMemcpyDeviceToDevice, st));
    if (any_grad) {
      if (launch_div_scalar(h->Graw_g.ptr, (long)p * h->ldG, (double)n, st)) return -1;
      if (h->has_P2) {
        if (launch_pt_v(h->P2.ptr, h->r2, h->r2, h->Graw_g.ptr, h->ldG, a, p, h->W.ptr, st)) return -1;
        if (allreduce(h, h->W.ptr, (size_t)h->r2 * a)) return -1;   // P2^T G sums over genes
        if (launch_sub_p_w(h->P2.ptr, h->r2, h->r2, h->Graw_g.ptr, h->ldG, a, p, h->W.ptr, st)) return -1;
      }
      static const bool nosort = getenv("CA_NOSORT") != nullptr;
      if (nosort ? launch_compact(h->gene_active.ptr, p, h->list_g.ptr, cnt + 1, st)
                 : launch_compact_by_key(h->gene_active.ptr, h->neval_g.ptr, p, h->list_g.ptr, cnt + 1, st))
        return -1;
      PrepArgs pa;
      pa.nrows = p;
      pa.rowlist = h->list_g.ptr;
      pa.count = cnt + 1;
      pa.kpad = kp;
      pa.ncols = n;
      pa.G = h->Graw_g.ptr;
      pa.ldG = h->ldG;
      pa.gc_lo = gc_lo;
      pa.gc_n = gc_n;
      pa.ball_lo = ball_lo;
      pa.ball_hi = ball_hi;
      pa.radius = 2.0 * C;
      pa.X = h->B.ptr;
      pa.lam = (a > 0 && h->has_lam) ? h->lam.ptr : nullptr;
      pa.ldlam = a;
      pa.prox_lo = d - a;
      pa.prox_hi = d;
      pa.ell0 = h->ell_g.ptr;
      pa.tys = h->tys_col.ptr;
      pa.g = h->g_g.ptr;
      pa.f0 = h->f0_g.ptr;
      pa.normg = h->ng_g.ptr;
      if (launch_prep_search(pa, st)) return -1;
      SearchLaunch s;
      s.X = h->B.ptr;
      s.g = h->g_g.ptr;
      s.f0 = h->f0_g.ptr;
      s.normg = h->ng_g.ptr;
      s.tys = h->tys_col.ptr;
      s.lam = pa.lam;
      s.ldlam = a;
      s.prox_lo = d - a;
      s.prox_hi = d;
      // columns that differ between candidates: where the direction is non-zero, plus the prox block
      // (the soft threshold moves a coefficient even where the direction is zero)
      s.mv_lo = pa.lam ? std::min(gc_lo, d - a) : gc_lo;
      s.mv_hi = pa.lam ? std::max(gc_lo + gc_n, d) : gc_lo + gc_n;
      s.alpha_row = h->has_alpha_gene ? h->alpha_gene.ptr : nullptr;
      s.alpha = alpha;
      s.beta = beta;
      s.tol = tol;
      s.max_iters = max_iters;
      s.rowlist = h->list_g.ptr;
      s.count = cnt + 1;
      s.nrows = p;
      s.ell_fin = h->ellfin_g.ptr;
      s.neval = h->neval_g.ptr;
      s.eval_total = h->evals.ptr + 1;
      s.pend = h->ls_pend.ptr;
      s.f01s = h->ls_f01.ptr;
      s.e01s = h->ls_e01.ptr;
      s.list2 = h->ls_list2.ptr;
      s.count2 = cnt + 6;
      if (s2) {
        S2Launch q;
        q.Ytr = h->s2_Ytr.ptr;
        q.Th = h->s2_Th.ptr;
        q.grp = h->s2_grp.ptr;
        q.ldtr = h->s2_ld;
        q.ntr = h->s2_ntr;
        q.a = a;
        q.kpad = kp;
        q.t_lo = d - a;
        q.ncols_total = n;
        q.X = h->B.ptr;
        q.g = h->g_g.ptr;
        q.lam = pa.lam;
        q.f0 = h->f0_g.ptr;
        q.normg = h->ng_g.ptr;
        q.tys = h->tys_col.ptr;
        q.ell0 = h->ell_g.ptr;
        q.alpha_row = s.alpha_row;
        q.alpha = alpha;
        q.beta = beta;
        q.tol = tol;
        q.max_iters = max_iters;
        q.rowlist = h->list_g.ptr;
        q.count = cnt + 1;
        q.nrows = p;
        q.nu = h->nu.ptr;
        q.inv_nu = h->inv_nu.ptr;
        q.log_nu = h->log_nu.ptr;
        q.family = h->family;
        q.thres = h->thres;
        q.ell_fin = h->ellfin_g.ptr;
        q.neval = h->neval_g.ptr;
        q.eval_total = h->evals.ptr + 1;
        if (launch_s2_onehot(q, st)) return -1;
      } else if (launch_search(rp, s, st)) {
        return -1;
      }
    }
    if (!h->has_P2 && r > 0) {
      // clip Gamma to [-10, 10] (gcate_opt.py:202-203); re-evaluate genes that were clipped
      if (launch_clip(h->B.ptr, p, kp, d, k, 10.0, h->list_clip.ptr, cnt + 2, st)) return -1;
      if (launch_rowpass(rp, h->B.ptr, h->list_clip.ptr, cnt + 2, p, false, 0, 0, h->ellfin_g.ptr, nullptr, 0, st))
        return -1;
    }
    if (launch_sum(h->ellfin_g.ptr, p, h->scal.ptr + 0, st)) return -1;
    if (allreduce(h, h->scal.ptr + 0, 1)) return -1;
  }
  return 0;
}

static double nll_from_sum(ca_gcate* h, double ell_sum, double tys_total) {
  return -(ell_sum + tys_total) / (double)h->n;
}

extern "C" int ca_gcate_update(ca_gcate_t h, double C, double alpha, double beta, int max_iters, double tol,
                               double* func_val) {
  CA_REQUIRE(h && h->counts_loaded && h->factors_set, "load counts and factors first");
  if (enqueue_update(h, C, alpha, beta, max_iters, tol)) return -1;
  double host[2];
  CA_CUDA(cudaMemcpyAsync(host, h->scal.ptr, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaMemcpyAsync(host + 1, h->tys_total.ptr, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  if (peer_check(h)) return -1;
  if (func_val) *func_val = nll_from_sum(h, host[0], host[1]);
  return 0;
}

extern "C" int ca_gcate_nll(ca_gcate_t h, double* out) {
  CA_REQUIRE(h && h->counts_loaded && h->factors_set && out, "load counts and factors first");
  RowProblem rp = gene_problem(h);
  if (launch_rowpass(rp, h->B.ptr, nullptr, nullptr, h->p, false, 0, 0, h->ell_g.ptr, nullptr, 0, h->st)) return -1;
  if (launch_sum(h->ell_g.ptr, h->p, h->scal.ptr + 1, h->st)) return -1;
  double host[2];
  CA_CUDA(cudaMemcpyAsync(host, h->scal.ptr + 1, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaMemcpyAsync(host + 1, h->tys_total.ptr, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  *out = nll_from_sum(h, host[0], host[1]);
  return 0;
}

extern "C" int ca_gcate_grad(ca_gcate_t h, int wrt_A, double* out) {
  CA_REQUIRE(h && h->counts_loaded && h->factors_set && out, "load counts and factors first");
  const int k = h->k;
  RowProblem rp = wrt_A ? cell_problem(h) : gene_problem(h);
  const int rows = wrt_A ? h->n : h->p;
  const double* X = wrt_A ? h->A.ptr : h->B.ptr;
  double* G = wrt_A ? h->Graw_c.ptr : h->Graw_g.ptr;
  double* ell = wrt_A ? h->ell_c.ptr : h->ell_g.ptr;
  if (launch_rowpass(rp, X, nullptr, nullptr, rows, true, 0, k, ell, G, h->ldG, h->st)) return -1;
  if (launch_div_scalar(G, (long)rows * h->ldG, (double)rp.ncols, h->st)) return -1;
  CA_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * k, G, sizeof(double) * h->ldG, sizeof(double) * k, rows,
                            cudaMemcpyDeviceToHost, h->st));
  CA_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

extern "C" int ca_gcate_alter_min(ca_gcate_t h, const ca_altmin_params* prm, ca_altmin_result* res, double* hist) {
  CA_REQUIRE(h && prm && res && hist, "null argument");
  CA_REQUIRE(h->counts_loaded && h->factors_set, "load counts and factors first");
  const int n = h->n, p = h->p, d = h->d, kp = h->kpad;
  cudaStream_t st = h->st;
  const int a = prm->a;
  CA_REQUIRE(a >= 0 && a <= d, "a must be in [0, d]");
  h->a = a;
  // stage-2 entry projection  B[:, d-a:d] -= P2 (P2^T B[:, d-a:d])   (gcate_opt.py:337-340)
  if (h->has_P2 && a > 0) {
    if (launch_pt_v(h->P2.ptr, h->r2, h->r2, h->B.ptr + (d - a), kp, a, p, h->W.ptr, st)) return -1;
    if (allreduce(h, h->W.ptr, (size_t)h->r2 * a)) return -1;
    if (launch_sub_p_w(h->P2.ptr, h->r2, h->r2, h->B.ptr + (d - a), kp, a, p, h->W.ptr, st)) return -1;
  }
  // adaptive lasso weights (gcate_opt.py:352-353)
  h->has_lam = false;
  if (a > 0) {
    if (h->lam.alloc((size_t)p * a) || h->l1part.alloc(l1_penalty_blocks(p))) return -1;
    if (launch_adaptive_weights(h->B.ptr, p, kp, d - a, d, prm->lam, h->lam.ptr, a, st)) return -1;
    h->has_lam = prm->lam != 0.0;
  }
  // per-gene initial step (gcate_opt.py:366-372)
  {
    std::vector<double> sp(p), ag(p);
    CA_CUDA(cudaMemcpyAsync(sp.data(), h->sparsity.ptr, sizeof(double) * p, cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < p; ++j) {
      ag[j] = prm->alpha;
      if (sp[j] > prm->sparsity_threshold) ag[j] *= (1.0 + prm->sparsity_boost * sp[j]);
    }
    CA_CUDA(cudaMemcpyAsync(h->alpha_gene.ptr, ag.data(), sizeof(double) * p, cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaStreamSynchronize(st));
    h->has_alpha_gene = true;
  }
  if (launch_fill_u8(h->gene_active.ptr, p, 1, st) || launch_fill_u8(h->cell_active.ptr, n, 1, st)) return -1;
  CA_CUDA(cudaMemsetAsync(h->evals.ptr, 0, h->evals.bytes(), st));

  double tys_total = 0.0;
  CA_CUDA(cudaMemcpyAsync(&tys_total, h->tys_total.ptr, sizeof(double), cudaMemcpyDeviceToHost, st));
  auto objective = [&](double ell_sum, double pen) { return (nll_from_sum(h, ell_sum, tys_total) + pen) / genes_total(h); };
  auto penalty_async = [&]() -> int {
    if (a > 0 && h->has_lam) {
      if (launch_l1_penalty(h->B.ptr, p, kp, d - a, d, h->lam.ptr, a, h->l1part.ptr, h->scal.ptr + 2, st)) return -1;
      return allreduce(h, h->scal.ptr + 2, 1);
    }
    CA_CUDA(cudaMemsetAsync(h->scal.ptr + 2, 0, sizeof(double), st));
    return 0;
  };
  double host[4];
  // initial objective
  {
    RowProblem rp = gene_problem(h);
    if (launch_rowpass(rp, h->B.ptr, nullptr, nullptr, p, false, 0, 0, h->ell_g.ptr, nullptr, 0, st)) return -1;
    if (launch_sum(h->ell_g.ptr, p, h->scal.ptr + 0, st)) return -1;
    if (allreduce(h, h->scal.ptr + 0, 1)) return -1;
    if (penalty_async()) return -1;
    CA_CUDA(cudaMemcpyAsync(host, h->scal.ptr, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
  }
  double f_pre = objective(host[0], host[2]);
  double f = f_pre;
  int nh = 0;
  hist[nh++] = f_pre;
  int64_t tot_g = 0, tot_c = 0;
  int n_act_g = p, n_act_c = n;

  // sparse-gene B-only warm-up (gcate_opt.py:382-398)
  if (prm->warmup_iters > 0) {
    std::vector<double> sp(p);
    CA_CUDA(cudaMemcpy(sp.data(), h->sparsity.ptr, sizeof(double) * p, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> gm(p);
    bool any = false;
    for (int j = 0; j < p; ++j) {
      gm[j] = sp[j] > prm->sparsity_threshold;
      any |= gm[j] != 0;
    }
    if (any) {
      CA_CUDA(cudaMemcpy(h->gene_active.ptr, gm.data(), p, cudaMemcpyHostToDevice));
      if (launch_fill_u8(h->cell_active.ptr, n, 0, st)) return -1;
      for (int w = 0; w < prm->warmup_iters; ++w)
        if (enqueue_update(h, prm->C, prm->alpha, prm->beta, prm->ls_max_iters, prm->tol)) return -1;
      if (penalty_async()) return -1;
      CA_CUDA(cudaMemcpyAsync(host, h->scal.ptr, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
      CA_CUDA(cudaStreamSynchronize(st));
      f_pre = objective(host[0], host[2]);
      f = f_pre;
      hist[0] = f_pre;
      if (launch_fill_u8(h->gene_active.ptr, p, 1, st) || launch_fill_u8(h->cell_active.ptr, n, 1, st)) return -1;
    }
  }

  // Early_Stopping state (causarray/utils.py:101-161)
  int es_step = -1, es_best_step = -1;
  double es_best = INFINITY;
  const bool use_mask = prm->tol_gene > 0 || prm->tol_cell > 0;
  int t = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (t = 0; t < prm->es_max_iters; ++t) {
    if (use_mask) {
      CA_CUDA(cudaMemcpyAsync(h->Bprev.ptr, h->B.ptr, h->B.bytes(), cudaMemcpyDeviceToDevice, st));
      CA_CUDA(cudaMemcpyAsync(h->Aprev.ptr, h->A.ptr, h->A.bytes(), cudaMemcpyDeviceToDevice, st));
    }
    if (enqueue_update(h, prm->C, prm->alpha, prm->beta, prm->ls_max_iters, prm->tol)) return -1;
    tot_g += n_act_g;
    tot_c += n_act_c;
    if (use_mask) {
      if (prm->recheck_interval > 0 && (t + 1) % prm->recheck_interval == 0) {
        if (launch_fill_u8(h->gene_active.ptr, p, 1, st) || launch_fill_u8(h->cell_active.ptr, n, 1, st)) return -1;
        n_act_g = p;
        n_act_c = n;
        CA_CUDA(cudaMemsetAsync(h->counts.ptr + 3, 0xff, 2 * sizeof(int), st));
      } else {
        if (prm->tol_gene > 0) {
          if (launch_step_mask(h->B.ptr, h->Bprev.ptr, p, kp, 0, kp, prm->tol_gene, h->gene_active.ptr,
                               h->counts.ptr + 4, st))
            return -1;
        }
        if (prm->tol_cell > 0) {
          if (launch_step_mask(h->A.ptr, h->Aprev.ptr, n, kp, d, kp, prm->tol_cell, h->cell_active.ptr,
                               h->counts.ptr + 3, st))
            return -1;
        }
      }
    }
    if (penalty_async()) return -1;
    int hc[2] = {-1, -1};
    CA_CUDA(cudaMemcpyAsync(host, h->scal.ptr, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaMemcpyAsync(hc, h->counts.ptr + 3, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    if (use_mask && !(prm->recheck_interval > 0 && (t + 1) % prm->recheck_interval == 0)) {
      if (prm->tol_cell > 0) n_act_c = hc[0];
      if (prm->tol_gene > 0) n_act_g = hc[1];
    }
    f = objective(host[0], host[2]);
    hist[nh++] = f;
    if (!std::isfinite(f) || f > std::fmax(1e3 * std::fabs(f_pre), 1e3)) {
      break;  // divergence guard (gcate_opt.py:442-444)
    }
    // Early_Stopping.__call__
    bool stop = false;
    es_step += 1;
    if (es_step >= prm->es_warmup) {
      const double thr = std::isfinite(es_best) ? prm->es_tolerance + prm->es_rel_tol * std::fabs(es_best) : 0.0;
      if (f < es_best - thr) {
        es_best = f;
        es_best_step = es_step;
      } else if (es_step - es_best_step > prm->es_patience) {
        stop = true;
      }
    }
    if (stop) break;
    f_pre = f;
  }
  if (t == prm->es_max_iters) t = prm->es_max_iters - 1;  // python's loop variable after exhaustion
  if (peer_check(h)) return -1;
  auto t1 = std::chrono::steady_clock::now();
  unsigned long long ev[8];
  CA_CUDA(cudaMemcpy(ev, h->evals.ptr, sizeof ev, cudaMemcpyDeviceToHost));
  res->n_iter = t;
  res->func_val = f;
  res->resid = f_pre - f;
  res->hist_len = nh;
  res->total_gene_updates = tot_g;
  res->total_cell_updates = tot_c;
  res->cell_evals = (int64_t)ev[0];
  res->gene_evals = (int64_t)ev[1];
  res->gene_evals_onehot = (int64_t)ev[4];
  res->onehot_treated_cells = h->s2_ok ? h->s2_ntr : 0;
  if (getenv("CA_TIMING")) {
    std::vector<int> ne(p);
    CA_CUDA(cudaMemcpy(ne.data(), h->neval_g.ptr, sizeof(int) * p, cudaMemcpyDeviceToHost));
    int hist[32] = {0};
    for (int j = 0; j < p; ++j) hist[ne[j] < 0 ? 0 : (ne[j] > 31 ? 31 : ne[j])]++;
    fprintf(stderr, "[ca] candidates per gene in the last line search:");
    for (int b = 0; b < 32; ++b)
      if (hist[b]) fprintf(stderr, " %d:%d", b, hist[b]);
    fprintf(stderr, "\n");
  }
  if (getenv("CA_TIMING")) fprintf(stderr, "[ca] line search: cell evals %llu (tile-lockstep %llu), gene evals %llu (tile-lockstep %llu)\n", ev[0], ev[2], ev[1], ev[3]);
  res->loop_seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}
