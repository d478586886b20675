// NCCL reached through dlopen (no link-time dependency: the library still loads on a box without
// NCCL, and inside a torch process it binds to the libnccl that torch has already loaded).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ca {

int comm_unique_id(unsigned char out[128]);                                     // 0 or -1 (error recorded)
int comm_init(void** comm, int rank, int world, const unsigned char id[128]);   // collective over the ranks
void comm_destroy(void* comm);
int comm_allreduce_sum(void* comm, double* buf, size_t count, cudaStream_t st); // in place, asynchronous on st

// Small all-reduces over NVLink peer memory (one kernel per call instead of an NCCL launch): every
// rank owns an arena (cudaMalloc + CUDA IPC, mapped by all peers) with two staging areas and a flag
// table.  k_peer_allreduce stages a block's slice locally, raises that block's flag in every peer's
// table (release, system scope), waits for the peers' flags and then sums the peers' staged slices
// -- read straight over NVLink -- in rank order, so every rank gets bit-identical sums.  The staging
// areas alternate by call (the epoch), which is what makes a trailing barrier unnecessary.
constexpr int PEER_MAX_WORLD = 8;
constexpr size_t PEER_MAX_DOUBLES = 512 * 1024;   // 4 MB per staging area; larger reductions go through NCCL
struct PeerArena {
  int rank = 0, world = 1;
  void* base = nullptr;                       // this rank's arena
  void* peer[PEER_MAX_WORLD] = {nullptr};     // every rank's arena in this process's address space
  unsigned long long epoch = 0;
};
// Collective over the ranks of `comm` (which exchanges the IPC handles).  On return *out is a usable
// arena or nullptr when peer mapping is not possible on ANY rank (all ranks agree); -1 only on a
// hard error.
int peer_setup(void* comm, int rank, int world, cudaStream_t st, PeerArena** out);
void peer_destroy(void* comm, PeerArena* a, cudaStream_t st);   // collective (ends with a barrier)
int peer_allreduce_sum(PeerArena* a, double* buf, size_t count, cudaStream_t st);   // count <= PEER_MAX_DOUBLES
int peer_allreduce_sum2(PeerArena* a, double* buf0, size_t count0, double* buf1, size_t count1, cudaStream_t st);
int peer_error(PeerArena* a, cudaStream_t st);   // synchronises st; 1 when a wait timed out since the last call

}  // namespace ca
