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

}  // namespace ca
