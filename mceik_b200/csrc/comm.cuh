// comm.cuh -- multi-GPU plumbing of libmceik_b200 (comm.cu): NCCL communicator of a context, field assignment.
#pragma once
#include <cstddef>
#include <vector>
#include <cuda_runtime.h>

namespace mceik {
namespace comm {

struct Comm;
void unique_id(void *out128);                                  // ncclGetUniqueId (128 bytes)
Comm *create(int world, int rank, const void *id128);          // ncclCommInitRank
void destroy(Comm *c);
int world(const Comm *c);
int rank(const Comm *c);
// in-place all-gather: rank r's slice is [r * bytes_per_rank, (r + 1) * bytes_per_rank) of d_all
void all_gather_inplace(Comm *c, void *d_all, size_t bytes_per_rank, cudaStream_t st);
// replicated buffer with one-sided puts over NVLink (CUDA IPC mappings of the peers' buffers)
void *alloc_replicated(Comm *c, size_t bytes, cudaStream_t st);  // collective
void free_replicated(Comm *c);
void *replicated_local(const Comm *c);
size_t replicated_bytes(const Comm *c);
void put_to_peers(Comm *c, size_t offset, size_t bytes, cudaStream_t st);
// rank and (rank-major) table row of every field; `slots` = rows per rank
void assign_fields(int nfields, const int *field_model, const int *cost, int world, std::vector<int> &rank_of, std::vector<int> &row_of,
                   int &slots);

}  // namespace comm
}  // namespace mceik
