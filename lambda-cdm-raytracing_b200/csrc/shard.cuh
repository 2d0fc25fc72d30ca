// shard.cuh -- NCCL-backed source all-gather owned by the context (shard.cu).
#pragma once
#include "common.cuh"

namespace b200 {

struct ShardState;

int shard_unique_id(unsigned char id[B200_SHARD_ID_BYTES]);
int shard_init(b200_ctx* ctx, const unsigned char id[B200_SHARD_ID_BYTES], int rank, int world);
int shard_finalize(b200_ctx* ctx);
int shard_info(const b200_ctx* ctx, int* rank, int* world);
int shard_allgather(b200_ctx* ctx, void* posm4_full, size_t n_total, cudaStream_t st);
int shard_allreduce_f64(b200_ctx* ctx, double* values, size_t count);
int shard_allgather_bytes(b200_ctx* ctx, const void* send, void* recv, size_t bytes, cudaStream_t st);
int shard_bcast_group(b200_ctx* ctx, int count, const void* const* send, void* const* recv, const size_t* bytes,
                      const int* root, cudaStream_t st);
const char* shard_error_string(int nccl_result);

}  // namespace b200
