#pragma once
#include "common.cuh"
namespace b200 {
int morton_keys(b200_ctx* ctx, const void* posm4, size_t n, float box, uint32_t* keys, cudaStream_t st);
int hilbert_keys(b200_ctx* ctx, const void* posm4, size_t n, float box, const float* root_dev, uint32_t* keys,
                 cudaStream_t st);
size_t sort_scratch_bytes(size_t n);
int sort_pairs(b200_ctx* ctx, const uint32_t* keys_in, size_t n, uint32_t* keys_out,
               int32_t* perm_out, int end_bit, void* scratch, cudaStream_t st);
}
