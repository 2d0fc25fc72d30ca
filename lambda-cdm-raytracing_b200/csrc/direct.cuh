// direct.cuh -- internal interface of the direct-sum kernels (direct.cu).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int DIRECT_TILE_J = 512;     // sources per shared-memory tile (8 KB tile-SoA)
constexpr int DIRECT_MAX_PARTS = 16;

// The source set as up to 16 tile-SoA buffers (local, or NVLink-mapped peers).
struct DirectSources {
    const float* tiles[DIRECT_MAX_PARTS];
    int tile_end[DIRECT_MAX_PARTS];    // cumulative tile count after part p
    int n_parts;
    int total_tiles;
};

size_t direct_tiles_bytes(size_t n_particles);
int direct_pack_tiles(b200_ctx* ctx, const void* posm4, size_t n, float* tiles, int* mass_diff,
                      cudaStream_t st);
int direct_forces(b200_ctx* ctx, const DirectSources& src, const void* targets4, size_t n_targets,
                  float eps, float box, void* acc3, const int* mass_diff, cudaStream_t st);
int direct_potential(b200_ctx* ctx, const DirectSources& src, const void* targets4, size_t n_targets,
                     float eps, float box, void* phi, const int* mass_diff, cudaStream_t st);
int energy_reduce(b200_ctx* ctx, const void* targets4, const void* vel3, const void* phi, size_t n,
                  double* out2, cudaStream_t st);

}  // namespace b200
