// ics.cuh -- Zel'dovich initial conditions on the device (ics.cu).
#pragma once
#include "common.cuh"

namespace b200 {
int zeldovich_ics(b200_ctx* ctx, const b200_ic_params* p, size_t n_particles, void* posm4, void* vel3,
                  double* stats_out, cudaStream_t st);
}
