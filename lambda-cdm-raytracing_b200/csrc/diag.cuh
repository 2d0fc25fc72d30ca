// diag.cuh -- device-side diagnostics (diag.cu).
#pragma once
#include "common.cuh"

namespace b200 {
int force_error(b200_ctx* ctx, const void* acc_test, const void* acc_ref, size_t n, double* avg, double* max,
                cudaStream_t st);
int power_spectrum(b200_ctx* ctx, const void* posm4, size_t n, int grid, float box, int mass_weighted,
                   int shot_noise_correction, float* k_out, float* p_out, int* count_out, cudaStream_t st);
}
