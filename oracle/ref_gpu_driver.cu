// ref_gpu_driver.cu -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
//
// A C-ABI handle on the reference's own CUDA kernels for the hot path, so that tests/ can run
// them on the B200 next to libb200grav.so: the tiled direct sum K2 / all-pairs K3 behind
// launch_force_computation (src/physics/lambda_cdm_kernels.cu:444-468), the leapfrog K4
// behind launch_leapfrog_update (:470-490) and the energy kernel K6 behind
// launch_energy_computation (:492-516).  The reference source is compiled where it lies
// (oracle/Makefile target `refgpu`, flags of the reference's CMakeLists.txt:93: -O3
// --use_fast_math --expt-relaxed-constexpr) for sm_100a into oracle/_ref/liblcdm_ref_gpu.so;
// nothing of it is copied into this repository.  "Those recompiled kernels are the baseline"
// -- tests/test_gpu_ref_kernels.py checks our results against theirs and records their speed.
#include <cuda_runtime.h>
#include "physics/lambda_cdm_kernels.hpp"

namespace {
int fail(cudaError_t e) { return 1000 + (int)e; }
#define TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(e_); } while (0)
}

extern "C" {

// forces3 (host, 3n) = launch_force_computation(posm4 (host, 4n)); ms_out[0] = mean device time of
// `reps` launches after `warmup` (CUDA events on the launching stream).
int refgpu_direct(const float* posm4, float* forces3, int n, float box, float softening, int warmup, int reps,
                  float* ms_out) {
    float4* d_pos = nullptr;
    float3* d_f = nullptr;
    TRY(cudaMalloc(&d_pos, (size_t)n * sizeof(float4)));
    TRY(cudaMalloc(&d_f, (size_t)n * sizeof(float3)));
    TRY(cudaMemcpy(d_pos, posm4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    TRY(cudaEventCreate(&e0));
    TRY(cudaEventCreate(&e1));
    for (int i = 0; i < warmup; ++i) physics::kernels::launch_force_computation(d_pos, d_f, n, box, softening, 0);
    TRY(cudaDeviceSynchronize());
    TRY(cudaEventRecord(e0, 0));
    for (int i = 0; i < reps; ++i) physics::kernels::launch_force_computation(d_pos, d_f, n, box, softening, 0);
    TRY(cudaEventRecord(e1, 0));
    TRY(cudaEventSynchronize(e1));
    TRY(cudaGetLastError());
    float ms = 0.f;
    TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_out) ms_out[0] = reps > 0 ? ms / (float)reps : 0.f;
    TRY(cudaMemcpy(forces3, d_f, (size_t)n * sizeof(float3), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_pos);
    cudaFree(d_f);
    return 0;
}

// One launch_leapfrog_update on host arrays (kick when kick != 0, else drift), in place.
int refgpu_leapfrog(float* posm4, float* vel3, const float* forces3, int n, float dt, float box, double a, int kick) {
    float4* d_pos = nullptr;
    float3 *d_v = nullptr, *d_f = nullptr;
    TRY(cudaMalloc(&d_pos, (size_t)n * sizeof(float4)));
    TRY(cudaMalloc(&d_v, (size_t)n * sizeof(float3)));
    TRY(cudaMalloc(&d_f, (size_t)n * sizeof(float3)));
    TRY(cudaMemcpy(d_pos, posm4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_v, vel3, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_f, forces3, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice));
    physics::kernels::launch_leapfrog_update(d_pos, d_v, d_f, n, dt, box, a, kick != 0, 0);
    TRY(cudaDeviceSynchronize());
    TRY(cudaGetLastError());
    TRY(cudaMemcpy(posm4, d_pos, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(vel3, d_v, (size_t)n * sizeof(float3), cudaMemcpyDeviceToHost));
    cudaFree(d_pos);
    cudaFree(d_v);
    cudaFree(d_f);
    return 0;
}

// launch_energy_computation on host arrays: kinetic and potential energy as the reference computes them
// (one thread per particle over j > i, float partial sums, float atomics across blocks).
int refgpu_energy(const float* posm4, const float* vel3, int n, float box, float softening, float* kinetic,
                  float* potential) {
    float4* d_pos = nullptr;
    float3* d_v = nullptr;
    float* d_e = nullptr;
    TRY(cudaMalloc(&d_pos, (size_t)n * sizeof(float4)));
    TRY(cudaMalloc(&d_v, (size_t)n * sizeof(float3)));
    TRY(cudaMalloc(&d_e, 2 * sizeof(float)));
    TRY(cudaMemcpy(d_pos, posm4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(d_v, vel3, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice));
    physics::kernels::launch_energy_computation(d_pos, d_v, d_e, d_e + 1, n, box, softening, 0);
    TRY(cudaDeviceSynchronize());
    TRY(cudaGetLastError());
    float h[2] = {0.f, 0.f};
    TRY(cudaMemcpy(h, d_e, sizeof h, cudaMemcpyDeviceToHost));
    *kinetic = h[0];
    *potential = h[1];
    cudaFree(d_pos);
    cudaFree(d_v);
    cudaFree(d_e);
    return 0;
}

}  // extern "C"
