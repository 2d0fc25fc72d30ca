// ref_morton.cu -- host-only call of the reference's __host__ __device__
// morton3D / expand_bits (include/forces/barnes_hut_tree.hpp:11-27).
// TEST INFRASTRUCTURE ONLY; compiled by nvcc because that header includes
// <cuda_runtime.h> next to core/math_types.hpp (float3 clash under plain g++).
#include <cmath>
#include <cstddef>
#include <cstdint>
#include "forces/barnes_hut_tree.hpp"

extern "C" {

uint32_t ref_expand_bits(uint32_t v) { return forces::expand_bits(v); }
uint32_t ref_morton3d(float x, float y, float z) { return forces::morton3D(x, y, z); }

// Host evaluation of compute_morton_codes_kernel's normalisation
// (src/forces/barnes_hut_tree.cu:41-54) followed by the reference's morton3D.
void ref_morton_keys(const float* pos3, size_t n, float box, uint32_t* keys) {
    for (size_t i = 0; i < n; ++i) {
        float x = pos3[3 * i + 0] / box, y = pos3[3 * i + 1] / box, z = pos3[3 * i + 2] / box;
        x = x - floorf(x); y = y - floorf(y); z = z - floorf(z);
        keys[i] = forces::morton3D(x, y, z);
    }
}

}
