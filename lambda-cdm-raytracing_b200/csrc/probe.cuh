#pragma once
#include "common.cuh"
namespace b200 {
int fp32_peak_probe(b200_ctx* ctx, int mode, int iters, double* tflops, float* ms);
}
