// probe.cu -- FP32 pipe probe: the measured denominator of the direct-sum
// roofline (MEASURED_PEAKS.json carries HBM and bf16 tensor numbers only).
// Register-resident FMA chains, 16 independent accumulators per thread, every
// SM filled with 2 x 1024 threads.  mode 0: fma.rn.f32 (FFMA); mode 1:
// fma.rn.f32x2 (FFMA2, two FMAs per lane-instruction).
#include "common.cuh"
#include "probe.cuh"

namespace b200 {
namespace {

constexpr int CHAINS = 16;
constexpr int INNER = 64;

__global__ void __launch_bounds__(1024, 2)
ffma_probe_kernel(float* out, int iters, float a, float b) {
    float acc[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) acc[k] = (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < INNER; ++u)
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += acc[k];
    if (s == 123.456f) out[0] = s;       // never true; keeps the chain alive
}

__global__ void __launch_bounds__(1024, 2)
ffma2_probe_kernel(float* out, int iters, float a, float b) {
    unsigned long long acc[CHAINS / 2], a2, b2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
    for (int k = 0; k < CHAINS / 2; ++k) {
        float x = (float)(threadIdx.x + k);
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(x));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < INNER; ++u)
#pragma unroll
            for (int k = 0; k < CHAINS / 2; ++k)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(a2), "l"(b2));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHAINS / 2; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

// modes 2..5: what else the inner loop issues -- FADD2, FMUL2, and FFMA2 with one
// MUFU.RSQ per 6 / per 3 packed ops (the direct-sum mix is 2 MUFU per 11-12 packed).
template <int MODE>
__global__ void __launch_bounds__(1024, 2)
mix_probe_kernel(float* out, int iters, float a, float b) {
    unsigned long long acc[CHAINS / 2], a2, b2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
    float mu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) mu[k] = 1.0f + threadIdx.x + k;
#pragma unroll
    for (int k = 0; k < CHAINS / 2; ++k) {
        float x = (float)(threadIdx.x + k);
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(x));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < INNER; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS / 2; ++k) {
                if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[k]) : "l"(b2));
                else if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc[k]) : "l"(a2));
                else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(a2), "l"(b2));
            }
            if (MODE == 4) {
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(mu[u & 3]));
                if ((u & 3) == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(mu[0]));   // ~1.25 per 8 packed
            }
            if (MODE == 5) {
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(mu[u & 3]));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(mu[(u + 1) & 3]));          // 2 per 8 packed
            }
        }
    }
    float s = mu[0] + mu[1] + mu[2] + mu[3];
#pragma unroll
    for (int k = 0; k < CHAINS / 2; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

// modes 6/7: FP64 pipe alone (DFMA chains) and FFMA2 + DFMA interleaved 1:1 -- are the
// FP32 and FP64 pipes independent, i.e. does the mix take max(t32, t64) or t32 + t64?
template <bool WITH_F32>
__global__ void __launch_bounds__(1024, 2)
dual_probe_kernel(float* out, int iters, float a, float b) {
    unsigned long long acc[CHAINS / 2], a2, b2;
    double dacc[CHAINS / 2];
    const double da = (double)a, db = (double)b;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
    for (int k = 0; k < CHAINS / 2; ++k) {
        float x = (float)(threadIdx.x + k);
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(x));
        dacc[k] = (double)x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < INNER; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS / 2; ++k) {
                if (WITH_F32) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(a2), "l"(b2));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dacc[k]) : "d"(da), "d"(db));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHAINS / 2; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi + (float)dacc[k];
    }
    if (s == 123.456f) out[0] = s;
}

// modes 8/9/10: register-operand pressure.  The peak probes above feed two of the three
// FMA operands from uniform registers; the direct-sum accumulate step (acc = f*d + acc)
// reads three DISTINCT per-thread registers (six 32-bit registers for the packed form).
// 8: scalar FFMA, 3 register operands; 9: FFMA2, 3 register-pair operands;
// 10: FFMA2 with 2 register-pair operands (a*a + c, the r^2 chain).
template <int MODE>
__global__ void __launch_bounds__(512, 2)      // 64 registers per thread: no spills
regs_probe_kernel(float* out, int iters) {
    float s = 0.f;
    if (MODE == 8) {
        float acc[CHAINS], x[CHAINS], y[CHAINS];
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            acc[k] = (float)(threadIdx.x + k);
            x[k] = 0.999f + 1e-6f * (float)(threadIdx.x + 3 * k);
            y[k] = 1e-3f * (float)(threadIdx.x + 5 * k);
        }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < INNER; ++u)
#pragma unroll
                for (int k = 0; k < CHAINS; ++k)
                    asm volatile("fma.rn.f32 %0, %1, %0, %2;" : "+f"(acc[k]) : "f"(x[k]), "f"(y[k]));
        }
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) s += acc[k];
    } else {
        unsigned long long acc[CHAINS / 2], x[CHAINS / 2], y[CHAINS / 2];
#pragma unroll
        for (int k = 0; k < CHAINS / 2; ++k) {
            float a = (float)(threadIdx.x + k), b = 0.999f + 1e-6f * (float)(threadIdx.x + 3 * k),
                  c = 1e-3f * (float)(threadIdx.x + 5 * k);
            asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(a));
            asm("mov.b64 %0, {%1, %1};" : "=l"(x[k]) : "f"(b));
            asm("mov.b64 %0, {%1, %1};" : "=l"(y[k]) : "f"(c));
        }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < INNER; ++u)
#pragma unroll
                for (int k = 0; k < CHAINS / 2; ++k) {
                    if (MODE == 9) asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(acc[k]) : "l"(x[k]), "l"(y[k]));
                    else asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc[k]) : "l"(x[k]));
                }
        }
#pragma unroll
        for (int k = 0; k < CHAINS / 2; ++k) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
            s += lo + hi;
        }
    }
    if (s == 123.456f) out[0] = s;
}

}  // namespace

int fp32_peak_probe(b200_ctx* ctx, int mode, int iters, double* tflops, float* ms_out) {
    if (iters <= 0 || mode < 0 || mode > 10) return B200_ERR_INVALID;
    B200_TRY(ctx->probe.reserve(256));
    cudaStream_t st = ctx->stream;
    const int grid = ctx->sm_count * 2, block = (mode >= 8) ? 512 : 1024;
    for (int rep = 0; rep < 2; ++rep) {      // rep 0 warms up, rep 1 is timed
        B200_CUDA(cudaEventRecord(ctx->ev0, st));
        if (mode == 0) ffma_probe_kernel<<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 1) ffma2_probe_kernel<<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 2) mix_probe_kernel<2><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 3) mix_probe_kernel<3><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 4) mix_probe_kernel<4><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 5) mix_probe_kernel<5><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 6) dual_probe_kernel<false><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 7) dual_probe_kernel<true><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters, 0.999f, 0.001f);
        else if (mode == 8) regs_probe_kernel<8><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters);
        else if (mode == 9) regs_probe_kernel<9><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters);
        else regs_probe_kernel<10><<<grid, block, 0, st>>>(ctx->probe.as<float>(), iters);
        B200_CUDA(cudaEventRecord(ctx->ev1, st));
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaEventSynchronize(ctx->ev1));
        ctx->launches += 1;
    }
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    // FMAs per thread: iters * INNER * CHAINS (mode 1: CHAINS/2 packed ops, 2 FMAs each)
    double fmas = (double)grid * block * (double)iters * INNER * CHAINS;
    if (tflops) *tflops = 2.0 * fmas / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return B200_OK;
}

}  // namespace b200
