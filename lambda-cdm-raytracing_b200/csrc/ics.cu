// ics.cu -- Zel'dovich initial conditions on the device (the step
// before the hot path).
//
// Follows the reference's InitialConditionsGenerator for everything it defines:
//   * P(k) = k^n_s T(k)^2 with the BBKS-form transfer function, Gamma = omega_m h
//     (src/physics/initial_conditions.cpp:83-105), normalised to sigma_8 at z = 0 by the
//     top-hat integral of :146-171 (1000 log-spaced midpoints over [1e-3, 1e2] h/Mpc);
//   * growth factor by the Carroll et al. (1992) form and growth rate Omega_m(a)^0.55
//     (include/physics/cosmology_model.hpp:79-97);
//   * particles start on the cell centres (i + 1/2) dx (:412-418), flat index
//     i*G*G + j*G + k (initial_conditions.hpp:143-151), x = q + D psi wrapped into
//     [0, box) (:279-298), v = a H(a) f(a) D psi (:334-355), unit masses and the
//     stride subsample of grid_to_particles (:358-380).
// It deliberately does NOT follow :304-332, where the reference uses the k-space value
// -i k delta_k / k^2 of a mode as the real-space displacement of the grid point with the
// same flat index (there is no inverse transform anywhere in that file), which yields an
// r.m.s. displacement of half a box instead of a Zel'dovich field.  Here the field is
//     w(x) ~ N(0,1) per cell (counter-based hash RNG: any cell, any order, any GPU)
//     W_k = FFT[w]                                                        (cuFFT R2C)
//     psi_k = i k / k^2 * sqrt(P(k) / (V G^3)) * W_k   (k = 0 and the axis' Nyquist plane dropped)
//     psi(x) = inverse FFT, one axis at a time                            (cuFFT C2R)
// so that <|delta_k|^2> = V P(k) in the reference's convention (:241-244) and
// div psi = -delta.  cuFFT (a plain library FFT, off the hot path) is bound at run time.
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "fft.cuh"
#include "ics.cuh"

namespace b200 {

namespace {

CufftApi g_fft;
std::once_flag g_fft_once;

void load_cufft() {
    const char* names[] = {"libcufft.so.11", "libcufft.so.12", "libcufft.so"};
    for (const char* nm : names) {
        g_fft.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_fft.handle) break;
    }
    if (!g_fft.handle) return;
#define B200_SYM(field, name) \
    *(void**)(&g_fft.field) = dlsym(g_fft.handle, name); \
    if (!g_fft.field) return;
    B200_SYM(Plan3d, "cufftPlan3d")
    B200_SYM(SetStream, "cufftSetStream")
    B200_SYM(ExecR2C, "cufftExecR2C")
    B200_SYM(ExecC2R, "cufftExecC2R")
    B200_SYM(Destroy, "cufftDestroy")
#undef B200_SYM
    g_fft.ok = true;
}

// ---- host scalars (double) ---------------------------------------------------
double transfer_bbks(double k, double gamma) {                  // initial_conditions.cpp:83-96
    const double q = k / gamma;
    return log(1.0 + 2.34 * q) / (2.34 * q) *
           pow(1.0 + 3.89 * q + pow(16.1 * q, 2) + pow(5.46 * q, 3) + pow(6.71 * q, 4), -0.25);
}
double power_unnormalised(double k, double n_s, double gamma) {  // :98-105
    const double t = transfer_bbks(k, gamma);
    return pow(k, n_s) * t * t;
}
double sigma8_of(double n_s, double gamma) {                    // :146-171
    const double R8 = 8.0;
    const int n = 1000;
    const double l0 = log(0.001), l1 = log(100.0), dl = (l1 - l0) / n;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const double k = exp(l0 + (i + 0.5) * dl), kr = k * R8;
        const double w = 3.0 * (sin(kr) - kr * cos(kr)) / (kr * kr * kr);
        s += power_unnormalised(k, n_s, gamma) * w * w * k * k * k * dl;
    }
    return sqrt(s / (2.0 * M_PI * M_PI));
}
double hubble_a(const b200_ic_params& p, double a) {             // cosmology_model.hpp:49-61
    return 100.0 * p.h * sqrt(p.omega_m * pow(a, -3) + p.omega_k * pow(a, -2) + p.omega_lambda);
}
double omega_m_a(const b200_ic_params& p, double a) {            // :99-103
    const double e2 = pow(hubble_a(p, a) / (100.0 * p.h), 2);
    return p.omega_m * pow(a, -3) / e2;
}
double growth_factor(const b200_ic_params& p, double a) {        // :79-91 (Carroll et al. 1992)
    const double e2 = pow(hubble_a(p, a) / (100.0 * p.h), 2);
    const double om = omega_m_a(p, a), ol = p.omega_lambda / e2;
    return a * pow(om, 0.6) / (pow(om, 0.6) + ol * (1.0 + om / 70.0));
}

// ---- device ---------------------------------------------------------------------
// splitmix64 finaliser over (seed, cell): one 64-bit draw per cell -> two 24-bit uniforms
__host__ __device__ inline unsigned long long ic_hash(unsigned long long cell, unsigned int seed) {
    unsigned long long z = (cell + 1ull) * 0x9E3779B97F4A7C15ull + (unsigned long long)seed * 0xD1B54A32D192ED03ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

__global__ void ic_noise_kernel(float* __restrict__ w, long long n_cells, unsigned int seed) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells;
         c += (long long)gridDim.x * blockDim.x) {
        const unsigned long long z = ic_hash((unsigned long long)c, seed);
        const double u1 = ((double)(z >> 40) + 0.5) * (1.0 / 16777216.0);          // (0,1)
        const double u2 = ((double)((z >> 16) & 0xFFFFFFull) + 0.5) * (1.0 / 16777216.0);
        w[c] = (float)(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));                    // Box-Muller
    }
}

// psi_k along `axis` from W_k.  Layout [G][G][G/2+1] (cuFFT R2C), z fastest.
__global__ void ic_psi_kernel(const float2* __restrict__ wk, float2* __restrict__ out, int G, int axis,
                              double dk, double gamma, double n_s, double norm /* P scale / (V G^3) */) {
    const int H = G / 2 + 1;
    const long long total = (long long)G * G * H;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int kz = (int)(t % H);
        const int jy = (int)((t / H) % G);
        const int ix = (int)(t / ((long long)H * G));
        const int nx = ix <= G / 2 ? ix : ix - G;          // initial_conditions.cpp:420-431
        const int ny = jy <= G / 2 ? jy : jy - G;
        const int nz = kz;
        const int na = axis == 0 ? nx : (axis == 1 ? ny : nz);
        float2 o = make_float2(0.f, 0.f);
        const long long n2 = (long long)nx * nx + (long long)ny * ny + (long long)nz * nz;
        const bool nyquist = (2 * (na < 0 ? -na : na) == G);
        if (n2 > 0 && !nyquist) {
            const double k = dk * sqrt((double)n2);
            const double q = k / gamma;
            const double tr = log(1.0 + 2.34 * q) / (2.34 * q) *
                              pow(1.0 + 3.89 * q + (16.1 * q) * (16.1 * q) + (5.46 * q) * (5.46 * q) * (5.46 * q) +
                                      (6.71 * q) * (6.71 * q) * (6.71 * q) * (6.71 * q), -0.25);
            const double amp = sqrt(pow(k, n_s) * tr * tr * norm) * (dk * na) / (k * k);
            const float2 v = wk[t];
            o.x = (float)(-amp * (double)v.y);             // i * (a + ib) = -b + ia
            o.y = (float)(amp * (double)v.x);
        }
        out[t] = o;
    }
}

// ---- second order (2LPT) ---------------------------------------------------------
// phi1,ab(k) = k_a k_b delta_k / k^2  (phi1 = inverse Laplacian of delta, psi1 = -grad phi1).  For a != b the
// Nyquist planes of a and b are dropped (k_a k_b is not odd there), as for psi.
__global__ void ic_d2_kernel(const float2* __restrict__ wk, float2* __restrict__ out, int G, int a, int b,
                             double dk, double gamma, double n_s, double norm) {
    const int H = G / 2 + 1;
    const long long total = (long long)G * G * H;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int kz = (int)(t % H);
        const int jy = (int)((t / H) % G);
        const int ix = (int)(t / ((long long)H * G));
        const int n3[3] = {ix <= G / 2 ? ix : ix - G, jy <= G / 2 ? jy : jy - G, kz};
        const long long n2 = (long long)n3[0] * n3[0] + (long long)n3[1] * n3[1] + (long long)n3[2] * n3[2];
        const bool nyq = a != b && (2 * (n3[a] < 0 ? -n3[a] : n3[a]) == G || 2 * (n3[b] < 0 ? -n3[b] : n3[b]) == G);
        float2 o = make_float2(0.f, 0.f);
        if (n2 > 0 && !nyq) {
            const double k = dk * sqrt((double)n2);
            const double q = k / gamma;
            const double tr = log(1.0 + 2.34 * q) / (2.34 * q) *
                              pow(1.0 + 3.89 * q + (16.1 * q) * (16.1 * q) + (5.46 * q) * (5.46 * q) * (5.46 * q) +
                                      (6.71 * q) * (6.71 * q) * (6.71 * q) * (6.71 * q), -0.25);
            const double amp = sqrt(pow(k, n_s) * tr * tr * norm) * (double)n3[a] * (double)n3[b] / (double)n2;
            const float2 v = wk[t];
            o.x = (float)(amp * (double)v.x);
            o.y = (float)(amp * (double)v.y);
        }
        out[t] = o;
    }
}
// S = sum_{a<b} (phi,aa phi,bb - phi,ab^2): the diagonal part, then one off-diagonal term at a time
__global__ void ic_s_diag_kernel(const float* __restrict__ xx, const float* __restrict__ yy,
                                 const float* __restrict__ zz, float* __restrict__ S, long long cells) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x)
        S[c] = xx[c] * yy[c] + xx[c] * zz[c] + yy[c] * zz[c];
}
__global__ void ic_s_offdiag_kernel(const float* __restrict__ ab, float* __restrict__ S, long long cells) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x)
        S[c] -= ab[c] * ab[c];
}
// psi2_k = grad of phi2, phi2 = inverse Laplacian of S: -i k_a S_k / k^2, with the 1/G^3 of the forward transform
__global__ void ic_psi2_kernel(const float2* __restrict__ sk, float2* __restrict__ out, int G, int axis, double dk) {
    const int H = G / 2 + 1;
    const long long total = (long long)G * G * H;
    const double inv_cells = 1.0 / ((double)G * G * G);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int kz = (int)(t % H);
        const int jy = (int)((t / H) % G);
        const int ix = (int)(t / ((long long)H * G));
        const int n3[3] = {ix <= G / 2 ? ix : ix - G, jy <= G / 2 ? jy : jy - G, kz};
        const long long n2 = (long long)n3[0] * n3[0] + (long long)n3[1] * n3[1] + (long long)n3[2] * n3[2];
        const int na = n3[axis];
        float2 o = make_float2(0.f, 0.f);
        if (n2 > 0 && 2 * (na < 0 ? -na : na) != G) {
            const double c = inv_cells * (double)na / (dk * (double)n2);      // k_a / k^2 / G^3
            const float2 v = sk[t];
            o.x = (float)(c * (double)v.y);                                   // -i (x + iy) = y - ix
            o.y = (float)(-c * (double)v.x);
        }
        out[t] = o;
    }
}

__global__ void ic_particles_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                    const float* __restrict__ pz, const float* __restrict__ p2 /* 3 planes or null */,
                                    float growth2, float vfac2, int G, long long n_particles,
                                    long long skip, float dx, float box, float growth, float vfac,
                                    float shift, float mass, float4* __restrict__ posm,
                                    float* __restrict__ vel3, double* __restrict__ stats /* sum psi^2, max |psi| bits */) {
    double s2 = 0.0;
    float mx = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_particles;
         p += (long long)gridDim.x * blockDim.x) {
        const long long c = p * skip;                      // grid_to_particles :371-380
        const int k = (int)(c % G), j = (int)((c / G) % G), i = (int)(c / ((long long)G * G));
        float dxs = growth * px[c], dys = growth * py[c], dzs = growth * pz[c];
        float vx = vfac * dxs, vy = vfac * dys, vz = vfac * dzs;                  // :350-354
        if (p2) {                                  // x = q + D1 psi1 + D2 psi2, v = a H (f1 D1 psi1 + f2 D2 psi2)
            const long long cells = (long long)G * G * G;
            const float ex = growth2 * p2[c], ey = growth2 * p2[cells + c], ez = growth2 * p2[2 * cells + c];
            dxs += ex; dys += ey; dzs += ez;
            vx += vfac2 * ex; vy += vfac2 * ey; vz += vfac2 * ez;
        }
        float x = (i + 0.5f) * dx + dxs, y = (j + 0.5f) * dx + dys, z = (k + 0.5f) * dx + dzs;
        while (x < 0.0f) x += box;                         // :287-292
        while (x >= box) x -= box;
        while (y < 0.0f) y += box;
        while (y >= box) y -= box;
        while (z < 0.0f) z += box;
        while (z >= box) z -= box;
        posm[p] = make_float4(x - shift, y - shift, z - shift, mass);
        vel3[3 * p + 0] = vx;
        vel3[3 * p + 1] = vy;
        vel3[3 * p + 2] = vz;
        const float d2 = dxs * dxs + dys * dys + dzs * dzs;
        s2 += (double)d2;
        mx = fmaxf(mx, d2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], s2);
        atomicMax((unsigned int*)&stats[1], __float_as_uint(mx));      // non-negative floats order as uints
    }
}

}  // namespace

const CufftApi* cufft() {
    std::call_once(g_fft_once, load_cufft);
    return g_fft.ok ? &g_fft : nullptr;
}

int zeldovich_ics(b200_ctx* ctx, const b200_ic_params* p, size_t n_particles, void* posm4, void* vel3,
                  double* stats_out, cudaStream_t st) {
    if (!p || p->grid < 4 || (p->grid & 1) || p->grid > 2048 || !(p->box > 0.f) || !(p->z_initial > -1.0) ||
        !(p->omega_m > 0.0) || !(p->h > 0.0) || !(p->sigma_8 > 0.0))
        return B200_ERR_INVALID;
    const int G = p->grid;
    const long long cells = (long long)G * G * G;
    if (n_particles == 0 || (long long)n_particles > cells) return B200_ERR_INVALID;
    const CufftApi* fft = cufft();
    if (!fft) return B200_ERR_UNSUPPORTED;

    // host scalars
    const double gamma = p->omega_m * p->h;
    const double pnorm = pow(p->sigma_8 / sigma8_of(p->n_s, gamma), 2);      // :131-144
    const double a = 1.0 / (1.0 + p->z_initial);
    const double D = growth_factor(*p, a);                                    // :395-398
    const double f = pow(omega_m_a(*p, a), 0.55);                             // :400-403
    const double vfac = a * hubble_a(*p, a) * f;                              // :340-343
    const double V = (double)p->box * p->box * p->box;
    const double dk = 2.0 * M_PI / p->box;

    const size_t real_bytes = (size_t)cells * sizeof(float);
    const size_t cplx_bytes = (size_t)G * G * (G / 2 + 1) * sizeof(float2);
    B200_TRY(ctx->ic_wk.reserve(cplx_bytes));
    B200_TRY(ctx->ic_tmp.reserve(cplx_bytes));          // psi_k, transformed in place to psi(x)
    const bool lpt2 = p->use_2lpt != 0;
    // planes: psi1 x,y,z [0..2]; 2LPT: psi2 x,y,z [3..5], phi,xx phi,yy phi,zz [6..8], one off-diagonal [9], S [10]
    B200_TRY(ctx->ic_psi.reserve((lpt2 ? 11 : 3) * real_bytes));
    B200_TRY(ctx->ic_stats.reserve(2 * sizeof(double)));
    B200_CUDA(cudaMemsetAsync(ctx->ic_stats.p, 0, 2 * sizeof(double), st));

    const int grid = ctx->sm_count * 8;
    float* w = ctx->ic_psi.as<float>();                  // white noise borrows the first psi plane
    ic_noise_kernel<<<grid, 256, 0, st>>>(w, cells, p->seed);
    B200_CUDA(cudaGetLastError());

    cufftHandle r2c = 0, c2r = 0;
    B200_FFT(fft->Plan3d(&r2c, G, G, G, CUFFT_R2C));
    cufftResult e = fft->Plan3d(&c2r, G, G, G, CUFFT_C2R);
    if (e != CUFFT_SUCCESS) { fft->Destroy(r2c); return 3000 + (int)e; }
    int status = B200_OK;
    do {
        if ((e = fft->SetStream(r2c, st)) != CUFFT_SUCCESS || (e = fft->SetStream(c2r, st)) != CUFFT_SUCCESS ||
            (e = fft->ExecR2C(r2c, w, (cufftComplex*)ctx->ic_wk.p)) != CUFFT_SUCCESS) {
            status = 3000 + (int)e;
            break;
        }
        for (int axis = 0; axis < 3 && status == B200_OK; ++axis) {
            ic_psi_kernel<<<grid, 256, 0, st>>>(ctx->ic_wk.as<float2>(), ctx->ic_tmp.as<float2>(), G, axis, dk,
                                                gamma, p->n_s, pnorm / (V * (double)cells));
            // out-of-place C2R into the axis' real plane (C2R overwrites its input)
            e = fft->ExecC2R(c2r, (cufftComplex*)ctx->ic_tmp.p, ctx->ic_psi.as<float>() + (size_t)axis * cells);
            if (e != CUFFT_SUCCESS) status = 3000 + (int)e;
        }
        if (status == B200_OK && lpt2) {
            // second order: phi1,ab to real space, S = sum_{a<b} (phi,aa phi,bb - phi,ab^2), psi2 = grad lap^-1 S
            float* planes = ctx->ic_psi.as<float>();
            float2* wk = ctx->ic_wk.as<float2>();
            float2* tmp = ctx->ic_tmp.as<float2>();
            const double nrm = pnorm / (V * (double)cells);
            for (int a = 0; a < 3 && status == B200_OK; ++a) {
                ic_d2_kernel<<<grid, 256, 0, st>>>(wk, tmp, G, a, a, dk, gamma, p->n_s, nrm);
                e = fft->ExecC2R(c2r, (cufftComplex*)tmp, planes + (size_t)(6 + a) * cells);
                if (e != CUFFT_SUCCESS) status = 3000 + (int)e;
            }
            float* S = planes + (size_t)10 * cells;
            if (status == B200_OK)
                ic_s_diag_kernel<<<grid, 256, 0, st>>>(planes + 6 * cells, planes + 7 * cells, planes + 8 * cells, S, cells);
            const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
            for (int q = 0; q < 3 && status == B200_OK; ++q) {
                ic_d2_kernel<<<grid, 256, 0, st>>>(wk, tmp, G, pairs[q][0], pairs[q][1], dk, gamma, p->n_s, nrm);
                e = fft->ExecC2R(c2r, (cufftComplex*)tmp, planes + (size_t)9 * cells);
                if (e != CUFFT_SUCCESS) { status = 3000 + (int)e; break; }
                ic_s_offdiag_kernel<<<grid, 256, 0, st>>>(planes + (size_t)9 * cells, S, cells);
            }
            if (status == B200_OK) {            // W_k is no longer needed: its buffer takes S_k
                e = fft->ExecR2C(r2c, S, (cufftComplex*)wk);
                if (e != CUFFT_SUCCESS) status = 3000 + (int)e;
            }
            for (int axis = 0; axis < 3 && status == B200_OK; ++axis) {
                ic_psi2_kernel<<<grid, 256, 0, st>>>(wk, tmp, G, axis, dk);
                e = fft->ExecC2R(c2r, (cufftComplex*)tmp, planes + (size_t)(3 + axis) * cells);
                if (e != CUFFT_SUCCESS) status = 3000 + (int)e;
            }
            ctx->launches += 13;
        }
    } while (0);
    if (status == B200_OK) {
        const long long skip = cells / (long long)n_particles > 1 ? cells / (long long)n_particles : 1;
        const float* psi = ctx->ic_psi.as<float>();
        // second-order growth: D2 = -3/7 D1^2 Omega_m(a)^(-1/143), f2 = 2 Omega_m(a)^(6/11) (Bouchet et al. 1995)
        const double om_a = omega_m_a(*p, a);
        const double D2 = -3.0 / 7.0 * D * D * pow(om_a, -1.0 / 143.0);
        const double vfac2 = a * hubble_a(*p, a) * 2.0 * pow(om_a, 6.0 / 11.0);
        ic_particles_kernel<<<grid, 256, 0, st>>>(psi, psi + cells, psi + 2 * cells, lpt2 ? psi + 3 * cells : nullptr,
                                                  (float)D2, (float)vfac2, G, (long long)n_particles, skip,
                                                  p->box / (float)G, p->box, (float)D, (float)vfac, p->origin_shift,
                                                  p->particle_mass > 0.f ? p->particle_mass : 1.0f, (float4*)posm4,
                                                  (float*)vel3, ctx->ic_stats.as<double>());
        if (cudaGetLastError() != cudaSuccess) status = B200_ERR_INVALID;
    }
    ctx->launches += 5;
    {   // blocking call: the FFT plans must outlive the work queued on them
        cudaError_t ce = cudaSuccess;
        double h[2] = {0.0, 0.0};
        if (status == B200_OK && stats_out)
            ce = cudaMemcpyAsync(h, ctx->ic_stats.p, sizeof h, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess && status == B200_OK) status = 1000 + (int)ce;
        if (status == B200_OK && stats_out) {
            unsigned int bits;
            memcpy(&bits, &h[1], sizeof bits);
            float mx2;
            memcpy(&mx2, &bits, sizeof mx2);
            stats_out[0] = sqrt(h[0] / (double)n_particles);       // r.m.s. displacement
            stats_out[1] = sqrt((double)mx2);                     // largest displacement
            stats_out[2] = D;
            stats_out[3] = vfac;
        }
    }
    fft->Destroy(r2c);
    fft->Destroy(c2r);
    return status;
}

}  // namespace b200
