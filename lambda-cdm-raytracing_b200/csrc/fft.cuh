// fft.cuh -- cuFFT bound at run time (dlopen), shared by the initial-conditions step (ics.cu) and
// the power-spectrum diagnostic (diag.cu).  cuFFT is a plain library FFT, off the force hot path.
#pragma once
#include <cufft.h>

namespace b200 {

struct CufftApi {
    void* handle = nullptr;
    cufftResult (*Plan3d)(cufftHandle*, int, int, int, cufftType) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecR2C)(cufftHandle, cufftReal*, cufftComplex*) = nullptr;
    cufftResult (*ExecC2R)(cufftHandle, cufftComplex*, cufftReal*) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool ok = false;
};
const CufftApi* cufft();     // nullptr if libcufft cannot be loaded

#define B200_FFT(call)                                           \
    do {                                                         \
        cufftResult r__ = (call);                                \
        if (r__ != CUFFT_SUCCESS) return 3000 + (int)r__;        \
    } while (0)

}  // namespace b200
