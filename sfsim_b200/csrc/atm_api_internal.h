// Helpers shared by the translation units behind the C ABI.
#pragma once

#include <string>

#include <cuda_runtime.h>

#include "sfsim_atmosphere.h"
#include "atm_device.cuh"

namespace atm {

int fail(const std::string &msg);
int fail_cuda(cudaError_t e, const char *what);
int ensure_init();
cudaStream_t stream();
// planet + scatter components -> Params (shapes left zero); the planet centre is NOT applied here
int make_planet_medium(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, Params &P);

}  // namespace atm

#define CUDA_TRY(expr)                                          \
  do {                                                          \
    cudaError_t e__ = (expr);                                   \
    if (e__ != cudaSuccess) return atm::fail_cuda(e__, #expr);  \
  } while (0)
