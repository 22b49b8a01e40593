// Error plumbing shared by the host-side translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/svin_b200.h"

namespace svin {
void set_error(const std::string& msg);
}

#define SVIN_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t svin_e__ = (expr);                                                                   \
    if (svin_e__ != cudaSuccess) {                                                                   \
      ::svin::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(svin_e__) + " (" +    \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                           \
      cudaGetLastError();                                                                            \
      return SVIN_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)
