// Stand-in for CUTLASS's examples/common/helper.h, which the reference's
// kernels include (kernels/batch_gather_gemm.cu:68) but which the CUTLASS
// headers bundled in this image do not ship.  Only the two checking macros
// the reference uses are provided.  Test infrastructure only.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "cutlass/cutlass.h"

#define CUTLASS_CHECK(status)                                                   \
  {                                                                             \
    cutlass::Status error_ = (status);                                          \
    if (error_ != cutlass::Status::kSuccess) {                                  \
      std::fprintf(stderr, "CUTLASS error: %s at %s:%d\n",                      \
                   cutlassGetStatusString(error_), __FILE__, __LINE__);         \
      std::exit(EXIT_FAILURE);                                                  \
    }                                                                           \
  }

#define CUDA_CHECK(status)                                                      \
  {                                                                             \
    cudaError_t error_ = (status);                                              \
    if (error_ != cudaSuccess) {                                                \
      std::fprintf(stderr, "CUDA error: %s at %s:%d\n",                         \
                   cudaGetErrorString(error_), __FILE__, __LINE__);             \
      std::exit(EXIT_FAILURE);                                                  \
    }                                                                           \
  }
