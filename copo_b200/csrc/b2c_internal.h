// Internal helpers shared by the translation units of libcopo_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "../../include/copo_b200.h"

int b2c_set_error(int code, const char* fmt, ...);

#define B2C_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return b2c_set_error(B2C_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
#define B2C_CUDA_OR(expr, cleanup)                                                                \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            cleanup;                                                                              \
            return b2c_set_error(B2C_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));           \
        }                                                                                         \
    } while (0)
