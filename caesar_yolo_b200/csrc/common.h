// Shared host-side helpers for the C-ABI layer: status codes, thread-local error text.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/caesar_b200.h"

namespace cy {
int set_error(int code, const char* fmt, ...);
const char* last_error();
}  // namespace cy

#define CY_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return cy::set_error(CY_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,     \
                                 cudaGetErrorString(_e));                                                \
    } while (0)
