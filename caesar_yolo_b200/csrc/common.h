// Shared host-side helpers for the C-ABI layer: status codes, thread-local error text.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/caesar_b200.h"

#include <atomic>

namespace cy {
int set_error(int code, const char* fmt, ...);
const char* last_error();

// Function attributes (dynamic shared memory opt-in), SM counts and memory-pool settings are PER DEVICE: a process
// that drives several GPUs (`--devices cuda:0,cuda:1` without torchrun) must set them once on each.  `mask` is a
// per-call-site bitmask over device ordinals; returns true exactly once per (call site, current device).
inline bool first_use_on_device(std::atomic<unsigned long long>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit) & bit) == 0;
}
inline int current_device_sms() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = cache[dev & 63];
    if (!c) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        c = n > 0 ? n : 148;
    }
    return c;
}
}  // namespace cy

#define CY_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return cy::set_error(CY_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,     \
                                 cudaGetErrorString(_e));                                                \
    } while (0)
