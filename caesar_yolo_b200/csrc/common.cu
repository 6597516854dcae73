#include "common.h"
#include <string.h>

namespace cy {
static thread_local char g_err[1024] = "";
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
const char* last_error() { return g_err; }
}  // namespace cy

extern "C" const char* cy_last_error(void) { return cy::last_error(); }
extern "C" int cy_version(void) { return 100; }
extern "C" int cy_device_check(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
        return cy::set_error(CY_ERR_CUDA, "no CUDA device");
    if (prop.major != 10) return cy::set_error(CY_ERR_CUDA, "device is sm_%d%d, need sm_100", prop.major, prop.minor);
    return CY_OK;
}
extern "C" int cy_memcpy_d2d(void* dst, const void* src, size_t nbytes, uintptr_t stream) {
    CY_CUDA_CHECK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CY_OK;
}
