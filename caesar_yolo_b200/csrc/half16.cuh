// 16-bit float storage helpers shared by the conv / model / preprocessing kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace cy {

// ---- 16-bit storage format of activations / weights: bf16 (default) or fp16 (f16 != 0), chosen per model at run time.
// Both are 2-byte floats moved by the same TMA maps and multiplied by the same tcgen05 kind::f16 / mma.sync m16n8k16
// instructions; only the pack / unpack / max conversions and the instruction's format field differ.
__device__ __forceinline__ uint32_t pack_h2(float a, float b, int f16) {
    if (f16) {
        const __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u, int f16) {
    if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ uint32_t hmax2_any(uint32_t a, uint32_t b, int f16) {
    if (f16) {
        const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
        return *reinterpret_cast<const uint32_t*>(&m);
    }
    const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&m);
}


}  // namespace cy
