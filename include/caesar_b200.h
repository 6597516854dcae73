/* caesar_b200.h — C ABI of libcaesar_b200.so, the B200 (sm_100a) implementation of caesar-yolo's tiled
 * source-finding hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer owned by the
 * caller unless the parameter name ends in _host.  Every function returns CY_OK (0) or a negative status and
 * never throws; cy_last_error() returns the message for the calling thread.  `stream` is a cudaStream_t
 * passed as uintptr_t (0 = legacy default stream).
 *
 * The reference (SKA-INAF/caesar-yolo) is pure Python and has no FFI; each entry point cites the reference
 * Python interface it replaces (file:line relative to the reference tree).
 */
#ifndef CAESAR_B200_H
#define CAESAR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CY_OK 0
#define CY_ERR_INVALID (-1)
#define CY_ERR_CUDA (-2)
#define CY_ERR_NOMEM (-3)
#define CY_ERR_STATE (-4)

const char* cy_last_error(void);
int cy_version(void);
/* CY_OK iff the current device is sm_100. */
int cy_device_check(void);

/* ---- convolution layer primitive (ultralytics Conv = Conv2d+BN+SiLU fused; reference model call at
 *      caesar_yolo/evaluation.py:181-193).  NHWC bf16 in, weights [cout_pad, k*k*cin] bf16 (K-major, BN folded),
 *      bias fp32[cout_pad]; optional residual added after the activation; output bf16 or fp32 written into a
 *      channel slice [out_coff, out_coff+cout) of an NHWC buffer with out_ctot channels. */
int cy_conv_block_n(int cout);
int cy_conv2d_nhwc(const void* in, int B, int Hin, int Win, int in_ctot, int in_coff, int cin, const void* w,
                   const float* bias, int cout, int cout_pad, int ksize, int stride, void* out, int out_ctot,
                   int out_coff, int out_f32, const void* res, int res_ctot, int res_coff, int act,
                   uintptr_t stream);

#ifdef __cplusplus
}
#endif
#endif
