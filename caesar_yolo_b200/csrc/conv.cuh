// Implicit-GEMM convolution on tcgen05/TMEM for NHWC bf16 activations (sm_100a only).
//
// Replaces the cuDNN conv calls that ultralytics' DetectionModel.forward issues for the reference
// (reference call site: caesar_yolo/evaluation.py:181-193 -> ultralytics Conv = Conv2d+BN+SiLU).
//
// GEMM view:  D[M = B*Hout*Wout, N = Cout] = sum over taps (kh,kw) and channel chunks of
//             A_tap[M, KC] * W_tap[N, KC]^T.
// A is never materialised: each K-step is one 4-D TMA box [KC channels, bw, bh, bn] of the NHWC input
// shifted by the tap offset (out-of-bounds rows/cols are zero-filled by TMA = the conv padding).
// Stride-2 convs use four parity-shifted tensor maps (even/odd rows x even/odd cols) so that every tap is
// again a dense box.  bw*bh*bn == 128 == UMMA M.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cy {

struct ConvKParams {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    const float* bias;             // [n_tiles*BLOCK_N], zero padded
    void* out;                     // bf16 or f32, NHWC with out_cstride channels per pixel
    const __nv_bfloat16* res;      // optional residual (added after activation), NHWC
    long long out_cstride, out_coff;
    long long res_cstride, res_coff;
    int B, H, W;                   // output extent
    int bw, bh, bn;                // box = M tile decomposition, bw*bh*bn == 128
    int tiles_w, tiles_h, tiles_n;
    int ntaps, cchunks, kc, cin;   // K loop: ntaps x cchunks steps of kc channels
    int cout_store;                // number of valid output columns (multiple of 8)
    int act;                       // 1 = SiLU
    int out_f32;                   // 1 = write fp32
    int stages;
    signed char tap_map[9], tap_dh[9], tap_dw[9];
};

// Host-side description of one convolution call.
struct ConvDesc {
    const __nv_bfloat16* in;  int in_ctot, in_coff, cin;  // input [B,Hin,Win,in_ctot], channel slice
    int B, Hin, Win;
    int ksize, stride;                                    // (1|3), (1|2); pad = ksize/2
    const __nv_bfloat16* w;   int cout_pad;               // [cout_pad, ksize*ksize*cin], K-major
    const float* bias;        int cout;                   // bias has cout_pad entries
    void* out;                int out_ctot, out_coff, out_f32;
    const __nv_bfloat16* res; int res_ctot, res_coff;
    int act;
};

struct ConvPlan {
    ConvKParams kp;
    dim3 grid;
    int block_n;
    size_t smem;
    double flops;
};

// returns 0 on success, fills plan (encodes tensor maps).  err gets a message on failure.
int conv_make_plan(const ConvDesc& d, ConvPlan* plan, char* err, size_t errlen);
int conv_launch(const ConvPlan& plan, cudaStream_t stream);

}  // namespace cy
