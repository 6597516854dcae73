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
//
// L2->SMEM traffic is what bounds this kernel (about 43 B/clk/SM against 8192 flop/clk/SM of tensor pipe), so a CTA
// computes up to M = 256 rows (two 128-row halves = two TMEM accumulators sharing every B tile), and 3x3 stride-1
// convs on maps that tile exactly use MODE 1: one K iteration = (channel chunk, horizontal tap dw) loads ONE input
// box of 16 x (8*halves + 2) pixels and the three weight tiles of the vertical taps; the vertical taps are the same
// shared-memory box read at +0/+1/+2 image rows through the UMMA descriptor start address (16 px * row bytes is a
// multiple of the swizzle atom, so the shift keeps the swizzle phase).
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
    int mode;                      // 0: one K iteration per (tap, chunk); 1: per (chunk, dw) with vertical tap reuse
    int halves;                    // 128-row accumulators per CTA (1 or 2)
    int n_mtiles;                  // number of 128-row M tiles (mode 0) / of CTA tiles (mode 1)
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
    int threads;
    int block_n;
    size_t smem;
    double flops;
};

// returns 0 on success, fills plan (encodes tensor maps).  err gets a message on failure.
int conv_make_plan(const ConvDesc& d, ConvPlan* plan, char* err, size_t errlen);
int conv_launch(const ConvPlan& plan, cudaStream_t stream);

}  // namespace cy
