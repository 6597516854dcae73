// Implicit-GEMM convolution on tcgen05/TMEM for NHWC bf16 activations (sm_100a only).
//
// Replaces the cuDNN conv calls that ultralytics' DetectionModel.forward issues for the reference
// (reference call site: caesar_yolo/evaluation.py:181-193 -> ultralytics Conv = Conv2d+BN+SiLU).
//
// GEMM view:  D[M = B*Hout*Wout, N = Cout] = sum over taps (kh,kw) and channel chunks of
//             A_tap[M, KC] * W_tap[N, KC]^T.
// A is never materialised: the K loop streams TMA boxes of the NHWC input (out-of-bounds rows/cols are zero-filled
// by TMA = the conv padding).  Stride-2 convs use four parity-shifted tensor maps (even/odd rows x even/odd cols) so
// that every tap is again a dense box.
//
// What bounds a tcgen05 conv (measured): operand delivery.  An SS-mode MMA of 128 x 128 x 16 reads 8 KB of
// shared memory per 64 clk of math (the whole 128 B/clk port), TMA writes fill the same memory, and L2 delivers about
// 42 B/clk to an SM.  So the design is about operand bytes per flop:
//  * persistent CTAs (one per SM), a work unit = up to two 128-row "half tiles" (two TMEM accumulators) that share
//    every weight tile, x one BLOCK_N column tile; accumulators are double-buffered in TMEM so the epilogue of unit i
//    overlaps the main loop of unit i+1;
//  * PAIR: two CTAs of a cluster run one unit with tcgen05 cta_group::2 (M = 256 per instruction): each CTA stages its
//    own A boxes and half of every weight tile, the leader issues the MMAs, commits are multicast to both CTAs;
//  * the K loop is a list of "A loads" (one TMA box per half tile) each serving a list of taps; a tap = one weight
//    tile + the MMAs that read the A box at a byte offset through the UMMA descriptor start address (the 128B-swizzle
//    XOR is a function of the absolute shared-memory address, so a shift by whole 128 B pixel rows is legal):
//      MODE 0: one A load per tap (any box shape bw x bh x bn = 128 rows): 1x1 convs, maps that 8 x 16 tiles cover badly;
//      MODE 1: 3x3 stride 1, conservative variant: three A loads per channel chunk (one per horizontal tap), box
//              8 x 18 pixels; the three vertical taps are the same box read at +0/+1/+2 image rows;
//      MODE 2: 3x3 stride 1: ONE A load per channel chunk, box 10 x 18 pixels (halo on all sides); all nine taps read
//              it at (kh*10 + kw) * 128 B, stride between 8-row groups = 10 px * 128 B;
//      MODE 3: 3x3 stride 2: four A loads per channel chunk, one per input parity class (odd/even rows x cols): the
//              9 taps fall into classes of 4 + 2 + 2 + 1 taps that read the same (17|16) x (9|8) pixel box of their
//              parity map at +0/+1 row / column offsets -- A traffic drops from 9 to 4.4 boxes per chunk;
//  * separate shared-memory rings for A boxes and weight tiles, each with its own TMA producer warp; all single-issuer
//    loops are warp-uniform with elect.sync around the issuing instruction (uniform-register UTCHMMA / UTMALDG);
//  * epilogue: TMEM -> bias + SiLU (tanh form) -> swizzled staging buffer (+ TMA-loaded residual, in place) -> one TMA
//    store per warp and 64-column group (TMA clips image borders and channel padding).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cy {

struct ConvALoad {
    signed char map, dw, dh, ntaps, tap0;
    uint32_t bytes;    // bytes one half-tile box of this load delivers
    uint32_t sbo;      // byte stride between 8-row groups of the A operand inside this box
};
struct ConvTap {
    int wtap;          // tap index in the weight K layout (K = wtap*cin + c)
    uint32_t a_off;    // byte offset of the tap's first row inside the A box
};

struct ConvKParams {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    CUtensorMap tmO;               // output (TMA store), box = [o_gw channels, 32-row sub box of the half tile]
    CUtensorMap tmR;               // residual (TMA load), same box
    const float* bias;             // [n_tiles*BLOCK_N], zero padded
    void* out;                     // bf16 or f32, NHWC with out_cstride channels per pixel
    const __nv_bfloat16* res;      // optional residual (added after activation), NHWC
    long long out_cstride, out_coff;
    long long res_cstride, res_coff;
    int B, H, W;                   // output extent
    int bw, bh, bn;                // half tile = bw x bh pixels x bn images, bw*bh*bn == 128 == UMMA M
    int tiles_w, tiles_h, tiles_n, n_half_tiles;
    int halves;                    // half tiles per work unit (1 or 2)
    int n_units_m, n_tiles_n, n_units;
    int block_n;                   // UMMA N
    int acc_stride;                // TMEM columns per accumulator
    int acc_bufs;                  // 1 or 2 accumulator sets
    int tmem_cols;                 // power of two >= 32
    int cchunks, kc, cin;          // K loop: cchunks chunks of kc channels
    int n_aloads;                  // A loads per channel chunk
    int a_stages, b_stages;
    uint32_t a_half_stride;        // distance between the halves inside an A stage (multiple of 1024)
    uint32_t a_stage_bytes;
    uint32_t b_tile_bytes, b_stage_bytes;
    int cout_store;                // number of valid output columns (multiple of 8)
    int o_gw;                      // output columns per epilogue group (one TMA store box)
    uint32_t o_row_bytes;          // o_gw * element size (32 / 64 / 128)
    uint32_t o_sw_mask;            // byte-address swizzle mask of the staging buffer: a ^= (a >> 3) & mask
    int o_esz;                     // 2 (bf16) or 4 (fp32)
    int act;                       // 1 = SiLU
    int out_f32;                   // 1 = write fp32
    int mode;
    int f16;                       // storage format of in / w / res / bf16-sized out: 0 bf16, 1 fp16
    int pair;                      // 1: CTA pairs (cluster of 2, tcgen05 cta_group::2), a unit = 2 x halves half tiles
    unsigned long long* dbg;       // optional per-CTA timeline (clock64 stamps), 8 words per unit, see conv_probe
    int dbg_units;
    int dbg_epi;                   // timeline slots 0..3 = epilogue phase sums instead of the MMA stamps
    ConvALoad aload[9];
    ConvTap tap[9];
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
    int f16;                                              // 0: bf16 tensors, 1: fp16 tensors
};

struct ConvPlan {
    ConvKParams kp;
    dim3 grid;
    int block_n;
    size_t smem;
    double flops;
    double bytes;   // algorithmic HBM bytes: input slice + weights + output slice (+ residual), each touched once
};

// returns 0 on success, fills plan (encodes tensor maps).  err gets a message on failure.
int conv_make_plan(const ConvDesc& d, ConvPlan* plan, char* err, size_t errlen);
int conv_launch(const ConvPlan& plan, cudaStream_t stream);
void conv_set_debug(unsigned long long* dev_buf, int units_per_cta);

}  // namespace cy
