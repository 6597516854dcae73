// YOLOv8 detection model runtime for B200: graph builder, BN folding, weight packing, activation buffers,
// per-shape execution plans (tensor maps cached), and the non-GEMM layer kernels (stem im2col gather, max-pool,
// nearest upsample).  All dense contractions go through the tcgen05 implicit-GEMM kernel in conv.cu.
//
// Replaces the `model(image, ...)` call of the reference (caesar_yolo/evaluation.py:181-193), i.e. ultralytics
// DetectionModel.forward with fused Conv+BN+SiLU / C2f / SPPF / Upsample / Concat / Detect convs
// (yolov8.yaml).  Concats are zero-copy: producers store into channel slices.
#include "model.h"
#include "common.h"
#include "half16.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace cy {

int conv_block_n(int cout);

// ------------------------------------------------------------------------------------------ small kernels

// Stem (model.0: 3x3 stride-2 pad-1 conv, Cin = 3, + bias + SiLU), one fused kernel: no im2col tensor in HBM.
// Bytes are the bound (read 8 B per input pixel once, write 2*cout B per output pixel once), and K = 27 is too thin
// for a tcgen05 tile pipeline (one K-step per 32 KB of output), so the contraction runs on warp-level mma.sync
// m16n8k16 bf16 with the weights resident in registers as B fragments:
//   * K layout: k = ky*16 + kx*4 + c (c < 4 input channels of the NHWC-4 model input, 4th channel and kx = 3 carry
//     zero weights) -> three K-steps, one per filter row; an A fragment register is then ONE aligned 32-bit shared
//     memory load (two channels of one input pixel) and a warp's 32 loads are 128 contiguous bytes (no conflicts);
//   * CTA = 8 warps = 8 output rows x 64 output columns; the 17 x 132 pixel input patch is staged in shared memory
//     once (halo overhead 7 %), warp w owns output row w as four 16-pixel M tiles;
//   * epilogue: bias + SiLU (same tanh form as the conv kernel) -> bf16 -> per-warp staging rows (padded, conflict
//     free) -> 16-byte stores; the 16 pixels of an M tile are 16*cout*2 contiguous bytes of the NHWC output.
//   * persistent CTAs (two per SM) walk the (image, row block, column block) tiles; the patch of the NEXT tile is
//     fetched with cp.async (8 bytes per pixel, zero fill outside the image = the conv padding) into the second
//     patch buffer while the current one is being multiplied, so the global-load latency never stalls the warps and
//     the weight fragments are read once per CTA.
template <int NT>
struct StemSmem {
    static constexpr int kRows = 17, kCols = 132, kCout = 8 * NT, kStageW = kCout / 2 + 4;  // staging stride in words
    uint2 patch[2][kRows][kCols];   // columns 2*ox0-2 .. 2*ox0+129: starts on an even pixel = 16-byte aligned chunks
    uint32_t stage[8][16 * kStageW];
    float bias[kCout];
};

template <int NT>  // cout = 8 * NT
__global__ void __launch_bounds__(256, (NT <= 8 ? 2 : 1))
stem_conv_kernel(const __nv_bfloat16* __restrict__ in,   // [B,H,W,4]
                 const uint32_t* __restrict__ wfrag,     // [3][NT][32][2] B fragments (see pack_stem_weights)
                 const float* __restrict__ bias,         // [8*NT]
                 __nv_bfloat16* __restrict__ out,        // [B,H/2,W/2,8*NT]
                 int B, int H, int W, int act, int f16) {
    using S = StemSmem<NT>;
    constexpr int kRows = S::kRows, kCols = S::kCols, kCout = S::kCout, kStageW = S::kStageW;
    extern __shared__ __align__(16) unsigned char stem_smem_raw[];
    S& sm = *reinterpret_cast<S*>(stem_smem_raw);
    const int Ho = H >> 1, Wo = W >> 1;
    const int nbx = (Wo + 63) >> 6, nby = (Ho + 7) >> 3;
    const long long ntiles = (long long)B * nby * nbx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint32_t wb[3][NT][2];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(wfrag) + (ky * NT + nt) * 32 + lane);
            wb[ky][nt][0] = q.x;
            wb[ky][nt][1] = q.y;
        }
    if (tid < kCout) sm.bias[tid] = bias[tid];

    // asynchronous fetch of one tile's input patch (rows 2*oy0-1 .. +16, columns 2*ox0-2 .. +129) in 16-byte chunks
    // (two pixels; W is even and chunks start on even columns, so a chunk is entirely inside or outside the image)
    auto fetch = [&](long long tile, int buf) {
        const int bx = (int)(tile % nbx), by = (int)((tile / nbx) % nby), b = (int)(tile / ((long long)nbx * nby));
        const int iy0 = 16 * by - 1, ix0 = 128 * bx - 2;
        const uint2* src = reinterpret_cast<const uint2*>(in) + (long long)b * H * W;
        constexpr int kChunks = kCols / 2;
        for (int i = tid; i < kRows * kChunks; i += 256) {
            const int r = i / kChunks, c = 2 * (i - r * kChunks);
            const int iy = iy0 + r, ix = ix0 + c;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
            const uint2* g = ok ? src + (long long)iy * W + ix : src;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.patch[buf][r][c]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int g = lane >> 2, t = lane & 3;
    uint32_t* st = sm.stage[warp];
    long long tile = blockIdx.x;
    int buf = 0;
    if (tile < ntiles) fetch(tile, 0);
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const long long next = tile + gridDim.x;
        if (next < ntiles) {
            fetch(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();   // patch[buf] (and, first time round, the bias) visible to every warp
        const int bx = (int)(tile % nbx), by = (int)((tile / nbx) % nby), b = (int)(tile / ((long long)nbx * nby));
        const int ox0 = 64 * bx, oy = 8 * by + warp;
        if (oy < Ho) {
            for (int mx = 0; mx < 4; ++mx) {
                const int ox = ox0 + 16 * mx;
                if (ox >= Wo) break;
                float acc[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    // input row 2*oy - 1 + ky = patch row 2*warp + ky; output pixel p of the M tile, tap kx: patch
                    // pixel 2*(16*mx + p) + kx + 1; word index inside the row = 2 * pixel + channel pair
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(&sm.patch[buf][2 * warp + ky][0]) + 64 * mx + lane + 2;
                    const uint32_t a0 = row[0], a1 = row[32], a2 = row[4], a3 = row[36];
                    if (f16) {
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            asm volatile(
                                "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
                                "{%8,%9}, {%0,%1,%2,%3};"
                                : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(wb[ky][nt][0]), "r"(wb[ky][nt][1]));
                    } else {
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            asm volatile(
                                "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
                                "{%8,%9}, {%0,%1,%2,%3};"
                                : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(wb[ky][nt][0]), "r"(wb[ky][nt][1]));
                    }
                }
                // epilogue: thread holds pixels g, g+8 x channels nt*8 + 2t, +1
                // (SiLU tanh form: h = (acc + bias) / 2 exactly as the conv kernel computes it — the halving is exact)
                float y[NT][4];
                const float sc = act == 1 ? 0.5f : 1.0f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 bb = *reinterpret_cast<const float2*>(&sm.bias[nt * 8 + 2 * t]);
                    const float b0 = bb.x * sc, b1 = bb.y * sc;   // sc is a power of two: fma(acc, sc, b*sc) == (acc + b) * sc
                    y[nt][0] = fmaf(acc[nt][0], sc, b0); y[nt][1] = fmaf(acc[nt][1], sc, b1);
                    y[nt][2] = fmaf(acc[nt][2], sc, b0); y[nt][3] = fmaf(acc[nt][3], sc, b1);
                }
                if (act == 1) {
                    float th[NT][4];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm("tanh.approx.f32 %0, %1;" : "=f"(th[nt][j]) : "f"(y[nt][j]));
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) y[nt][j] = fmaf(y[nt][j], th[nt][j], y[nt][j]);
                } else if (act == 2) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int j = 0; j < 4; ++j) y[nt][j] = __fdividef(y[nt][j], 1.0f + __expf(-y[nt][j]));
                }
                __syncwarp();
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    st[g * kStageW + nt * 4 + t] = pack_h2(y[nt][0], y[nt][1], f16);
                    st[(g + 8) * kStageW + nt * 4 + t] = pack_h2(y[nt][2], y[nt][3], f16);
                }
                __syncwarp();
                uint4* dst = reinterpret_cast<uint4*>(out + (((long long)b * Ho + oy) * Wo + ox) * kCout);
#pragma unroll
                for (int v = lane; v < 16 * NT; v += 32) {
                    const int p = v / NT, ch = v - p * NT;
                    dst[v] = *reinterpret_cast<const uint4*>(&st[p * kStageW + ch * 4]);
                }
            }
        }
        __syncthreads();   // every warp is done with patch[buf] before the fetch of the tile after next refills it
    }
}

// B fragments of the stem weights for mma.m16n8k16 (col operand): lane (g = lane / 4, t = lane % 4) of n-tile nt and
// K-step ky holds {B[2t][g], B[2t+1][g]} and {B[2t+8][g], B[2t+9][g]} with B[k][n] = w[n][c][ky][kx] * bn_scale[n],
// k = kx*4 + c (zero for c = 3 and kx = 3).
static inline unsigned short to_h16(float x, int f16) {
    if (f16) {
        const __half h = __float2half_rn(x);
        return *reinterpret_cast<const unsigned short*>(&h);
    }
    const __nv_bfloat16 h = __float2bfloat16(x);
    return *reinterpret_cast<const unsigned short*>(&h);
}
static inline float round_h16(float x, int f16) {
    return f16 ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16(x));
}

static void pack_stem_weights(const std::vector<float>& w, const std::vector<float>& scale, int cout,
                              std::vector<uint32_t>& frag, int f16) {
    const int NT = cout / 8;
    frag.assign((size_t)3 * NT * 32 * 2, 0u);
    auto wt = [&](int n, int ky, int k) -> uint32_t {
        const int kx = k >> 2, c = k & 3;
        if (kx > 2 || c > 2) return 0u;
        return (uint32_t)to_h16(w[((n * 3 + c) * 3 + ky) * 3 + kx] * scale[n], f16);
    };
    for (int ky = 0; ky < 3; ++ky)
        for (int nt = 0; nt < NT; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3, n = nt * 8 + g;
                uint32_t* f = &frag[((size_t)(ky * NT + nt) * 32 + lane) * 2];
                f[0] = wt(n, ky, 2 * t) | (wt(n, ky, 2 * t + 1) << 16);
                f[1] = wt(n, ky, 2 * t + 8) | (wt(n, ky, 2 * t + 9) << 16);
            }
}

template <int NT>
static int launch_stem(const void* in, const ConvW& w, void* out, int B, int H, int W, int act, int f16, cudaStream_t st) {
    static int ctas_dev[64] = {0};   // persistent grid: resident CTAs per SM x SMs, per device ordinal
    const int smem = (int)sizeof(StemSmem<NT>);
    int dev = 0;
    cudaGetDevice(&dev);
    int& ctas = ctas_dev[dev & 63];
    if (!ctas) {
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stem_conv_kernel<NT>, 256, smem);
        if (e != cudaSuccess || per_sm < 1 || sms < 1) return (int)(e != cudaSuccess ? e : cudaErrorLaunchOutOfResources);
        ctas = per_sm * sms;
    }
    const long long ntiles = (long long)B * ((H / 2 + 7) / 8) * ((W / 2 + 63) / 64);
    const unsigned grid = (unsigned)std::min<long long>(ntiles, ctas);
    stem_conv_kernel<NT><<<grid, 256, smem, st>>>((const __nv_bfloat16*)in, (const uint32_t*)w.w, w.b,
                                                  (__nv_bfloat16*)out, B, H, W, act, f16);
    return 0;
}

static int launch_stem_any(const void* in, const ConvW& w, void* out, int B, int H, int W, int act, int f16, cudaStream_t st) {
    switch (w.cout / 8) {
        case 2: return launch_stem<2>(in, w, out, B, H, W, act, f16, st);
        case 4: return launch_stem<4>(in, w, out, B, H, W, act, f16, st);
        case 6: return launch_stem<6>(in, w, out, B, H, W, act, f16, st);
        case 8: return launch_stem<8>(in, w, out, B, H, W, act, f16, st);
        case 10: return launch_stem<10>(in, w, out, B, H, W, act, f16, st);
        case 12: return launch_stem<12>(in, w, out, B, H, W, act, f16, st);
        default: return (int)cudaErrorInvalidValue;
    }
}

// MaxPool2d(5, stride 1, pad 2) on a channel slice; thread = (pixel, 8-channel group).
__global__ void maxpool5_kernel(const __nv_bfloat16* __restrict__ in, int in_ctot, int in_coff,
                                __nv_bfloat16* __restrict__ out, int out_ctot, int out_coff, int B, int H, int W,
                                int C) {
    const int groups = C / 8;
    const long long total = (long long)B * H * W * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    long long pix = idx / groups;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int dh = -2; dh <= 2; ++dh) {
        const int ih = h + dh;
        if (ih < 0 || ih >= H) continue;
        for (int dw = -2; dw <= 2; ++dw) {
            const int iw = w + dw;
            if (iw < 0 || iw >= W) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(
                in + (((long long)b * H + ih) * W + iw) * in_ctot + in_coff + g * 8));
            const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __bfloat1622float2(p[q]);
                m[2 * q] = fmaxf(m[2 * q], f.x);
                m[2 * q + 1] = fmaxf(m[2 * q + 1], f.y);
            }
        }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) o2[q] = __floats2bfloat162_rn(m[2 * q], m[2 * q + 1]);
    *reinterpret_cast<uint4*>(out + pix * out_ctot + out_coff + g * 8) = o;
}

// SPPF pooling chain in one kernel: y1 = m(x), y2 = m(y1), y3 = m(y2) with m = MaxPool2d(5, 1, 2), x = channel slice 0
// of the SPPF concat buffer, y_k = slice k.  CTA = (image, 32 channels): the whole H x W map of those channels lives in
// shared memory and every m is done separably (5-tap row max, then 5-tap column max; out-of-range taps are skipped =
// the -inf padding), so x is read once from HBM and each y_k written once — instead of three 25-tap gathers through
// L1.  bf16 max is exact: results are bit-identical to the per-stage kernel.
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, const uint4 b) {
    __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int q = 0; q < 4; ++q) pa[q] = __hmax2(pa[q], pb[q]);
    return a;
}

__global__ void __launch_bounds__(256) sppf_pool3_kernel(__nv_bfloat16* __restrict__ buf, int ctot, int C, int H, int W) {
    extern __shared__ __align__(16) uint4 sppf_smem[];
    const int n = H * W * 4;   // 16-byte vectors per buffer: (pixel, group of 8 channels), 4 groups = 32 channels
    uint4 *cur = sppf_smem, *tmp = sppf_smem + n, *nxt = sppf_smem + 2 * n;
    __nv_bfloat16* base = buf + (long long)blockIdx.y * H * W * ctot + blockIdx.x * 32;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        cur[i] = __ldg(reinterpret_cast<const uint4*>(base + (long long)(i >> 2) * ctot + (i & 3) * 8));
    __syncthreads();
    for (int stage = 1; stage <= 3; ++stage) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pix = i >> 2, x = pix % W;
            uint4 m = cur[i];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && x + d >= 0 && x + d < W) m = bf16x8_max(m, cur[i + 4 * d]);
            tmp[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pix = i >> 2, y = pix / W;
            uint4 m = tmp[i];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && y + d >= 0 && y + d < H) m = bf16x8_max(m, tmp[i + 4 * d * W]);
            nxt[i] = m;
            *reinterpret_cast<uint4*>(base + (long long)pix * ctot + stage * C + (i & 3) * 8) = m;
        }
        __syncthreads();
        uint4* t = cur; cur = nxt; nxt = t;
    }
}
static constexpr int kSppfMaxSmem = 200 * 1024;

// nn.Upsample(scale_factor=2, mode='nearest') into a channel slice; thread = (INPUT pixel, 8-channel group): one
// 16-byte load, four 16-byte stores (the 2x2 output block), so the source is read exactly once.
__global__ void __launch_bounds__(256) upsample2_kernel(const __nv_bfloat16* __restrict__ in, int in_ctot, int in_coff,
                                                        __nv_bfloat16* __restrict__ out, int out_ctot, int out_coff,
                                                        int B, int H, int W, int C) {  // H, W = input extent
    const int groups = C / 8;
    const long long total = (long long)B * H * W * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const long long pix = idx / groups;           // (b*H + h)*W + w
    const int w = (int)(pix % W);
    const long long bh = pix / W;                 // b*H + h
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + pix * in_ctot + in_coff + g * 8));
    const int Wo = 2 * W;
    __nv_bfloat16* o = out + ((2 * bh) * Wo + 2 * w) * (long long)out_ctot + out_coff + g * 8;
    const long long row = (long long)Wo * out_ctot;
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + out_ctot) = v;
    *reinterpret_cast<uint4*>(o + row) = v;
    *reinterpret_cast<uint4*>(o + row + out_ctot) = v;
}

// Depthwise 3x3 conv (stride 1, pad 1) + bias (+ SiLU) (+ residual) on NHWC bf16; thread = (4 pixels of a row, 8
// channels): the 3 x 6 input window of the four outputs is loaded once (18 16-byte loads for 4 outputs instead of 36)
// and the 9 x 8 weights stay in registers.  yolo11: DWConv of the Detect class branch and the `pe` positional conv of
// the attention block (which reads v straight out of the qkv tensor: logical channel c -> input channel
// in_off + (c / gs) * gst + c % gs).  HBM-bound.
__global__ void __launch_bounds__(256, 3) dwconv3x3_kernel(const __nv_bfloat16* __restrict__ in, int in_ctot, int in_off,
                                                        int gs, int gst, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                        int out_ctot, int out_off, const __nv_bfloat16* __restrict__ res,
                                                        int res_ctot, int res_off, int H, int W, int C, int act, int f16) {
    const int groups = C / 8;
    const int x0 = 4 * (blockIdx.x * (blockDim.x / groups) + threadIdx.x / groups);   // blockDim.x: multiple of groups
    const int g = threadIdx.x % groups;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x0 >= W) return;
    const int c0 = g * 8;
    const int ci = in_off + (c0 / gs) * gst + (c0 % gs);
    float acc[4][8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc[p][0] = b0.x; acc[p][1] = b0.y; acc[p][2] = b0.z; acc[p][3] = b0.w;
            acc[p][4] = b1.x; acc[p][5] = b1.y; acc[p][6] = b1.z; acc[p][7] = b1.w;
        }
    }
    // branch-free taps: out-of-image rows / columns are read at a clamped address and zeroed, so the compiler can
    // issue all 18 window loads of the thread back to back
    uint4 win[3][6];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        const bool rok = iy >= 0 && iy < H;
        const __nv_bfloat16* rowp = in + (((long long)b * H + min(max(iy, 0), H - 1)) * W) * in_ctot + ci;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int ix = x0 - 1 + j;
            win[ky][j] = __ldg(reinterpret_cast<const uint4*>(rowp + (long long)min(max(ix, 0), W - 1) * in_ctot));
        }
        (void)rok;
    }
    // masks in a second sweep: a select next to its load made every load wait for the previous one (in-order issue:
    // 53 % of the kernel's stall samples sat on that select)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        const bool rok = iy >= 0 && iy < H;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int ix = x0 - 1 + j;
            if (!(rok && ix >= 0 && ix < W)) win[ky][j] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        float wk[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float* wt = w + (ky * 3 + kx) * C + c0;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt)), w1 = __ldg(reinterpret_cast<const float4*>(wt + 4));
            wk[kx][0] = w0.x; wk[kx][1] = w0.y; wk[kx][2] = w0.z; wk[kx][3] = w0.w;
            wk[kx][4] = w1.x; wk[kx][5] = w1.y; wk[kx][6] = w1.z; wk[kx][7] = w1.w;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {   // input column x0 - 1 + j feeds output p = j - kx for kx = 0..2
            const uint32_t* pv = reinterpret_cast<const uint32_t*>(&win[ky][j]);
            float f[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 t = unpack_h2(pv[q], f16);
                f[2 * q] = t.x;
                f[2 * q + 1] = t.y;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int p = j - kx;
                if (p >= 0 && p < 4) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[p][e] = fmaf(f[e], wk[kx][e], acc[p][e]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (x >= W) break;
        if (act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (act == 1) {   // the conv kernel's SiLU: h + h * tanh(h), h = y / 2
                    const float h = 0.5f * acc[p][j];
                    float t;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
                    acc[p][j] = fmaf(h, t, h);
                } else {
                    acc[p][j] = __fdividef(acc[p][j], 1.0f + __expf(-acc[p][j]));
                }
            }
        }
        const long long pix = ((long long)b * H + y) * W + x;
        if (res) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(res + pix * res_ctot + res_off + c0));
            const uint32_t* pr = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = unpack_h2(pr[q], f16);
                acc[p][2 * q] += f.x;
                acc[p][2 * q + 1] += f.y;
            }
        }
        uint4 o;
        uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) po[q] = pack_h2(acc[p][2 * q], acc[p][2 * q + 1], f16);
        *reinterpret_cast<uint4*>(out + pix * out_ctot + out_off + c0) = o;
    }
}

// Multi-head self-attention of the yolo11 PSA block over the N = H*W positions of one image (N = 400 at imgsz 640):
// qkv [B,N,nh*128] bf16 with per head 32 q | 32 k | 64 v channels -> out [B,N,nh*64] bf16 slice,
// out[n, head*64 + d] = sum_m softmax_m(q_n . k_m / sqrt(32)) v[m, d].  CTA = (64 queries, head, image): K and V of the
// head sit in shared memory (K rows padded to 17 words: conflict-free when 32 lanes read 32 different keys), one warp
// per query: scores -> shared memory, warp-shuffle max / sum, then every lane accumulates two of the 64 output
// channels.  fp32 arithmetic on CUDA cores: 0.12 GFLOP per image of the 87 GFLOP model, so it is kept simple.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attention_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_ctot,
                                                                __nv_bfloat16* __restrict__ out, int out_ctot, int out_off,
                                                                int N, int f16) {
    extern __shared__ __align__(16) uint32_t attn_smem[];
    uint32_t* sk = attn_smem;                 // [N][17] bf16x2 words (16 used)
    uint32_t* sv = sk + (size_t)N * 17;       // [N][32] bf16x2 words
    float* sp = reinterpret_cast<float*>(sv + (size_t)N * 32);   // [WARPS][N]
    const int head = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __nv_bfloat16* base = qkv + (long long)b * N * qkv_ctot + head * 128;
    for (int i = threadIdx.x; i < N * 16; i += WARPS * 32) {
        const int m = i >> 4, wd = i & 15;
        sk[m * 17 + wd] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)m * qkv_ctot + 32) + wd);
    }
    for (int i = threadIdx.x; i < N * 32; i += WARPS * 32) {
        const int m = i >> 5, wd = i & 31;
        sv[i] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)m * qkv_ctot + 64) + wd);
    }
    __syncthreads();
    float* p = sp + (size_t)warp * N;
    const float scale = 0.17677669529663687f;   // 32^-0.5
    const int q_end = min(N, (int)(blockIdx.x + 1) * 64);
    for (int n = blockIdx.x * 64 + warp; n < q_end; n += WARPS) {
        float2 q[16];
        {
            const uint32_t* qp = reinterpret_cast<const uint32_t*>(base + (long long)n * qkv_ctot);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t u = __ldg(qp + j);
                q[j] = unpack_h2(u, f16);
            }
        }
        float mx = -INFINITY;
        for (int m = lane; m < N; m += 32) {
            const uint32_t* kr = sk + m * 17;
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t u = kr[j];
                const float2 kf = unpack_h2(u, f16);
                sacc = fmaf(q[j].x, kf.x, sacc);
                sacc = fmaf(q[j].y, kf.y, sacc);
            }
            sacc *= scale;
            p[m] = sacc;
            mx = fmaxf(mx, sacc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int m = lane; m < N; m += 32) {
            const float e = __expf(p[m] - mx);
            p[m] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int m = 0; m < N; ++m) {
            const float pm = p[m];
            const uint32_t u = sv[m * 32 + lane];
            const float2 vf = unpack_h2(u, f16);
            o0 = fmaf(pm, vf.x, o0);
            o1 = fmaf(pm, vf.y, o1);
        }
        const float inv = 1.0f / sum;
        *reinterpret_cast<uint32_t*>(out + ((long long)b * N + n) * out_ctot + out_off + head * 64 + 2 * lane) =
            pack_h2(o0 * inv, o1 * inv, f16);
        __syncwarp();
    }
}

// Tensor-core version of the attention above (the one the plan launches; the CUDA-core kernel stays as the fallback
// for maps whose K/V do not fit shared memory in this layout and as a second implementation for the parity tests:
// CY_ATTN_SIMPLE=1).  Flash-style: CTA = (64 queries, head, image), 4 warps x 16 queries; K [Npad][32] and V
// [Npad][64] of the head in shared memory (rows padded to 80 / 144 bytes: conflict-free fragment loads); per chunk of
// 64 keys S = Q K^T with mma.sync.m16n8k16 (bf16, fp32 accumulate), online softmax in registers (exp2, running max /
// sum per query row, quad shuffles), P re-packed as the A operand and O += P V with V fragments from
// ldmatrix.x4.trans.  48 MMAs per warp and chunk; 2.7 ms -> see profiles/ for the measured time.
template <bool F16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    return pack_h2(lo, hi, F16 ? 1 : 0);
}

template <bool F16>
__global__ void __launch_bounds__(128) attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_ctot,
                                                            __nv_bfloat16* __restrict__ out, int out_ctot, int out_off,
                                                            int N, int Npad) {
    extern __shared__ __align__(16) unsigned char attn_mma_smem[];
    unsigned char* sK = attn_mma_smem;                       // [Npad] rows of 80 bytes (64 used)
    unsigned char* sV = attn_mma_smem + (size_t)Npad * 80;   // [Npad] rows of 144 bytes (128 used)
    const int head = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* base = qkv + (long long)b * N * qkv_ctot + head * 128;
    for (int i = threadIdx.x; i < Npad * 4; i += 128) {
        const int m = i >> 2, ch = i & 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (m < N) v = __ldg(reinterpret_cast<const uint4*>(base + (long long)m * qkv_ctot + 32) + ch);
        *reinterpret_cast<uint4*>(sK + (size_t)m * 80 + ch * 16) = v;
    }
    for (int i = threadIdx.x; i < Npad * 8; i += 128) {
        const int m = i >> 3, ch = i & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (m < N) v = __ldg(reinterpret_cast<const uint4*>(base + (long long)m * qkv_ctot + 64) + ch);
        *reinterpret_cast<uint4*>(sV + (size_t)m * 144 + ch * 16) = v;
    }
    // Q fragments of this warp's 16 queries (rows past N are clamped; their results are not stored)
    const int q0 = blockIdx.x * 64 + warp * 16;
    uint32_t qa[2][4];
    {
        const int r0 = min(q0 + g, N - 1), r1 = min(q0 + g + 8, N - 1);
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base + (long long)r0 * qkv_ctot);
        const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base + (long long)r1 * qkv_ctot);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            qa[ks][0] = __ldg(p0 + ks * 8 + t);
            qa[ks][1] = __ldg(p1 + ks * 8 + t);
            qa[ks][2] = __ldg(p0 + ks * 8 + t + 4);
            qa[ks][3] = __ldg(p1 + ks * 8 + t + 4);
        }
    }
    __syncthreads();
    const float c = 0.17677669529663687f * 1.4426950408889634f;   // 32^-0.5 * log2(e): softmax through exp2
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const uint32_t sV_addr = (uint32_t)__cvta_generic_to_shared(sV);
    for (int kc = 0; kc < Npad; kc += 64) {
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
            const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + (size_t)(kc + nt * 8 + g) * 80);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) mma16816<F16>(sc[nt], qa[ks], kr[ks * 8 + t], kr[ks * 8 + t + 4]);
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int key = kc + nt * 8 + 2 * t;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool valid = key + (j & 1) < N;
                sc[nt][j] = valid ? sc[nt][j] * c : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: key 0 of chunk 0 is always valid
        const float al0 = exp2f(m0 - mn0), al1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pa[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(sc[nt][0] - mn0), p1 = exp2f(sc[nt][1] - mn0);
            const float p2 = exp2f(sc[nt][2] - mn1), p3 = exp2f(sc[nt][3] - mn1);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2<F16>(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2<F16>(p2, p3);
        }
        l0 = l0 * al0 + rs0;    // per-lane partial sums of the row; the quad is reduced once at the end
        l1 = l1 * al1 + rs1;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            o[d][0] *= al0; o[d][1] *= al0;
            o[d][2] *= al1; o[d][3] *= al1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // lane L addresses row (L % 8) of 8x8 matrix L / 8: matrices 0/1 = keys +0 / +8 of dim tile d, 2/3 of d + 1
            const int key = kc + 16 * j + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                const uint32_t addr = sV_addr + (uint32_t)key * 144u + (uint32_t)(2 * dp + (lane >> 4)) * 16u;
                uint32_t r0, r1, r2, r3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
                mma16816<F16>(o[2 * dp], pa[j], r0, r1);
                mma16816<F16>(o[2 * dp + 1], pa[j], r2, r3);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const int col = out_off + head * 64 + d * 8 + 2 * t;
        if (r0 < N)
            *reinterpret_cast<uint32_t*>(out + ((long long)b * N + r0) * out_ctot + col) = pack_bf16x2<F16>(o[d][0] * i0, o[d][1] * i0);
        if (r1 < N)
            *reinterpret_cast<uint32_t*>(out + ((long long)b * N + r1) * out_ctot + col) = pack_bf16x2<F16>(o[d][2] * i1, o[d][3] * i1);
    }
}

// Stand-alone stem launch for the parity tests (cy_stem_conv_nhwc4): w fp32 [cout,3,3,3] (OIHW), bias fp32 [cout].
int stem_conv_run(const void* in, int B, int H, int W, const float* w_host, const float* bias_host, int cout, int act,
                  void* out, cudaStream_t st) {
    if (cout % 16 != 0 || cout < 16 || cout > 96)
        return set_error(CY_ERR_INVALID, "stem: cout must be 16, 32, 48, 64, 80 or 96 (got %d)", cout);
    if (H % 2 != 0 || W % 32 != 0) return set_error(CY_ERR_INVALID, "stem: H must be even and W a multiple of 32");
    std::vector<float> w(w_host, w_host + (size_t)cout * 27), scale(cout, 1.f);
    std::vector<uint32_t> frag;
    pack_stem_weights(w, scale, cout, frag, 0);
    ConvW cw;
    cw.cout = cw.cout_pad = cout;
    CY_CUDA_CHECK(cudaMalloc(&cw.w, frag.size() * sizeof(uint32_t)));
    CY_CUDA_CHECK(cudaMalloc(&cw.b, cout * sizeof(float)));
    CY_CUDA_CHECK(cudaMemcpyAsync(cw.w, frag.data(), frag.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CY_CUDA_CHECK(cudaMemcpyAsync(cw.b, bias_host, cout * sizeof(float), cudaMemcpyHostToDevice, st));
    const int lrc = launch_stem_any(in, cw, out, B, H, W, act, 0, st);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(st);
    cudaFree(cw.w);
    cudaFree(cw.b);
    if (lrc) e = (cudaError_t)lrc;
    if (e != cudaSuccess) return set_error(CY_ERR_CUDA, "stem launch failed: %s", cudaGetErrorString(e));
    return CY_OK;
}

// ------------------------------------------------------------------------------------------ model

static int make_divisible(double x, int d) { return (int)(ceil(x / d) * d); }

int Model::init(const char* variant, int nc_) {
    double depth, width;
    int maxc;
    if (variant[0] == '1' && variant[1] == '1') {   // yolo11{n,s,m,l,x}: cfg/models/11/yolo11.yaml scales
        switch (variant[2]) {
            case 'n': depth = 0.50; width = 0.25; maxc = 1024; break;
            case 's': depth = 0.50; width = 0.50; maxc = 1024; break;
            case 'm': depth = 0.50; width = 1.00; maxc = 512; break;
            case 'l': depth = 1.00; width = 1.00; maxc = 512; break;
            case 'x': depth = 1.00; width = 1.50; maxc = 512; break;
            default: return set_error(CY_ERR_INVALID, "unknown YOLO11 variant '%s'", variant);
        }
        auto ch11 = [&](int c) { return make_divisible(std::min(c, maxc) * width, 8); };
        family = 11;
        this->variant = variant[2];
        w64 = ch11(64); w128 = ch11(128); w256 = ch11(256); w512 = ch11(512); w1024 = ch11(1024);
        n11 = std::max((int)lround(2 * depth), 1);
        c3k11 = variant[2] == 'm' || variant[2] == 'l' || variant[2] == 'x';
        nc = nc_;
        if (nc < 1 || nc > 16) return set_error(CY_ERR_INVALID, "nc must be in [1,16] (got %d)", nc);
        c1 = w64; c3 = w256; c4 = w512; c5 = w1024; c2 = w128;
        cb = std::max(16, std::max(w256 / 4, 64));
        cc = std::max(w256, std::min(nc, 100));
        return CY_OK;
    }
    switch (variant[0]) {
        case 'n': depth = 0.33; width = 0.25; maxc = 1024; break;
        case 's': depth = 0.33; width = 0.50; maxc = 1024; break;
        case 'm': depth = 0.67; width = 0.75; maxc = 768; break;
        case 'l': depth = 1.00; width = 1.00; maxc = 512; break;
        case 'x': depth = 1.00; width = 1.25; maxc = 512; break;
        default: return set_error(CY_ERR_INVALID, "unknown YOLOv8 variant '%s'", variant);
    }
    auto ch = [&](int c) { return make_divisible(std::min(c, maxc) * width, 8); };
    // python round(): banker's rounding; the products here (0.99, 1.98, 2.01, 4.02, ...) never hit .5
    auto rep = [&](int n) { return std::max((int)lround(n * depth), 1); };
    c1 = ch(64); c2 = ch(128); c3 = ch(256); c4 = ch(512); c5 = ch(1024);
    n2 = rep(3); n4 = rep(6); n6 = rep(6); n8 = rep(3); nh = rep(3);
    nc = nc_;
    if (nc < 1 || nc > 16) return set_error(CY_ERR_INVALID, "nc must be in [1,16] (got %d)", nc);
    cb = std::max(16, std::max(c3 / 4, 64));
    cc = std::max(c3, std::min(nc, 100));
    this->variant = variant[0];
    return CY_OK;
}

int Model::set_tensor(const char* name, const float* data, long long numel) {
    raw[name] = std::vector<float>(data, data + numel);
    return CY_OK;
}

const std::vector<float>* Model::get(const std::string& k) const {
    auto it = raw.find(k);
    return it == raw.end() ? nullptr : &it->second;
}

static inline float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int Model::add_conv(const std::string& p, int cin, int cout, int k, bool bn, int cin_pad) {
    const std::vector<float>* w = get(bn ? p + ".conv.weight" : p + ".weight");
    if (!w) return set_error(CY_ERR_STATE, "missing tensor %s", (p + (bn ? ".conv.weight" : ".weight")).c_str());
    if ((long long)w->size() != (long long)cout * cin * k * k)
        return set_error(CY_ERR_INVALID, "tensor %s has %zu elements, expected %lld", p.c_str(), w->size(),
                         (long long)cout * cin * k * k);
    std::vector<float> scale(cout, 1.f), bias(cout, 0.f);
    if (bn) {
        const auto *g = get(p + ".bn.weight"), *b = get(p + ".bn.bias"), *m = get(p + ".bn.running_mean"),
                   *v = get(p + ".bn.running_var");
        if (!g || !b || !m || !v) return set_error(CY_ERR_STATE, "missing BN tensors for %s", p.c_str());
        for (int o = 0; o < cout; ++o) {
            const float s = (*g)[o] / sqrtf((*v)[o] + 1e-3f);
            scale[o] = s;
            bias[o] = (*b)[o] - (*m)[o] * s;
        }
    } else {
        const auto* b = get(p + ".bias");
        if (!b) return set_error(CY_ERR_STATE, "missing bias for %s", p.c_str());
        for (int o = 0; o < cout; ++o) bias[o] = (*b)[o];
    }
    ConvW cw;
    cw.cin = cin; cw.cout = cout; cw.k = k;
    if (cin == 3) {  // stem: mma.sync B fragments (pack_stem_weights), consumed by stem_conv_kernel
        if (k != 3 || cout % 16 != 0 || cout > 96)
            return set_error(CY_ERR_INVALID, "stem %s: expected a 3x3 conv with cout %% 16 == 0, cout <= 96", p.c_str());
        cw.cout_pad = cout;
        std::vector<uint32_t> frag;
        pack_stem_weights(*w, scale, cout, frag, f16);
        CY_CUDA_CHECK(cudaMalloc(&cw.w, frag.size() * sizeof(uint32_t)));
        CY_CUDA_CHECK(cudaMemcpy(cw.w, frag.data(), frag.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    } else {
        // cin_pad > cin: the input buffer carries zero-filled padding channels (the conv kernel needs cin % 16 == 0;
        // yolo11n has 8-channel bottleneck intermediates) and their weights are zero
        const int cinp = cin_pad > cin ? cin_pad : cin;
        const int bn_ = conv_block_n(cout);
        cw.cout_pad = (cout + bn_ - 1) / bn_ * bn_;
        const size_t K = (size_t)k * k * cinp;
        std::vector<unsigned short> hw((size_t)cw.cout_pad * K, (unsigned short)0);
        for (int o = 0; o < cout; ++o)
            for (int i = 0; i < cin; ++i)
                for (int kh = 0; kh < k; ++kh)
                    for (int kw = 0; kw < k; ++kw)
                        hw[(size_t)o * K + (size_t)(kh * k + kw) * cinp + i] =
                            to_h16((*w)[(((size_t)o * cin + i) * k + kh) * k + kw] * scale[o], f16);
        cw.cin = cinp;
        CY_CUDA_CHECK(cudaMalloc(&cw.w, hw.size() * sizeof(unsigned short)));
        CY_CUDA_CHECK(cudaMemcpy(cw.w, hw.data(), hw.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    }
    std::vector<float> hb(cw.cout_pad, 0.f);
    for (int o = 0; o < cout; ++o) hb[o] = bias[o];
    CY_CUDA_CHECK(cudaMalloc(&cw.b, hb.size() * sizeof(float)));
    CY_CUDA_CHECK(cudaMemcpy(cw.b, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
    nparams += (long long)cout * cin * k * k;
    convs[p] = cw;
    return CY_OK;
}

int Model::add_c2f(const std::string& p, int cin, int cout, int n) {
    const int c = cout / 2;
    int r;
    if ((r = add_conv(p + ".cv1", cin, 2 * c, 1, true))) return r;
    if ((r = add_conv(p + ".cv2", (2 + n) * c, cout, 1, true))) return r;
    for (int i = 0; i < n; ++i) {
        const std::string m = p + ".m." + std::to_string(i);
        if ((r = add_conv(m + ".cv1", c, c, 3, true))) return r;
        if ((r = add_conv(m + ".cv2", c, c, 3, true))) return r;
    }
    return CY_OK;
}

// Depthwise 3x3 Conv (+BN): fp32 [9][C] weights (BN folded, rounded to bf16 like every other weight) + fp32 bias.
int Model::add_dwconv(const std::string& p, int c) {
    const auto *w = get(p + ".conv.weight"), *g = get(p + ".bn.weight"), *b = get(p + ".bn.bias"),
               *m = get(p + ".bn.running_mean"), *v = get(p + ".bn.running_var");
    if (!w || !g || !b || !m || !v) return set_error(CY_ERR_STATE, "missing tensors for depthwise conv %s", p.c_str());
    if ((long long)w->size() != (long long)c * 9) return set_error(CY_ERR_INVALID, "tensor %s: expected [%d,1,3,3]", p.c_str(), c);
    std::vector<float> hw((size_t)9 * c), hb(c);
    for (int o = 0; o < c; ++o) {
        const float sc = (*g)[o] / sqrtf((*v)[o] + 1e-3f);
        hb[o] = (*b)[o] - (*m)[o] * sc;
        for (int t = 0; t < 9; ++t) hw[(size_t)t * c + o] = round_h16((*w)[(size_t)o * 9 + t] * sc, f16);
    }
    ConvW cw;
    cw.cin = cw.cout = cw.cout_pad = c;
    cw.k = 3;
    CY_CUDA_CHECK(cudaMalloc(&cw.w, hw.size() * sizeof(float)));
    CY_CUDA_CHECK(cudaMemcpy(cw.w, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice));
    CY_CUDA_CHECK(cudaMalloc(&cw.b, hb.size() * sizeof(float)));
    CY_CUDA_CHECK(cudaMemcpy(cw.b, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
    nparams += (long long)c * 9;
    convs[p] = cw;
    return CY_OK;
}

static inline int pad16(int c) { return (c + 15) / 16 * 16; }

// C3k2(cin, cout, n11, c3k, e): C2f skeleton whose inner blocks are C3k(c, c, 2) or Bottleneck(c, c, e = 0.5).
int Model::add_c3k2(const std::string& p, int cin, int cout, bool c3k, double e) {
    const int c = (int)(cout * e);
    int r;
    if ((r = add_conv(p + ".cv1", cin, 2 * c, 1, true))) return r;
    if ((r = add_conv(p + ".cv2", (2 + n11) * c, cout, 1, true))) return r;
    for (int i = 0; i < n11; ++i) {
        const std::string m = p + ".m." + std::to_string(i);
        if (c3k) {
            const int c_ = c / 2;
            if ((r = add_conv(m + ".cv1", c, c_, 1, true))) return r;
            if ((r = add_conv(m + ".cv2", c, c_, 1, true))) return r;
            if ((r = add_conv(m + ".cv3", 2 * c_, c, 1, true))) return r;
            for (int j = 0; j < 2; ++j) {
                const std::string b = m + ".m." + std::to_string(j);
                if ((r = add_conv(b + ".cv1", c_, c_, 3, true))) return r;
                if ((r = add_conv(b + ".cv2", c_, c_, 3, true))) return r;
            }
        } else {
            if ((r = add_conv(m + ".cv1", c, c / 2, 3, true))) return r;
            if ((r = add_conv(m + ".cv2", c / 2, c, 3, true, pad16(c / 2)))) return r;
        }
    }
    return CY_OK;
}

int Model::finalize11() {
    int r;
#define TRY(x) if ((r = (x))) return r
    TRY(add_conv("model.0", 3, w64, 3, true));
    TRY(add_conv("model.1", w64, w128, 3, true));
    TRY(add_c3k2("model.2", w128, w256, c3k11, 0.25));
    TRY(add_conv("model.3", w256, w256, 3, true));
    TRY(add_c3k2("model.4", w256, w512, c3k11, 0.25));
    TRY(add_conv("model.5", w512, w512, 3, true));
    TRY(add_c3k2("model.6", w512, w512, true, 0.5));
    TRY(add_conv("model.7", w512, w1024, 3, true));
    TRY(add_c3k2("model.8", w1024, w1024, true, 0.5));
    TRY(add_conv("model.9.cv1", w1024, w1024 / 2, 1, true));
    TRY(add_conv("model.9.cv2", w1024 * 2, w1024, 1, true));
    const int c = w1024 / 2, nh = c / 64;
    if (c % 64) return set_error(CY_ERR_INVALID, "C2PSA width %d is not a multiple of 64", c);
    TRY(add_conv("model.10.cv1", w1024, 2 * c, 1, true));
    TRY(add_conv("model.10.cv2", 2 * c, w1024, 1, true));
    for (int i = 0; i < n11; ++i) {
        const std::string m = "model.10.m." + std::to_string(i);
        TRY(add_conv(m + ".attn.qkv", c, c + 2 * nh * 32, 1, true));
        TRY(add_dwconv(m + ".attn.pe", c));
        TRY(add_conv(m + ".attn.proj", c, c, 1, true));
        TRY(add_conv(m + ".ffn.0", c, 2 * c, 1, true));
        TRY(add_conv(m + ".ffn.1", 2 * c, c, 1, true));
    }
    TRY(add_c3k2("model.13", w1024 + w512, w512, c3k11, 0.5));
    TRY(add_c3k2("model.16", w512 + w512, w256, c3k11, 0.5));
    TRY(add_conv("model.17", w256, w256, 3, true));
    TRY(add_c3k2("model.19", w256 + w512, w512, c3k11, 0.5));
    TRY(add_conv("model.20", w512, w512, 3, true));
    TRY(add_c3k2("model.22", w512 + w1024, w1024, true, 0.5));
    const int cl[3] = {w256, w512, w1024};
    for (int l = 0; l < 3; ++l) {
        const std::string b = "model.23.cv2." + std::to_string(l), cs = "model.23.cv3." + std::to_string(l);
        TRY(add_conv(b + ".0", cl[l], cb, 3, true));
        TRY(add_conv(b + ".1", cb, cb, 3, true));
        TRY(add_conv(b + ".2", cb, 64, 1, false));
        TRY(add_dwconv(cs + ".0.0", cl[l]));
        TRY(add_conv(cs + ".0.1", cl[l], cc, 1, true));
        TRY(add_dwconv(cs + ".1.0", cc));
        TRY(add_conv(cs + ".1.1", cc, cc, 1, true));
        TRY(add_conv(cs + ".2", cc, nc, 1, false));
    }
#undef TRY
    raw.clear();
    finalized = true;
    return CY_OK;
}

int Model::finalize() {
    if (family == 11) return finalize11();
    int r;
#define TRY(x) if ((r = (x))) return r
    TRY(add_conv("model.0", 3, c1, 3, true));
    TRY(add_conv("model.1", c1, c2, 3, true));
    TRY(add_c2f("model.2", c2, c2, n2));
    TRY(add_conv("model.3", c2, c3, 3, true));
    TRY(add_c2f("model.4", c3, c3, n4));
    TRY(add_conv("model.5", c3, c4, 3, true));
    TRY(add_c2f("model.6", c4, c4, n6));
    TRY(add_conv("model.7", c4, c5, 3, true));
    TRY(add_c2f("model.8", c5, c5, n8));
    TRY(add_conv("model.9.cv1", c5, c5 / 2, 1, true));
    TRY(add_conv("model.9.cv2", c5 * 2, c5, 1, true));
    TRY(add_c2f("model.12", c5 + c4, c4, nh));
    TRY(add_c2f("model.15", c4 + c3, c3, nh));
    TRY(add_conv("model.16", c3, c3, 3, true));
    TRY(add_c2f("model.18", c3 + c4, c4, nh));
    TRY(add_conv("model.19", c4, c4, 3, true));
    TRY(add_c2f("model.21", c4 + c5, c5, nh));
    const int cl[3] = {c3, c4, c5};
    for (int l = 0; l < 3; ++l) {
        const std::string b = "model.22.cv2." + std::to_string(l), c = "model.22.cv3." + std::to_string(l);
        TRY(add_conv(b + ".0", cl[l], cb, 3, true));
        TRY(add_conv(b + ".1", cb, cb, 3, true));
        TRY(add_conv(b + ".2", cb, 64, 1, false));
        TRY(add_conv(c + ".0", cl[l], cc, 3, true));
        TRY(add_conv(c + ".1", cc, cc, 3, true));
        TRY(add_conv(c + ".2", cc, nc, 1, false));
    }
#undef TRY
    raw.clear();
    finalized = true;
    return CY_OK;
}

Model::~Model() {
    for (auto& kv : convs) {
        cudaFree(kv.second.w);
        cudaFree(kv.second.b);
    }
    for (auto& kv : plans) delete kv.second;
}

Plan::~Plan() {
    for (void* p : allocs) cudaFree(p);
}

// ------------------------------------------------------------------------------------------ plan building

struct PlanBuilder {
    Model& m;
    Plan& pl;
    int B;
    char err[256];
    PlanBuilder(Model& m_, Plan& p_, int B_) : m(m_), pl(p_), B(B_) { err[0] = 0; }

    Buf alloc(int H, int W, int C, bool f32 = false) {
        Buf b;
        b.H = H; b.W = W; b.C = C;
        const size_t bytes = (size_t)B * H * W * C * (f32 ? 4 : 2);
        if (cudaMalloc(&b.p, bytes) != cudaSuccess) {
            snprintf(err, sizeof(err), "cudaMalloc of %zu bytes failed", bytes);
            b.p = nullptr;
            return b;
        }
        pl.allocs.push_back(b.p);
        pl.bytes += bytes;
        return b;
    }
    // conv from slice of `in` to slice of `out`
    int conv(const std::string& name, const Buf& in, int in_off, const Buf& out, int out_off, int stride,
             bool act = true, const Buf* res = nullptr, int res_off = 0, bool out_f32 = false) {
        auto it = m.convs.find(name);
        if (it == m.convs.end()) {
            snprintf(err, sizeof(err), "conv %s not loaded", name.c_str());
            return -1;
        }
        const ConvW& w = it->second;
        Op op;
        op.type = Op::CONV;
        op.name = name;
        ConvDesc d;
        d.in = (const __nv_bfloat16*)in.p; d.in_ctot = in.C; d.in_coff = in_off; d.cin = w.cin;
        d.B = B; d.Hin = in.H; d.Win = in.W; d.ksize = w.k; d.stride = stride;
        d.w = (const __nv_bfloat16*)w.w; d.cout_pad = w.cout_pad; d.bias = w.b; d.cout = w.cout;
        d.out = out.p; d.out_ctot = out.C; d.out_coff = out_off; d.out_f32 = out_f32 ? 1 : 0;
        d.res = res ? (const __nv_bfloat16*)res->p : nullptr; d.res_ctot = res ? res->C : 0; d.res_coff = res_off;
        d.act = act ? 1 : 0;
        d.f16 = m.f16;
        if (conv_make_plan(d, &op.conv, err, sizeof(err)) != 0) return -1;
        pl.flops += op.conv.flops;
        pl.ops.push_back(op);
        return 0;
    }
    // C2f: in slice -> out slice
    int c2f(const std::string& p, const Buf& in, int cout, int n, bool shortcut, const Buf& out, int out_off) {
        const int c = cout / 2;
        Buf cat = alloc(in.H, in.W, (2 + n) * c);
        Buf tmp = alloc(in.H, in.W, c);
        if (!cat.p || !tmp.p) return -1;
        if (conv(p + ".cv1", in, 0, cat, 0, 1)) return -1;
        for (int i = 0; i < n; ++i) {
            const std::string mm = p + ".m." + std::to_string(i);
            if (conv(mm + ".cv1", cat, (1 + i) * c, tmp, 0, 1)) return -1;
            if (conv(mm + ".cv2", tmp, 0, cat, (2 + i) * c, 1, true, shortcut ? &cat : nullptr, (1 + i) * c)) return -1;
        }
        return conv(p + ".cv2", cat, 0, out, out_off, 1);
    }
    // zero-filled buffer (padding channels of 8-channel intermediates must read as 0)
    Buf alloc_zero(int H, int W, int C) {
        Buf b = alloc(H, W, C);
        if (b.p && cudaMemset(b.p, 0, (size_t)B * H * W * C * 2) != cudaSuccess) {
            snprintf(err, sizeof(err), "cudaMemset failed");
            b.p = nullptr;
        }
        return b;
    }
    int dwconv(const std::string& name, const Buf& in, int in_off, int gs, int gst, const Buf& out, int out_off, int C,
               bool act, const Buf* res = nullptr, int res_off = 0) {
        auto it = m.convs.find(name);
        if (it == m.convs.end() || it->second.cout != C) {
            snprintf(err, sizeof(err), "depthwise conv %s not loaded", name.c_str());
            return -1;
        }
        if (C % 8 || gs % 8 || in_off % 8 || out_off % 8 || C / 8 > 256) {
            snprintf(err, sizeof(err), "depthwise conv %s: unsupported channel layout", name.c_str());
            return -1;
        }
        Op op;
        op.type = Op::DWCONV; op.name = name;
        op.in = in; op.in_off = in_off; op.out = out; op.out_off = out_off; op.C = C;
        op.dw_w = (const float*)it->second.w; op.dw_b = it->second.b;
        op.act = act ? 1 : 0; op.gs = gs; op.gst = gst;
        if (res) { op.res = *res; op.res_off = res_off; }
        op.conv.flops = 2.0 * B * in.H * in.W * C * 9;
        pl.ops.push_back(op);
        return 0;
    }
    // yolo11 C3k2 (see Model::add_c3k2): in (all channels) -> out slice
    int c3k2(const std::string& p, const Buf& in, int cout, bool c3k, double e, const Buf& out, int out_off) {
        const int n = m.n11, c = (int)(cout * e);
        Buf cat = alloc(in.H, in.W, (2 + n) * c);
        if (!cat.p) return -1;
        if (conv(p + ".cv1", in, 0, cat, 0, 1)) return -1;
        for (int i = 0; i < n; ++i) {
            const std::string mm = p + ".m." + std::to_string(i);
            const int src = (1 + i) * c, dst = (2 + i) * c;
            if (c3k) {   // C3k: cv3(cat(m(cv1(x)), cv2(x))), m = two Bottleneck(c/2, c/2) with shortcut
                const int c_ = c / 2;
                Buf cat2 = alloc(in.H, in.W, 2 * c_), a0 = alloc(in.H, in.W, c_), a1 = alloc(in.H, in.W, c_),
                    t = alloc(in.H, in.W, c_);
                if (!cat2.p || !a0.p || !a1.p || !t.p) return -1;
                if (conv(mm + ".cv1", cat, src, a0, 0, 1)) return -1;
                if (conv(mm + ".cv2", cat, src, cat2, c_, 1)) return -1;
                if (conv(mm + ".m.0.cv1", a0, 0, t, 0, 1)) return -1;
                if (conv(mm + ".m.0.cv2", t, 0, a1, 0, 1, true, &a0, 0)) return -1;
                if (conv(mm + ".m.1.cv1", a1, 0, t, 0, 1)) return -1;
                if (conv(mm + ".m.1.cv2", t, 0, cat2, 0, 1, true, &a1, 0)) return -1;
                if (conv(mm + ".cv3", cat2, 0, cat, dst, 1)) return -1;
            } else {     // Bottleneck(c, c, e = 0.5) with shortcut; the c/2 intermediate is padded to 16 channels
                const int ch = c / 2;
                Buf t = (ch % 16) ? alloc_zero(in.H, in.W, (ch + 15) / 16 * 16) : alloc(in.H, in.W, ch);
                if (!t.p) return -1;
                if (conv(mm + ".cv1", cat, src, t, 0, 1)) return -1;
                if (conv(mm + ".cv2", t, 0, cat, dst, 1, true, &cat, src)) return -1;
            }
        }
        return conv(p + ".cv2", cat, 0, out, out_off, 1);
    }
    // yolo11 C2PSA(c1, c1, n, e = 0.5): cv2(cat(a, PSA^n(b))), (a, b) = split(cv1(x))
    int c2psa(const std::string& p, const Buf& in, const Buf& out, int out_off) {
        const int c = in.C / 2, nh = c / 64, H = in.H, W = in.W;
        Buf y = alloc(H, W, 2 * c), qkv = alloc(H, W, nh * 128), ao = alloc(H, W, c), pe = alloc(H, W, c),
            t1 = alloc(H, W, c), f = alloc(H, W, 2 * c);
        if (!y.p || !qkv.p || !ao.p || !pe.p || !t1.p || !f.p) return -1;
        if (conv(p + ".cv1", in, 0, y, 0, 1)) return -1;
        for (int i = 0; i < m.n11; ++i) {
            const std::string mm = p + ".m." + std::to_string(i);
            if (conv(mm + ".attn.qkv", y, c, qkv, 0, 1, false)) return -1;
            Op op;
            op.type = Op::ATTN; op.name = mm + ".attn";
            op.in = qkv; op.out = ao; op.out_off = 0; op.nh = nh;
            op.conv.flops = 2.0 * B * nh * (double)(H * W) * (H * W) * (32 + 64);
            pl.ops.push_back(op);
            // pe(v) + attention output: v = channels [64, 128) of every head's 128-channel group of qkv
            if (dwconv(mm + ".attn.pe", qkv, 64, 64, 128, pe, 0, c, false, &ao, 0)) return -1;
            if (conv(mm + ".attn.proj", pe, 0, t1, 0, 1, false, &y, c)) return -1;       // b + attn(b)
            if (conv(mm + ".ffn.0", t1, 0, f, 0, 1)) return -1;
            if (conv(mm + ".ffn.1", f, 0, y, c, 1, false, &t1, 0)) return -1;            // + ffn(.) back into y[:, c:]
        }
        return conv(p + ".cv2", y, 0, out, out_off, 1);
    }
};

// Plans (activation buffers + tensor maps) are cached per (batch, extent).  A large batch of YOLOv8l holds 148 MB per
// tile, so when building a new plan runs out of device memory the other cached plans are dropped and the build is
// retried once (the caller synchronises: plans of earlier batches are not in use once their results were consumed).
int Model::get_plan(int B, int Sh, int Sw, Plan** out) {
    if (!finalized) return set_error(CY_ERR_STATE, "model not finalized");
    if (Sh % 32 || Sw % 32 || Sh <= 0 || Sw <= 0) return set_error(CY_ERR_INVALID, "input extent must be a multiple of 32");
    const auto key = std::make_tuple(B, Sh, Sw);
    auto it = plans.find(key);
    if (it != plans.end()) {
        *out = it->second;
        return CY_OK;
    }
    int rc = build_plan(B, Sh, Sw, out);
    if (rc != CY_OK && !plans.empty()) {
        cudaDeviceSynchronize();
        for (auto& kv : plans) delete kv.second;
        plans.clear();
        cudaGetLastError();
        rc = build_plan(B, Sh, Sw, out);
    }
    return rc;
}

int Model::build_plan(int B, int Sh, int Sw, Plan** out) {
    if (family == 11) return build_plan11(B, Sh, Sw, out);
    const auto key = std::make_tuple(B, Sh, Sw);
    Plan* pl = new Plan();
    pl->B = B; pl->Sh = Sh; pl->Sw = Sw;
    PlanBuilder pb(*this, *pl, B);
#define PB(x)                                                            \
    if ((x)) {                                                           \
        int rc = set_error(CY_ERR_INVALID, "plan build failed: %s", pb.err); \
        delete pl;                                                       \
        return rc;                                                       \
    }
    const int H2 = Sh / 2, W2 = Sw / 2, H4 = Sh / 4, W4 = Sw / 4, H8 = Sh / 8, W8 = Sw / 8, H16 = Sh / 16,
              W16 = Sw / 16, H32 = Sh / 32, W32 = Sw / 32;
    Buf x0 = pb.alloc(H2, W2, c1);
    Buf x1 = pb.alloc(H4, W4, c2);
    Buf x2 = pb.alloc(H4, W4, c2);
    Buf x3 = pb.alloc(H8, W8, c3);
    Buf cat14 = pb.alloc(H8, W8, c4 + c3);   // [up(x12) | x4]
    Buf x5 = pb.alloc(H16, W16, c4);
    Buf cat11 = pb.alloc(H16, W16, c5 + c4); // [up(x9) | x6]
    Buf x7 = pb.alloc(H32, W32, c5);
    Buf x8 = pb.alloc(H32, W32, c5);
    Buf sp = pb.alloc(H32, W32, 2 * c5);     // SPPF concat: 4 x c5/2
    Buf cat20 = pb.alloc(H32, W32, c4 + c5); // [x19 | x9]
    Buf cat17 = pb.alloc(H16, W16, c3 + c4); // [x16 | x12]
    Buf x15 = pb.alloc(H8, W8, c3);
    Buf x18 = pb.alloc(H16, W16, c4);
    Buf x21 = pb.alloc(H32, W32, c5);
    PB(!x0.p || !x1.p || !x2.p || !x3.p || !cat14.p || !x5.p || !cat11.p || !x7.p || !x8.p || !sp.p || !cat20.p ||
       !cat17.p || !x15.p || !x18.p || !x21.p);
    {   // model.0 stem: fused 3x3 stride-2 conv + bias + SiLU straight from the model input (stem_conv_kernel)
        auto it0 = convs.find("model.0");
        if (it0 == convs.end() || it0->second.cout != c1) snprintf(pb.err, sizeof(pb.err), "stem model.0 not loaded");
        PB(it0 == convs.end() || it0->second.cout != c1);
        Op op;
        op.type = Op::STEM; op.name = "model.0";
        op.out = x0;
        op.conv.flops = 2.0 * B * H2 * W2 * c1 * 27;
        pl->flops += op.conv.flops;
        pl->ops.push_back(op);
    }
    Buf inb; inb.p = nullptr;
    PB(pb.conv("model.1", x0, 0, x1, 0, 2));
    PB(pb.c2f("model.2", x1, c2, n2, true, x2, 0));
    PB(pb.conv("model.3", x2, 0, x3, 0, 2));
    PB(pb.c2f("model.4", x3, c3, n4, true, cat14, c4));
    {   // x4 lives in cat14[:, c4:]; model.5 reads that slice
        Buf x4v = cat14;
        PB(pb.conv("model.5", x4v, c4, x5, 0, 2));
    }
    PB(pb.c2f("model.6", x5, c4, n6, true, cat11, c5));
    PB(pb.conv("model.7", cat11, c5, x7, 0, 2));
    PB(pb.c2f("model.8", x7, c5, n8, true, x8, 0));
    // SPPF
    PB(pb.conv("model.9.cv1", x8, 0, sp, 0, 1));
    if ((size_t)3 * H32 * W32 * 64 <= (size_t)kSppfMaxSmem && (c5 / 2) % 32 == 0) {
        Op op;   // the three max-pools of SPPF in one kernel (map resident in shared memory)
        op.type = Op::SPPF_POOL; op.name = "model.9.m";
        op.in = sp; op.out = sp; op.C = c5 / 2;
        pl->ops.push_back(op);
    } else {
        for (int i = 0; i < 3; ++i) {
            Op op;
            op.type = Op::MAXPOOL; op.name = "model.9.m";
            op.in = sp; op.in_off = i * (c5 / 2); op.out = sp; op.out_off = (i + 1) * (c5 / 2); op.C = c5 / 2;
            pl->ops.push_back(op);
        }
    }
    PB(pb.conv("model.9.cv2", sp, 0, cat20, c4, 1));  // x9 -> cat20[:, c4:]
    {   // upsample x9 -> cat11[:, :c5]
        Op op;
        op.type = Op::UPSAMPLE; op.name = "model.10";
        op.in = cat20; op.in_off = c4; op.out = cat11; op.out_off = 0; op.C = c5;
        pl->ops.push_back(op);
    }
    PB(pb.c2f("model.12", cat11, c4, nh, false, cat17, c3));  // x12 -> cat17[:, c3:]
    {   // upsample x12 -> cat14[:, :c4]
        Op op;
        op.type = Op::UPSAMPLE; op.name = "model.13";
        op.in = cat17; op.in_off = c3; op.out = cat14; op.out_off = 0; op.C = c4;
        pl->ops.push_back(op);
    }
    PB(pb.c2f("model.15", cat14, c3, nh, false, x15, 0));
    PB(pb.conv("model.16", x15, 0, cat17, 0, 2));
    PB(pb.c2f("model.18", cat17, c4, nh, false, x18, 0));
    PB(pb.conv("model.19", x18, 0, cat20, 0, 2));
    PB(pb.c2f("model.21", cat20, c5, nh, false, x21, 0));
    // Detect
    const Buf* feats[3] = {&x15, &x18, &x21};
    for (int l = 0; l < 3; ++l) {
        const Buf& f = *feats[l];
        Buf tb1 = pb.alloc(f.H, f.W, cb), tb2 = pb.alloc(f.H, f.W, cb);
        Buf tc1 = pb.alloc(f.H, f.W, cc), tc2 = pb.alloc(f.H, f.W, cc);
        Buf head = pb.alloc(f.H, f.W, kHeadC, true);
        PB(!tb1.p || !tb2.p || !tc1.p || !tc2.p || !head.p);
        const std::string b = "model.22.cv2." + std::to_string(l), c = "model.22.cv3." + std::to_string(l);
        PB(pb.conv(b + ".0", f, 0, tb1, 0, 1));
        PB(pb.conv(b + ".1", tb1, 0, tb2, 0, 1));
        PB(pb.conv(b + ".2", tb2, 0, head, 0, 1, false, nullptr, 0, true));
        PB(pb.conv(c + ".0", f, 0, tc1, 0, 1));
        PB(pb.conv(c + ".1", tc1, 0, tc2, 0, 1));
        PB(pb.conv(c + ".2", tc2, 0, head, 64, 1, false, nullptr, 0, true));
        pl->head[l] = head;
    }
#undef PB
    plans[key] = pl;
    *out = pl;
    return CY_OK;
}

// yolo11 (cfg/models/11/yolo11.yaml): 0-1 Conv, 2 C3k2, 3 Conv, 4 C3k2, 5 Conv, 6 C3k2, 7 Conv, 8 C3k2, 9 SPPF,
// 10 C2PSA, 11-13 up + cat(6) + C3k2, 14-16 up + cat(4) + C3k2, 17-19 Conv + cat(13) + C3k2, 20-22 Conv + cat(10) +
// C3k2, 23 Detect(16, 19, 22).  Concats are channel slices of shared buffers like in the yolov8 plan.
int Model::build_plan11(int B, int Sh, int Sw, Plan** out) {
    const auto key = std::make_tuple(B, Sh, Sw);
    Plan* pl = new Plan();
    pl->B = B; pl->Sh = Sh; pl->Sw = Sw;
    PlanBuilder pb(*this, *pl, B);
#define PB(x)                                                            \
    if ((x)) {                                                           \
        int rc = set_error(CY_ERR_INVALID, "plan build failed: %s", pb.err); \
        delete pl;                                                       \
        return rc;                                                       \
    }
    const int H2 = Sh / 2, W2 = Sw / 2, H4 = Sh / 4, W4 = Sw / 4, H8 = Sh / 8, W8 = Sw / 8, H16 = Sh / 16,
              W16 = Sw / 16, H32 = Sh / 32, W32 = Sw / 32;
    Buf x0 = pb.alloc(H2, W2, w64), x1 = pb.alloc(H4, W4, w128), x2 = pb.alloc(H4, W4, w256);
    Buf x3 = pb.alloc(H8, W8, w256);
    Buf cat15 = pb.alloc(H8, W8, w512 + w512);     // [up(x13) | x4]
    Buf x5 = pb.alloc(H16, W16, w512);
    Buf cat12 = pb.alloc(H16, W16, w1024 + w512);  // [up(x10) | x6]
    Buf x7 = pb.alloc(H32, W32, w1024), x8 = pb.alloc(H32, W32, w1024);
    Buf sp = pb.alloc(H32, W32, 2 * w1024), x9 = pb.alloc(H32, W32, w1024);
    Buf cat21 = pb.alloc(H32, W32, w512 + w1024);  // [x20 | x10]
    Buf cat18 = pb.alloc(H16, W16, w256 + w512);   // [x17 | x13]
    Buf x16 = pb.alloc(H8, W8, w256), x19 = pb.alloc(H16, W16, w512), x22 = pb.alloc(H32, W32, w1024);
    PB(!x0.p || !x1.p || !x2.p || !x3.p || !cat15.p || !x5.p || !cat12.p || !x7.p || !x8.p || !sp.p || !x9.p ||
       !cat21.p || !cat18.p || !x16.p || !x19.p || !x22.p);
    {
        auto it0 = convs.find("model.0");
        if (it0 == convs.end() || it0->second.cout != w64) snprintf(pb.err, sizeof(pb.err), "stem model.0 not loaded");
        PB(it0 == convs.end() || it0->second.cout != w64);
        Op op;
        op.type = Op::STEM; op.name = "model.0";
        op.out = x0;
        op.conv.flops = 2.0 * B * H2 * W2 * w64 * 27;
        pl->flops += op.conv.flops;
        pl->ops.push_back(op);
    }
    PB(pb.conv("model.1", x0, 0, x1, 0, 2));
    PB(pb.c3k2("model.2", x1, w256, c3k11, 0.25, x2, 0));
    PB(pb.conv("model.3", x2, 0, x3, 0, 2));
    PB(pb.c3k2("model.4", x3, w512, c3k11, 0.25, cat15, w512));      // x4 -> cat15[:, w512:]
    PB(pb.conv("model.5", cat15, w512, x5, 0, 2));
    PB(pb.c3k2("model.6", x5, w512, true, 0.5, cat12, w1024));       // x6 -> cat12[:, w1024:]
    PB(pb.conv("model.7", cat12, w1024, x7, 0, 2));
    PB(pb.c3k2("model.8", x7, w1024, true, 0.5, x8, 0));
    PB(pb.conv("model.9.cv1", x8, 0, sp, 0, 1));
    if ((size_t)3 * H32 * W32 * 64 <= (size_t)kSppfMaxSmem && (w1024 / 2) % 32 == 0) {
        Op op;
        op.type = Op::SPPF_POOL; op.name = "model.9.m";
        op.in = sp; op.out = sp; op.C = w1024 / 2;
        pl->ops.push_back(op);
    } else {
        for (int i = 0; i < 3; ++i) {
            Op op;
            op.type = Op::MAXPOOL; op.name = "model.9.m";
            op.in = sp; op.in_off = i * (w1024 / 2); op.out = sp; op.out_off = (i + 1) * (w1024 / 2); op.C = w1024 / 2;
            pl->ops.push_back(op);
        }
    }
    PB(pb.conv("model.9.cv2", sp, 0, x9, 0, 1));
    PB(pb.c2psa("model.10", x9, cat21, w512));                        // x10 -> cat21[:, w512:]
    {
        Op op;
        op.type = Op::UPSAMPLE; op.name = "model.11";
        op.in = cat21; op.in_off = w512; op.out = cat12; op.out_off = 0; op.C = w1024;
        pl->ops.push_back(op);
    }
    PB(pb.c3k2("model.13", cat12, w512, c3k11, 0.5, cat18, w256));    // x13 -> cat18[:, w256:]
    {
        Op op;
        op.type = Op::UPSAMPLE; op.name = "model.14";
        op.in = cat18; op.in_off = w256; op.out = cat15; op.out_off = 0; op.C = w512;
        pl->ops.push_back(op);
    }
    PB(pb.c3k2("model.16", cat15, w256, c3k11, 0.5, x16, 0));
    PB(pb.conv("model.17", x16, 0, cat18, 0, 2));
    PB(pb.c3k2("model.19", cat18, w512, c3k11, 0.5, x19, 0));
    PB(pb.conv("model.20", x19, 0, cat21, 0, 2));
    PB(pb.c3k2("model.22", cat21, w1024, true, 0.5, x22, 0));
    const Buf* feats[3] = {&x16, &x19, &x22};
    for (int l = 0; l < 3; ++l) {
        const Buf& f = *feats[l];
        Buf tb1 = pb.alloc(f.H, f.W, cb), tb2 = pb.alloc(f.H, f.W, cb);
        Buf d1 = pb.alloc(f.H, f.W, f.C), tc1 = pb.alloc(f.H, f.W, cc), d2 = pb.alloc(f.H, f.W, cc),
            tc2 = pb.alloc(f.H, f.W, cc);
        Buf head = pb.alloc(f.H, f.W, kHeadC, true);
        PB(!tb1.p || !tb2.p || !d1.p || !tc1.p || !d2.p || !tc2.p || !head.p);
        const std::string b = "model.23.cv2." + std::to_string(l), c = "model.23.cv3." + std::to_string(l);
        PB(pb.conv(b + ".0", f, 0, tb1, 0, 1));
        PB(pb.conv(b + ".1", tb1, 0, tb2, 0, 1));
        PB(pb.conv(b + ".2", tb2, 0, head, 0, 1, false, nullptr, 0, true));
        PB(pb.dwconv(c + ".0.0", f, 0, f.C, 0, d1, 0, f.C, true));
        PB(pb.conv(c + ".0.1", d1, 0, tc1, 0, 1));
        PB(pb.dwconv(c + ".1.0", tc1, 0, cc, 0, d2, 0, cc, true));
        PB(pb.conv(c + ".1.1", d2, 0, tc2, 0, 1));
        PB(pb.conv(c + ".2", tc2, 0, head, 64, 1, false, nullptr, 0, true));
        pl->head[l] = head;
    }
#undef PB
    plans[key] = pl;
    *out = pl;
    return CY_OK;
}

int Model::launch_op(const Op& op, const void* in, int B, int Sh, int Sw, cudaStream_t st) {
    switch (op.type) {
        case Op::STEM: {
            const ConvW& w = convs.at("model.0");
            const char* ex = getenv("CY_CONV_SILU_EXACT");   // same switch as the conv kernel (1: tanh form, 2: ex2 + rcp)
            const int act = (ex && atoi(ex)) ? 2 : 1;
            const int e = launch_stem_any(in, w, op.out.p, B, Sh, Sw, act, f16, st);
            if (e) return set_error(CY_ERR_CUDA, "stem launch failed: %s", cudaGetErrorString((cudaError_t)e));
            break;
        }
        case Op::CONV: {
            int e = conv_launch(op.conv, st);
            if (e) return set_error(CY_ERR_CUDA, "conv %s launch failed: %s", op.name.c_str(),
                                    cudaGetErrorString((cudaError_t)e));
            break;
        }
        case Op::MAXPOOL: {
            const long long total = (long long)B * op.in.H * op.in.W * (op.C / 8);
            maxpool5_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                (const __nv_bfloat16*)op.in.p, op.in.C, op.in_off, (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, B,
                op.in.H, op.in.W, op.C);
            break;
        }
        case Op::SPPF_POOL: {
            const int smem = 3 * op.in.H * op.in.W * 64;
            static std::atomic<unsigned long long> attr_done{0};
            if (first_use_on_device(attr_done)) {
                cudaError_t e = cudaFuncSetAttribute(sppf_pool3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSppfMaxSmem);
                if (e != cudaSuccess) return set_error(CY_ERR_CUDA, "sppf_pool3 attribute: %s", cudaGetErrorString(e));
            }
            sppf_pool3_kernel<<<dim3((unsigned)(op.C / 32), (unsigned)B), 256, smem, st>>>(
                (__nv_bfloat16*)op.in.p, op.in.C, op.C, op.in.H, op.in.W);
            break;
        }
        case Op::DWCONV: {
            const int groups = op.C / 8, ppb = std::max(1, 256 / groups);
            const char* ex = getenv("CY_CONV_SILU_EXACT");
            const int act = op.act ? ((ex && atoi(ex)) ? 2 : 1) : 0;
            dwconv3x3_kernel<<<dim3((unsigned)((op.in.W + 4 * ppb - 1) / (4 * ppb)), (unsigned)op.in.H, (unsigned)B), groups * ppb, 0, st>>>(
                (const __nv_bfloat16*)op.in.p, op.in.C, op.in_off, op.gs, op.gst, op.dw_w, op.dw_b,
                (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, (const __nv_bfloat16*)op.res.p, op.res.C, op.res_off,
                op.in.H, op.in.W, op.C, act, f16);
            break;
        }
        case Op::ATTN: {
            const int N = op.in.H * op.in.W;
            const size_t kv = (size_t)N * (17 + 32) * 4;
            static std::atomic<unsigned long long> attr_done{0};
            if (first_use_on_device(attr_done)) {
                cudaError_t e1 = cudaFuncSetAttribute(attention_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                cudaError_t e2 = cudaFuncSetAttribute(attention_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                if (e1 != cudaSuccess || e2 != cudaSuccess) return set_error(CY_ERR_CUDA, "attention attribute failed");
            }
            const dim3 grid((unsigned)((N + 63) / 64), (unsigned)op.nh, (unsigned)B);
            const int Npad = (N + 63) / 64 * 64;
            const size_t mma_smem = (size_t)Npad * (80 + 144);
            const char* simple = getenv("CY_ATTN_SIMPLE");
            if (!(simple && atoi(simple)) && mma_smem <= 227 * 1024) {
                static std::atomic<unsigned long long> mma_attr{0};
                if (first_use_on_device(mma_attr)) {
                    if (cudaFuncSetAttribute(attention_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
                        cudaFuncSetAttribute(attention_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                        return set_error(CY_ERR_CUDA, "attention attribute failed");
                }
                if (f16)
                    attention_mma_kernel<true><<<grid, 128, mma_smem, st>>>((const __nv_bfloat16*)op.in.p, op.in.C,
                                                                           (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, N, Npad);
                else
                    attention_mma_kernel<false><<<grid, 128, mma_smem, st>>>((const __nv_bfloat16*)op.in.p, op.in.C,
                                                                            (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, N, Npad);
            } else if (kv + (size_t)8 * N * 4 <= 200 * 1024)
                attention_kernel<8><<<grid, 256, kv + (size_t)8 * N * 4, st>>>(
                    (const __nv_bfloat16*)op.in.p, op.in.C, (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, N, f16);
            else if (kv + (size_t)4 * N * 4 <= 220 * 1024)
                attention_kernel<4><<<grid, 128, kv + (size_t)4 * N * 4, st>>>(
                    (const __nv_bfloat16*)op.in.p, op.in.C, (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, N, f16);
            else
                return set_error(CY_ERR_INVALID, "attention over %d positions does not fit shared memory (imgsz > 1024)", N);
            break;
        }
        case Op::UPSAMPLE: {
            const long long total = (long long)B * op.in.H * op.in.W * (op.C / 8);
            upsample2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                (const __nv_bfloat16*)op.in.p, op.in.C, op.in_off, (__nv_bfloat16*)op.out.p, op.out.C, op.out_off, B,
                op.in.H, op.in.W, op.C);
            break;
        }
    }
    return CY_OK;
}

int Model::forward(const void* in, int B, int Sh, int Sw, cudaStream_t st, Plan** plan_out) {
    Plan* pl;
    int r = get_plan(B, Sh, Sw, &pl);
    if (r) return r;
    for (const Op& op : pl->ops)
        if ((r = launch_op(op, in, B, Sh, Sw, st))) return r;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(CY_ERR_CUDA, "forward launch failed: %s", cudaGetErrorString(e));
    if (plan_out) *plan_out = pl;
    return CY_OK;
}

// Per-op device timing (CUDA events on `st`), used by bench.py for the per-layer tensor-pipe report.
int Model::profile(const void* in, int B, int Sh, int Sw, int cap, const char** names, float* ms, double* flops,
                   int* nops, cudaStream_t st) {
    Plan* pl;
    int r = forward(in, B, Sh, Sw, st, &pl);  // warm-up (also builds the plan)
    if (r) return r;
    CY_CUDA_CHECK(cudaStreamSynchronize(st));
    const int n = (int)pl->ops.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CY_CUDA_CHECK(cudaEventCreate(&e));
    CY_CUDA_CHECK(cudaEventRecord(ev[0], st));
    for (int i = 0; i < n; ++i) {
        if ((r = launch_op(pl->ops[i], in, B, Sh, Sw, st))) return r;
        CY_CUDA_CHECK(cudaEventRecord(ev[i + 1], st));
    }
    CY_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int i = 0; i < n && i < cap; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
        if (ms) ms[i] = t;
        if (names) names[i] = pl->ops[i].name.c_str();
        if (flops) flops[i] = (pl->ops[i].type == Op::CONV || pl->ops[i].type == Op::STEM) ? pl->ops[i].conv.flops : 0.0;   // tensor-pipe ops only
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (nops) *nops = n;
    return CY_OK;
}

}  // namespace cy
