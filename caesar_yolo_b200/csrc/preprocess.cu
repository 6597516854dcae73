// Preprocessing chain on the device: tile cut-out (FITS byte order, NaN -> 0), the stage list of
// caesar_yolo/preprocessing.py, and the ultralytics predictor preprocess (letterbox resize, channel flip, /255).
//
// Reference: DataPreprocessor (caesar_yolo/preprocessing.py:47-67), BkgSubtractor (:591-658), SigmaClipShifter (:664-717),
// SigmaClipper (:723-771), ChanResizer (:1077-1133), ZScaleTransformer (:934-971), Chan3Trasformer (:1020-1072),
// HistEqualizer (:977-1012), MinMaxNormalizer (:75-111), AbsMinMaxNormalizer (:116-146), MaxScaler (:152-176),
// AbsMaxScaler (:182-226), ChanMaxScaler (:232-288), MinShifter (:294-327), Shifter (:333-363), Standardizer (:369-402),
// NegativeDataFixer (:408-440), LogStretcher (:480-538), BorderMasker (:544-586); stage order of scripts/run.py:272-293;
// Analyzer.predict front part (caesar_yolo/evaluation.py:146-176); astropy sigma_clip / ZScaleInterval, skimage
// equalize_hist and the ultralytics LetterBox semantics are restated in SURVEY.md App. A.1-A.4.
//
// Design.  Every stage is a monotone non-decreasing scalar map of the pixel value, and "0 means masked" is sticky.  So a
// channel is a short op list (scalars only) applied to the ORIGINAL pixel value in fp64; no intermediate image exists.
// What the stages need from the pixels are order statistics (medians of shrinking value intervals), ranks of thresholds,
// moment sums over value intervals and min / max -- none of which needs the full sorted order (the first version sorted
// every tile with a 4-pass radix sort).  Three kernels, the tile is read from HBM twice and nothing else of its size is
// written except the 16-bit model input:
//   pp_bucket_kernel: one CTA per tile.  A 1024-sample sort fixes a robust centre / scale of the tile; the pixels then
//       stream through shared memory in chunks of 16384 (TMA 2-D boxes of the mosaic, byte swap + NaN -> 0 on the
//       shared-memory read) and every chunk is counting-sorted into 1024 value bins (linear over +-8 robust sigma,
//       logarithmic tails; optionally split in "inside / outside the central box"): chunk-local bucketed values,
//       per-chunk bin offsets, and per-bin count / min / max / fp64 moment sums about a pivot.
//   pp_chain_kernel: one CTA per tile, runs the stage list on the bin tables.  A rank query is a scan of the 1024 bin
//       maxima + a count inside ONE bin; an order statistic is a gather + sort of ONE bin (cached); moment sums of a
//       clipped interval come from the per-bin sums in closed form (the maps are affine + clamp) + exact sums over
//       the few bins a boundary cuts.  zscale sorts its 1000 positional samples in shared memory and runs the <= 5 line
//       fits in fp64; histogram equalisation bins whole value bins at once.  Exact zeros (= masked) are tracked as value
//       intervals so masks match the reference bit for bit.
//   pp_final_kernel: evaluates the three final channel maps per pixel and, fused, does the letterbox bilinear resize
//       (cv2's double-precision source coordinates), pad 114, channel reversal, /255 into the bf16 / fp16 NHWC(4) model
//       input.  The fp32 HWC chain image is an optional parity output.
#include "common.h"
#include "half16.cuh"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace cy {

static constexpr int kPPThreads = 512;     // chain kernel: latency-bound, several CTAs (tiles) per SM
static constexpr int kPPWarps = kPPThreads / 32;
static constexpr int kBkThreads = 1024;    // bucket kernel: one full-width CTA per SM
static constexpr int kNB = 1024;           // logical value bins
static constexpr int kChunk = 16384;       // pixels per bucketing chunk (16 per thread)
static constexpr int kGather = 4096;       // largest bin that is gathered + sorted in shared memory
static constexpr int kMaxOps = 12;
static constexpr int kMaxZero = 8;

enum OpKind { OP_SUB = 1, OP_SHIFT = 2, OP_CLAMP = 3, OP_ZSCALE = 4, OP_HISTEQ = 5, OP_MINMAX = 6, OP_DIV = 7,
              OP_STD = 8, OP_LOG = 9 };

struct Chan {
    int nops;
    int hid;  // history id: channels with equal hid hold identical data
    int kind[kMaxOps];
    double p0[kMaxOps], p1[kMaxOps], p2[kMaxOps], p3[kMaxOps];
    int nz;  // zero (masked) sets, sorted + merged: rank ranges [z0, z1) of the all-pixels view = value intervals
    int z0[kMaxZero], z1[kMaxZero];
    float zx0[kMaxZero], zx1[kMaxZero];   // [zx0, zx1] of the ORIGINAL pixel value (closed)
};

// Compiled form of a channel's op list for the full passes: every op except HISTEQ / LOG is affine + clamp with a
// positive slope, so the composition collapses to v = clamp(a*x + b, l, h) (one optional HISTEQ in the middle splits it
// into A and B).  Differs from the sequential fp64 evaluation only by rounding (~1e-16 relative); exact-zero (= masked)
// decisions never come from it: they are the value intervals of the channel.
struct Comp {
    double a0, b0, l0, h0;
    double a1, b1, l1, h1;
    int has_he;
    int ok;                      // 0: not representable (non-finite / non-positive slope, LOG) -> interpreter
    int nzx;
    float zx0[kMaxZero], zx1[kMaxZero];
};

struct HistEq {  // skimage.exposure.equalize_hist tables (one histogram equalisation per chain)
    double edges[257];
    double cdf[256];
    double center[256];   // (edges[k] + edges[k+1]) / 2
    double slope[256];    // (cdf[k+1] - cdf[k]) / (center[k+1] - center[k]), k < 255: the division np.interp does
    double inv_step;      // 255 / (center[255] - center[0]) (0 when the range is empty): bracket guess of the fast lookup
};

// Final state of a tile's chain: what pp_final_kernel needs to evaluate the three channel maps per pixel.
struct TileFinal {
    Comp cc[3];
    Chan ch[3];       // op lists: interpreter fallback for maps the compiled form cannot represent (cc.ok == 0)
    HistEq he;        // valid when use_he
    int same01, same02, same12;
    int use_he;
    int status;       // 0 ok, < 0: tile rejected
    int valid;        // maps are valid (a rejected tile whose chain ran still has its chain image, evaluation.py:171-176)
    int pad_[2];
};
static_assert(sizeof(TileFinal) % 16 == 0 && offsetof(TileFinal, he) % 16 == 0 && sizeof(HistEq) % 16 == 0,
              "TileFinal is copied in 16-byte words");

struct TileHdr {      // written by pp_bucket_kernel
    int n;            // live (non-masked) pixels
    int n_sub[2];     // live pixels outside / inside the box (nsub == 2)
    float c, inv_s;   // binning centre and 1 / scale
    int pad_;
    double pivot;     // pivot of the per-bin moment sums
};

struct PPParams {
    cy_pp_chain chain;
    const uint32_t* img;
    long long row_stride;
    int big_endian;
    const int* x0;
    const int* y0;
    int B, Ty, Tx;
    // bucketing geometry
    int nsub;               // 1, or 2: every value bin split into (outside, inside) the central box
    int bx0, bx1, by0, by1; // the box [by0,by1) x [bx0,bx1) (BkgSubtractor / AbsMaxScaler / BorderMasker geometry)
    int border_mask;        // 1: pixels outside the box read as 0 (leading BorderMasker)
    int rows_per_chunk, nchunks, nbs;   // nbs = kNB * nsub sub-bins
    int use_tma, pw, npanels;           // TMA path: boxes of pw columns x rows_per_chunk rows
    // scratch (per tile: index b)
    TileHdr* hdr;
    float* vals;            // [B][N] bucketed live values, chunk-major
    unsigned short* off;    // [B][nchunks][nbs + 1] chunk-local start offsets of the sub-bins
    int* cbase;             // [B][nchunks + 1] start of every chunk inside vals
    int* cnt;               // [B][nbs]
    float* bmin;            // [B][nbs]
    float* bmax;            // [B][nbs]
    double* m1;             // [B][nbs] sum (x - pivot)
    double* m2;             // [B][nbs] sum (x - pivot)^2
    TileFinal* fin;         // [B]
    float* chain_out;       // [B][N][3] (optional parity output, written by pp_final_kernel)
    int* status;            // [B]
};

// ------------------------------------------------------------------------------------------ small device helpers

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum of up to three doubles; result valid in all threads.  red: shared double[3*32+3].
__device__ void block_sum3(double& a, double& b, double& c, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) {
        red[w] = a;
        red[32 + w] = b;
        red[64 + w] = c;
    }
    __syncthreads();
    if (w == 0) {
        double x = lane < kPPWarps ? red[lane] : 0.0, y = lane < kPPWarps ? red[32 + lane] : 0.0,
               z = lane < kPPWarps ? red[64 + lane] : 0.0;
        x = warp_sum(x);
        y = warp_sum(y);
        z = warp_sum(z);
        if (lane == 0) {
            red[96] = x;
            red[97] = y;
            red[98] = z;
        }
    }
    __syncthreads();
    a = red[96];
    b = red[97];
    c = red[98];
}
__device__ void block_minmax(double& mn, double& mx, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __syncthreads();
    if (lane == 0) {
        red[w] = mn;
        red[32 + w] = mx;
    }
    __syncthreads();
    if (w == 0) {
        double x = lane < kPPWarps ? red[lane] : INFINITY, y = lane < kPPWarps ? red[32 + lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
            y = fmax(y, __shfl_xor_sync(0xffffffffu, y, o));
        }
        if (lane == 0) {
            red[96] = x;
            red[97] = y;
        }
    }
    __syncthreads();
    mn = red[96];
    mx = red[97];
}

// ascending bitonic sort of n (power of two) floats in shared memory by NT threads; ends with a barrier
template <int NT>
__device__ void bitonic_sort_f(float* a, int n) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int e = threadIdx.x; e < n; e += NT) {
                const int q = e ^ j;
                if (q > e) {
                    const float x = a[e], y = a[q];
                    const bool asc = ((e & k) == 0);
                    if (asc ? (x > y) : (x < y)) {
                        a[e] = y;
                        a[q] = x;
                    }
                }
            }
            __syncthreads();
        }
}

// Value bin of pixel value x: linear bins of 1/48 robust sigma over [-8, 8) sigma around the tile's robust centre
// (bins 128..895), 8 logarithmic bins per octave beyond (0..127 below, 896..1023 above).  Monotone non-decreasing in x.
__device__ __forceinline__ int value_bin(float x, float c, float inv_s) {
    const float t = (x - c) * inv_s;
    if (t >= 8.f) {
        const uint32_t k = (__float_as_uint(t * 0.125f) - 0x3F800000u) >> 20;
        return 896 + (int)min(k, 127u);
    }
    if (t <= -8.f) {
        const uint32_t k = (__float_as_uint(-t * 0.125f) - 0x3F800000u) >> 20;
        return 127 - (int)min(k, 127u);
    }
    const int b = 128 + (int)floorf((t + 8.f) * 48.f);
    return min(max(b, 128), 895);
}

__device__ __forceinline__ float decode_pixel(uint32_t raw, int big_endian) {
    if (big_endian) raw = __byte_perm(raw, 0, 0x0123);
    const float f = __uint_as_float(raw);
    return isfinite(f) ? f : 0.0f;  // utils.py:219,394
}
__device__ __forceinline__ bool in_box(const PPParams& p, int y, int x) {
    return y >= p.by0 && y < p.by1 && x >= p.bx0 && x < p.bx1;
}
// pixel (y, x) of tile b as the chain sees it
__device__ __forceinline__ float load_pixel_yx(const PPParams& p, int b, int y, int x) {
    const float f = decode_pixel(p.img[(long long)(p.y0[b] + y) * p.row_stride + p.x0[b] + x], p.big_endian);
    return (p.border_mask && !in_box(p, y, x)) ? 0.0f : f;
}
__device__ __forceinline__ float load_pixel(const PPParams& p, int b, int idx) {
    const int y = idx / p.Tx;
    return load_pixel_yx(p, b, y, idx - y * p.Tx);
}

// ------------------------------------------------------------------------------------------ kernel 1: bucketing

struct BkSmem {
    uint32_t raw[kChunk];        // chunk as it lies in the mosaic (TMA destination; panels of pw columns)
    float srt[kChunk];           // chunk in sub-bin order
    int hist[2 * kNB];           // sub-bin counts of the chunk
    int hoff[2 * kNB + 1];       // exclusive scan of hist
    int tot[2 * kNB];
    float bmin[2 * kNB], bmax[2 * kNB];
    double m1[2 * kNB], m2[2 * kNB];
    int wsum[32];
    unsigned long long mbar;
    float c, inv_s;
    int nsamp;
};

__global__ void __launch_bounds__(kBkThreads, 1)
pp_bucket_kernel(const __grid_constant__ PPParams p, const __grid_constant__ CUtensorMap tm) {
    extern __shared__ __align__(128) unsigned char bk_smem_raw[];
    BkSmem& sm = *reinterpret_cast<BkSmem*>(bk_smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Ty = p.Ty, Tx = p.Tx, N = Ty * Tx;
    const int nbs = p.nbs, R = p.rows_per_chunk;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(&sm.mbar);
    const int tx0 = p.x0[b], ty0 = p.y0[b];

    auto issue_tma = [&](int ch) {   // one elected thread: boxes of the chunk's rows
        const uint32_t bytes = (uint32_t)R * Tx * 4u;
        mbar_arrive_expect_tx(mbar, bytes);
        for (int q = 0; q < p.npanels; ++q)
            tma_load_2d(&sm.raw[q * R * p.pw], &tm, mbar, tx0 + q * p.pw, ty0 + ch * R);
    };
    // chunks loaded by TMA (a partial last chunk is read directly): the box start must be 16-byte aligned, i.e. the
    // tile's first column a multiple of 4 pixels (pitch and base are checked on the host)
    const int full_chunks = (p.use_tma && (tx0 & 3) == 0) ? Ty / R : 0;
    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < nbs; i += kBkThreads) {
        sm.hist[i] = 0;
        sm.tot[i] = 0;
        sm.bmin[i] = INFINITY;
        sm.bmax[i] = -INFINITY;
        sm.m1[i] = 0.0;
        sm.m2[i] = 0.0;
    }
    __syncthreads();
    if (tid == 0 && full_chunks > 0) issue_tma(0);      // overlaps the sample sort below

    // ---- robust centre / scale from <= 1024 positional samples (sorted in srt[0..1024))
    {
        const int stride = max(1, N / 1024);
        const int ns = min(1024, (N + stride - 1) / stride);
        float v = INFINITY;
        if (tid < ns) {
            v = load_pixel(p, b, tid * stride);
            if (v == 0.0f) v = INFINITY;
        }
        sm.srt[tid] = v;
        const int live = __syncthreads_count(v != INFINITY);
        bitonic_sort_f<kBkThreads>(sm.srt, 1024);
        if (tid == 0) {
            float c = 0.f, s = 0.f;
            if (live > 0) {
                c = sm.srt[live / 2];
                s = 0.5f * (sm.srt[min(live - 1, (int)(0.84f * live))] - sm.srt[(int)(0.16f * live)]);
                if (!(s > 0.f)) s = (sm.srt[live - 1] - sm.srt[0]) * 0.0625f;
            }
            if (!(s > 0.f) || !isfinite(s)) s = fmaxf(fabsf(c) * 1e-3f, 1e-30f);
            float inv = 1.0f / s;
            if (!isfinite(inv) || !(inv > 0.f)) inv = 1.0f;
            sm.c = c;
            sm.inv_s = inv;
        }
        __syncthreads();
    }
    const float bc = sm.c, binv = sm.inv_s;
    const double pivot = (double)bc;

    int gbase = 0;               // live pixels written so far
    int nsub0 = 0, nsub1 = 0;
    float* vals = p.vals + (long long)b * N;
    unsigned short* off = p.off + (long long)b * p.nchunks * (nbs + 1);
    int* cbase = p.cbase + (long long)b * (p.nchunks + 1);
    uint32_t phase = 0;
    for (int ch = 0; ch < p.nchunks; ++ch) {
        const int r0 = ch * R, rows = min(R, Ty - r0), len = rows * Tx;
        const bool via_tma = ch < full_chunks;
        if (via_tma) {
            mbar_wait(mbar, phase);
            phase ^= 1;
        } else {
            // direct coalesced reads of the chunk's rows into the same buffer (panel layout with pw = Tx)
            for (int i = tid; i < len; i += kBkThreads) {
                const int y = i / Tx, x = i - y * Tx;
                sm.raw[i] = p.img[(long long)(ty0 + r0 + y) * p.row_stride + tx0 + x];
            }
            __syncthreads();
        }
        // ---- registers <- shared memory (byte swap, NaN -> 0), sub-bin of every live pixel
        const int pw = via_tma ? p.pw : Tx, rpw = R * pw;
        const bool need_xy = p.nsub == 2 || p.border_mask;   // pixel coordinates only when a stage uses the box
        float xv[kChunk / kBkThreads];
        int sb[kChunk / kBkThreads];
#pragma unroll
        for (int k = 0; k < kChunk / kBkThreads; ++k) {
            const int i = tid + k * kBkThreads;
            sb[k] = -1;
            xv[k] = 0.f;
            if (i < len) {
                float f = decode_pixel(sm.raw[i], p.big_endian);
                bool inb = true;
                if (need_xy) {
                    int y, x;
                    if (via_tma) {
                        const int q = i / rpw, rem = i - q * rpw;
                        y = rem / pw;
                        x = q * pw + (rem - y * pw);
                    } else {
                        y = i / Tx;
                        x = i - y * Tx;
                    }
                    inb = in_box(p, r0 + y, x);
                }
                if (p.border_mask && !inb) f = 0.0f;
                if (f != 0.0f) {
                    xv[k] = f;
                    const int vb = value_bin(f, bc, binv);
                    sb[k] = p.nsub == 2 ? 2 * vb + (inb ? 1 : 0) : vb;
                }
            }
        }
        __syncthreads();                                   // every thread has its pixels: the raw buffer is free
        if (tid == 0 && ch + 1 < full_chunks) {
            fence_proxy_async_smem();
            issue_tma(ch + 1);                             // the next chunk lands while this one is bucketed
        }
        // ---- slot inside the sub-bin (shared-memory atomics), then chunk-local counting sort
#pragma unroll
        for (int k = 0; k < kChunk / kBkThreads; ++k)
            if (sb[k] >= 0) sb[k] |= atomicAdd(&sm.hist[sb[k]], 1) << 12;     // nbs <= 2048 < 2^12, slot < 2^15
        __syncthreads();
        {   // exclusive scan of hist[0..nbs) -> hoff: thread t owns entries [t*e, (t+1)*e), e = nbs / 1024
            const int e = nbs / kBkThreads;
            int loc[2], s = 0;
            for (int q = 0; q < e; ++q) {
                loc[q] = sm.hist[tid * e + q];
                s += loc[q];
            }
            int inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) sm.wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int w = sm.wsum[lane], wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += y;
                }
                sm.wsum[lane] = wi - w;
            }
            __syncthreads();
            int run = sm.wsum[warp] + inc - s;
            for (int q = 0; q < e; ++q) {
                sm.hoff[tid * e + q] = run;
                run += loc[q];
            }
            if (tid == kBkThreads - 1) sm.hoff[nbs] = run;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kChunk / kBkThreads; ++k)
            if (sb[k] >= 0) sm.srt[sm.hoff[sb[k] & 0xfff] + (sb[k] >> 12)] = xv[k];
        __syncthreads();
        const int nlive = sm.hoff[nbs];
        // ---- per sub-bin statistics of the chunk (thread per sub-bin, no atomics) + write-out
        for (int j = tid; j < nbs; j += kBkThreads) {
            const int s = sm.hoff[j], e = sm.hoff[j + 1];
            if (e > s) {
                float mn = sm.bmin[j], mx = sm.bmax[j];
                double a1 = 0.0, a2 = 0.0;
                for (int i = s; i < e; ++i) {
                    const float x = sm.srt[i];
                    mn = fminf(mn, x);
                    mx = fmaxf(mx, x);
                    const double d = (double)x - pivot;
                    a1 += d;
                    a2 = fma(d, d, a2);
                }
                sm.bmin[j] = mn;
                sm.bmax[j] = mx;
                sm.m1[j] += a1;
                sm.m2[j] += a2;
                sm.tot[j] += e - s;
            }
        }
        for (int i = tid; i < nlive; i += kBkThreads) vals[gbase + i] = sm.srt[i];
        unsigned short* offc = off + (long long)ch * (nbs + 1);
        for (int j = tid; j <= nbs; j += kBkThreads) offc[j] = (unsigned short)sm.hoff[j];
        if (tid == 0) cbase[ch] = gbase;
        gbase += nlive;
        __syncthreads();
        for (int i = tid; i < nbs; i += kBkThreads) sm.hist[i] = 0;
        __syncthreads();
    }
    // ---- per-tile tables
    {
        int* cnt = p.cnt + (long long)b * nbs;
        float* bmin = p.bmin + (long long)b * nbs;
        float* bmax = p.bmax + (long long)b * nbs;
        double* m1 = p.m1 + (long long)b * nbs;
        double* m2 = p.m2 + (long long)b * nbs;
        int s0 = 0, s1 = 0;
        for (int j = tid; j < nbs; j += kBkThreads) {
            cnt[j] = sm.tot[j];
            bmin[j] = sm.bmin[j];
            bmax[j] = sm.bmax[j];
            m1[j] = sm.m1[j];
            m2[j] = sm.m2[j];
            if (p.nsub == 2) {
                if (j & 1) s1 += sm.tot[j];
                else s0 += sm.tot[j];
            }
        }
        if (p.nsub == 2) {
            // block totals of the two subsets
            s0 = __reduce_add_sync(0xffffffffu, s0);
            s1 = __reduce_add_sync(0xffffffffu, s1);
            __syncthreads();
            if (lane == 0) {
                sm.wsum[warp] = s0;
                sm.hist[warp] = s1;
            }
            __syncthreads();
            if (tid == 0) {
                for (int w = 0; w < 32; ++w) {
                    nsub0 += sm.wsum[w];
                    nsub1 += sm.hist[w];
                }
            }
        }
        if (tid == 0) {
            cbase[p.nchunks] = gbase;
            TileHdr h;
            h.n = gbase;
            h.n_sub[0] = nsub0;
            h.n_sub[1] = nsub1;
            h.c = bc;
            h.inv_s = binv;
            h.pad_ = 0;
            h.pivot = pivot;
            p.hdr[b] = h;
        }
    }
}

// ------------------------------------------------------------------------------------------ op lists

// np.interp(v, centers, cdf) with centers = (edges[:-1] + edges[1:]) / 2 (skimage equalize_hist, App. A.3)
__device__ double histeq_interp(const HistEq& h, double v) {
    const double c0 = (h.edges[0] + h.edges[1]) / 2.0, c255 = (h.edges[255] + h.edges[256]) / 2.0;
    if (!(v > c0)) return h.cdf[0];
    if (v >= c255) return h.cdf[255];
    int lo = 0, hi = 255;  // invariant: center[lo] <= v < center[hi]
    while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        const double cm = (h.edges[m] + h.edges[m + 1]) / 2.0;
        if (cm <= v) lo = m; else hi = m;
    }
    const double xl = (h.edges[lo] + h.edges[lo + 1]) / 2.0, xr = (h.edges[lo + 1] + h.edges[lo + 2]) / 2.0;
    const double slope = __ddiv_rn(__dsub_rn(h.cdf[lo + 1], h.cdf[lo]), __dsub_rn(xr, xl));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(v, xl)), h.cdf[lo]);
}

// Apply the first `nops` ops of channel c to the original pixel value x (fp64, reference operation order, no FMA
// contraction).  STICKY: a value that is (or becomes) exactly 0 stays 0 — the reference's `out[~cond] = 0`.
// !STICKY: the plain monotone composition (used to locate ranks among live elements).
template <bool STICKY>
__device__ double eval_ops(const Chan& c, int nops, const HistEq& he, double v) {
    for (int k = 0; k < nops; ++k) {
        if (STICKY && v == 0.0) return 0.0;
        switch (c.kind[k]) {
            case OP_SUB: v = __dsub_rn(v, c.p0[k]); break;
            case OP_SHIFT:
                v = __dsub_rn(v, c.p0[k]);
                if (v < 0.0) v = 0.0;
                break;
            case OP_CLAMP:
                if (v < c.p0[k]) v = c.p0[k];
                if (v > c.p1[k]) v = c.p1[k];
                break;
            case OP_ZSCALE:
                v = __dsub_rn(v, c.p0[k]);
                if (c.p1[k] != 0.0) v = __ddiv_rn(v, c.p1[k]);
                v = fmin(fmax(v, 0.0), 1.0);
                break;
            case OP_HISTEQ: v = histeq_interp(he, v); break;
            case OP_MINMAX:
                v = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(v, c.p0[k]), c.p1[k]), c.p2[k]), c.p3[k]);
                break;
            case OP_DIV: v = __ddiv_rn(v, c.p0[k]); break;
            case OP_STD: v = __ddiv_rn(__dsub_rn(v, c.p0[k]), c.p1[k]); break;
            case OP_LOG:   // LogStretcher with minmaxnorm: log10 of positive pixels, others at the smallest log
                v = v > 0.0 ? log10(v) : c.p0[k];
                v = __ddiv_rn(__dsub_rn(v, c.p1[k]), c.p2[k]);
                if (c.p3[k] != 0.0 && v < 0.0) v = 0.0;
                break;
        }
    }
    return v;
}

// np.interp on the uniform histogram centres with an O(1) bracket (same interpolation arithmetic as histeq_interp).
__device__ __forceinline__ double histeq_interp_fast(const HistEq& h, double v) {
    const double c0 = h.center[0], c255 = h.center[255];
    if (!(v > c0)) return h.cdf[0];
    if (v >= c255) return h.cdf[255];
    int lo = (int)((v - c0) * h.inv_step);   // a guess only: the two loops below make the bracket exact
    lo = max(0, min(254, lo));
    while (lo > 0 && h.center[lo] > v) --lo;
    while (lo < 254 && h.center[lo + 1] <= v) ++lo;
    return __dadd_rn(__dmul_rn(h.slope[lo], __dsub_rn(v, h.center[lo])), h.cdf[lo]);
}

// Value of a NON-masked pixel through the compiled op list (callers decide masking by value interval).
__device__ __forceinline__ double eval_fast(const Comp& cc, const Chan& c, const HistEq& he, double x) {
    if (!cc.ok) return eval_ops<true>(c, c.nops, he, x);
    double v = fmin(fmax(fma(cc.a0, x, cc.b0), cc.l0), cc.h0);
    if (cc.has_he) {
        v = histeq_interp_fast(he, v);
        v = fmin(fmax(fma(cc.a1, v, cc.b1), cc.l1), cc.h1);
    }
    return v;
}
template <class Z>
__device__ __forceinline__ bool in_zero_x(const Z& z, int nz, float x) {
    bool m = false;
    for (int k = 0; k < nz; ++k) m = m || (x >= z.zx0[k] && x <= z.zx1[k]);
    return m;
}

// (thread 0) rebuild the compiled form of channel c after its op list / zero sets changed
__device__ void compile_chan(const Chan& c, Comp& cc) {
    double a = 1.0, b = 0.0, l = -INFINITY, h = INFINITY;
    cc.has_he = 0;
    cc.ok = 1;
    for (int k = 0; k < c.nops; ++k) {
        switch (c.kind[k]) {
            case OP_SUB:
                b -= c.p0[k]; l -= c.p0[k]; h -= c.p0[k];
                break;
            case OP_SHIFT:
                b -= c.p0[k];
                l = fmax(l - c.p0[k], 0.0);
                h = fmax(h - c.p0[k], l);
                break;
            case OP_CLAMP:
                l = fmin(fmax(l, c.p0[k]), c.p1[k]);
                h = fmin(fmax(h, c.p0[k]), c.p1[k]);
                break;
            case OP_ZSCALE: {
                const double sc = c.p1[k] != 0.0 ? 1.0 / c.p1[k] : 1.0;
                if (!(sc > 0.0) || !isfinite(sc)) cc.ok = 0;
                a *= sc; b = (b - c.p0[k]) * sc; l = (l - c.p0[k]) * sc; h = (h - c.p0[k]) * sc;
                l = fmin(fmax(l, 0.0), 1.0);
                h = fmin(fmax(h, 0.0), 1.0);
                break;
            }
            case OP_MINMAX: {
                const double sc = c.p2[k] / c.p1[k];
                if (!(sc > 0.0) || !isfinite(sc)) cc.ok = 0;
                a *= sc; b = (b - c.p0[k]) * sc + c.p3[k]; l = (l - c.p0[k]) * sc + c.p3[k];
                h = (h - c.p0[k]) * sc + c.p3[k];
                break;
            }
            case OP_DIV: {
                const double sc = 1.0 / c.p0[k];
                if (!(sc > 0.0) || !isfinite(sc)) cc.ok = 0;
                a *= sc; b *= sc; l *= sc; h *= sc;
                break;
            }
            case OP_STD: {
                const double sc = 1.0 / c.p1[k];
                if (!(sc > 0.0) || !isfinite(sc)) cc.ok = 0;
                a *= sc; b = (b - c.p0[k]) * sc; l = (l - c.p0[k]) * sc; h = (h - c.p0[k]) * sc;
                break;
            }
            case OP_HISTEQ:
                if (cc.has_he) cc.ok = 0;
                cc.a0 = a; cc.b0 = b; cc.l0 = l; cc.h0 = h;
                cc.has_he = 1;
                a = 1.0; b = 0.0; l = -INFINITY; h = INFINITY;
                break;
            default: cc.ok = 0;
        }
    }
    if (cc.has_he) {
        cc.a1 = a; cc.b1 = b; cc.l1 = l; cc.h1 = h;
    } else {
        cc.a0 = a; cc.b0 = b; cc.l0 = l; cc.h0 = h;
        cc.a1 = 1.0; cc.b1 = 0.0; cc.l1 = -INFINITY; cc.h1 = INFINITY;
    }
    if (!isfinite(a) || !isfinite(b) || !isfinite(cc.a0) || !isfinite(cc.b0) || isnan(l) || isnan(h)) cc.ok = 0;
    cc.nzx = c.nz;
    for (int k = 0; k < c.nz; ++k) {
        cc.zx0[k] = c.zx0[k];
        cc.zx1[k] = c.zx1[k];
    }
}

// (single thread) add the zero set [h0, h1) = [x0, x1] to channel c: insertion sort + merge of ranks and values
__device__ void add_zero_range(Chan& c, int h0, int h1, float x0, float x1) {
    if (h1 <= h0) return;
    int n = c.nz;
    if (n < kMaxZero) {
        c.z0[n] = h0; c.z1[n] = h1; c.zx0[n] = x0; c.zx1[n] = x1;
        ++n;
    } else {  // table full: widen the last set (never expected: every stage adds at most one contiguous set)
        if (h1 > c.z1[n - 1]) { c.z1[n - 1] = h1; c.zx1[n - 1] = x1; }
        if (h0 < c.z0[n - 1]) { c.z0[n - 1] = h0; c.zx0[n - 1] = x0; }
    }
    for (int i = n - 1; i > 0 && c.z0[i] < c.z0[i - 1]; --i) {
        const int t0 = c.z0[i], t1 = c.z1[i];
        const float u0 = c.zx0[i], u1 = c.zx1[i];
        c.z0[i] = c.z0[i - 1]; c.z1[i] = c.z1[i - 1]; c.zx0[i] = c.zx0[i - 1]; c.zx1[i] = c.zx1[i - 1];
        c.z0[i - 1] = t0; c.z1[i - 1] = t1; c.zx0[i - 1] = u0; c.zx1[i - 1] = u1;
    }
    int m = 0;
    for (int i = 1; i < n; ++i) {
        if (c.z0[i] <= c.z1[m]) {
            if (c.z1[i] > c.z1[m]) { c.z1[m] = c.z1[i]; c.zx1[m] = c.zx1[i]; }
        } else {
            ++m;
            c.z0[m] = c.z0[i]; c.z1[m] = c.z1[i]; c.zx0[m] = c.zx0[i]; c.zx1[m] = c.zx1[i];
        }
    }
    c.nz = n ? m + 1 : 0;
}
__device__ int live_count(const Chan& c, int a, int b) {
    int n = b - a;
    for (int i = 0; i < c.nz; ++i) {
        const int lo = max(a, c.z0[i]), hi = min(b, c.z1[i]);
        if (hi > lo) n -= hi - lo;
    }
    return n;
}
// rank of the k-th (0-based) live element at or after rank a
__device__ int kth_live(const Chan& c, int a, int k) {
    int pos = a;
    for (int i = 0; i < c.nz; ++i) {
        if (c.z1[i] <= pos) continue;
        if (c.z0[i] <= pos) {
            pos = c.z1[i];
            continue;
        }
        const int gap = c.z0[i] - pos;
        if (k < gap) return pos + k;
        k -= gap;
        pos = c.z1[i];
    }
    return pos + k;
}

// ------------------------------------------------------------------------------------------ bin-table view of a tile

struct Shared {
    Chan ch[3];
    Comp cc[3];
    Chan tmp;                // channel with its zero sets re-ranked for a subset view
    HistEq he;
    double red[3 * 32 + 4];
    double zs[1024];         // zscale samples / flat residuals
    unsigned char zbad[1024];
    unsigned char zbad2[1024];
    int hist[256];
    int next_hid;
    int fail;                // tile status
    int sidx;                // block-wide search scratch
    int he_used;             // a histogram equalisation owns `he`
    // active view of the bucketed tile: -1 all pixels, 0 / 1 = outside / inside the box
    int view, n;
    int prefix[kNB + 1];     // rank of the first element of every logical bin
    float bmin[kNB], bmax[kNB];
    float gath[kGather];     // elements of bin gath_bin as gathered (gath_mode 2: sorted)
    float gath2[kGather];    // the same in sub-bucket order (gath_mode 1)
    int gsub[256], gpre[257];
    int gath_bin, gath_n, gath_mode;
    int plist[kNB], nplist;  // bins a query has to walk element by element
    int wtmp[32];
    float fres[4];
};

struct TileView {            // global tables of one tile
    const float* vals;
    const unsigned short* off;
    const int* cbase;
    const int* cnt;
    const float* bmin;
    const float* bmax;
    const double* m1;
    const double* m2;
    int nchunks, nbs, nsub;
    double pivot;
};
__device__ __forceinline__ TileView tile_view(const PPParams& p, int b) {
    TileView v;
    const long long N = (long long)p.Ty * p.Tx;
    v.vals = p.vals + b * N;
    v.off = p.off + (long long)b * p.nchunks * (p.nbs + 1);
    v.cbase = p.cbase + (long long)b * (p.nchunks + 1);
    v.cnt = p.cnt + (long long)b * p.nbs;
    v.bmin = p.bmin + (long long)b * p.nbs;
    v.bmax = p.bmax + (long long)b * p.nbs;
    v.m1 = p.m1 + (long long)b * p.nbs;
    v.m2 = p.m2 + (long long)b * p.nbs;
    v.nchunks = p.nchunks;
    v.nbs = p.nbs;
    v.nsub = p.nsub;
    v.pivot = p.hdr[b].pivot;
    return v;
}
// sub-bin range [s0, s1) of logical bin j in the active view
__device__ __forceinline__ void sub_range(const TileView& tv, int view, int j, int& s0, int& s1) {
    if (tv.nsub == 1) { s0 = j; s1 = j + 1; }
    else if (view < 0) { s0 = 2 * j; s1 = 2 * j + 2; }
    else { s0 = 2 * j + view; s1 = s0 + 1; }
}
// Visits every element of logical bin j: warp w walks chunks w, w + 16, ...; lanes stride over the chunk's segment.
template <class F>
__device__ __forceinline__ void for_bin_elements(const TileView& tv, int view, int j, F f) {
    int s0, s1;
    sub_range(tv, view, j, s0, s1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < tv.nchunks; c += kPPWarps) {
        const unsigned short* oc = tv.off + (long long)c * (tv.nbs + 1);
        const int base = tv.cbase[c];
        const int a = base + oc[s0], e = base + oc[s1];
        for (int i = a + lane; i < e; i += 32) f(tv.vals[i]);
    }
}

// Visits every element of the bins listed in sh.plist[0 .. np) (entry: low 10 bits = logical bin, the rest is passed
// to f).  The work items are (bin, chunk) pairs; a warp fetches the segment bounds of 32 items at once (one per lane: the
// three dependent global loads of an item overlap across lanes) and then walks the segments one after the other.
template <class F>
__device__ __forceinline__ void for_list_elements(const Shared& sh, const TileView& tv, int np, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = np * tv.nchunks;
    for (int base = warp * 32; base < total; base += kPPWarps * 32) {
        const int t = base + lane;
        int a = 0, e = 0, ent = 0;
        if (t < total) {
            const int i = t / tv.nchunks, c = t - i * tv.nchunks;
            ent = sh.plist[i];
            int s0, s1;
            sub_range(tv, sh.view, ent & 1023, s0, s1);
            const unsigned short* oc = tv.off + (long long)c * (tv.nbs + 1);
            const int cb = tv.cbase[c];
            a = cb + oc[s0];
            e = cb + oc[s1];
        }
        const int nitems = min(32, total - base);
        for (int q = 0; q < nitems; ++q) {
            const int aa = __shfl_sync(0xffffffffu, a, q), ee = __shfl_sync(0xffffffffu, e, q);
            const int en = __shfl_sync(0xffffffffu, ent, q);
            // four loads in flight per lane: the walk is bound by the latency of these global loads
            int k = aa + lane;
            for (; k + 96 < ee; k += 128) {
                const float v0 = tv.vals[k], v1 = tv.vals[k + 32], v2 = tv.vals[k + 64], v3 = tv.vals[k + 96];
                f(v0, en);
                f(v1, en);
                f(v2, en);
                f(v3, en);
            }
            for (; k < ee; k += 32) f(tv.vals[k], en);
        }
    }
}

// (all threads) make `view` the active view: logical-bin counts -> prefix ranks, minima / maxima
__device__ void set_view(Shared& sh, const TileView& tv, int view) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __syncthreads();
    int c2[2];
    for (int q = 0; q < 2; ++q) {    // thread t owns logical bins 2t, 2t + 1
        const int j = 2 * t + q;
        int s0, s1;
        sub_range(tv, view, j, s0, s1);
        int cn = 0;
        float mn = INFINITY, mx = -INFINITY;
        for (int s = s0; s < s1; ++s) {
            cn += tv.cnt[s];
            mn = fminf(mn, tv.bmin[s]);
            mx = fmaxf(mx, tv.bmax[s]);
        }
        c2[q] = cn;
        sh.bmin[j] = mn;
        sh.bmax[j] = mx;
    }
    const int s = c2[0] + c2[1];
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sh.wtmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < kPPWarps ? sh.wtmp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        if (lane < kPPWarps) sh.wtmp[lane] = wi - w;
    }
    __syncthreads();
    const int run = sh.wtmp[warp] + inc - s;
    sh.prefix[2 * t] = run;
    sh.prefix[2 * t + 1] = run + c2[0];
    if (t == kPPThreads - 1) {
        sh.prefix[kNB] = run + s;
        sh.n = run + s;
        sh.view = view;
        sh.gath_bin = -1;
    }
    __syncthreads();
}

// Value at rank r of the active view (exact order statistic): the elements of r's bin are gathered and sorted in shared
// memory (kept until another bin is asked for); bins larger than kGather take a 4-pass radix selection instead.
__device__ float value_at(Shared& sh, const TileView& tv, int r) {
    // bin of rank r: largest j with prefix[j] <= r (uniform binary search in shared memory)
    int lo = 0, hi = kNB;
    while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (sh.prefix[m] <= r) lo = m; else hi = m;
    }
    const int j = lo, k = r - sh.prefix[j], m = sh.prefix[j + 1] - sh.prefix[j];
    __syncthreads();
    if (m <= kGather) {
        // The bin's elements are gathered ONCE and counting-sorted into 256 linear sub-buckets of [bmin, bmax] (kept
        // until another bin is asked for); a query then ranks the few elements of ONE sub-bucket.  gath_mode: 1 =
        // sub-bucketed (sh.gath2 in sub-bucket order, sh.gpre = starts), 2 = fully sorted (a sub-bucket was too large:
        // many equal values).
        if (sh.gath_bin != j) {
            const float bmn = sh.bmin[j], bmx = sh.bmax[j];
            const float sc = bmx > bmn ? 256.0f / (bmx - bmn) : 0.0f;
            __syncthreads();
            if (threadIdx.x == 0) sh.gath_n = 0;
            for (int i = threadIdx.x; i < 256; i += kPPThreads) sh.gsub[i] = 0;
            __syncthreads();
            for_bin_elements(tv, sh.view, j, [&](float x) {
                sh.gath[atomicAdd(&sh.gath_n, 1)] = x;
                atomicAdd(&sh.gsub[min(255, (int)((x - bmn) * sc))], 1);
            });
            __syncthreads();
            if (threadIdx.x < 32) {        // exclusive scan of the 256 counts by one warp (8 per lane)
                const int l = threadIdx.x;
                int loc[8], sum = 0, mxc = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    loc[q] = sh.gsub[8 * l + q];
                    sum += loc[q];
                    mxc = max(mxc, loc[q]);
                }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (l >= o) inc += y;
                }
                int run = inc - sum;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    sh.gpre[8 * l + q] = run;
                    sh.gsub[8 * l + q] = run;      // cursor of the scatter below
                    run += loc[q];
                }
                if (l == 31) sh.gpre[256] = run;
                mxc = __reduce_max_sync(0xffffffffu, mxc);
                if (l == 0) sh.gath_mode = mxc <= 256 ? 1 : 2;
            }
            __syncthreads();
            if (sh.gath_mode == 1) {
                for (int i = threadIdx.x; i < m; i += kPPThreads) {
                    const float x = sh.gath[i];
                    sh.gath2[atomicAdd(&sh.gsub[min(255, (int)((x - bmn) * sc))], 1)] = x;
                }
            } else {
                int np2 = 32;
                while (np2 < m) np2 <<= 1;
                for (int i = m + threadIdx.x; i < np2; i += kPPThreads) sh.gath[i] = INFINITY;
                __syncthreads();
                bitonic_sort_f<kPPThreads>(sh.gath, np2);
            }
            if (threadIdx.x == 0) sh.gath_bin = j;
            __syncthreads();
        }
        if (sh.gath_mode == 2) return sh.gath[k];
        // sub-bucket of rank k, then the (k - start)-th smallest of its elements (ties broken by position)
        int sl = 0, sr = 256;
        while (sr - sl > 1) {
            const int mm = (sl + sr) >> 1;
            if (sh.gpre[mm] <= k) sl = mm; else sr = mm;
        }
        const int s0 = sh.gpre[sl], cntb = sh.gpre[sl + 1] - s0, want = k - s0;
        if (threadIdx.x < cntb) {
            const float me = sh.gath2[s0 + threadIdx.x];
            int rank = 0;
            for (int q = 0; q < cntb; ++q) {
                const float o = sh.gath2[s0 + q];
                rank += (o < me) || (o == me && q < (int)threadIdx.x);
            }
            if (rank == want) sh.fres[0] = me;
        }
        __syncthreads();
        return sh.fres[0];
    }
    // radix selection on the order-preserving key, most significant byte first
    uint32_t prefix_key = 0, mask = 0;
    int kk = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += kPPThreads) sh.hist[i] = 0;
        __syncthreads();
        for_bin_elements(tv, sh.view, j, [&](float x) {
            const uint32_t u = __float_as_uint(x);
            const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            if ((key & mask) == prefix_key) atomicAdd(&sh.hist[(key >> shift) & 255u], 1);
        });
        __syncthreads();
        int d = 0, acc = 0;
        for (; d < 256; ++d) {      // every thread walks the same 256 counters
            const int h = sh.hist[d];
            if (kk < acc + h) break;
            acc += h;
        }
        kk -= acc;
        prefix_key |= (uint32_t)d << shift;
        mask |= 0xffu << shift;
        __syncthreads();
    }
    const uint32_t u = (prefix_key & 0x80000000u) ? (prefix_key & 0x7fffffffu) : ~prefix_key;
    return __uint_as_float(u);
}

// First rank r in [a, b) of the active view with P(x) true, P(x) = f(x) >= t (STRICT: f(x) > t), f = the first nops ops
// of c (plain composition, monotone); b if none.  *xbelow / *xabove: largest pixel value with !P / smallest with P over
// the whole view (-inf / +inf if none) -- the value form of the rank boundary.
// P(x) of lower_index.  `fast` (optional): the compiled form of exactly these ops; it decides when the value is further
// from t than its rounding error can be (|v - t| > 1e-9 of the magnitudes involved; the compiled form differs from the
// sequential evaluation by a few fp64 ulps), otherwise the reference operation order does.
template <bool STRICT>
__device__ __forceinline__ bool pred_ge(const Chan& c, int nops, const HistEq& he, const Comp* fast, double x, double t) {
    if (fast) {
        const double ax = fast->a0 * x;
        const double v = fmin(fmax(ax + fast->b0, fast->l0), fast->h0);
        const double margin = 1e-9 * (fabs(ax) + fabs(fast->b0) + fabs(t));
        if (v > t + margin) return true;
        if (v < t - margin) return false;
    }
    const double v = eval_ops<false>(c, nops, he, x);
    return STRICT ? (v > t) : (v >= t);
}

template <bool STRICT>
__device__ int lower_index(Shared& sh, const TileView& tv, const Chan& c, int nops, int a, int b, double t,
                           float* xbelow, float* xabove, const Comp* fast = nullptr) {
    if (fast && (!fast->ok || fast->has_he || nops != c.nops)) fast = nullptr;
    __syncthreads();
    if (threadIdx.x == 0) sh.sidx = kNB;
    __syncthreads();
    for (int j = threadIdx.x; j < kNB; j += kPPThreads) {
        if (sh.prefix[j + 1] > sh.prefix[j]) {
            if (pred_ge<STRICT>(c, nops, sh.he, fast, (double)sh.bmax[j], t)) atomicMin(&sh.sidx, j);
        }
    }
    __syncthreads();
    const int jb = sh.sidx;
    int r;
    float xb = -INFINITY, xa = INFINITY;
    if (jb == kNB) {
        r = sh.n;
    } else {
        int cntb = 0;
        float mxb = -INFINITY, mna = INFINITY;
        for_bin_elements(tv, sh.view, jb, [&](float x) {
            if (pred_ge<STRICT>(c, nops, sh.he, fast, (double)x, t)) mna = fminf(mna, x);
            else { ++cntb; mxb = fmaxf(mxb, x); }
        });
        double d0 = (double)cntb, d1 = 0.0, d2 = 0.0, mn = (double)mna, mx = (double)mxb;
        block_sum3(d0, d1, d2, sh.red);
        block_minmax(mn, mx, sh.red);
        r = sh.prefix[jb] + (int)d0;
        xb = (float)mx;
        xa = (float)mn;
    }
    if (xb == -INFINITY) {   // nothing below inside bin jb: the largest element of the previous non-empty bin
        for (int j = min(jb, kNB) - 1; j >= 0; --j)
            if (sh.prefix[j + 1] > sh.prefix[j]) {
                xb = sh.bmax[j];
                break;
            }
    }
    if (xbelow) *xbelow = xb;
    if (xabove) *xabove = xa;
    return min(max(r, a), b);
}

// Moment sums of the live values f(x) over the pixels with xa <= x <= xb of the active view:  n, sum(v-p), sum((v-p)^2)
// in fp64.  Bins entirely inside the interval whose map is affine (or constant) over the bin come from the per-bin
// sums in closed form; bins cut by an interval end, a masked interval or a clamp bound are walked element by element.
__device__ void range_sums(Shared& sh, const TileView& tv, const Chan& c, const Comp& cc, float xa, float xb, double pv,
                           double& s0, double& s1, double& s2) {
    double n = 0.0, u = 0.0, q = 0.0;
    const bool simple = cc.ok && !cc.has_he;
    __syncthreads();
    if (threadIdx.x == 0) sh.nplist = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < kNB; j += kPPThreads) {
        const int cn = sh.prefix[j + 1] - sh.prefix[j];
        if (cn == 0) continue;
        const float lo = sh.bmin[j], hi = sh.bmax[j];
        if (hi < xa || lo > xb) continue;
        bool partial = !(lo >= xa && hi <= xb) || !simple;
        bool masked = false, touches = false;
        for (int k = 0; k < c.nz; ++k) {
            if (hi < c.zx0[k] || lo > c.zx1[k]) continue;
            if (lo >= c.zx0[k] && hi <= c.zx1[k]) masked = true; else partial = touches = true;
        }
        if (masked) continue;
        if (!partial) {
            const double flo = fma(cc.a0, (double)lo, cc.b0), fhi = fma(cc.a0, (double)hi, cc.b0);
            if (flo >= cc.l0 && fhi <= cc.h0) {
                int q0, q1;
                sub_range(tv, sh.view, j, q0, q1);
                double m1 = 0.0, m2 = 0.0;
                for (int s = q0; s < q1; ++s) {
                    m1 += tv.m1[s];
                    m2 += tv.m2[s];
                }
                const double off = fma(cc.a0, tv.pivot, cc.b0) - pv;     // f(x) - p = a (x - pivot) + off
                n += (double)cn;
                u += cc.a0 * m1 + (double)cn * off;
                q += cc.a0 * cc.a0 * m2 + 2.0 * cc.a0 * off * m1 + (double)cn * off * off;
            } else if (fhi <= cc.l0 || flo >= cc.h0) {
                const double d = (fhi <= cc.l0 ? cc.l0 : cc.h0) - pv;
                n += (double)cn;
                u += (double)cn * d;
                q += (double)cn * d * d;
            } else {
                partial = true;
            }
        }
        if (partial) sh.plist[atomicAdd(&sh.nplist, 1)] = j | (touches ? 1 << 10 : 0);
    }
    __syncthreads();
    const int np = sh.nplist;
    for_list_elements(sh, tv, np, [&](float x, int ent) {
        if (x < xa || x > xb) return;
        if ((ent >> 10) && in_zero_x(c, c.nz, x)) return;     // only bins that touch a masked interval test for it
        const double d = eval_fast(cc, c, sh.he, (double)x) - pv;
        n += 1.0;
        u += d;
        q = fma(d, d, q);
    });
    block_sum3(n, u, q, sh.red);
    s0 = n;
    s1 = u;
    s2 = q;
}

__device__ double live_median(Shared& sh, const TileView& tv, const Chan& c, int a, int cnt) {
    // numpy median: middle element, or mean of the two middle elements
    if (cnt & 1) return eval_ops<true>(c, c.nops, sh.he, (double)value_at(sh, tv, kth_live(c, a, cnt >> 1)));
    const double x = eval_ops<true>(c, c.nops, sh.he, (double)value_at(sh, tv, kth_live(c, a, (cnt >> 1) - 1)));
    const double y = eval_ops<true>(c, c.nops, sh.he, (double)value_at(sh, tv, kth_live(c, a, cnt >> 1)));
    return (x + y) / 2.0;
}

// astropy SigmaClip (axis=None, median/std, maxiters 5; App. A.1) over the live values of channel c in the active view.
// Outputs the bounds of the last iteration and the (mean, std) of the survivors.  Returns false if the set is empty.
// Every iteration keeps a contiguous rank range = value interval [xa, xb]; the moment sums are taken about the initial
// median (mean-pivot = O(std): no cancellation): mean = p + S1/n, std = sqrt(S2/n - (S1/n)^2)  (== numpy's two-pass
// definition up to fp64 rounding).
__device__ bool sigma_clip(Shared& sh, const TileView& tv, const Chan& c, const Comp& cc, double sig_lo, double sig_hi,
                           double& lo, double& hi, double& mean, double& sd) {
    int a = 0, b = sh.n;
    float xa = -INFINITY, xb = INFINITY;
    int cnt = live_count(c, a, b);
    if (cnt <= 0) return false;
    const double p = live_median(sh, tv, c, a, cnt);
    double s0, s1, s2;
    range_sums(sh, tv, c, cc, xa, xb, p, s0, s1, s2);
    if ((int)s0 != cnt) {  // rank bookkeeping and evaluation disagree: never expected
        if (threadIdx.x == 0) sh.fail = -5;
        cnt = (int)s0;
        if (cnt <= 0) return false;
    }
    lo = hi = 0.0;
    for (int it = 0; it < 5; ++it) {
        const double m1 = s1 / s0;
        mean = p + m1;
        sd = sqrt(fmax(s2 / s0 - m1 * m1, 0.0));
        const double med = live_median(sh, tv, c, a, cnt);
        lo = med - sd * sig_lo;
        hi = med + sd * sig_hi;
        float xbel, xabv;
        int na = a, nb = b;
        float nxa = xa, nxb = xb;
        const int r1 = lower_index<false>(sh, tv, c, c.nops, a, b, lo, &xbel, &xabv, &cc);   // first f >= lo
        if (r1 > a) { na = r1; nxa = xabv; }
        const int r2 = lower_index<true>(sh, tv, c, c.nops, na, b, hi, &xbel, &xabv, &cc);   // first f > hi
        if (r2 < b) { nb = r2; nxb = xbel; }
        const int ncnt = live_count(c, na, nb);
        if (ncnt == cnt) break;
        if (ncnt <= 0) return false;
        a = na; b = nb; xa = nxa; xb = nxb; cnt = ncnt;
        range_sums(sh, tv, c, cc, xa, xb, p, s0, s1, s2);
        if ((int)s0 != cnt && threadIdx.x == 0) sh.fail = -5;
        if (s0 <= 0.0) return false;
    }
    const double m1 = s1 / s0;
    mean = p + m1;
    sd = sqrt(fmax(s2 / s0 - m1 * m1, 0.0));
    return true;
}

// Append an op to channel ci and register the pixels it newly maps to exactly 0 (they become masked).  View: all.
__device__ void push_op(Shared& sh, const TileView& tv, int ci, int kind, double p0, double p1, double p2, double p3) {
    Chan& c = sh.ch[ci];
    __syncthreads();
    if (threadIdx.x == 0) {
        const int k = c.nops;
        if (k < kMaxOps) {
            c.kind[k] = kind;
            c.p0[k] = p0; c.p1[k] = p1; c.p2[k] = p2; c.p3[k] = p3;
            c.nops = k + 1;
        } else {
            sh.fail = -4;
        }
        compile_chan(c, sh.cc[ci]);     // the compiled form of the new op list: fast path of the two searches below
    }
    __syncthreads();
    // zero set of the plain composition is a contiguous rank range (monotone): [first f >= 0, first f > 0)
    float xb0, xa0, xb1, xa1;
    const int h0 = lower_index<false>(sh, tv, c, c.nops, 0, sh.n, 0.0, &xb0, &xa0, &sh.cc[ci]);
    const int h1 = lower_index<true>(sh, tv, c, c.nops, h0, sh.n, 0.0, &xb1, &xa1, &sh.cc[ci]);
    __syncthreads();
    if (threadIdx.x == 0) {
        add_zero_range(c, h0, h1, xa0, xb1);
        c.hid = sh.next_hid++;
        compile_chan(c, sh.cc[ci]);
    }
    __syncthreads();
}

__device__ void copy_chan(Shared& sh, int dst, int src) {
    __syncthreads();
    if (threadIdx.x == 0) {
        sh.ch[dst] = sh.ch[src];
        sh.cc[dst] = sh.cc[src];
    }
    __syncthreads();
}

// smallest / largest live value of channel c after its ops (NaN-free); false if the channel has no live pixel
__device__ bool live_min_max(Shared& sh, const TileView& tv, const Chan& c, double& mn, double& mx) {
    const int nl = live_count(c, 0, sh.n);
    if (nl <= 0) return false;
    mn = eval_ops<true>(c, c.nops, sh.he, (double)value_at(sh, tv, kth_live(c, 0, 0)));
    mx = eval_ops<true>(c, c.nops, sh.he, (double)value_at(sh, tv, kth_live(c, 0, nl - 1)));
    return true;
}

// (all threads) channel c with its zero sets expressed in ranks of the active (subset) view -> sh.tmp
__device__ void rerank_zero_sets(Shared& sh, const TileView& tv, const Chan& c) {
    __syncthreads();
    if (threadIdx.x == 0) sh.tmp = c;
    __syncthreads();
    const int nz = c.nz;
    for (int k = 0; k < nz; ++k) {
        const int z0 = lower_index<false>(sh, tv, c, 0, 0, sh.n, (double)c.zx0[k], nullptr, nullptr);
        const int z1 = lower_index<true>(sh, tv, c, 0, 0, sh.n, (double)c.zx1[k], nullptr, nullptr);
        if (threadIdx.x == 0) {
            sh.tmp.z0[k] = z0;
            sh.tmp.z1[k] = z1;
        }
    }
    __syncthreads();
}
static_assert(kNB == 2 * kPPThreads, "set_view: two logical bins per thread");

// ------------------------------------------------------------------------------------------ stages

// astropy ZScaleInterval.get_limits (App. A.2) on channel ci (positional samples of the current image, zeros
// included), then the OP_ZSCALE op.
__device__ void zscale_stage(Shared& sh, const PPParams& p, int b, const TileView& tv, int ci, double contrast) {
    const Chan& c = sh.ch[ci];
    const int t = threadIdx.x;
    const int N = p.Ty * p.Tx;
    constexpr int kZS = 1024;                     // sample slots (>= the 1000 samples of ZScaleInterval)
    constexpr int kZE = kZS / kPPThreads;         // slots per thread: e = t + q * kPPThreads
    const int stride = (int)fmax(1.0, (double)N / 1000.0);
    int npix = (N + stride - 1) / stride;
    if (npix > 1000) npix = 1000;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kZE; ++q) {
        const int e = t + q * kPPThreads;
        sh.zs[e] = (e < npix) ? eval_ops<true>(c, c.nops, sh.he, (double)load_pixel(p, b, e * stride)) : INFINITY;
    }
    __syncthreads();
    // bitonic sort ascending, 1024 elements, one compare-exchange per element pair
    for (int k = 2; k <= kZS; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int q = 0; q < kZE; ++q) {
                const int e = t + q * kPPThreads;
                const int pp = e ^ j;
                if (pp > e) {
                    const double x = sh.zs[e], y = sh.zs[pp];
                    const bool asc = ((e & k) == 0);
                    if (asc ? (x > y) : (x < y)) {
                        sh.zs[e] = y;
                        sh.zs[pp] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    double y[kZE], x[kZE], flat[kZE];
#pragma unroll
    for (int q = 0; q < kZE; ++q) {
        const int e = t + q * kPPThreads;
        y[q] = e < npix ? sh.zs[e] : 0.0;
        x[q] = (double)e;
        sh.zbad[e] = 0;
    }
    double vmin = sh.zs[0], vmax = sh.zs[npix - 1];
    const int minpix = max(5, (int)(npix * 0.5));
    int ngood = npix, last = npix + 1;
    const int ngrow = max(1, (int)(npix * 0.01));
    double slope = 0.0, icpt = 0.0;
    bool fitted = false;
    __syncthreads();
    for (int it = 0; it < 5; ++it) {
        if (ngood >= last || ngood < minpix) break;
        bool good[kZE];
        // weighted (0/1) least squares line, centred for stability (== np.polyfit deg 1 up to rounding)
        double sw = 0.0, sx = 0.0, sy = 0.0;
#pragma unroll
        for (int q = 0; q < kZE; ++q) {
            const int e = t + q * kPPThreads;
            good[q] = (e < npix) && !sh.zbad[e];
            if (good[q]) {
                sw += 1.0;
                sx += x[q];
                sy += y[q];
            }
        }
        block_sum3(sw, sx, sy, sh.red);
        const double xm = sx / sw, ym = sy / sw;
        double sxx = 0.0, sxy = 0.0, zz = 0.0;
#pragma unroll
        for (int q = 0; q < kZE; ++q)
            if (good[q]) {
                sxx += (x[q] - xm) * (x[q] - xm);
                sxy += (x[q] - xm) * (y[q] - ym);
            }
        block_sum3(sxx, sxy, zz, sh.red);
        slope = sxy / sxx;
        icpt = ym - slope * xm;
        fitted = true;
        double sf = 0.0, z1 = 0.0, z2 = 0.0;
#pragma unroll
        for (int q = 0; q < kZE; ++q) {
            flat[q] = y[q] - (slope * x[q] + icpt);
            if (good[q]) sf += flat[q];
        }
        block_sum3(sf, z1, z2, sh.red);
        const double fm = sf / sw;
        double sq = 0.0;
#pragma unroll
        for (int q = 0; q < kZE; ++q)
            if (good[q]) sq += (flat[q] - fm) * (flat[q] - fm);
        z1 = z2 = 0.0;
        block_sum3(sq, z1, z2, sh.red);
        const double thr = 2.5 * sqrt(sq / sw);
#pragma unroll
        for (int q = 0; q < kZE; ++q) {
            const int e = t + q * kPPThreads;
            if (e < npix && (flat[q] < -thr || flat[q] > thr)) sh.zbad[e] = 1;
        }
        __syncthreads();
        // np.convolve(badpix, ones(ngrow), 'same') on bools: out[i] = OR bad[i - ngrow/2 .. i + (ngrow-1)/2]
        double g = 0.0;
#pragma unroll
        for (int q = 0; q < kZE; ++q) {
            const int e = t + q * kPPThreads;
            unsigned char nb = 0;
            if (e < npix) {
                const int l0 = max(0, e - ngrow / 2), l1 = min(npix - 1, e + (ngrow - 1) / 2);
                for (int r = l0; r <= l1; ++r) nb |= sh.zbad[r];
            }
            sh.zbad2[e] = nb;
            if (e < npix && !nb) g += 1.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kZE; ++q) {
            const int e = t + q * kPPThreads;
            sh.zbad[e] = sh.zbad2[e];
        }
        z1 = z2 = 0.0;
        block_sum3(g, z1, z2, sh.red);
        last = ngood;
        ngood = (int)g;
    }
    if (ngood >= minpix && fitted) {
        double sl = slope;
        if (contrast > 0) sl = sl / contrast;
        const int center = (npix - 1) / 2;
        const double med = (npix & 1) ? sh.zs[npix / 2] : (sh.zs[npix / 2 - 1] + sh.zs[npix / 2]) / 2.0;
        vmin = fmax(vmin, med - (double)(center - 1) * sl);
        vmax = fmin(vmax, med + (double)(npix - center) * sl);
    }
    push_op(sh, tv, ci, OP_ZSCALE, vmin, vmax - vmin, 0.0, 0.0);
}

// histogram bin of value v among the 256 bins of sh.he.edges (numpy: edges[k] <= v < edges[k+1], last bin closed);
// -1 below the first edge
__device__ __forceinline__ int he_bin(const HistEq& h, double v) {
    if (v < h.edges[0]) return -1;
    int lo = 0, hi = 256;   // invariant: edges[lo] <= v, (hi == 256 or v < edges[hi])
    while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (h.edges[m] <= v) lo = m; else hi = m;
    }
    return min(lo, 255);
}

// skimage equalize_hist (App. A.3) on channel ci: 256-bin histogram over [min,max] of ALL pixels (masked zeros
// included), CDF, then OP_HISTEQ.  The channel is a monotone map of the pixel value, so min / max are its first / last
// live element (and 0 if any pixel is masked); value bins that fall into ONE histogram bin are counted at once, the
// bins a histogram edge (or a masked interval) cuts are walked element by element.
__device__ bool histeq_stage(Shared& sh, const PPParams& p, const TileView& tv, int ci) {
    const Chan& c = sh.ch[ci];
    const int N = p.Ty * p.Tx;
    if (sh.he_used) {   // one table per chain
        if (threadIdx.x == 0) sh.fail = -6;
        return false;
    }
    const int nlive = live_count(c, 0, sh.n);
    const int nzero = N - nlive;
    double mn = INFINITY, mx = -INFINITY;
    if (nlive > 0) live_min_max(sh, tv, c, mn, mx);
    if (nzero > 0) {
        mn = fmin(mn, 0.0);
        mx = fmax(mx, 0.0);
    }
    double first = mn, last = mx;
    if (first == last) {  // numpy _get_outer_edges
        first -= 0.5;
        last += 0.5;
    }
    __syncthreads();
    // np.linspace(first, last, 257): arange * step + first, endpoint forced
    const double step = (last - first) / 256.0;
    for (int i = threadIdx.x; i < 257; i += kPPThreads)
        sh.he.edges[i] = i == 256 ? last : __dadd_rn(__dmul_rn((double)i, step), first);
    for (int i = threadIdx.x; i < 256; i += kPPThreads) sh.hist[i] = 0;
    if (threadIdx.x == 0) sh.nplist = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < kNB; j += kPPThreads) {
        const int cn = sh.prefix[j + 1] - sh.prefix[j];
        if (cn == 0) continue;
        const float lo = sh.bmin[j], hi = sh.bmax[j];
        bool partial = false, masked = false;
        for (int k = 0; k < c.nz; ++k) {
            if (hi < c.zx0[k] || lo > c.zx1[k]) continue;
            if (lo >= c.zx0[k] && hi <= c.zx1[k]) masked = true; else partial = true;
        }
        if (masked) continue;
        // histogram bins of the value bin's ends through the plain (monotone) composition: every live element of the
        // bin falls into [b0, b1] (live elements never pass through an exact zero, so plain == sticky for them)
        const int b0 = he_bin(sh.he, eval_ops<false>(c, c.nops, sh.he, (double)lo));
        const int b1 = he_bin(sh.he, eval_ops<false>(c, c.nops, sh.he, (double)hi));
        if (!partial && b0 == b1) {
            if (b0 >= 0) atomicAdd(&sh.hist[b0], cn);
        } else {
            sh.plist[atomicAdd(&sh.nplist, 1)] = j | ((b0 + 1) << 10) | ((b1 + 1) << 19) | (partial ? 1 << 28 : 0);
        }
    }
    __syncthreads();
    const int np = sh.nplist;
    const Comp& cc = sh.cc[ci];
    const bool fast = cc.ok && !cc.has_he;
    {
        for_list_elements(sh, tv, np, [&](float x, int e) {
            const int b0 = ((e >> 10) & 511) - 1, b1 = ((e >> 19) & 511) - 1;
            if ((e >> 28) && in_zero_x(c, c.nz, x)) return;   // only bins that touch a masked interval test for it
            int hb = -2;
            if (fast && b0 >= 0) {
                // compiled map + a search restricted to [b0, b1]; a value closer to an edge than the compiled form's
                // rounding error can reach is decided by the reference operation order below
                const double ax = cc.a0 * (double)x;
                const double v = fmin(fmax(ax + cc.b0, cc.l0), cc.h0);
                int lo = b0, hi = b1 + 1;          // edges[lo] <= v, (hi == 256 or v < edges[hi])
                while (hi - lo > 1) {
                    const int m = (lo + hi) >> 1;
                    if (sh.he.edges[m] <= v) lo = m; else hi = m;
                }
                const double margin = 1e-9 * (fabs(ax) + fabs(cc.b0) + fabs(v));
                const bool near_lo = lo > 0 && v - sh.he.edges[lo] < margin;
                const bool near_hi = lo < 255 && sh.he.edges[lo + 1] - v < margin;
                if (!near_lo && !near_hi) hb = min(lo, 255);
            }
            if (hb == -2) hb = he_bin(sh.he, eval_ops<true>(c, c.nops, sh.he, (double)x));
            if (hb >= 0) atomicAdd(&sh.hist[hb], 1);
        });
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (nzero > 0) {  // the masked pixels (value 0): numpy histogram fast path for uniform bins
            const double v = 0.0, denom = last - first;
            int idx = (int)(__dmul_rn(__ddiv_rn(__dsub_rn(v, first), denom), 256.0));
            if (idx == 256) idx = 255;
            if (idx < 0) idx = 0;
            if (idx > 255) idx = 255;
            if (v < sh.he.edges[idx]) --idx;
            else if (idx != 255 && v >= sh.he.edges[idx + 1]) ++idx;
            if (idx < 0) idx = 0;
            sh.hist[idx] += nzero;
        }
        long long run = 0;
        for (int i = 0; i < 256; ++i) {
            run += sh.hist[i];
            sh.he.cdf[i] = (double)run;
        }
        const double tot = (double)run;
        for (int i = 0; i < 256; ++i) sh.he.cdf[i] = sh.he.cdf[i] / tot;
        sh.he_used = 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kPPThreads) sh.he.center[i] = (sh.he.edges[i] + sh.he.edges[i + 1]) / 2.0;
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kPPThreads)
        sh.he.slope[i] = i < 255 ? __ddiv_rn(__dsub_rn(sh.he.cdf[i + 1], sh.he.cdf[i]),
                                            __dsub_rn(sh.he.center[i + 1], sh.he.center[i]))
                                 : 0.0;
    if (threadIdx.x == 0) {
        const double st = (sh.he.center[255] - sh.he.center[0]) / 255.0;
        sh.he.inv_step = st > 0.0 ? 1.0 / st : 0.0;
    }
    __syncthreads();
    push_op(sh, tv, ci, OP_HISTEQ, 0.0, 0.0, 0.0, 0.0);
    return true;
}

// SigmaClipper._clip (preprocessing.py:735-751) on channel ci
__device__ bool sigma_clipper_stage(Shared& sh, const TileView& tv, int ci, double s_lo, double s_hi) {
    // astropy: sigma_lower = sigma_lower or sigma(=3.0): a falsy 0 falls back to 3 (App. A.1 / B#1)
    const double slo = s_lo != 0.0 ? s_lo : 3.0, shi = s_hi != 0.0 ? s_hi : 3.0;
    double lo, hi, mean, sd;
    if (!sigma_clip(sh, tv, sh.ch[ci], sh.cc[ci], slo, shi, lo, hi, mean, sd)) return false;
    push_op(sh, tv, ci, OP_CLAMP, lo, hi, 0.0, 0.0);
    return true;
}

// sigma clipping of channel ci over the pixels outside (sub 0) / inside (sub 1) the box: the view is switched, the
// channel's masked sets are re-ranked for it, and the all-pixels view is restored
__device__ bool sigma_clip_subset(Shared& sh, const TileView& tv, int ci, int sub, double sig, double& mean, double& sd) {
    set_view(sh, tv, sub);
    rerank_zero_sets(sh, tv, sh.ch[ci]);
    double lo, hi;
    const bool ok = sigma_clip(sh, tv, sh.tmp, sh.cc[ci], sig, sig, lo, hi, mean, sd);
    set_view(sh, tv, -1);
    return ok;
}
// largest live value of channel ci inside the box (AbsMaxScaler / ChanMaxScaler with use_mask_box)
__device__ bool box_max(Shared& sh, const TileView& tv, int ci, double& mx) {
    set_view(sh, tv, 1);
    rerank_zero_sets(sh, tv, sh.ch[ci]);
    double mn;
    const bool ok = live_min_max(sh, tv, sh.tmp, mn, mx);
    set_view(sh, tv, -1);
    return ok;
}

// ------------------------------------------------------------------------------------------ kernel 2: the chain

__device__ __forceinline__ bool chan_selected(int chid, int c) { return chid == -1 || chid == c; }

__global__ void __launch_bounds__(kPPThreads, 2) pp_chain_kernel(const __grid_constant__ PPParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
    const int b = blockIdx.x;
    const cy_pp_chain& chain = p.chain;
    TileView tv;
    memset(&tv, 0, sizeof(tv));
    if (chain.nstages > 0) tv = tile_view(p, b);
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) {
            sh.ch[c].nops = 0;
            sh.ch[c].hid = 0;
            sh.ch[c].nz = 0;
        }
        sh.next_hid = 1;
        sh.fail = 0;
        sh.he_used = 0;
        sh.n = 0;
        sh.view = -1;
        sh.gath_bin = -1;
        for (int c = 0; c < 3; ++c) compile_chan(sh.ch[c], sh.cc[c]);
    }
    __syncthreads();
    if (chain.nstages > 0) set_view(sh, tv, -1);
    bool ok = chain.reject_all == 0;  // uniform across the block

    // Channel de-duplication: a stage applied with equal parameters to channels that hold identical data (equal
    // history id) gives identical results, so it is computed once and the channel state is copied.  `in_hid[c]` is
    // the history id channel c had when the current stage processed it (-1: not processed).
    int in_hid[3];
    auto find_same = [&](int c) {
        for (int q = 0; q < c; ++q)
            if (in_hid[q] >= 0 && in_hid[q] == sh.ch[c].hid) return q;
        return -1;
    };

    for (int si = 0; si < chain.nstages && ok; ++si) {
        const cy_pp_stage& st = chain.st[si];
        in_hid[0] = in_hid[1] = in_hid[2] = -1;
        switch (st.type) {
            case CY_PP_BKG_SUB:      // x - clipped mean (preprocessing.py:591-658)
                for (int c = 0; c < 3 && ok; ++c) {
                    if (!chan_selected(st.chid, c)) continue;
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) {
                        copy_chan(sh, c, q);
                    } else {
                        double lo, hi, mean, sd;
                        const double sg = st.p[0] != 0 ? st.p[0] : 3.0;
                        if (st.flag) ok = sigma_clip_subset(sh, tv, c, 0, sg, mean, sd);
                        else ok = sigma_clip(sh, tv, sh.ch[c], sh.cc[c], sg, sg, lo, hi, mean, sd);
                        if (ok) push_op(sh, tv, c, OP_SUB, mean, 0, 0, 0);
                    }
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_CLIP_SHIFT:   // x - (clipmean + sigma*std), negatives -> 0 (preprocessing.py:664-717)
                for (int c = 0; c < 3 && ok; ++c) {
                    if (!chan_selected(st.chid, c)) continue;
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) {
                        copy_chan(sh, c, q);
                    } else {
                        double lo, hi, mean, sd;
                        const double sg = st.p[0] != 0 ? st.p[0] : 3.0;
                        ok = sigma_clip(sh, tv, sh.ch[c], sh.cc[c], sg, sg, lo, hi, mean, sd);
                        if (ok) push_op(sh, tv, c, OP_SHIFT, mean + st.p[0] * sd, 0, 0, 0);
                    }
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_SIGMA_CLIP:   // preprocessing.py:723-771
                for (int c = 0; c < 3 && ok; ++c) {
                    if (!chan_selected(st.chid, c)) continue;
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) copy_chan(sh, c, q);
                    else ok = sigma_clipper_stage(sh, tv, c, st.p[0], st.p[1]);
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_CHAN_RESIZE:  // the cube already has 3 channels (evaluation.py:146-154): no-op for 1 or 3
                break;
            case CY_PP_ZSCALE:       // preprocessing.py:934-971
                for (int c = 0; c < 3; ++c) {
                    int q = -1;
                    for (int k = 0; k < c; ++k)
                        if (in_hid[k] == sh.ch[c].hid && st.p[k] == st.p[c]) q = k;
                    const int hid = sh.ch[c].hid;
                    if (q >= 0) copy_chan(sh, c, q);
                    else zscale_stage(sh, p, b, tv, c, st.p[c]);
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_CHAN3: {      // preprocessing.py:1020-1072: p0 baseline, p1 low, p2 up, p3 contrast
                const int hid0 = sh.ch[0].hid, hid1 = sh.ch[1].hid;
                ok = sigma_clipper_stage(sh, tv, 0, st.p[0], st.p[2]);
                if (ok) zscale_stage(sh, p, b, tv, 0, st.p[3]);
                if (ok) {
                    const double blo = st.p[0] != 0.0 ? st.p[0] : 3.0;
                    const double llo = st.p[1] != 0.0 ? st.p[1] : 3.0;
                    if (hid0 == hid1 && blo == llo) {
                        copy_chan(sh, 1, 0);
                    } else {
                        ok = sigma_clipper_stage(sh, tv, 1, st.p[1], st.p[2]);
                        if (ok) zscale_stage(sh, p, b, tv, 1, st.p[3]);
                    }
                }
                if (ok) ok = histeq_stage(sh, p, tv, 2);
                break;
            }
            case CY_PP_HISTEQ: {     // preprocessing.py:977-1012: every channel; one table: channels must be identical
                const int hid0 = sh.ch[0].hid;
                if (sh.ch[1].hid != hid0 || sh.ch[2].hid != hid0) {
                    if (threadIdx.x == 0) sh.fail = -6;
                    ok = false;
                    break;
                }
                ok = histeq_stage(sh, p, tv, 0);
                if (ok) {
                    copy_chan(sh, 1, 0);
                    copy_chan(sh, 2, 0);
                }
                break;
            }
            case CY_PP_MINMAX:       // preprocessing.py:75-111
                for (int c = 0; c < 3 && ok; ++c) {
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) {
                        copy_chan(sh, c, q);
                    } else {
                        double mn, mx;
                        if (!live_min_max(sh, tv, sh.ch[c], mn, mx)) ok = false;  // no non-zero pixel -> None (:101-103)
                        else push_op(sh, tv, c, OP_MINMAX, mn, mx - mn, st.p[1] - st.p[0], st.p[0]);
                    }
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_ABS_MINMAX: { // preprocessing.py:116-146: one min / max over all channels
                double mn = INFINITY, mx = -INFINITY;
                bool any = false;
                for (int c = 0; c < 3; ++c) {
                    double a, z;
                    if (c > 0 && sh.ch[c].hid == sh.ch[c - 1].hid) continue;
                    if (live_min_max(sh, tv, sh.ch[c], a, z)) {
                        mn = fmin(mn, a);
                        mx = fmax(mx, z);
                        any = true;
                    }
                }
                if (!any) break;     // everything masked: the output stays all zero
                for (int c = 0; c < 3; ++c) {
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) copy_chan(sh, c, q);
                    else push_op(sh, tv, c, OP_MINMAX, mn, mx - mn, st.p[1] - st.p[0], st.p[0]);
                    in_hid[c] = hid;
                }
                break;
            }
            case CY_PP_MAX_SCALE:    // preprocessing.py:152-176: x / max of the channel
                for (int c = 0; c < 3 && ok; ++c) {
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) {
                        copy_chan(sh, c, q);
                    } else {
                        double mn, mx;
                        if (!live_min_max(sh, tv, sh.ch[c], mn, mx)) { in_hid[c] = hid; continue; }  // all masked: stays 0
                        if (!(mx > 0.0)) {   // order-reversing division: not representable as a monotone map
                            if (threadIdx.x == 0) sh.fail = -6;
                            ok = false;
                        } else {
                            push_op(sh, tv, c, OP_DIV, mx, 0, 0, 0);
                        }
                    }
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_ABS_MAX_SCALE:      // preprocessing.py:182-226: x / max over all channels (inside the box)
            case CY_PP_CHAN_MAX_SCALE: {   // preprocessing.py:232-288: x / max of channel chref; None if a channel max <= 0
                double mxs[3];
                bool have[3];
                for (int c = 0; c < 3; ++c) {
                    if (c > 0 && sh.ch[c].hid == sh.ch[c - 1].hid) {
                        mxs[c] = mxs[c - 1];
                        have[c] = have[c - 1];
                        continue;
                    }
                    double mn;
                    have[c] = st.flag ? box_max(sh, tv, c, mxs[c]) : live_min_max(sh, tv, sh.ch[c], mn, mxs[c]);
                }
                double d;
                if (st.type == CY_PP_ABS_MAX_SCALE) {
                    d = -INFINITY;
                    bool any = false;
                    for (int c = 0; c < 3; ++c)
                        if (have[c]) {
                            d = fmax(d, mxs[c]);
                            any = true;
                        }
                    if (!any) break;
                } else {
                    for (int c = 0; c < 3; ++c)
                        if (!have[c] || !(mxs[c] > 0.0) || !isfinite(mxs[c])) ok = false;   // returns None (:276-279)
                    if (!ok) break;
                    d = mxs[min(max(st.n, 0), 2)];
                }
                if (!(d > 0.0)) {
                    if (threadIdx.x == 0) sh.fail = -6;
                    ok = false;
                    break;
                }
                for (int c = 0; c < 3; ++c) {
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) copy_chan(sh, c, q);
                    else push_op(sh, tv, c, OP_DIV, d, 0, 0, 0);
                    in_hid[c] = hid;
                }
                break;
            }
            case CY_PP_MIN_SHIFT:    // preprocessing.py:294-327
            case CY_PP_NEG_FIX:      // preprocessing.py:408-440: the same for channels without a positive pixel
                for (int c = 0; c < 3 && ok; ++c) {
                    if (st.type == CY_PP_MIN_SHIFT && !chan_selected(st.chid, c)) continue;
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) {
                        copy_chan(sh, c, q);
                    } else {
                        double mn, mx;
                        if (live_min_max(sh, tv, sh.ch[c], mn, mx)) {
                            if (st.type == CY_PP_MIN_SHIFT || !(mx > 0.0)) push_op(sh, tv, c, OP_SUB, mn, 0, 0, 0);
                        } else {
                            ok = false;   // numpy: min() of an empty array raises; the reference run dies -> reject
                        }
                    }
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_SHIFT:        // preprocessing.py:333-363
                for (int c = 0; c < 3; ++c) {
                    int q = -1;
                    for (int k = 0; k < c; ++k)
                        if (in_hid[k] == sh.ch[c].hid && st.p[k] == st.p[c]) q = k;
                    const int hid = sh.ch[c].hid;
                    if (q >= 0) copy_chan(sh, c, q);
                    else push_op(sh, tv, c, OP_SUB, st.p[c], 0, 0, 0);
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_STANDARDIZE:  // preprocessing.py:369-402
                for (int c = 0; c < 3 && ok; ++c) {
                    if (!(st.p[3 + c] > 0.0)) {
                        if (threadIdx.x == 0) sh.fail = -6;
                        ok = false;
                        break;
                    }
                    int q = -1;
                    for (int k = 0; k < c; ++k)
                        if (in_hid[k] == sh.ch[c].hid && st.p[k] == st.p[c] && st.p[3 + k] == st.p[3 + c]) q = k;
                    const int hid = sh.ch[c].hid;
                    if (q >= 0) copy_chan(sh, c, q);
                    else push_op(sh, tv, c, OP_STD, st.p[c], st.p[3 + c], 0, 0);
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_LOG_STRETCH:  // preprocessing.py:480-538 (minmaxnorm form): chid = EXCLUDED channel
                for (int c = 0; c < 3 && ok; ++c) {
                    if (st.chid != -1 && c == st.chid) continue;
                    const int q = find_same(c), hid = sh.ch[c].hid;
                    if (q >= 0) {
                        copy_chan(sh, c, q);
                    } else {
                        const Chan& cch = sh.ch[c];
                        const int i0 = lower_index<true>(sh, tv, cch, cch.nops, 0, sh.n, 0.0, nullptr, nullptr);
                        const int npos = live_count(cch, i0, sh.n);
                        if (npos <= 0) {
                            ok = false;       // no positive pixel: returns None (:513-516)
                        } else {
                            const double vmin = eval_ops<true>(cch, cch.nops, sh.he,
                                                               (double)value_at(sh, tv, kth_live(cch, i0, 0)));
                            push_op(sh, tv, c, OP_LOG, log10(vmin), st.p[0], st.p[1] - st.p[0], (st.flag & 2) ? 1.0 : 0.0);
                        }
                    }
                    in_hid[c] = hid;
                }
                break;
            case CY_PP_BORDER_MASK:  // applied when the pixels are read (leading stage only; checked on the host)
                break;
            default:
                if (threadIdx.x == 0) sh.fail = -6;
                ok = false;
        }
    }

    // ---- hand the final per-channel maps to pp_final_kernel (which evaluates them per pixel, fused with the letterbox
    // resize) + the reference's degenerate-image check on rows 0..2 (evaluation.py:171-176)
    TileFinal& tf = p.fin[b];
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) {
            const int stt = sh.fail ? sh.fail : -1;
            p.status[b] = stt;
            tf.status = stt;
            tf.valid = 0;
        }
        return;
    }
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(&tf);
        const uint32_t* s_cc = reinterpret_cast<const uint32_t*>(&sh.cc[0]);
        const uint32_t* s_ch = reinterpret_cast<const uint32_t*>(&sh.ch[0]);
        const uint32_t* s_he = reinterpret_cast<const uint32_t*>(&sh.he);
        constexpr int n_cc = sizeof(Comp) * 3 / 4, n_ch = sizeof(Chan) * 3 / 4, n_he = sizeof(HistEq) / 4;
        for (int i = threadIdx.x; i < n_cc; i += kPPThreads) dst[offsetof(TileFinal, cc) / 4 + i] = s_cc[i];
        for (int i = threadIdx.x; i < n_ch; i += kPPThreads) dst[offsetof(TileFinal, ch) / 4 + i] = s_ch[i];
        if (sh.he_used)
            for (int i = threadIdx.x; i < n_he; i += kPPThreads) dst[offsetof(TileFinal, he) / 4 + i] = s_he[i];
        if (threadIdx.x == 0) {
            tf.same01 = sh.ch[0].hid == sh.ch[1].hid;
            tf.same02 = sh.ch[0].hid == sh.ch[2].hid;
            tf.same12 = sh.ch[1].hid == sh.ch[2].hid;
            tf.use_he = sh.he_used;
            tf.valid = 1;
        }
    }
    int bad = 0;
    for (int r = 0; r < 3 && r < p.Ty; ++r) {
        double mn = INFINITY, mx = -INFINITY;
        for (int i = threadIdx.x; i < p.Tx; i += kPPThreads) {
            const double x = (double)load_pixel_yx(p, b, r, i);
            for (int c = 0; c < 3; ++c) {
                const double v = eval_ops<true>(sh.ch[c], sh.ch[c].nops, sh.he, x);
                mn = fmin(mn, v);
                mx = fmax(mx, v);
            }
        }
        block_minmax(mn, mx, sh.red);
        if (mn == mx) bad = 1;
    }
    if (threadIdx.x == 0) {
        const int stt = sh.fail ? sh.fail : (bad ? -1 : 0);
        p.status[b] = stt;
        tf.status = stt;
    }
}

// ------------------------------------------------------------------------------------------ kernel 3: final maps + letterbox
//
// pp_final_kernel evaluates the three final channel maps of a tile per pixel and -- fused -- does the ultralytics
// predictor preprocess on them (LetterBox half-pixel bilinear resize with cv2's double-precision source coordinates,
// pad 114, channel reversal, /255) straight into the 16-bit NHWC(4) model input.  The fp32 HWC chain image never exists
// in HBM (it was a 3 MB write + 3 MB read per 512^2 tile); tests that want it ask for the optional `chain_out`.
// One CTA walks bands of `band_h` output rows of its tile: the raw input rows a band needs (band_h * Ty / new_h + 2)
// are staged in shared memory once, every DISTINCT channel map is evaluated over them with its parameters in registers
// (planes of fp32 in shared memory), then every thread interpolates whole output columns from the planes.
// emit_only: bands are input rows and the planes are written to chain_out as they are (parity surface, and the path of
// tiles whose bands do not fit shared memory, which then go through pp_resize_kernel).
struct FinalParams {
    PPParams p;
    void* out16;         // [B,Sh,Sw,4] bf16 / fp16
    float* out_f32;      // optional [B,3,Sh,Sw]
    int Sh, Sw, new_h, new_w, top, left;
    double scale_y, scale_x;   // 1 / (dst/src), as cv2.resize computes it
    int band_h, rows_cap, nbands;
    int f16, emit_only;
};

// source row / column of cv2.resize INTER_LINEAR for destination index d: i0, i1 and the weight of i1
__device__ __forceinline__ void src_coord(int d, double scale, int n, int& i0, int& i1, float& w) {
    const double f = ((double)d + 0.5) * scale - 0.5;
    i0 = (int)floor(f);
    w = (float)(f - (double)i0);
    if (i0 < 0) { i0 = 0; w = 0.f; }
    if (i0 >= n - 1) { i0 = n - 1; w = 0.f; }
    i1 = min(i0 + 1, n - 1);
}

static constexpr int kFinThreads = 512;

// one channel map over the staged rows: raw[] -> plane[]
__device__ void eval_plane(const TileFinal& tf, int ci, const float* __restrict__ raw, float* __restrict__ plane, int n) {
    const Comp& cc = tf.cc[ci];
    const int nz = cc.nzx;
    if (cc.ok && !cc.has_he && nz <= 2) {
        // the common case: affine + clamp with at most two masked intervals, everything in registers
        const double a = cc.a0, b = cc.b0, l = cc.l0, h = cc.h0;
        const float z00 = nz > 0 ? cc.zx0[0] : INFINITY, z01 = nz > 0 ? cc.zx1[0] : -INFINITY;
        const float z10 = nz > 1 ? cc.zx0[1] : INFINITY, z11 = nz > 1 ? cc.zx1[1] : -INFINITY;
#pragma unroll 4
        for (int i = threadIdx.x; i < n; i += kFinThreads) {
            const float x = raw[i];
            const bool masked = (x == 0.0f) || (x >= z00 && x <= z01) || (x >= z10 && x <= z11);
            const double v = fmin(fmax(fma(a, (double)x, b), l), h);
            plane[i] = masked ? 0.0f : (float)v;
        }
        return;
    }
    for (int i = threadIdx.x; i < n; i += kFinThreads) {
        const float x = raw[i];
        const bool masked = (x == 0.0f) || in_zero_x(cc, nz, x);
        plane[i] = masked ? 0.0f : (float)eval_fast(cc, tf.ch[ci], tf.he, (double)x);
    }
}

__global__ void __launch_bounds__(kFinThreads, 2) pp_final_kernel(const __grid_constant__ FinalParams r) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    TileFinal& tf = *reinterpret_cast<TileFinal*>(fin_smem);
    int* xi0 = reinterpret_cast<int*>(fin_smem + sizeof(TileFinal));           // [Sw] source column of tap 0 (-1: padding)
    int* xi1 = xi0 + r.Sw;                                                      // [Sw] source column of tap 1
    float* xw = reinterpret_cast<float*>(xi1 + r.Sw);                           // [Sw]
    int* ry0t = reinterpret_cast<int*>(xw + r.Sw);                              // [16] source rows of the band's output rows
    int* ry1t = ry0t + 16;
    float* rwt = reinterpret_cast<float*>(ry1t + 16);
    float* raw = reinterpret_cast<float*>(fin_smem + sizeof(TileFinal) + (((size_t)r.Sw * 12 + 192 + 15) & ~(size_t)15));
    const int b = blockIdx.x, tid = threadIdx.x;
    const PPParams& p = r.p;
    const int Ty = p.Ty, Tx = p.Tx;
    const int plane_sz = r.rows_cap * Tx;
    float* pl0 = raw + plane_sz;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&p.fin[b]);
        uint4* dst = reinterpret_cast<uint4*>(&tf);
        // the histogram-equalisation tables are the bulk of the state: skip them when no channel uses them
        const int n_head = (int)(offsetof(TileFinal, he) / 16), n_all = (int)(sizeof(TileFinal) / 16);
        const int n_tail0 = (int)((offsetof(TileFinal, he) + sizeof(HistEq)) / 16);
        for (int i = n_tail0 + tid; i < n_all; i += kFinThreads) dst[i] = src[i];
        __syncthreads();
        if (tf.valid) {
            for (int i = tid; i < n_head; i += kFinThreads) dst[i] = src[i];
            if (tf.use_he)
                for (int i = n_head + tid; i < n_tail0; i += kFinThreads) dst[i] = src[i];
        }
        if (!r.emit_only)
            for (int ox = tid; ox < r.Sw; ox += kFinThreads) {
                const int rx = ox - r.left;
                int a = -1, c = -1;
                float w = 0.f;
                if (rx >= 0 && rx < r.new_w) src_coord(rx, r.scale_x, Tx, a, c, w);
                xi0[ox] = a;
                xi1[ox] = c;
                xw[ox] = w;
            }
        __syncthreads();
    }
    const bool valid = tf.valid != 0;
    // distinct planes: channel c reads plane pidx[c]
    const int pidx1 = tf.same01 ? 0 : 1;
    const int pidx2 = tf.same02 ? 0 : (tf.same12 ? pidx1 : pidx1 + 1);
    const float* pc0 = pl0;
    const float* pc1 = pl0 + pidx1 * plane_sz;
    const float* pc2 = pl0 + pidx2 * plane_sz;
    float* chain = p.chain_out ? p.chain_out + (long long)b * Ty * Tx * 3 : nullptr;
    const uint32_t* gimg = p.img + (long long)p.y0[b] * p.row_stride + p.x0[b];
    for (int band = blockIdx.y; band < r.nbands; band += gridDim.y) {
        int yin0, nrows, oy0 = 0, oy1 = 0;
        if (r.emit_only) {
            yin0 = band * r.band_h;
            nrows = min(Ty, yin0 + r.band_h) - yin0;
        } else {
            oy0 = band * r.band_h;
            oy1 = min(r.Sh, oy0 + r.band_h);
            const int ry0 = max(oy0 - r.top, 0), ry1 = min(oy1 - r.top, r.new_h);   // resized-image rows [ry0, ry1)
            yin0 = 0;
            nrows = 0;
            if (ry0 < ry1) {
                int a0, a1, c0, c1;
                float w;
                src_coord(ry0, r.scale_y, Ty, a0, a1, w);
                src_coord(ry1 - 1, r.scale_y, Ty, c0, c1, w);
                yin0 = a0;
                nrows = c1 - a0 + 1;
            }
        }
        const int n = nrows * Tx;
        if (!r.emit_only && tid < oy1 - oy0) {
            const int ry = oy0 + tid - r.top;
            int y0 = -1, y1 = -1;
            float wy = 0.f;
            if (ry >= 0 && ry < r.new_h) src_coord(ry, r.scale_y, Ty, y0, y1, wy);
            ry0t[tid] = y0;
            ry1t[tid] = y1;
            rwt[tid] = wy;
        }
        // raw rows of the band -> shared memory (byte swap, NaN -> 0, border mask)
        for (int i = tid; i < n; i += kFinThreads) {
            const int y = i / Tx, x = i - y * Tx;
            float f = decode_pixel(gimg[(long long)(yin0 + y) * p.row_stride + x], p.big_endian);
            if (p.border_mask && !in_box(p, yin0 + y, x)) f = 0.0f;
            raw[i] = f;
        }
        __syncthreads();
        if (valid) {
            eval_plane(tf, 0, raw, pl0, n);
            if (!tf.same01) eval_plane(tf, 1, raw, pl0 + pidx1 * plane_sz, n);
            if (!tf.same02 && !tf.same12) eval_plane(tf, 2, raw, pl0 + pidx2 * plane_sz, n);
        } else {
            for (int i = tid; i < n; i += kFinThreads) pl0[i] = 0.0f;   // same01 / same02 are garbage: see pc* below
        }
        __syncthreads();
        const float* q0 = pc0;
        const float* q1 = valid ? pc1 : pc0;
        const float* q2 = valid ? pc2 : pc0;
        if (r.emit_only) {
            if (chain)
                for (int i = tid; i < n; i += kFinThreads) {
                    float* o = chain + ((long long)yin0 * Tx + i) * 3;
                    o[0] = q0[i];
                    o[1] = q1[i];
                    o[2] = q2[i];
                }
        } else {
            // every thread owns output columns; the row geometry is shared by the whole band
            for (int ox = tid; ox < r.Sw; ox += kFinThreads) {
                const int x0 = xi0[ox], x1 = xi1[ox];
                const float wx = xw[ox], ux = 1.f - wx;
                for (int oy = oy0; oy < oy1; ++oy) {
                    float v0 = 114.f, v1 = 114.f, v2 = 114.f;   // cv2.copyMakeBorder value
                    const int y0 = ry0t[oy - oy0];
                    if (y0 >= 0 && x0 >= 0) {
                        const int y1 = ry1t[oy - oy0];
                        const float wy = rwt[oy - oy0];
                        const float uy = 1.f - wy;
                        const int i00 = (y0 - yin0) * Tx + x0, i01 = (y0 - yin0) * Tx + x1;
                        const int i10 = (y1 - yin0) * Tx + x0, i11 = (y1 - yin0) * Tx + x1;
                        v0 = (q0[i00] * ux + q0[i01] * wx) * uy + (q0[i10] * ux + q0[i11] * wx) * wy;
                        v1 = (q1[i00] * ux + q1[i01] * wx) * uy + (q1[i10] * ux + q1[i11] * wx) * wy;
                        v2 = (q2[i00] * ux + q2[i01] * wx) * uy + (q2[i10] * ux + q2[i11] * wx) * wy;
                    }
                    // predictor.preprocess: im[..., ::-1] (channel reversal), float32, /255
                    const float m0 = v2 / 255.f, m1 = v1 / 255.f, m2 = v0 / 255.f;
                    const long long idx = ((long long)b * r.Sh + oy) * r.Sw + ox;
                    uint2 pk;
                    pk.x = pack_h2(m0, m1, r.f16);
                    pk.y = pack_h2(m2, 0.f, r.f16);
                    *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(r.out16) + idx * 4) = pk;
                    if (r.out_f32) {
                        const long long plane = (long long)r.Sh * r.Sw;
                        float* of = r.out_f32 + (long long)b * 3 * plane + (long long)oy * r.Sw + ox;
                        of[0] = m0;
                        of[plane] = m1;
                        of[2 * plane] = m2;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ kernel 3b: fused final (production)
//
// pp_fused_kernel is the production form of the final pass (pp_final_kernel stays as the parity-output / fallback path).
// The round-2 profile of pp_final_kernel showed an issue-bound kernel: 4.2 M warp instructions per 512^2 tile, 67 % of the
// issue slots busy, 2.4 x the compulsory DRAM traffic.  What this kernel does differently:
//   * the geometry (source columns / rows and weights of every output column / row, the source-row range of every band)
//     is computed ONCE per call by pp_geom_kernel (fp64 source coordinates as cv2 does) and read from L1/L2;
//   * the evaluated planes live in a RING of source rows in shared memory: consecutive bands overlap by two source rows,
//     the ring keeps them, so every source pixel is evaluated exactly once per distinct plane (was 1.4 x);
//   * raw rows of the NEXT band arrive by cp.async (16 bytes per copy when the tile is 16-byte aligned) while the current
//     band is resampled: two barriers per band, no exposed global-load latency;
//   * one thread owns one output column and walks the band's rows: the horizontal interpolation of a source row is
//     reused by the next output row (1.6 instead of 4 shared-memory loads per output value at scale 0.8);
//   * affine maps: fp64 FMA, then the clamp in fp32 (rounding is monotone, so float(clamp(v)) == clamp(float(v)));
//     x / 255 by reciprocal + one FMA correction (correctly rounded).
struct FusedParams {
    PPParams p;
    void* out16;         // [B,Sh,Sw,4] bf16 / fp16
    float* out_f32;      // optional [B,3,Sh,Sw]
    int Sh, Sw, new_h, new_w, top, left;
    double scale_y, scale_x;
    int band_h, ring, rawrows, nbands, chunks, f16;
    unsigned magic_tx;   // ceil(2^32 / Tx): i / Tx == __umulhi(i, magic_tx) for i < 2^32 / Tx
    unsigned magic_tx4;  // the same for Tx / 4 (16-byte copies)
    // geometry tables (device, written by pp_geom_kernel)
    int* gx0; int* gx1; float* gxw;       // [Sw]  source columns of tap 0 / 1 (-1: padding column), weight of tap 1
    int* gy0; int* gy1; float* gwy;       // [Sh]  source rows (-1: padding row)
    int* gs0; int* gs1;                    // [Sh]  ring offsets (row % ring) * Tx of the two source rows
    int* gband;                            // [nbands][2] first / last source row a band needs (first > last: none)
};

__global__ void pp_geom_kernel(const FusedParams r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r.Sw) {
        const int rx = i - r.left;
        int a = -1, c = -1;
        float w = 0.f;
        if (rx >= 0 && rx < r.new_w) src_coord(rx, r.scale_x, r.p.Tx, a, c, w);
        r.gx0[i] = a; r.gx1[i] = c; r.gxw[i] = w;
    }
    if (i < r.Sh) {
        const int ry = i - r.top;
        int a = -1, c = -1;
        float w = 0.f;
        if (ry >= 0 && ry < r.new_h) src_coord(ry, r.scale_y, r.p.Ty, a, c, w);
        r.gy0[i] = a; r.gy1[i] = c; r.gwy[i] = w;
        r.gs0[i] = a >= 0 ? (a % r.ring) * r.p.Tx : 0;
        r.gs1[i] = c >= 0 ? (c % r.ring) * r.p.Tx : 0;
    }
    if (i < r.nbands) {
        const int oy0 = i * r.band_h, oy1 = min(r.Sh, oy0 + r.band_h);
        const int ry0 = max(oy0 - r.top, 0), ry1 = min(oy1 - r.top, r.new_h);
        int lo = 1, hi = 0;
        if (ry0 < ry1) {
            int a0, a1, c0, c1;
            float w;
            src_coord(ry0, r.scale_y, r.p.Ty, a0, a1, w);
            src_coord(ry1 - 1, r.scale_y, r.p.Ty, c0, c1, w);
            lo = a0; hi = c1;
        }
        r.gband[2 * i] = lo;
        r.gband[2 * i + 1] = hi;
    }
}

__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// x / 255 correctly rounded: q = x * r, then one FMA correction with the exact remainder
__device__ __forceinline__ float div255(float x) {
    const float rcp = 1.0f / 255.0f;
    const float q = x * rcp;
    const float e = fmaf(-q, 255.0f, x);
    return fmaf(e, rcp, q);
}

static constexpr int kFuMaxThreads = 640;

__global__ void __launch_bounds__(kFuMaxThreads, 2) pp_fused_kernel(const __grid_constant__ FusedParams r) {
    extern __shared__ __align__(16) unsigned char fu_smem[];
    TileFinal& tf = *reinterpret_cast<TileFinal*>(fu_smem);
    int* rowtab = reinterpret_cast<int*>(fu_smem + sizeof(TileFinal));     // [16][3]: ring offsets of the two source rows, wy
    float* hef = reinterpret_cast<float*>(fu_smem + sizeof(TileFinal) + 256);   // [768] float tables of the equalisation
    uint32_t* raw = reinterpret_cast<uint32_t*>(fu_smem + sizeof(TileFinal) + 256 + 3072);
    const PPParams& p = r.p;
    const int Ty = p.Ty, Tx = p.Tx, NT = blockDim.x, tid = threadIdx.x, b = blockIdx.x;
    float* planes = reinterpret_cast<float*>(raw + (size_t)r.rawrows * Tx);
    const int plane_sz = r.ring * Tx;
    {   // per-tile state (the histogram-equalisation tables are the bulk: skipped when no channel uses them)
        const uint4* src = reinterpret_cast<const uint4*>(&p.fin[b]);
        uint4* dst = reinterpret_cast<uint4*>(&tf);
        const int n_head = (int)(offsetof(TileFinal, he) / 16), n_all = (int)(sizeof(TileFinal) / 16);
        const int n_tail0 = (int)((offsetof(TileFinal, he) + sizeof(HistEq)) / 16);
        for (int i = n_tail0 + tid; i < n_all; i += NT) dst[i] = src[i];
        __syncthreads();
        if (tf.valid) {
            for (int i = tid; i < n_head; i += NT) dst[i] = src[i];
            if (tf.use_he)
                for (int i = n_head + tid; i < n_tail0; i += NT) dst[i] = src[i];
        }
        __syncthreads();
    }
    const bool valid = tf.valid != 0;
    if (valid && tf.use_he) {              // float copies of the interpolation tables: centres, slopes, cdf
        for (int i = tid; i < 256; i += NT) {
            hef[i] = (float)(tf.he.center[i] - tf.he.center[0]);
            hef[256 + i] = (float)tf.he.slope[i];
            hef[512 + i] = (float)tf.he.cdf[i];
        }
    }
    // distinct planes: channel c reads plane pidx[c]
    const int pidx1 = (!valid || tf.same01) ? 0 : 1;
    const int pidx2 = (!valid || tf.same02) ? 0 : (tf.same12 ? pidx1 : pidx1 + 1);
    const int np = max(pidx1, pidx2) + 1;
    int chan_of_plane[3] = {0, 0, 0};     // a channel whose map plane q holds
    chan_of_plane[pidx1] = pidx1 ? 1 : 0;
    if (pidx2 > pidx1) chan_of_plane[pidx2] = 2;
    const long long g_row0 = (long long)p.y0[b] * p.row_stride + p.x0[b];
    const uint32_t* gimg = p.img + g_row0;
    const bool vec16 = ((reinterpret_cast<uintptr_t>(p.img) & 15) == 0) && ((p.row_stride & 3) == 0) && ((g_row0 & 3) == 0) &&
                       ((Tx & 3) == 0);

    const int bnd0 = (int)(((long long)blockIdx.y * r.nbands) / gridDim.y);
    const int bnd1 = (int)(((long long)(blockIdx.y + 1) * r.nbands) / gridDim.y);
    int have = -1;                         // source rows < have are in the ring (as far as the ring reaches back)
    // rows the next Phase A has to evaluate: [nlo, nhi]; staged in `raw` by cp.async
    int nlo = 1, nhi = 0;
    auto stage_rows = [&](int band) {
        nlo = 1; nhi = 0;
        if (band < bnd1) {
            const int lo = __ldg(&r.gband[2 * band]), hi = __ldg(&r.gband[2 * band + 1]);
            if (lo <= hi) {
                nlo = max(lo, have);
                nhi = hi;
                if (have < 0) nlo = lo;
            }
        }
        if (nlo > nhi) return;
        const int m = nhi - nlo + 1;
        if (vec16) {
            const int w4 = Tx >> 2, n4 = m * w4;
            for (int i = tid; i < n4; i += NT) {
                const int q = (int)__umulhi((unsigned)i, r.magic_tx4);
                const int x4 = i - q * w4;
                cp_async16(raw + (size_t)q * Tx + 4 * x4, gimg + (long long)(nlo + q) * p.row_stride + 4 * x4);
            }
        } else {
            const int n = m * Tx;
            for (int i = tid; i < n; i += NT) {
                const int q = (int)__umulhi((unsigned)i, r.magic_tx);
                const int x = i - q * Tx;
                cp_async4(raw + i, gimg + (long long)(nlo + q) * p.row_stride + x);
            }
        }
    };
    stage_rows(bnd0);
    cp_async_commit();

    // my output column (NT >= Sw in the common case: the column geometry stays in registers)
    for (int band = bnd0; band < bnd1; ++band) {
        cp_async_wait_all();
        __syncthreads();                   // raw rows of this band visible; Phase B of the previous band has finished
        if (tid < r.band_h) {              // row geometry of this band's output rows (read in Phase B)
            const int oy = band * r.band_h + tid;
            int o0 = -1, o1 = -1, wyb = 0;
            if (oy < r.Sh && __ldg(&r.gy0[oy]) >= 0) {
                o0 = __ldg(&r.gs0[oy]);
                o1 = __ldg(&r.gs1[oy]);
                wyb = __float_as_int(__ldg(&r.gwy[oy]));
            }
            rowtab[3 * tid] = o0;
            rowtab[3 * tid + 1] = o1;
            rowtab[3 * tid + 2] = wyb;
        }
        // ---- Phase A: new source rows -> planes (ring slots)
        if (nlo <= nhi) {
            const int m = nhi - nlo + 1, n = m * Tx;
            const int slot0 = nlo % r.ring;
            for (int q = 0; q < np; ++q) {
                float* pl = planes + (size_t)q * plane_sz;
                if (!valid) {
                    for (int i = tid; i < n; i += NT) {
                        const int rr = (int)__umulhi((unsigned)i, r.magic_tx);
                        int slot = slot0 + rr;
                        if (slot >= r.ring) slot -= r.ring;
                        pl[slot * Tx + (i - rr * Tx)] = 0.0f;
                    }
                    continue;
                }
                const int ci = chan_of_plane[q];
                const Comp& cc = tf.cc[ci];
                const int nz = cc.nzx;
                if (cc.ok && nz <= 2) {
                    // affine + clamp (fp64 FMA, clamp in fp32) with at most two masked intervals in registers; a
                    // histogram equalisation in the middle is interpolated in fp32 from float copies of the tables
                    // (np.interp is continuous: a bracket that is off by one ulp of v changes nothing measurable)
                    const bool he = cc.has_he != 0;
                    // with an equalisation the first map is taken relative to the first histogram centre (folded into
                    // the fp64 FMA), so the fp32 interpolation works on offsets of at most 256 bin widths
                    const double c0d = he ? tf.he.center[0] : 0.0;
                    const double a = cc.a0, bb = cc.b0 - c0d;
                    const float l = (float)(cc.l0 - c0d), h = (float)(cc.h0 - c0d);
                    const float z00 = nz > 0 ? cc.zx0[0] : INFINITY, z01 = nz > 0 ? cc.zx1[0] : -INFINITY;
                    const float z10 = nz > 1 ? cc.zx0[1] : INFINITY, z11 = nz > 1 ? cc.zx1[1] : -INFINITY;
                    const float a1 = (float)cc.a1, b1 = (float)cc.b1, l1 = (float)cc.l1, h1 = (float)cc.h1;
                    const float hstep = (float)tf.he.inv_step;
#pragma unroll 4
                    for (int i = tid; i < n; i += NT) {
                        const int rr = (int)__umulhi((unsigned)i, r.magic_tx);
                        const int x = i - rr * Tx;
                        int slot = slot0 + rr;
                        if (slot >= r.ring) slot -= r.ring;
                        float f = decode_pixel(raw[i], p.big_endian);
                        if (p.border_mask && !in_box(p, nlo + rr, x)) f = 0.0f;
                        const bool masked = (f == 0.0f) || (f >= z00 && f <= z01) || (f >= z10 && f <= z11);
                        float v = fminf(fmaxf((float)fma(a, (double)f, bb), l), h);
                        if (he) {
                            // np.interp(v, centers, cdf): bracket guess from the uniform spacing, corrected by <= 2 steps
                            int lo = min(254, max(0, (int)(v * hstep)));
                            lo -= (lo > 0 && hef[lo] > v);
                            lo -= (lo > 0 && hef[lo] > v);
                            lo += (lo < 254 && hef[lo + 1] <= v);
                            lo += (lo < 254 && hef[lo + 1] <= v);
                            float u = fmaf(hef[256 + lo], v - hef[lo], hef[512 + lo]);
                            if (!(v > 0.0f)) u = hef[512];
                            if (v >= hef[255]) u = hef[512 + 255];
                            v = fminf(fmaxf(fmaf(a1, u, b1), l1), h1);
                        }
                        pl[slot * Tx + x] = masked ? 0.0f : v;
                    }
                } else {
                    for (int i = tid; i < n; i += NT) {
                        const int rr = (int)__umulhi((unsigned)i, r.magic_tx);
                        const int x = i - rr * Tx;
                        int slot = slot0 + rr;
                        if (slot >= r.ring) slot -= r.ring;
                        float f = decode_pixel(raw[i], p.big_endian);
                        if (p.border_mask && !in_box(p, nlo + rr, x)) f = 0.0f;
                        const bool masked = (f == 0.0f) || in_zero_x(cc, nz, f);
                        pl[slot * Tx + x] = masked ? 0.0f : (float)eval_fast(cc, tf.ch[ci], tf.he, (double)f);
                    }
                }
            }
            have = nhi + 1;
        }
        __syncthreads();                   // planes ready; raw free
        stage_rows(band + 1);
        cp_async_commit();
        // ---- Phase B: one thread = one output column, walking the band's output rows (row geometry from shared memory)
        const int oy0 = band * r.band_h, oy1 = min(r.Sh, oy0 + r.band_h);
        const float* P0 = planes;
        const float* P1 = planes + (size_t)(np > 1 ? 1 : 0) * plane_sz;
        const float* P2 = planes + (size_t)(np > 2 ? 2 : 0) * plane_sz;
        const float pad = div255(114.f);              // cv2.copyMakeBorder value, /255
        for (int ox = tid; ox < r.Sw; ox += NT) {
            const int x0 = __ldg(&r.gx0[ox]), x1 = __ldg(&r.gx1[ox]);
            const float wx = __ldg(&r.gxw[ox]), ux = 1.f - wx;
            unsigned short* out = reinterpret_cast<unsigned short*>(r.out16) + (((long long)b * r.Sh + oy0) * r.Sw + ox) * 4;
            for (int k = 0; k < oy1 - oy0; ++k, out += (long long)r.Sw * 4) {
                float m0 = pad, m1 = pad, m2 = pad;
                const int o0 = rowtab[3 * k];
                if (o0 >= 0 && x0 >= 0) {
                    const int o1 = rowtab[3 * k + 1];
                    const float wy = __int_as_float(rowtab[3 * k + 2]), uy = 1.f - wy;
                    const int i00 = o0 + x0, i01 = o0 + x1, i10 = o1 + x0, i11 = o1 + x1;
                    const float t0 = (P0[i00] * ux + P0[i01] * wx) * uy + (P0[i10] * ux + P0[i11] * wx) * wy;
                    float t1 = t0, t2 = t0;
                    if (np > 1) t1 = (P1[i00] * ux + P1[i01] * wx) * uy + (P1[i10] * ux + P1[i11] * wx) * wy;
                    if (np > 2) t2 = (P2[i00] * ux + P2[i01] * wx) * uy + (P2[i10] * ux + P2[i11] * wx) * wy;
                    const float v1 = pidx1 == 0 ? t0 : t1;
                    const float v2 = pidx2 == 0 ? t0 : (pidx2 == 1 ? t1 : t2);
                    // predictor.preprocess: im[..., ::-1] (channel reversal), float32, /255
                    m0 = div255(v2);
                    m1 = div255(v1);
                    m2 = div255(t0);
                }
                uint2 pk;
                pk.x = pack_h2(m0, m1, r.f16);
                pk.y = pack_h2(m2, 0.f, r.f16);
                *reinterpret_cast<uint2*>(out) = pk;
                if (r.out_f32) {
                    const long long plane = (long long)r.Sh * r.Sw;
                    float* of = r.out_f32 + (long long)b * 3 * plane + (long long)(oy0 + k) * r.Sw + ox;
                    of[0] = m0;
                    of[plane] = m1;
                    of[2 * plane] = m2;
                }
            }
        }
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------ kernel 3c: fused final, interleaved planes
//
// pp_fused4_kernel: the same pass as pp_fused_kernel with half the instructions.  The second profile (r02k) showed
// pp_fused_kernel issue-bound at 81 % of the issue slots: 63 instructions per pixel and plane in Phase A (the pixel was
// decoded and indexed once per plane) and 95 per output pixel in Phase B (twelve 4-byte shared-memory loads with their
// own addresses).  Here the three maps of a pixel are evaluated in ONE pass (per-plane constants in registers) and stored
// as one float4 (p0, p1, p2, -) in the ring, so a bilinear tap is ONE 16-byte load; a thread owns up to five output
// columns (their geometry in registers) and every fourth row of the band.
static constexpr int kFu4Threads = 512;

struct PlaneK {              // constants of one distinct plane (registers)
    double a, b;             // first affine map (relative to the first histogram centre when he)
    float l, h;
    float z00, z01, z10, z11;
    float a1, b1, l1, h1;    // second affine map (after the equalisation)
    int mode;                // 0: affine + clamp, 1: + equalisation, 2: generic (eval_fast)
};

__device__ __forceinline__ float plane_value(const PlaneK& k, const TileFinal& tf, int ci, const float* __restrict__ hef,
                                             float hstep, float f) {
    if (k.mode == 2) {
        const Comp& cc = tf.cc[ci];
        const bool masked = (f == 0.0f) || in_zero_x(cc, cc.nzx, f);
        return masked ? 0.0f : (float)eval_fast(cc, tf.ch[ci], tf.he, (double)f);
    }
    const bool masked = (f == 0.0f) || (f >= k.z00 && f <= k.z01) || (f >= k.z10 && f <= k.z11);
    float v = fminf(fmaxf((float)fma(k.a, (double)f, k.b), k.l), k.h);
    if (k.mode == 1) {
        int lo = min(254, max(0, (int)(v * hstep)));
        lo -= (lo > 0 && hef[lo] > v);
        lo -= (lo > 0 && hef[lo] > v);
        lo += (lo < 254 && hef[lo + 1] <= v);
        lo += (lo < 254 && hef[lo + 1] <= v);
        float u = fmaf(hef[256 + lo], v - hef[lo], hef[512 + lo]);
        if (!(v > 0.0f)) u = hef[512];
        if (v >= hef[255]) u = hef[512 + 255];
        v = fminf(fmaxf(fmaf(k.a1, u, k.b1), k.l1), k.h1);
    }
    return masked ? 0.0f : v;
}

__device__ __forceinline__ PlaneK make_plane(const TileFinal& tf, int ci, bool valid) {
    PlaneK k;
    const Comp& cc = tf.cc[ci];
    const int nz = cc.nzx;
    k.mode = (cc.ok && nz <= 2) ? (cc.has_he ? 1 : 0) : 2;
    if (!valid) k.mode = 0;
    const double c0d = (valid && k.mode == 1) ? tf.he.center[0] : 0.0;
    k.a = valid ? cc.a0 : 0.0;
    k.b = valid ? cc.b0 - c0d : 0.0;
    k.l = valid ? (float)(cc.l0 - c0d) : 0.0f;
    k.h = valid ? (float)(cc.h0 - c0d) : 0.0f;
    k.z00 = (valid && nz > 0) ? cc.zx0[0] : INFINITY;
    k.z01 = (valid && nz > 0) ? cc.zx1[0] : -INFINITY;
    k.z10 = (valid && nz > 1) ? cc.zx0[1] : INFINITY;
    k.z11 = (valid && nz > 1) ? cc.zx1[1] : -INFINITY;
    k.a1 = (float)cc.a1; k.b1 = (float)cc.b1; k.l1 = (float)cc.l1; k.h1 = (float)cc.h1;
    return k;
}

__global__ void __launch_bounds__(kFu4Threads, 2) pp_fused4_kernel(const __grid_constant__ FusedParams r) {
    extern __shared__ __align__(16) unsigned char fu_smem[];
    TileFinal& tf = *reinterpret_cast<TileFinal*>(fu_smem);
    int* rowtab = reinterpret_cast<int*>(fu_smem + sizeof(TileFinal));          // [16][3]
    float* hef = reinterpret_cast<float*>(fu_smem + sizeof(TileFinal) + 256);   // [768]
    uint32_t* raw = reinterpret_cast<uint32_t*>(fu_smem + sizeof(TileFinal) + 256 + 3072);
    const PPParams& p = r.p;
    const int Tx = p.Tx, NT = kFu4Threads, tid = threadIdx.x, b = blockIdx.x;
    float4* ring = reinterpret_cast<float4*>(raw + (size_t)r.rawrows * Tx);     // [ring rows][Tx] (p0, p1, p2, -)
    {
        const uint4* src = reinterpret_cast<const uint4*>(&p.fin[b]);
        uint4* dst = reinterpret_cast<uint4*>(&tf);
        const int n_head = (int)(offsetof(TileFinal, he) / 16), n_all = (int)(sizeof(TileFinal) / 16);
        const int n_tail0 = (int)((offsetof(TileFinal, he) + sizeof(HistEq)) / 16);
        for (int i = n_tail0 + tid; i < n_all; i += NT) dst[i] = src[i];
        __syncthreads();
        if (tf.valid) {
            for (int i = tid; i < n_head; i += NT) dst[i] = src[i];
            if (tf.use_he)
                for (int i = n_head + tid; i < n_tail0; i += NT) dst[i] = src[i];
        }
        __syncthreads();
    }
    const bool valid = tf.valid != 0;
    if (valid && tf.use_he) {
        for (int i = tid; i < 256; i += NT) {
            hef[i] = (float)(tf.he.center[i] - tf.he.center[0]);
            hef[256 + i] = (float)tf.he.slope[i];
            hef[512 + i] = (float)tf.he.cdf[i];
        }
    }
    const float hstep = (valid && tf.use_he) ? (float)tf.he.inv_step : 0.0f;
    const int pidx1 = (!valid || tf.same01) ? 0 : 1;
    const int pidx2 = (!valid || tf.same02) ? 0 : (tf.same12 ? pidx1 : pidx1 + 1);
    const int np = max(pidx1, pidx2) + 1;
    const int c1 = pidx1 ? 1 : 0, c2 = pidx2 > pidx1 ? 2 : (pidx2 ? c1 : 0);   // a channel of plane 1 / plane 2
    const PlaneK k0 = make_plane(tf, 0, valid);
    const PlaneK k1 = make_plane(tf, np > 1 ? (pidx1 == 1 ? 1 : 2) : 0, valid);
    const PlaneK k2 = make_plane(tf, np > 2 ? 2 : 0, valid);
    const int ci1 = np > 1 ? (pidx1 == 1 ? 1 : 2) : 0, ci2 = np > 2 ? 2 : 0;
    (void)c1; (void)c2;
    const long long g_row0 = (long long)p.y0[b] * p.row_stride + p.x0[b];
    const uint32_t* gimg = p.img + g_row0;
    const bool vec16 = ((reinterpret_cast<uintptr_t>(p.img) & 15) == 0) && ((p.row_stride & 3) == 0) && ((g_row0 & 3) == 0) &&
                       ((Tx & 3) == 0);
    const int bnd0 = (int)(((long long)blockIdx.y * r.nbands) / gridDim.y);
    const int bnd1 = (int)(((long long)(blockIdx.y + 1) * r.nbands) / gridDim.y);
    int have = -1, nlo = 1, nhi = 0;
    auto stage_rows = [&](int band) {
        nlo = 1; nhi = 0;
        if (band < bnd1) {
            const int lo = __ldg(&r.gband[2 * band]), hi = __ldg(&r.gband[2 * band + 1]);
            if (lo <= hi) {
                nlo = have < 0 ? lo : max(lo, have);
                nhi = hi;
            }
        }
        if (nlo > nhi) return;
        const int m = nhi - nlo + 1;
        if (vec16) {
            const int w4 = Tx >> 2, n4 = m * w4;
            for (int i = tid; i < n4; i += NT) {
                const int q = (int)__umulhi((unsigned)i, r.magic_tx4);
                const int x4 = i - q * w4;
                cp_async16(raw + (size_t)q * Tx + 4 * x4, gimg + (long long)(nlo + q) * p.row_stride + 4 * x4);
            }
        } else {
            const int n = m * Tx;
            for (int i = tid; i < n; i += NT) {
                const int q = (int)__umulhi((unsigned)i, r.magic_tx);
                cp_async4(raw + i, gimg + (long long)(nlo + q) * p.row_stride + (i - q * Tx));
            }
        }
    };
    stage_rows(bnd0);
    cp_async_commit();

    // Phase B ownership: thread = (column lane tid & 127, row group tid >> 7); columns lane, lane + 128, ... (<= 5 kept in
    // registers), rows group, group + 4, ...
    const int lane_c = tid & 127, grp = tid >> 7;
    constexpr int kCols = 5;
    int cx0[kCols], cx1[kCols];
    float cwx[kCols];
#pragma unroll
    for (int c = 0; c < kCols; ++c) {
        const int ox = lane_c + 128 * c;
        cx0[c] = -2;                       // -2: no such column, -1: padding column
        cx1[c] = 0;
        cwx[c] = 0.f;
        if (ox < r.Sw) {
            cx0[c] = __ldg(&r.gx0[ox]);
            cx1[c] = __ldg(&r.gx1[ox]);
            cwx[c] = __ldg(&r.gxw[ox]);
        }
    }
    const float pad = div255(114.f);
    for (int band = bnd0; band < bnd1; ++band) {
        cp_async_wait_all();
        __syncthreads();
        if (tid < r.band_h) {
            const int oy = band * r.band_h + tid;
            int o0 = -1, o1 = -1, wyb = 0;
            if (oy < r.Sh && __ldg(&r.gy0[oy]) >= 0) {
                o0 = __ldg(&r.gs0[oy]);
                o1 = __ldg(&r.gs1[oy]);
                wyb = __float_as_int(__ldg(&r.gwy[oy]));
            }
            rowtab[3 * tid] = o0;
            rowtab[3 * tid + 1] = o1;
            rowtab[3 * tid + 2] = wyb;
        }
        // ---- Phase A: every new source pixel once, its three plane values as one float4
        if (nlo <= nhi) {
            const int n = (nhi - nlo + 1) * Tx;
            const int slot0 = nlo % r.ring;
#pragma unroll 2
            for (int i = tid; i < n; i += NT) {
                const int rr = (int)__umulhi((unsigned)i, r.magic_tx);
                const int x = i - rr * Tx;
                int slot = slot0 + rr;
                if (slot >= r.ring) slot -= r.ring;
                float f = decode_pixel(raw[i], p.big_endian);
                if (p.border_mask && !in_box(p, nlo + rr, x)) f = 0.0f;
                float4 v;
                v.x = plane_value(k0, tf, 0, hef, hstep, f);
                v.y = np > 1 ? plane_value(k1, tf, ci1, hef, hstep, f) : v.x;
                v.z = np > 2 ? plane_value(k2, tf, ci2, hef, hstep, f) : v.x;
                v.w = 0.f;
                ring[slot * Tx + x] = v;
            }
            have = nhi + 1;
        }
        __syncthreads();
        stage_rows(band + 1);
        cp_async_commit();
        // ---- Phase B
        const int oy0 = band * r.band_h, nrows = min(r.Sh, oy0 + r.band_h) - oy0;
        for (int k = grp; k < nrows; k += kFu4Threads / 128) {
            const int o0 = rowtab[3 * k], o1 = rowtab[3 * k + 1];
            const float wy = __int_as_float(rowtab[3 * k + 2]), uy = 1.f - wy;
            unsigned short* orow = reinterpret_cast<unsigned short*>(r.out16) + (((long long)b * r.Sh + oy0 + k) * r.Sw) * 4;
#pragma unroll
            for (int c = 0; c < kCols; ++c) {
                if (cx0[c] == -2) continue;
                const int ox = lane_c + 128 * c;
                float m0 = pad, m1 = pad, m2 = pad;
                if (o0 >= 0 && cx0[c] >= 0) {
                    const float wx = cwx[c], ux = 1.f - wx;
                    const float4 q00 = ring[o0 + cx0[c]], q01 = ring[o0 + cx1[c]];
                    const float4 q10 = ring[o1 + cx0[c]], q11 = ring[o1 + cx1[c]];
                    const float t0 = (q00.x * ux + q01.x * wx) * uy + (q10.x * ux + q11.x * wx) * wy;
                    float t1 = t0, t2 = t0;
                    if (np > 1) t1 = (q00.y * ux + q01.y * wx) * uy + (q10.y * ux + q11.y * wx) * wy;
                    if (np > 2) t2 = (q00.z * ux + q01.z * wx) * uy + (q10.z * ux + q11.z * wx) * wy;
                    const float v1 = pidx1 == 0 ? t0 : t1;
                    const float v2 = pidx2 == 0 ? t0 : (pidx2 == 1 ? t1 : t2);
                    m0 = div255(v2);      // predictor.preprocess: channel reversal, /255
                    m1 = div255(v1);
                    m2 = div255(t0);
                }
                uint2 pk;
                pk.x = pack_h2(m0, m1, r.f16);
                pk.y = pack_h2(m2, 0.f, r.f16);
                *reinterpret_cast<uint2*>(orow + (long long)ox * 4) = pk;
                if (r.out_f32) {
                    const long long plane = (long long)r.Sh * r.Sw;
                    float* of = r.out_f32 + (long long)b * 3 * plane + (long long)(oy0 + k) * r.Sw + ox;
                    of[0] = m0;
                    of[plane] = m1;
                    of[2 * plane] = m2;
                }
            }
        }
        // columns beyond 5 x 128 (Sw > 640): plain loop
        for (int ox = 128 * kCols + lane_c; ox < r.Sw; ox += 128) {
            const int x0 = __ldg(&r.gx0[ox]), x1 = __ldg(&r.gx1[ox]);
            const float wx = __ldg(&r.gxw[ox]), ux = 1.f - wx;
            for (int k = grp; k < nrows; k += kFu4Threads / 128) {
                const int o0 = rowtab[3 * k], o1 = rowtab[3 * k + 1];
                const float wy = __int_as_float(rowtab[3 * k + 2]), uy = 1.f - wy;
                float m0 = pad, m1 = pad, m2 = pad;
                if (o0 >= 0 && x0 >= 0) {
                    const float4 q00 = ring[o0 + x0], q01 = ring[o0 + x1], q10 = ring[o1 + x0], q11 = ring[o1 + x1];
                    const float t0 = (q00.x * ux + q01.x * wx) * uy + (q10.x * ux + q11.x * wx) * wy;
                    const float t1 = (q00.y * ux + q01.y * wx) * uy + (q10.y * ux + q11.y * wx) * wy;
                    const float t2 = (q00.z * ux + q01.z * wx) * uy + (q10.z * ux + q11.z * wx) * wy;
                    m0 = div255(pidx2 == 0 ? t0 : (pidx2 == 1 ? t1 : t2));
                    m1 = div255(pidx1 == 0 ? t0 : t1);
                    m2 = div255(t0);
                }
                uint2 pk;
                pk.x = pack_h2(m0, m1, r.f16);
                pk.y = pack_h2(m2, 0.f, r.f16);
                *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(r.out16) +
                                          (((long long)b * r.Sh + oy0 + k) * r.Sw + ox) * 4) = pk;
                if (r.out_f32) {
                    const long long plane = (long long)r.Sh * r.Sw;
                    float* of = r.out_f32 + (long long)b * 3 * plane + (long long)(oy0 + k) * r.Sw + ox;
                    of[0] = m0;
                    of[plane] = m1;
                    of[2 * plane] = m2;
                }
            }
        }
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------ letterbox resize of an HWC image

struct ResizeParams {
    const float* chain;  // [B,Ty,Tx,3]
    __nv_bfloat16* out;  // [B,Sh,Sw,4]
    float* out_f32;      // optional [B,3,Sh,Sw]
    int B, Ty, Tx, Sh, Sw;
    int new_h, new_w, top, left;
    double scale_y, scale_x;  // 1 / (dst/src), as cv2.resize computes it
    int f16;                  // 16-bit output format: 0 bf16, 1 fp16
};

__global__ void __launch_bounds__(256) pp_resize_kernel(const ResizeParams r) {
    // grid = (column blocks, output rows, tiles): no 64-bit div/mod per pixel
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y, b = blockIdx.z;
    if (ox >= r.Sw) return;
    const long long idx = ((long long)b * r.Sh + oy) * r.Sw + ox;
    float v[3] = {114.f, 114.f, 114.f};  // cv2.copyMakeBorder value
    const int ry = oy - r.top, rx = ox - r.left;
    if (ry >= 0 && ry < r.new_h && rx >= 0 && rx < r.new_w) {
        int y0, y1, x0, x1;
        float wy, wx;
        src_coord(ry, r.scale_y, r.Ty, y0, y1, wy);
        src_coord(rx, r.scale_x, r.Tx, x0, x1, wx);
        const float* base = r.chain + (long long)b * r.Ty * r.Tx * 3;
        const float* p00 = base + ((long long)y0 * r.Tx + x0) * 3;
        const float* p01 = base + ((long long)y0 * r.Tx + x1) * 3;
        const float* p10 = base + ((long long)y1 * r.Tx + x0) * 3;
        const float* p11 = base + ((long long)y1 * r.Tx + x1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t0 = __ldg(p00 + c) * (1.f - wx) + __ldg(p01 + c) * wx;
            const float t1 = __ldg(p10 + c) * (1.f - wx) + __ldg(p11 + c) * wx;
            v[c] = t0 * (1.f - wy) + t1 * wy;
        }
    }
    // predictor.preprocess: im[..., ::-1] (channel reversal), float32, /255
    const float m0 = v[2] / 255.f, m1 = v[1] / 255.f, m2 = v[0] / 255.f;
    uint2 pk;
    pk.x = pack_h2(m0, m1, r.f16);
    pk.y = pack_h2(m2, 0.f, r.f16);
    *reinterpret_cast<uint2*>(r.out + idx * 4) = pk;
    if (r.out_f32) {
        const long long plane = (long long)r.Sh * r.Sw;
        float* o = r.out_f32 + (long long)b * 3 * plane + (long long)oy * r.Sw + ox;
        o[0] = m0;
        o[plane] = m1;
        o[2 * plane] = m2;
    }
}

}  // namespace cy

// ------------------------------------------------------------------------------------------ C ABI

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void add_stage(cy_pp_chain* ch, int type, int chid, int flag, int n, double p0 = 0, double p1 = 0, double p2 = 0,
                      double p3 = 0) {
    if (ch->nstages >= CY_PP_MAX_STAGES) return;
    cy_pp_stage& s = ch->st[ch->nstages++];
    memset(&s, 0, sizeof(s));
    s.type = type; s.chid = chid; s.flag = flag; s.n = n;
    s.p[0] = p0; s.p[1] = p1; s.p[2] = p2; s.p[3] = p3;
}

extern "C" int cy_pp_chain_from_config(const cy_pp_config* cfg, cy_pp_chain* ch) {
    if (!cfg || !ch) return cy::set_error(CY_ERR_INVALID, "cy_pp_chain_from_config: null argument");
    memset(ch, 0, sizeof(*ch));
    ch->out_f16 = cfg->out_f16 ? 1 : 0;
    if (!cfg->enabled) return CY_OK;
    // scripts/run.py:272-293
    if (cfg->subtract_bkg)
        add_stage(ch, CY_PP_BKG_SUB, cfg->bkg_chid, cfg->use_box_mask_in_bkg ? 1 : 0, 0, cfg->sigma_bkg, cfg->bkg_box_mask_fract);
    if (cfg->clip_shift_data) add_stage(ch, CY_PP_CLIP_SHIFT, cfg->clip_chid, 0, 0, cfg->sigma_clip);
    if (cfg->clip_data) add_stage(ch, CY_PP_SIGMA_CLIP, cfg->clip_chid, 0, 0, cfg->sigma_clip_low, cfg->sigma_clip_up);
    if (cfg->nchannels > 1) add_stage(ch, CY_PP_CHAN_RESIZE, -1, 0, cfg->nchannels);
    if (cfg->zscale_stretch)
        add_stage(ch, CY_PP_ZSCALE, -1, 0, 3, cfg->zscale_contrasts[0], cfg->zscale_contrasts[1], cfg->zscale_contrasts[2]);
    if (cfg->chan3_preproc)
        add_stage(ch, CY_PP_CHAN3, -1, 0, 0, cfg->sigma_clip_baseline, cfg->sigma_clip_low, cfg->sigma_clip_up,
                  cfg->zscale_contrasts[0]);
    if (cfg->normalize_minmax) add_stage(ch, CY_PP_MINMAX, -1, 0, 0, cfg->norm_min, cfg->norm_max);
    if (cfg->nchannels != 1 && cfg->nchannels != 3)
        return cy::set_error(CY_ERR_INVALID, "nchannels must be 1 or 3 (the model takes 3-channel images)");
    if (cfg->chan3_preproc && cfg->nchannels != 3)
        return cy::set_error(CY_ERR_INVALID, "chan3_preproc requires nchannels == 3 (scripts/run.py:253-256)");
    return cy_pp_chain_validate(ch);
}

static bool stage_uses_box_stats(const cy_pp_stage& s) {
    return (s.type == CY_PP_BKG_SUB || s.type == CY_PP_ABS_MAX_SCALE || s.type == CY_PP_CHAN_MAX_SCALE) && s.flag;
}
static double stage_box_fract(const cy_pp_stage& s) {
    return s.type == CY_PP_BKG_SUB ? s.p[1] : s.p[0];
}

extern "C" int cy_pp_chain_validate(cy_pp_chain* ch) {
    using namespace cy;
    if (!ch) return set_error(CY_ERR_INVALID, "cy_pp_chain_validate: null argument");
    if (ch->nstages < 0 || ch->nstages > CY_PP_MAX_STAGES)
        return set_error(CY_ERR_INVALID, "a chain holds at most %d stages", CY_PP_MAX_STAGES);
    ch->reject_all = 0;
    bool have_box = false, stats_seen = false;
    double fract = 0;
    int nhe = 0;
    for (int i = 0; i < ch->nstages; ++i) {
        const cy_pp_stage& s = ch->st[i];
        if (s.chid < -1 || s.chid > 2) return set_error(CY_ERR_INVALID, "stage %d: chid must be -1, 0, 1 or 2", i);
        const bool box_stage = stage_uses_box_stats(s) || s.type == CY_PP_BORDER_MASK;
        if (box_stage) {
            const double f = stage_box_fract(s);
            if (have_box && f != fract)
                return set_error(CY_ERR_INVALID, "stage %d: one box geometry (mask_fract) per chain in this build", i);
            have_box = true;
            fract = f;
        }
        switch (s.type) {
            case CY_PP_BKG_SUB: case CY_PP_CLIP_SHIFT: case CY_PP_SIGMA_CLIP: case CY_PP_MINMAX: case CY_PP_ABS_MINMAX:
            case CY_PP_MAX_SCALE: case CY_PP_ABS_MAX_SCALE: case CY_PP_MIN_SHIFT: case CY_PP_NEG_FIX:
                stats_seen = true;
                break;
            case CY_PP_CHAN_MAX_SCALE:
                if (s.n < 0 || s.n > 2) return set_error(CY_ERR_INVALID, "stage %d: chref must be 0, 1 or 2", i);
                stats_seen = true;
                break;
            case CY_PP_CHAN_RESIZE:
                if (s.n != 1 && s.n != 3)
                    return set_error(CY_ERR_INVALID, "ChanResizer: nchans must be 1 or 3 (the model takes 3-channel images)");
                break;
            case CY_PP_ZSCALE:
                if (s.n < 3) ch->reject_all = 1;      // the reference returns None (preprocessing.py:955-957)
                stats_seen = true;
                break;
            case CY_PP_CHAN3: case CY_PP_HISTEQ:
                if (++nhe > 1)
                    return set_error(CY_ERR_INVALID, "stage %d: one histogram equalisation per chain in this build", i);
                stats_seen = true;
                break;
            case CY_PP_SHIFT: case CY_PP_STANDARDIZE:
                if (s.n != 3) ch->reject_all = 1;     // length check of the reference fails -> None (:344-348, :381-389)
                break;
            case CY_PP_LOG_STRETCH:
                if (!(s.flag & 1))
                    return set_error(CY_ERR_INVALID, "LogStretcher without minmaxnorm turns masked pixels into non-zero "
                                                     "values: not implemented in this build");
                if (!(s.p[1] > s.p[0])) return set_error(CY_ERR_INVALID, "LogStretcher: data_norm_max must exceed data_norm_min");
                stats_seen = true;
                break;
            case CY_PP_BORDER_MASK:
                if (stats_seen)
                    return set_error(CY_ERR_INVALID, "BorderMasker after a stage that computes image statistics is not "
                                                     "implemented in this build (put it first)");
                break;
            default:
                return set_error(CY_ERR_INVALID, "stage %d: unknown stage type %d", i, s.type);
        }
    }
    return CY_OK;
}

// Letterbox geometry shared by the fused final kernel and cy_letterbox_resize (ultralytics LetterBox, App. A.4).
struct LbGeom {
    int Sh, Sw, new_h, new_w, top, left;
    double scale_y, scale_x;
};
static int lb_geometry(int Ty, int Tx, int imgsz, LbGeom* g) {
    cy_letterbox lb;
    int rc = cy_letterbox_shape(Ty, Tx, imgsz, &g->Sh, &g->Sw, &lb);
    if (rc) return rc;
    const double rr = fmin((double)imgsz / Ty, (double)imgsz / Tx);
    g->new_w = (int)nearbyint(Tx * rr);
    g->new_h = (int)nearbyint(Ty * rr);
    g->top = (int)nearbyint(((imgsz - g->new_h) % 32) / 2.0 - 0.1);
    g->left = (int)nearbyint(((imgsz - g->new_w) % 32) / 2.0 - 0.1);
    g->scale_x = 1.0 / ((double)g->new_w / (double)Tx);
    g->scale_y = 1.0 / ((double)g->new_h / (double)Ty);
    return CY_OK;
}
// host copy of the kernel's source-coordinate rule
static void src_coord_host(int d, double scale, int n, int* i0, int* i1) {
    const double f = ((double)d + 0.5) * scale - 0.5;
    int a = (int)floor(f);
    if (a < 0) a = 0;
    if (a >= n - 1) a = n - 1;
    *i0 = a;
    *i1 = a + 1 < n - 1 ? a + 1 : n - 1;
}
static constexpr int kFinBandH = 8;                 // output rows per band of the fused final kernel (<= 16)
static constexpr size_t kFinSmemMax = 110 * 1024;   // two CTAs per SM

static size_t fin_smem_bytes(int Sw, int rows, int Tx) {
    return sizeof(cy::TileFinal) + (((size_t)Sw * 12 + 192 + 15) & ~(size_t)15) + (size_t)rows * Tx * 16;
}
// rows of shared memory the fused bands need (0: does not fit -> chain image + pp_resize_kernel path)
static int fin_rows_cap(const LbGeom& g, int Ty, int Tx) {
    int cap = 1;
    for (int oy0 = 0; oy0 < g.Sh; oy0 += kFinBandH) {
        const int oy1 = oy0 + kFinBandH < g.Sh ? oy0 + kFinBandH : g.Sh;
        const int ry0 = oy0 - g.top > 0 ? oy0 - g.top : 0, ry1 = oy1 - g.top < g.new_h ? oy1 - g.top : g.new_h;
        if (ry0 >= ry1) continue;
        int a0, a1, c0, c1;
        src_coord_host(ry0, g.scale_y, Ty, &a0, &a1);
        src_coord_host(ry1 - 1, g.scale_y, Ty, &c0, &c1);
        if (c1 - a0 + 1 > cap) cap = c1 - a0 + 1;
    }
    return fin_smem_bytes(g.Sw, cap, Tx) <= kFinSmemMax ? cap : 0;
}

static constexpr size_t kFuTableBytes = 256 * 1024;   // geometry tables of the fused final kernel (part of the scratch)
static constexpr size_t kFuSmemMax = 112 * 1024;      // two CTAs per SM
static constexpr size_t kFu4SmemMax = 200 * 1024;     // interleaved form: 110 KB for 512^2 tiles (two CTAs per SM)

struct BkGeom {
    int nsub, R, nchunks, nbs;
};
static int bk_geometry(const cy_pp_chain* ch, int Ty, int Tx, BkGeom* g) {
    if (Tx > cy::kChunk) return cy::set_error(CY_ERR_INVALID, "cy_preprocess: tiles wider than %d pixels are not supported", cy::kChunk);
    g->nsub = 1;
    for (int i = 0; i < ch->nstages; ++i)
        if (stage_uses_box_stats(ch->st[i])) g->nsub = 2;
    g->R = cy::kChunk / Tx;
    if (g->R > 256) g->R = 256;
    if (g->R < 1) g->R = 1;
    g->nchunks = (Ty + g->R - 1) / g->R;
    g->nbs = cy::kNB * g->nsub;
    return CY_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 pp_get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

extern "C" size_t cy_preprocess_chain_scratch_bytes(const cy_pp_chain* ch, int B, int Ty, int Tx) {
    if (!ch || B <= 0 || Ty <= 0 || Tx <= 0) return 0;
    BkGeom g;
    if (bk_geometry(ch, Ty, Tx, &g)) return 0;
    const size_t N = (size_t)Ty * Tx, b = (size_t)B;
    size_t t = align256(b * sizeof(cy::TileFinal)) + 256 + kFuTableBytes;
    if (ch->nstages > 0)
        t += align256(b * sizeof(cy::TileHdr)) + align256(b * N * 4) + align256(b * g.nchunks * (g.nbs + 1) * 2) +
             align256(b * (g.nchunks + 1) * 4) + 3 * align256(b * g.nbs * 4) + 2 * align256(b * g.nbs * 8);
    return t;
}

extern "C" size_t cy_preprocess_scratch_bytes(const cy_pp_config* cfg, int B, int Ty, int Tx) {
    cy_pp_chain ch;
    if (!cfg || cy_pp_chain_from_config(cfg, &ch)) return 0;
    return cy_preprocess_chain_scratch_bytes(&ch, B, Ty, Tx);
}

extern "C" int cy_preprocess(const cy_pp_config* cfg, const void* img, long long row_stride, int big_endian,
                             const int32_t* tile_x0, const int32_t* tile_y0, int B, int Ty, int Tx, int imgsz,
                             float* chain_out, void* model_in, float* model_in_f32, int32_t* status, void* scratch,
                             uintptr_t stream) {
    if (!cfg) return cy::set_error(CY_ERR_INVALID, "cy_preprocess: null argument");
    cy_pp_chain ch;
    int rc = cy_pp_chain_from_config(cfg, &ch);
    if (rc) return rc;
    return cy_preprocess_chain(&ch, img, row_stride, big_endian, tile_x0, tile_y0, B, Ty, Tx, imgsz, chain_out, model_in,
                               model_in_f32, status, scratch, stream);
}

extern "C" int cy_preprocess_chain(const cy_pp_chain* chain_in, const void* img, long long row_stride, int big_endian,
                                   const int32_t* tile_x0, const int32_t* tile_y0, int B, int Ty, int Tx, int imgsz,
                                   float* chain_out, void* model_in, float* model_in_f32, int32_t* status, void* scratch,
                                   uintptr_t stream) {
    using namespace cy;
    if (!chain_in || !img || !tile_x0 || !tile_y0 || !status || !scratch)
        return set_error(CY_ERR_INVALID, "cy_preprocess: null argument");
    if (!chain_out && !model_in) return set_error(CY_ERR_INVALID, "cy_preprocess: no output requested");
    if (B <= 0 || Ty <= 0 || Tx <= 0) return set_error(CY_ERR_INVALID, "cy_preprocess: invalid shape");
    if ((long long)Ty * Tx >= (1ll << 30)) return set_error(CY_ERR_INVALID, "cy_preprocess: tile too large");
    if (B > 65535) return set_error(CY_ERR_INVALID, "cy_preprocess: at most 65535 tiles per call");
    cy_pp_chain chain = *chain_in;
    int rc = cy_pp_chain_validate(&chain);
    if (rc) return rc;
    BkGeom g;
    if ((rc = bk_geometry(&chain, Ty, Tx, &g))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PPParams p;
    memset(&p, 0, sizeof(p));
    p.chain = chain;
    p.img = (const uint32_t*)img;
    p.row_stride = row_stride;
    p.big_endian = big_endian;
    p.x0 = tile_x0;
    p.y0 = tile_y0;
    p.B = B; p.Ty = Ty; p.Tx = Tx;
    p.nsub = g.nsub; p.rows_per_chunk = g.R; p.nchunks = g.nchunks; p.nbs = g.nbs;
    // box geometry (preprocessing.py:610-621): xc = int(W/2), dx = int(W * fract / 2), rows / cols [c - d, c + d)
    p.bx0 = p.by0 = 0; p.bx1 = Tx; p.by1 = Ty;
    for (int i = 0; i < chain.nstages; ++i) {
        const cy_pp_stage& s = chain.st[i];
        if (stage_uses_box_stats(s) || s.type == CY_PP_BORDER_MASK) {
            const double f = stage_box_fract(s);
            const int xc = Tx / 2, yc = Ty / 2, dy = (int)(Ty * f / 2.0), dx = (int)(Tx * f / 2.0);
            p.bx0 = xc - dx; p.bx1 = xc + dx; p.by0 = yc - dy; p.by1 = yc + dy;
            if (s.type == CY_PP_BORDER_MASK) p.border_mask = 1;
        }
    }
    const size_t N = (size_t)Ty * Tx, b = (size_t)B;
    char* s = (char*)scratch;
    p.fin = (TileFinal*)s; s += align256(b * sizeof(TileFinal));
    if (chain.nstages > 0) {
        p.hdr = (TileHdr*)s; s += align256(b * sizeof(TileHdr));
        p.vals = (float*)s; s += align256(b * N * 4);
        p.off = (unsigned short*)s; s += align256(b * g.nchunks * (g.nbs + 1) * 2);
        p.cbase = (int*)s; s += align256(b * (g.nchunks + 1) * 4);
        p.cnt = (int*)s; s += align256(b * g.nbs * 4);
        p.bmin = (float*)s; s += align256(b * g.nbs * 4);
        p.bmax = (float*)s; s += align256(b * g.nbs * 4);
        p.m1 = (double*)s; s += align256(b * g.nbs * 8);
        p.m2 = (double*)s; s += align256(b * g.nbs * 8);
    }
    char* fu_tables = s;
    s += kFuTableBytes;
    p.chain_out = chain_out;
    p.status = status;
    static std::atomic<unsigned long long> attr_done{0};
    if (first_use_on_device(attr_done)) {
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared)));
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BkSmem)));
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemMax));
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFuSmemMax));
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_fused4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFu4SmemMax));
    }
    if (chain.nstages > 0) {
        // TMA staging of the chunks: 2-D boxes of pw x R pixels of the mosaic (16-byte aligned base / pitch / box rows)
        CUtensorMap tm;
        memset(&tm, 0, sizeof(tm));
        p.pw = Tx <= 256 ? Tx : 256;
        p.npanels = Tx / p.pw;
        const char* no_tma = getenv("CY_PP_NO_TMA");
        p.use_tma = !(no_tma && atoi(no_tma)) && ((uintptr_t)img % 16 == 0) && ((row_stride * 4) % 16 == 0) && (Tx % 4 == 0) &&
                    (Tx % p.pw == 0) && Ty >= g.R && row_stride >= Tx;
        if (p.use_tma) {
            auto enc = pp_get_encode();
            if (!enc) {
                p.use_tma = 0;
            } else {
                const cuuint64_t dims[2] = {(cuuint64_t)row_stride, (cuuint64_t)1 << 30};
                const cuuint64_t str[1] = {(cuuint64_t)row_stride * 4};
                const cuuint32_t box[2] = {(cuuint32_t)p.pw, (cuuint32_t)g.R};
                const cuuint32_t estr[2] = {1, 1};
                CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)img, dims, str, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) p.use_tma = 0;
            }
        }
        if (!p.use_tma) { p.pw = Tx; p.npanels = 1; }
        pp_bucket_kernel<<<B, kBkThreads, sizeof(BkSmem), st>>>(p, tm);
    }
    pp_chain_kernel<<<B, kPPThreads, sizeof(Shared), st>>>(p);
    CY_CUDA_CHECK(cudaGetLastError());

    LbGeom lg;
    if ((rc = lb_geometry(Ty, Tx, imgsz, &lg))) return rc;
    FinalParams r;
    memset(&r, 0, sizeof(r));
    r.p = p;
    r.out16 = model_in;
    r.out_f32 = model_in_f32;
    r.Sh = lg.Sh; r.Sw = lg.Sw; r.new_h = lg.new_h; r.new_w = lg.new_w; r.top = lg.top; r.left = lg.left;
    r.scale_y = lg.scale_y; r.scale_x = lg.scale_x;
    r.f16 = chain.out_f16 ? 1 : 0;
    const int sms = current_device_sms();
    const int cap = model_in ? fin_rows_cap(lg, Ty, Tx) : 0;
    auto emit_chain = [&](float* dst) -> int {   // evaluated maps as an fp32 HWC image (parity output / resize input)
        FinalParams e = r;
        e.p.chain_out = dst;
        e.emit_only = 1;
        const int rows_fit = (int)((kFinSmemMax - sizeof(TileFinal) - 256) / ((size_t)Tx * 16));
        if (rows_fit < 1) return set_error(CY_ERR_INVALID, "cy_preprocess: tile rows of %d pixels do not fit shared memory", Tx);
        e.band_h = rows_fit < 8 ? rows_fit : 8;
        e.rows_cap = e.band_h;
        e.nbands = (Ty + e.band_h - 1) / e.band_h;
        e.Sw = 0;
        int chunks = (2 * sms + B - 1) / B;
        chunks = chunks < 1 ? 1 : (chunks > e.nbands ? e.nbands : chunks);
        pp_final_kernel<<<dim3((unsigned)B, (unsigned)chunks), kFinThreads, fin_smem_bytes(0, e.band_h, Tx), st>>>(e);
        return CY_OK;
    };
    if (chain_out && (rc = emit_chain(chain_out))) return rc;
    bool fused_done = false;
    if (model_in && cap > 0 && Tx >= 8 && !(getenv("CY_PP_OLD_FINAL") && atoi(getenv("CY_PP_OLD_FINAL")))) {
        // production path: pp_geom_kernel (once per call) + pp_fused_kernel
        FusedParams f;
        memset(&f, 0, sizeof(f));
        f.p = p;
        f.out16 = model_in;
        f.out_f32 = model_in_f32;
        f.Sh = lg.Sh; f.Sw = lg.Sw; f.new_h = lg.new_h; f.new_w = lg.new_w; f.top = lg.top; f.left = lg.left;
        f.scale_y = lg.scale_y; f.scale_x = lg.scale_x;
        f.f16 = r.f16;
        f.band_h = kFinBandH;
        f.ring = cap;
        f.rawrows = cap;
        f.nbands = (lg.Sh + kFinBandH - 1) / kFinBandH;
        f.magic_tx = (unsigned)(((1ull << 32) + (unsigned)Tx - 1) / (unsigned)Tx);
        f.magic_tx4 = (Tx % 4 == 0) ? (unsigned)(((1ull << 32) + (unsigned)(Tx / 4) - 1) / (unsigned)(Tx / 4)) : 0u;
        const size_t tab = ((size_t)lg.Sw * 12 + (size_t)lg.Sh * 20 + (size_t)f.nbands * 8 + 255) & ~(size_t)255;
        const size_t smem = sizeof(TileFinal) + 256 + 3072 + (size_t)cap * Tx * 4 * 4;
        if (tab <= kFuTableBytes && smem <= kFuSmemMax && (long long)cap * Tx < (1ll << 32) / Tx) {
            char* t = fu_tables;
            f.gx0 = (int*)t; t += (size_t)lg.Sw * 4;
            f.gx1 = (int*)t; t += (size_t)lg.Sw * 4;
            f.gxw = (float*)t; t += (size_t)lg.Sw * 4;
            f.gy0 = (int*)t; t += (size_t)lg.Sh * 4;
            f.gy1 = (int*)t; t += (size_t)lg.Sh * 4;
            f.gwy = (float*)t; t += (size_t)lg.Sh * 4;
            f.gs0 = (int*)t; t += (size_t)lg.Sh * 4;
            f.gs1 = (int*)t; t += (size_t)lg.Sh * 4;
            f.gband = (int*)t;
            int chunks = (2 * sms + B - 1) / B;
            chunks = chunks < 1 ? 1 : (chunks > f.nbands ? f.nbands : chunks);
            f.chunks = chunks;
            const int gmax = lg.Sw > lg.Sh ? lg.Sw : lg.Sh;
            pp_geom_kernel<<<(gmax + 255) / 256, 256, 0, st>>>(f);
            // interleaved-plane form (pp_fused4_kernel): staging rows = the most NEW source rows a band brings (all of
            // a band's rows when a CTA starts in the middle of the tile, i.e. chunks > 1)
            int max_new = 1, prev_hi = -1;
            for (int bnd = 0; bnd < f.nbands; ++bnd) {
                const int oy0 = bnd * kFinBandH, oy1 = oy0 + kFinBandH < lg.Sh ? oy0 + kFinBandH : lg.Sh;
                const int ry0 = oy0 - lg.top > 0 ? oy0 - lg.top : 0, ry1 = oy1 - lg.top < lg.new_h ? oy1 - lg.top : lg.new_h;
                if (ry0 >= ry1) continue;
                int a0, a1, c0, c1;
                src_coord_host(ry0, lg.scale_y, Ty, &a0, &a1);
                src_coord_host(ry1 - 1, lg.scale_y, Ty, &c0, &c1);
                const int lo = a0 > prev_hi + 1 ? a0 : prev_hi + 1;
                if (c1 - lo + 1 > max_new) max_new = c1 - lo + 1;
                prev_hi = c1;
            }
            const int raw4 = chunks > 1 ? cap : (max_new < cap ? max_new : cap);
            const size_t smem4 = sizeof(TileFinal) + 256 + 3072 + (size_t)raw4 * Tx * 4 + (size_t)cap * Tx * 16;
            const char* v1 = getenv("CY_PP_FUSED_V1");
            if (smem4 <= kFu4SmemMax && !(v1 && atoi(v1))) {
                f.rawrows = raw4;
                pp_fused4_kernel<<<dim3((unsigned)B, (unsigned)chunks), kFu4Threads, smem4, st>>>(f);
            } else {
                int nt = (lg.Sw + 31) & ~31;
                nt = nt > kFuMaxThreads ? kFuMaxThreads : (nt < 128 ? 128 : nt);
                pp_fused_kernel<<<dim3((unsigned)B, (unsigned)chunks), nt, smem, st>>>(f);
            }
            fused_done = true;
        }
    }
    if (model_in && !fused_done) {
        if (cap > 0) {
            r.emit_only = 0;
            r.band_h = kFinBandH;
            r.rows_cap = cap;
            r.nbands = (lg.Sh + kFinBandH - 1) / kFinBandH;
            int chunks = (2 * sms + B - 1) / B;
            chunks = chunks < 1 ? 1 : (chunks > r.nbands ? r.nbands : chunks);
            pp_final_kernel<<<dim3((unsigned)B, (unsigned)chunks), kFinThreads, fin_smem_bytes(lg.Sw, cap, Tx), st>>>(r);
        } else {
            // bands of this tile shape do not fit shared memory: materialise the fp32 HWC image once, then resize it
            float* hwc = chain_out;
            if (!hwc) {
                if (cudaMallocAsync(&hwc, b * N * 12, st) != cudaSuccess)
                    return set_error(CY_ERR_NOMEM, "cy_preprocess: %zu bytes for the chain image", b * N * 12);
                if ((rc = emit_chain(hwc))) return rc;
            }
            rc = cy_letterbox_resize_fmt(hwc, B, Ty, Tx, imgsz, model_in, model_in_f32, r.f16, stream);
            if (!chain_out) cudaFreeAsync(hwc, st);
            if (rc) return rc;
        }
    }
    CY_CUDA_CHECK(cudaGetLastError());
    return CY_OK;
}

extern "C" int cy_letterbox_resize(const float* chain, int B, int Ty, int Tx, int imgsz, void* model_in,
                                   float* model_in_f32, uintptr_t stream) {
    return cy_letterbox_resize_fmt(chain, B, Ty, Tx, imgsz, model_in, model_in_f32, 0, stream);
}

extern "C" int cy_letterbox_resize_fmt(const float* chain, int B, int Ty, int Tx, int imgsz, void* model_in,
                                       float* model_in_f32, int out_f16, uintptr_t stream) {
    using namespace cy;
    if (!chain || !model_in || B <= 0) return set_error(CY_ERR_INVALID, "cy_letterbox_resize: invalid argument");
    LbGeom g;
    int rc = lb_geometry(Ty, Tx, imgsz, &g);
    if (rc) return rc;
    ResizeParams r;
    r.chain = chain;
    r.out = (__nv_bfloat16*)model_in;
    r.out_f32 = model_in_f32;
    r.B = B; r.Ty = Ty; r.Tx = Tx; r.Sh = g.Sh; r.Sw = g.Sw;
    r.new_w = g.new_w; r.new_h = g.new_h; r.top = g.top; r.left = g.left;
    r.scale_x = g.scale_x; r.scale_y = g.scale_y;
    r.f16 = out_f16 ? 1 : 0;
    if (g.Sh > 65535 || B > 65535) return set_error(CY_ERR_INVALID, "cy_letterbox_resize: batch or image too large");
    pp_resize_kernel<<<dim3((unsigned)((g.Sw + 255) / 256), (unsigned)g.Sh, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(r);
    CY_CUDA_CHECK(cudaGetLastError());
    return CY_OK;
}
