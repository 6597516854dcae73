// Preprocessing chain on the device: tile cut-out (FITS byte order, NaN -> 0), the run.py stage list of
// caesar_yolo/preprocessing.py, and the ultralytics predictor preprocess (letterbox resize, channel flip, /255).
//
// Reference: BkgSubtractor (caesar_yolo/preprocessing.py:591-658), SigmaClipShifter (:664-717), SigmaClipper
// (:723-771), ChanResizer (:1077-1133), ZScaleTransformer (:934-971), Chan3Trasformer (:1020-1072), HistEqualizer
// (:977-1012), MinMaxNormalizer (:75-111), stage order scripts/run.py:272-293, Analyzer.predict front part
// (caesar_yolo/evaluation.py:146-176); astropy sigma_clip / ZScaleInterval, skimage equalize_hist and the
// ultralytics LetterBox semantics are restated in SURVEY.md App. A.1-A.4.
//
// Design.  Every stage of the chain is a monotone non-decreasing scalar map of the pixel value, and "0 means masked"
// is sticky.  So each channel is represented by a short op list (scalars only) applied to the ORIGINAL pixel value
// in fp64, never by a materialised fp64 image:
//   kernel 1 (pp_sort_kernel):   one CTA per tile; cut-out + byte swap + NaN->0 into a native fp32 tile buffer, then
//                                an in-L2 LSD radix sort of the non-masked pixels (order-preserving keys).
//   kernel 2 (pp_chain_kernel):  one CTA per tile; runs the stage list.  Order statistics (medians, clip bounds,
//                                histogram edges) are O(log n) lookups in the sorted array through the monotone op
//                                list; means/standard deviations are two-pass fp64 block reductions (numpy's
//                                definition) with warp-shuffle trees; zscale sorts its 1000 positional samples in
//                                shared memory and does the iterative line fit in fp64.  Writes the chain output
//                                (fp32 HWC) with one fused evaluation of the final op lists.
//   kernel 3 (pp_resize_kernel): half-pixel bilinear letterbox resize (double-precision source coordinates like
//                                cv2.resize on float64), pad 114, channel reversal, /255, bf16 NHWC(4) store.
#include "common.h"
#include "half16.cuh"
#include <stddef.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

namespace cy {

static constexpr int kPPThreads = 512;    // two CTAs (tiles) per SM: the kernels are latency-bound, the second CTA
                                          // fills the first one's barrier and dependent-load stalls
static constexpr int kPPWarps = kPPThreads / 32;
static constexpr int kSortThreads = 1024;  // the sort is instruction-bound: one full-width CTA per SM
static constexpr int kSortWarps = kSortThreads / 32;
static constexpr int kMaxOps = 12;
static constexpr int kMaxZero = 8;

enum OpKind { OP_SUB = 1, OP_SHIFT = 2, OP_CLAMP = 3, OP_ZSCALE = 4, OP_HISTEQ = 5, OP_MINMAX = 6 };

struct Chan {
    int nops;
    int hid;  // history id: channels with equal hid hold identical data
    int kind[kMaxOps];
    double p0[kMaxOps], p1[kMaxOps], p2[kMaxOps], p3[kMaxOps];
    int nz;  // zero (masked) index ranges in the sorted array, sorted + merged
    int z0[kMaxZero], z1[kMaxZero];
};

// Compiled form of a channel's op list for the full passes: every op except HISTEQ is affine + clamp with a positive
// slope, so the composition collapses to v = clamp(a*x + b, l, h) (one optional HISTEQ in the middle splits it into
// A and B).  Differs from the sequential fp64 evaluation only by rounding (~1e-16 relative); exact-zero (= masked)
// decisions never come from it: they are the index ranges Chan::z0/z1 of the sorted array, i.e. x intervals.
struct Comp {
    double a0, b0, l0, h0;
    double a1, b1, l1, h1;
    int has_he;
    int ok;                      // 0: not representable (non-finite / non-positive slope) -> interpreter
    int nzx;
    float zx0[8], zx1[8];        // zero intervals [zx0, zx1] of the ORIGINAL pixel value
};

struct HistEq {  // skimage.exposure.equalize_hist tables (one HistEqualizer per chain: Chan3 channel 2)
    double edges[257];
    double cdf[256];
    double center[256];   // (edges[k] + edges[k+1]) / 2
    double slope[256];    // (cdf[k+1] - cdf[k]) / (center[k+1] - center[k]), k < 255: the division np.interp does
    double inv_step;      // 255 / (center[255] - center[0]) (0 when the range is empty): bracket guess of the fast lookup
};

struct TileFinal;
struct PPParams {
    cy_pp_config cfg;
    const uint32_t* img;
    long long row_stride;
    int big_endian;
    const int* x0;
    const int* y0;
    int B, Ty, Tx;
    float* tilebuf;     // [B][N] native fp32, non-finite -> 0
    uint32_t* keysA;    // [B][N] sorted live values (fp32 bits after the sort kernel)
    uint32_t* keysB;    // [B][N] ping-pong
    uint32_t* keysC;    // [B][N] sorted live values outside the bkg box (only when use_box_mask_in_bkg)
    int* nlive;         // [B]
    int* nbox;          // [B]
    float* chain_out;   // [B][N][3] (optional parity output, written by pp_final_kernel)
    int* status;        // [B]
    struct TileFinal* fin;   // [B] final per-channel maps of every tile (chain kernel -> final kernel)
};

// ------------------------------------------------------------------------------------------ small device helpers

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

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
// !STICKY: the plain monotone composition (used to locate index ranges among live elements).
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
        }
    }
    return v;
}

// first index i in [a,b) with f*(S[i]) >= t (STRICT: > t); b if none.  Executed redundantly by every thread.
template <bool STRICT>
__device__ int lower_index(const Chan& c, int nops, const HistEq& he, const float* S, int a, int b, double t) {
    int lo = a, hi = b;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        const double v = eval_ops<false>(c, nops, he, (double)S[m]);
        const bool ok = STRICT ? (v > t) : (v >= t);
        if (ok) hi = m; else lo = m + 1;
    }
    return lo;
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

// Value of a NON-masked pixel through the compiled op list (callers decide masking by index / x interval).
__device__ __forceinline__ double eval_fast(const Comp& cc, const Chan& c, const HistEq& he, double x) {
    if (!cc.ok) return eval_ops<true>(c, c.nops, he, x);
    double v = fmin(fmax(fma(cc.a0, x, cc.b0), cc.l0), cc.h0);
    if (cc.has_he) {
        v = histeq_interp_fast(he, v);
        v = fmin(fmax(fma(cc.a1, v, cc.b1), cc.l1), cc.h1);
    }
    return v;
}
__device__ __forceinline__ bool in_zero_x(const Comp& cc, float x) {
    bool z = (x == 0.0f);
    for (int k = 0; k < cc.nzx; ++k) z = z || (x >= cc.zx0[k] && x <= cc.zx1[k]);
    return z;
}
__device__ __forceinline__ bool in_zero_idx(const Chan& c, int i) {
    bool z = false;
    for (int k = 0; k < c.nz; ++k) z = z || (i >= c.z0[k] && i < c.z1[k]);
    return z;
}

// (thread 0) rebuild the compiled form of channel c after its op list / zero ranges changed
__device__ void compile_chan(const Chan& c, Comp& cc, const float* S) {
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
        cc.zx0[k] = S[c.z0[k]];
        cc.zx1[k] = S[c.z1[k] - 1];
    }
}

__device__ void add_zero_range(Chan& c, int h0, int h1) {  // single thread
    if (h1 <= h0) return;
    int n = c.nz;
    if (n < kMaxZero) {
        c.z0[n] = h0;
        c.z1[n] = h1;
        ++n;
    }
    // insertion sort + merge
    for (int i = n - 1; i > 0 && c.z0[i] < c.z0[i - 1]; --i) {
        const int t0 = c.z0[i], t1 = c.z1[i];
        c.z0[i] = c.z0[i - 1]; c.z1[i] = c.z1[i - 1];
        c.z0[i - 1] = t0; c.z1[i - 1] = t1;
    }
    int m = 0;
    for (int i = 1; i < n; ++i) {
        if (c.z0[i] <= c.z1[m]) {
            if (c.z1[i] > c.z1[m]) c.z1[m] = c.z1[i];
        } else {
            ++m;
            c.z0[m] = c.z0[i];
            c.z1[m] = c.z1[i];
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
// index of the k-th (0-based) live element at or after a
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

struct Shared {
    Chan ch[3];
    Comp cc[3];
    HistEq he;
    double red[3 * 32 + 4];
    double zs[1024];         // zscale samples / flat residuals
    unsigned char zbad[1024];
    unsigned char zbad2[1024];
    int hist[256];
    int next_hid;
    int fail;                // tile status
    int sidx;                // block-wide search scratch
};

// Final state of a tile's chain: what pp_final_kernel needs to evaluate the three channel maps per pixel.
struct TileFinal {
    Comp cc[3];
    Chan ch[3];       // op lists: interpreter fallback for maps the compiled form cannot represent (cc.ok == 0)
    HistEq he;        // valid when use_he
    int same01, same02, same12;
    int use_he;
    int status;       // 0 ok, < 0: tile rejected (model input zero-filled)
    int pad_[3];
};
static_assert(sizeof(TileFinal) % 16 == 0 && offsetof(TileFinal, he) % 16 == 0 && sizeof(HistEq) % 16 == 0,
              "TileFinal is copied in 16-byte words");

// Block-cooperative version of lower_index (all threads call it with the same arguments): every round probes
// kPPThreads equally spaced elements at once, so 2^18 elements need two rounds of one load each instead of 18
// dependent loads per thread.
template <bool STRICT>
__device__ int lower_index_coop(Shared& sh, const Chan& c, int nops, const float* S, int a, int b, double t) {
    int lo = a, hi = b;
    while (hi > lo) {
        const int len = hi - lo;
        const int step = (len + kPPThreads - 1) / kPPThreads;
        const long long mi = (long long)lo + (long long)threadIdx.x * step;
        bool ok = true;  // probes at or beyond hi count as "true"
        if (mi < hi) {
            const double v = eval_ops<false>(c, nops, sh.he, (double)S[mi]);
            ok = STRICT ? (v > t) : (v >= t);
        }
        __syncthreads();
        if (threadIdx.x == 0) sh.sidx = kPPThreads;
        __syncthreads();
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && bal) atomicMin(&sh.sidx, (int)(threadIdx.x & ~31u) + __ffs(bal) - 1);
        __syncthreads();
        const int T = sh.sidx;  // first probing thread whose predicate holds (kPPThreads: none)
        const long long mT = (long long)lo + (long long)T * step;
        const int nhi = (int)(mT < hi ? mT : hi);
        const int nlo = T == 0 ? lo : (int)(mT - step + 1 < hi ? mT - step + 1 : hi);
        if (step == 1 || T == 0) return nhi;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

// Pivoted moment sums of the live values f(S[i]), i in [a,b) U [a2,b2):  n, sum(v-p), sum((v-p)^2)   (fp64).
// Values come from the compiled op list; masked elements are the zero index ranges of the channel, so each input
// range is cut into its live segments once and the element loop carries no mask test.
__device__ void range_sums(const Chan& c, const Comp& cc, const HistEq& he, const float* S, int a, int b, int a2,
                           int b2, double p, double* red, double& s0, double& s1, double& s2) {
    double n = 0.0, u = 0.0, q = 0.0;
    const bool simple = cc.ok && !cc.has_he;
    const double bp = cc.b0 - p, lp = cc.l0 - p, hp = cc.h0 - p;   // clamp(a*x + b, l, h) - p
    for (int part = 0; part < 2; ++part) {
        int pos = part ? a2 : a;
        const int end = part ? b2 : b;
        int zi = 0;
        while (pos < end) {
            // next live segment [pos, seg_end)
            while (zi < c.nz && c.z1[zi] <= pos) ++zi;
            if (zi < c.nz && c.z0[zi] <= pos) {
                pos = c.z1[zi];
                continue;
            }
            const int seg_end = (zi < c.nz && c.z0[zi] < end) ? c.z0[zi] : end;
            if (threadIdx.x == 0) n += (double)(seg_end - pos);
            if (simple) {
                // two independent accumulator pairs: the DADD / DFMA dependency chains were the stall of this loop
                double u1 = 0.0, q1 = 0.0;
                int i = pos + threadIdx.x;
                for (; i + kPPThreads < seg_end; i += 2 * kPPThreads) {
                    const float x0 = S[i], x1 = S[i + kPPThreads];
                    const double d0 = fmin(fmax(fma(cc.a0, (double)x0, bp), lp), hp);
                    const double d1 = fmin(fmax(fma(cc.a0, (double)x1, bp), lp), hp);
                    u += d0;
                    q = fma(d0, d0, q);
                    u1 += d1;
                    q1 = fma(d1, d1, q1);
                }
                if (i < seg_end) {
                    const double d = fmin(fmax(fma(cc.a0, (double)S[i], bp), lp), hp);
                    u += d;
                    q = fma(d, d, q);
                }
                u += u1;
                q += q1;
            } else {
#pragma unroll 2
                for (int i = pos + threadIdx.x; i < seg_end; i += kPPThreads) {
                    const double d = eval_fast(cc, c, he, (double)S[i]) - p;
                    u += d;
                    q = fma(d, d, q);
                }
            }
            pos = seg_end;
        }
    }
    block_sum3(n, u, q, red);
    s0 = n;
    s1 = u;
    s2 = q;
}

__device__ double live_median(const Chan& c, const HistEq& he, const float* S, int a, int cnt) {
    // numpy median: middle element, or mean of the two middle elements
    if (cnt & 1) return eval_ops<true>(c, c.nops, he, (double)S[kth_live(c, a, cnt >> 1)]);
    const double x = eval_ops<true>(c, c.nops, he, (double)S[kth_live(c, a, (cnt >> 1) - 1)]);
    const double y = eval_ops<true>(c, c.nops, he, (double)S[kth_live(c, a, cnt >> 1)]);
    return (x + y) / 2.0;
}

// astropy SigmaClip (axis=None, median/std, maxiters 5; App. A.1) over the live values of channel c in S[0..n).
// Outputs the bounds of the last iteration and the (mean, std) of the survivors.  Returns false if the set is empty.
// One full pass builds the moment sums about a pivot (the initial median, so mean-pivot = O(std): no cancellation);
// every iteration keeps a contiguous index range of the sorted array, so it only subtracts the sums of the clipped
// tails: mean = p + S1/n, std = sqrt(S2/n - (S1/n)^2)  (== numpy's two-pass definition up to fp64 rounding).
__device__ bool sigma_clip(Shared& sh, const Chan& c, const Comp& cc, const float* S, int n, double sig_lo,
                           double sig_hi, double& lo, double& hi, double& mean, double& sd) {
    int a = 0, b = n;
    int cnt = live_count(c, a, b);
    if (cnt <= 0) return false;
    const double p = live_median(c, sh.he, S, a, cnt);
    double s0, s1, s2;
    range_sums(c, cc, sh.he, S, a, b, 0, 0, p, sh.red, s0, s1, s2);
    if ((int)s0 != cnt) {  // index-range bookkeeping and evaluation disagree: never expected
        if (threadIdx.x == 0) sh.fail = -5;
        cnt = (int)s0;
        if (cnt <= 0) return false;
    }
    lo = hi = 0.0;
    for (int it = 0; it < 5; ++it) {
        const double m1 = s1 / s0;
        mean = p + m1;
        sd = sqrt(fmax(s2 / s0 - m1 * m1, 0.0));
        const double med = live_median(c, sh.he, S, a, cnt);
        lo = med - sd * sig_lo;
        hi = med + sd * sig_hi;
        const int na = lower_index_coop<false>(sh, c, c.nops, S, a, b, lo);   // first f >= lo
        const int nb = lower_index_coop<true>(sh, c, c.nops, S, na, b, hi);   // first f > hi
        const int ncnt = live_count(c, na, nb);
        if (ncnt == cnt) break;
        if (ncnt <= 0) return false;
        double t0, t1, t2;
        range_sums(c, cc, sh.he, S, a, na, nb, b, p, sh.red, t0, t1, t2);  // the clipped tails only
        s0 -= t0;
        s1 -= t1;
        s2 -= t2;
        a = na;
        b = nb;
        cnt = ncnt;
    }
    const double m1 = s1 / s0;
    mean = p + m1;
    sd = sqrt(fmax(s2 / s0 - m1 * m1, 0.0));
    return true;
}

// Append an op to channel c and register the pixels it newly maps to exactly 0 (they become masked).
__device__ void push_op(Shared& sh, int ci, const float* S, int n, int kind, double p0, double p1, double p2, double p3) {
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
    }
    __syncthreads();
    // zero set of the plain composition is a contiguous index range (monotone): [first >= 0, first > 0)
    const int h0 = lower_index_coop<false>(sh, c, c.nops, S, 0, n, 0.0);
    const int h1 = lower_index_coop<true>(sh, c, c.nops, S, h0, n, 0.0);
    __syncthreads();
    if (threadIdx.x == 0) {
        add_zero_range(c, h0, h1);
        c.hid = sh.next_hid++;
        compile_chan(c, sh.cc[ci], S);
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

// astropy ZScaleInterval.get_limits (App. A.2) on channel c (positional samples of the current image, zeros
// included), then the OP_ZSCALE op.
__device__ void zscale_stage(Shared& sh, int ci, const float* tile, int N, const float* S, int n, double contrast) {
    const Chan& c = sh.ch[ci];
    const int t = threadIdx.x;
    constexpr int kZS = 1024;                     // sample slots (>= the 1000 samples of ZScaleInterval)
    constexpr int kZE = kZS / kPPThreads;         // slots per thread: e = t + q * kPPThreads
    const int stride = (int)fmax(1.0, (double)N / 1000.0);
    int npix = (N + stride - 1) / stride;
    if (npix > 1000) npix = 1000;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kZE; ++q) {
        const int e = t + q * kPPThreads;
        sh.zs[e] = (e < npix) ? eval_ops<true>(c, c.nops, sh.he, (double)tile[(long long)e * stride]) : INFINITY;
    }
    __syncthreads();
    // bitonic sort ascending, 1024 elements, one compare-exchange per element pair
    for (int k = 2; k <= kZS; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int q = 0; q < kZE; ++q) {
                const int e = t + q * kPPThreads;
                const int p = e ^ j;
                if (p > e) {
                    const double x = sh.zs[e], y = sh.zs[p];
                    const bool asc = ((e & k) == 0);
                    if (asc ? (x > y) : (x < y)) {
                        sh.zs[e] = y;
                        sh.zs[p] = x;
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
    push_op(sh, ci, S, n, OP_ZSCALE, vmin, vmax - vmin, 0.0, 0.0);
}

// skimage equalize_hist (App. A.3) on channel c: 256-bin histogram over [min,max] of ALL pixels (masked zeros
// included), CDF, then OP_HISTEQ.  The channel is a monotone map of the sorted array, so min / max are its first / last
// live element (and 0 if any pixel is masked) and the bin counts are differences of 257 binary searches at the bin
// edges (numpy assigns v to the bin with edges[k] <= v < edges[k+1], last bin closed) -- no pass over the pixels.
__device__ void histeq_stage(Shared& sh, int ci, const float* tile, int N, const float* S, int n) {
    const Chan& c = sh.ch[ci];
    const int nlive = live_count(c, 0, n);
    const int nzero = N - nlive;
    double mn = INFINITY, mx = -INFINITY;
    if (nlive > 0) {
        mn = eval_ops<true>(c, c.nops, sh.he, (double)S[kth_live(c, 0, 0)]);
        mx = eval_ops<true>(c, c.nops, sh.he, (double)S[kth_live(c, 0, nlive - 1)]);
    }
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
    __syncthreads();
    // cum[k] = number of live elements with value < edges[k]  (k = 256: all of them, the last bin is closed)
    int* cum = reinterpret_cast<int*>(sh.zs);
    for (int k = threadIdx.x; k < 257; k += kPPThreads)
        cum[k] = k == 256 ? nlive : live_count(c, 0, lower_index<false>(c, c.nops, sh.he, S, 0, n, sh.he.edges[k]));
    __syncthreads();
    for (int k = threadIdx.x; k < 256; k += kPPThreads) sh.hist[k] = cum[k + 1] - cum[k];
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
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kPPThreads) sh.he.center[i] = (sh.he.edges[i] + sh.he.edges[i + 1]) / 2.0;
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kPPThreads)
        sh.he.slope[i] = i < 255 ? __ddiv_rn(__dsub_rn(sh.he.cdf[i + 1], sh.he.cdf[i]),
                                            __dsub_rn(sh.he.center[i + 1], sh.he.center[i]))
                                 : 0.0;
    if (threadIdx.x == 0) {
        const double step = (sh.he.center[255] - sh.he.center[0]) / 255.0;
        sh.he.inv_step = step > 0.0 ? 1.0 / step : 0.0;
    }
    __syncthreads();
    push_op(sh, ci, S, n, OP_HISTEQ, 0.0, 0.0, 0.0, 0.0);
}

// SigmaClipper._clip (preprocessing.py:735-751) on channel ci
__device__ bool sigma_clipper_stage(Shared& sh, int ci, const float* S, int n, double s_lo, double s_hi) {
    // astropy: sigma_lower = sigma_lower or sigma(=3.0): a falsy 0 falls back to 3 (App. A.1 / B#1)
    const double slo = s_lo != 0.0 ? s_lo : 3.0, shi = s_hi != 0.0 ? s_hi : 3.0;
    double lo, hi, mean, sd;
    if (!sigma_clip(sh, sh.ch[ci], sh.cc[ci], S, n, slo, shi, lo, hi, mean, sd)) return false;
    push_op(sh, ci, S, n, OP_CLAMP, lo, hi, 0.0, 0.0);
    return true;
}

// ------------------------------------------------------------------------------------------ kernel 1: cut-out + sort

__device__ __forceinline__ float load_pixel(const PPParams& p, int b, int idx) {
    const int y = idx / p.Tx, x = idx - y * p.Tx;
    uint32_t raw = p.img[(long long)(p.y0[b] + y) * p.row_stride + p.x0[b] + x];
    if (p.big_endian) raw = __byte_perm(raw, 0, 0x0123);
    const float f = __uint_as_float(raw);
    return isfinite(f) ? f : 0.0f;  // utils.py:219,394
}

// LSD radix sort (4 x 8 bit) of `n` keys in global memory with one CTA.  Every pass streams the keys in chunks of 8 keys per thread
// through shared memory: warp-striped loads, stable ranks inside the chunk from __match_any_sync + per-warp digit
// counters, a counting sort of the chunk in shared memory, then a write-out in which consecutive threads carry
// consecutive keys of one digit to consecutive addresses (full-sector stores; the direct per-key scatter this replaces
// wrote 4 bytes per 32-byte sector).  The histogram of the NEXT digit is built during the write-out, so a pass reads
// and writes every key once; passes whose digit is the same for all keys are skipped.
static constexpr int kSortChunk = 8 * kSortThreads;   // 8 keys per thread, warp-striped
struct SortSmem {
    int hist[256][kSortWarps + 1];  // [digit][warp] counters -> exclusive warp bases inside the chunk (+1: bank skew)
    uint32_t sorted[kSortChunk];  // the chunk in digit order
    int dtot[256], dbase[256], gcur[256];
    int nh[2][256];               // digit histograms of the whole array (current / next pass)
    int flag;
};

// lanes of `msk` whose 8-bit digit equals mine (eight independent ballots; MATCH.ANY is far slower than this)
__device__ __forceinline__ uint32_t digit_peers(uint32_t msk, uint32_t d) {
    uint32_t peers = msk;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool on = (d >> bit) & 1u;
        const uint32_t b = __ballot_sync(msk, on);
        peers &= on ? b : ~b;
    }
    return peers;
}

// exclusive scan of a[0..256) into out[0..256) by warp 0 (8 entries per lane); returns nothing
__device__ __forceinline__ void scan256_warp0(const int* a, int* out) {
    const int lane = threadIdx.x & 31;
    int v[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        v[q] = a[lane * 8 + q];
        s += v[q];
    }
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    int run = inc - s;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        out[lane * 8 + q] = run;
        run += v[q];
    }
}

__device__ long long* g_sort_dbg = nullptr;   // optional phase timing (clock64 sums of thread 0 of block 0), tools only
#define SORT_T(k) do { if (dbg) { const long long _c = clock64(); dbg[k] += _c - tprev; tprev = _c; } } while (0)

// Returns the buffer (a or b) that holds the sorted keys.  *as_float is set when the top-digit pass moved the keys (it
// then wrote them back as fp32 bit patterns, key2f fused into its write-out: no separate conversion pass).
__device__ uint32_t* block_radix_sort(uint32_t* a, uint32_t* b, int n, SortSmem& sm, bool* as_float) {
    *as_float = false;
    long long* dbg = (g_sort_dbg && blockIdx.x == 0 && threadIdx.x == 0) ? g_sort_dbg : nullptr;
    long long tprev = dbg ? clock64() : 0;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    uint32_t* src = a;
    uint32_t* dst = b;
    for (int i = t; i < 512; i += kSortThreads) (&sm.nh[0][0])[i] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 4 * kSortThreads) {
        uint32_t k[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * kSortThreads + t;
            k[q] = i < n ? src[i] : 0xffffffffu;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (i0 + q * kSortThreads + t < n) atomicAdd(&sm.nh[0][k[q] & 255u], 1);
    }
    __syncthreads();
    SORT_T(0);
    int cur = 0;
    for (int shift = 0; shift < 32; shift += 8, cur ^= 1) {
        int* nh = sm.nh[cur];
        int* nh_next = sm.nh[cur ^ 1];
        if (t < 32) {
            scan256_warp0(nh, sm.gcur);
            int one = 0;
            for (int q = 0; q < 8; ++q) one |= (nh[lane * 8 + q] == n);
            one = __any_sync(0xffffffffu, one);
            if (lane == 0) sm.flag = one;
        }
        for (int i = t; i < 256; i += kSortThreads) nh_next[i] = 0;
        __syncthreads();
        const bool trivial = sm.flag != 0;
        const bool more = shift < 24;
        if (trivial) {  // nothing moves; only the next digit's histogram is needed
            if (more)
                for (int i = t; i < n; i += kSortThreads) atomicAdd(&nh_next[(src[i] >> (shift + 8)) & 255u], 1);
            __syncthreads();
            continue;
        }
        for (int c0 = 0; c0 < n; c0 += kSortChunk) {
            const int m = min(kSortChunk, n - c0);
            for (int i = t; i < 256 * (kSortWarps + 1); i += kSortThreads) (&sm.hist[0][0])[i] = 0;
            uint32_t key[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // independent loads first
                const int li = w * 256 + j * 32 + lane;
                key[j] = li < m ? src[c0 + li] : 0u;
            }
            __syncthreads();
            SORT_T(1);
            int rnk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int li = w * 256 + j * 32 + lane;
                const bool v = li < m;
                rnk[j] = -1;
                const uint32_t msk = __ballot_sync(0xffffffffu, v);
                if (v) {
                    const uint32_t d = (key[j] >> shift) & 255u;
                    const uint32_t peers = digit_peers(msk, d);
                    const int r = __popc(peers & ((1u << lane) - 1u));
                    const int pre = sm.hist[d][w];
                    __syncwarp(msk);
                    if (r == 0) sm.hist[d][w] = pre + __popc(peers);
                    rnk[j] = pre + r;
                }
                __syncwarp();
            }
            __syncthreads();
            SORT_T(2);
            // per digit: exclusive scan over the warps' counters (one warp per digit, 256 / kSortWarps digits per warp)
#pragma unroll
            for (int q = 0; q < 256 / kSortWarps; ++q) {
                const int d = w * (256 / kSortWarps) + q;
                const int v = lane < kSortWarps ? sm.hist[d][lane] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                if (lane < kSortWarps) sm.hist[d][lane] = inc - v;
                if (lane == 31) sm.dtot[d] = inc;
            }
            __syncthreads();
            SORT_T(3);
            if (t < 32) scan256_warp0(sm.dtot, sm.dbase);
            __syncthreads();
            SORT_T(4);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (rnk[j] >= 0) {
                    const uint32_t d = (key[j] >> shift) & 255u;
                    sm.sorted[sm.dbase[d] + sm.hist[d][w] + rnk[j]] = key[j];
                }
            __syncthreads();
            SORT_T(5);
#pragma unroll
            for (int q = 0; q < kSortChunk / kSortThreads; ++q) {
                const int j = q * kSortThreads + t;
                if (j < m) {
                    const uint32_t k = sm.sorted[j];
                    const uint32_t d = (k >> shift) & 255u;
                    dst[sm.gcur[d] + (j - sm.dbase[d])] = more ? k : __float_as_uint(key2f(k));
                    if (more) atomicAdd(&nh_next[(k >> (shift + 8)) & 255u], 1);
                }
            }
            __syncthreads();
            SORT_T(6);
            if (t < 256) sm.gcur[t] += sm.dtot[t];
        }
        if (!more) *as_float = true;
        __syncthreads();
        uint32_t* x = src;
        src = dst;
        dst = x;
    }
    return src;
}

__global__ void __launch_bounds__(kSortThreads, 2) pp_sort_kernel(const __grid_constant__ PPParams p) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(sort_smem_raw);
    __shared__ int s_n, s_nb;
    const int b = blockIdx.x;
    const int N = p.Ty * p.Tx;
    float* tile = p.tilebuf + (long long)b * N;
    uint32_t* A = p.keysA + (long long)b * N;
    uint32_t* Bk = p.keysB + (long long)b * N;
    uint32_t* C = p.keysC ? p.keysC + (long long)b * N : nullptr;
    if (threadIdx.x == 0) {
        s_n = 0;
        s_nb = 0;
    }
    __syncthreads();
    // box of BkgSubtractor (preprocessing.py:610-621)
    const int xc = p.Tx / 2, yc = p.Ty / 2;
    const int dy = (int)(p.Ty * p.cfg.bkg_box_mask_fract / 2.0), dx = (int)(p.Tx * p.cfg.bkg_box_mask_fract / 2.0);
    const int lane = threadIdx.x & 31;
    for (int i00 = 0; i00 < N; i00 += 4 * kSortThreads) {
        float fq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // four independent loads in flight per thread
            const int i = i00 + q * kSortThreads + threadIdx.x;
            fq[q] = i < N ? load_pixel(p, b, i) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i00 + q * kSortThreads + threadIdx.x;
            const float f = fq[q];
            if (i < N) tile[i] = f;
            const bool live = (i < N) && (f != 0.0f);
            const uint32_t m = __ballot_sync(0xffffffffu, live);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) A[base + __popc(m & ((1u << lane) - 1u))] = f2key(f);
            if (C) {
                const int y = i / p.Tx, x = i - y * p.Tx;
                const bool inbox = (y >= yc - dy && y < yc + dy && x >= xc - dx && x < xc + dx);
                const bool lb = live && !inbox;
                const uint32_t m2 = __ballot_sync(0xffffffffu, lb);
                int base2 = 0;
                if (lane == 0 && m2) base2 = atomicAdd(&s_nb, __popc(m2));
                base2 = __shfl_sync(0xffffffffu, base2, 0);
                if (lb) C[base2 + __popc(m2 & ((1u << lane) - 1u))] = f2key(f);
            }
        }
    }
    __syncthreads();
    const int n = s_n, nb = s_nb;
    // sorted keys -> fp32 values in `out`; usually nothing is left to do (four moving passes end in `out`, already
    // converted by the last one)
    auto finish = [&](const uint32_t* r, bool as_float, uint32_t* out, int cnt) {
        if (as_float && r == out) return;
        for (int i = threadIdx.x; i < cnt; i += kSortThreads) {
            const uint32_t k = r[i];
            out[i] = as_float ? k : __float_as_uint(key2f(k));
        }
    };
    bool as_float;
    {
        const uint32_t* r = block_radix_sort(A, Bk, n, sm, &as_float);
        finish(r, as_float, A, n);
    }
    if (C) {
        __syncthreads();
        const uint32_t* r = block_radix_sort(C, Bk, nb, sm, &as_float);
        finish(r, as_float, C, nb);
    }
    if (threadIdx.x == 0) {
        p.nlive[b] = n;
        p.nbox[b] = nb;
    }
}

// ------------------------------------------------------------------------------------------ kernel 2: the chain

__device__ __forceinline__ bool chan_selected(int chid, int c) { return chid == -1 || chid == c; }

__global__ void __launch_bounds__(kPPThreads, 2) pp_chain_kernel(const __grid_constant__ PPParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
    const int b = blockIdx.x;
    const int N = p.Ty * p.Tx;
    const float* tile = p.tilebuf + (long long)b * N;
    const float* S = reinterpret_cast<const float*>(p.keysA + (long long)b * N);
    const float* Sbox = p.keysC ? reinterpret_cast<const float*>(p.keysC + (long long)b * N) : nullptr;
    const int n = p.nlive[b];
    const cy_pp_config& cfg = p.cfg;
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) {
            sh.ch[c].nops = 0;
            sh.ch[c].hid = 0;
            sh.ch[c].nz = 0;
        }
        sh.next_hid = 1;
        sh.fail = 0;
        for (int c = 0; c < 3; ++c) compile_chan(sh.ch[c], sh.cc[c], S);
    }
    __syncthreads();
    bool ok = true;  // uniform across the block

    // Channel de-duplication: a stage applied with equal parameters to channels that hold identical data (equal
    // history id) gives identical results, so it is computed once and the channel state is copied.  `in_hid[c]` is
    // the history id channel c had when the current stage processed it (-1: not processed).
    int in_hid[3];
    auto find_same = [&](int c) {
        for (int q = 0; q < c; ++q)
            if (in_hid[q] >= 0 && in_hid[q] == sh.ch[c].hid) return q;
        return -1;
    };

    if (cfg.enabled) {
        // ---- BkgSubtractor (preprocessing.py:591-658): x - clipped mean
        if (cfg.subtract_bkg) {
            in_hid[0] = in_hid[1] = in_hid[2] = -1;
            for (int c = 0; c < 3 && ok; ++c) {
                if (!chan_selected(cfg.bkg_chid, c)) continue;
                const int q = find_same(c), hid = sh.ch[c].hid;
                if (q >= 0) {
                    copy_chan(sh, c, q);
                } else {
                    double lo, hi, mean, sd;
                    const float* Sb = cfg.use_box_mask_in_bkg ? Sbox : S;
                    const int nb = cfg.use_box_mask_in_bkg ? p.nbox[b] : n;
                    const double sg = cfg.sigma_bkg != 0 ? cfg.sigma_bkg : 3.0;
                    ok = sigma_clip(sh, sh.ch[c], sh.cc[c], Sb, nb, sg, sg, lo, hi, mean, sd);
                    if (ok) push_op(sh, c, S, n, OP_SUB, mean, 0, 0, 0);
                }
                in_hid[c] = hid;
            }
        }
        // ---- SigmaClipShifter (preprocessing.py:664-717): x - (clipmean + sigma*std), negatives -> 0
        if (cfg.clip_shift_data && ok) {
            in_hid[0] = in_hid[1] = in_hid[2] = -1;
            for (int c = 0; c < 3 && ok; ++c) {
                if (!chan_selected(cfg.clip_chid, c)) continue;
                const int q = find_same(c), hid = sh.ch[c].hid;
                if (q >= 0) {
                    copy_chan(sh, c, q);
                } else {
                    double lo, hi, mean, sd;
                    const double sg = cfg.sigma_clip != 0 ? cfg.sigma_clip : 3.0;
                    ok = sigma_clip(sh, sh.ch[c], sh.cc[c], S, n, sg, sg, lo, hi, mean, sd);
                    if (ok) push_op(sh, c, S, n, OP_SHIFT, mean + cfg.sigma_clip * sd, 0, 0, 0);
                }
                in_hid[c] = hid;
            }
        }
        // ---- SigmaClipper (preprocessing.py:723-771)
        if (cfg.clip_data && ok) {
            in_hid[0] = in_hid[1] = in_hid[2] = -1;
            for (int c = 0; c < 3 && ok; ++c) {
                if (!chan_selected(cfg.clip_chid, c)) continue;
                const int q = find_same(c), hid = sh.ch[c].hid;
                if (q >= 0) copy_chan(sh, c, q);
                else ok = sigma_clipper_stage(sh, c, S, n, cfg.sigma_clip_low, cfg.sigma_clip_up);
                in_hid[c] = hid;
            }
        }
        // ---- ChanResizer: the cube already has 3 channels (evaluation.py:146-154) -> no-op for nchannels 1 or 3
        // ---- ZScaleTransformer (preprocessing.py:934-971)
        if (cfg.zscale_stretch && ok) {
            in_hid[0] = in_hid[1] = in_hid[2] = -1;
            for (int c = 0; c < 3; ++c) {
                int q = -1;
                for (int k = 0; k < c; ++k)
                    if (in_hid[k] == sh.ch[c].hid && cfg.zscale_contrasts[k] == cfg.zscale_contrasts[c]) q = k;
                const int hid = sh.ch[c].hid;
                if (q >= 0) copy_chan(sh, c, q);
                else zscale_stage(sh, c, tile, N, S, n, cfg.zscale_contrasts[c]);
                in_hid[c] = hid;
            }
        }
        // ---- Chan3Trasformer (preprocessing.py:1020-1072)
        if (cfg.chan3_preproc && ok) {
            const int hid0 = sh.ch[0].hid, hid1 = sh.ch[1].hid;
            ok = sigma_clipper_stage(sh, 0, S, n, cfg.sigma_clip_baseline, cfg.sigma_clip_up);
            if (ok) zscale_stage(sh, 0, tile, N, S, n, cfg.zscale_contrasts[0]);
            if (ok) {
                const double blo = cfg.sigma_clip_baseline != 0.0 ? cfg.sigma_clip_baseline : 3.0;
                const double llo = cfg.sigma_clip_low != 0.0 ? cfg.sigma_clip_low : 3.0;
                if (hid0 == hid1 && blo == llo) {
                    copy_chan(sh, 1, 0);
                } else {
                    ok = sigma_clipper_stage(sh, 1, S, n, cfg.sigma_clip_low, cfg.sigma_clip_up);
                    if (ok) zscale_stage(sh, 1, tile, N, S, n, cfg.zscale_contrasts[0]);
                }
            }
            if (ok) histeq_stage(sh, 2, tile, N, S, n);
        }
        // ---- MinMaxNormalizer (preprocessing.py:75-111)
        if (cfg.normalize_minmax && ok) {
            in_hid[0] = in_hid[1] = in_hid[2] = -1;
            for (int c = 0; c < 3 && ok; ++c) {
                const int q = find_same(c), hid = sh.ch[c].hid;
                if (q >= 0) {
                    copy_chan(sh, c, q);
                } else {
                    // min / max over the non-zero pixels = first / last live element of the sorted array
                    const Chan& ch = sh.ch[c];
                    double mn = INFINITY, mx = -INFINITY;
                    const int nl = live_count(ch, 0, n);
                    if (nl > 0) {
                        mn = eval_ops<true>(ch, ch.nops, sh.he, (double)S[kth_live(ch, 0, 0)]);
                        mx = eval_ops<true>(ch, ch.nops, sh.he, (double)S[kth_live(ch, 0, nl - 1)]);
                    }
                    if (!(mn <= mx)) ok = false;  // no non-zero pixel -> None (preprocessing.py:101-103)
                    else push_op(sh, c, S, n, OP_MINMAX, mn, mx - mn, cfg.norm_max - cfg.norm_min, cfg.norm_min);
                }
                in_hid[c] = hid;
            }
        }
    }

    // ---- hand the final per-channel maps to pp_final_kernel (which evaluates them per pixel, fused with the letterbox
    // resize) + the reference's degenerate-image check on rows 0..2 (evaluation.py:171-176)
    TileFinal& tf = p.fin[b];
    if (!ok) {
        if (threadIdx.x == 0) {
            const int stt = sh.fail ? sh.fail : -1;
            p.status[b] = stt;
            tf.status = stt;
        }
        return;
    }
    {
        __syncthreads();
        uint32_t* dst = reinterpret_cast<uint32_t*>(&tf);
        const uint32_t* s_cc = reinterpret_cast<const uint32_t*>(&sh.cc[0]);
        const uint32_t* s_ch = reinterpret_cast<const uint32_t*>(&sh.ch[0]);
        const uint32_t* s_he = reinterpret_cast<const uint32_t*>(&sh.he);
        constexpr int n_cc = sizeof(Comp) * 3 / 4, n_ch = sizeof(Chan) * 3 / 4, n_he = sizeof(HistEq) / 4;
        for (int i = threadIdx.x; i < n_cc; i += kPPThreads) dst[offsetof(TileFinal, cc) / 4 + i] = s_cc[i];
        for (int i = threadIdx.x; i < n_ch; i += kPPThreads) dst[offsetof(TileFinal, ch) / 4 + i] = s_ch[i];
        const bool any_he = sh.cc[0].has_he || sh.cc[1].has_he || sh.cc[2].has_he || !sh.cc[0].ok || !sh.cc[1].ok || !sh.cc[2].ok;
        if (any_he)
            for (int i = threadIdx.x; i < n_he; i += kPPThreads) dst[offsetof(TileFinal, he) / 4 + i] = s_he[i];
        if (threadIdx.x == 0) {
            tf.same01 = sh.ch[0].hid == sh.ch[1].hid;
            tf.same02 = sh.ch[0].hid == sh.ch[2].hid;
            tf.same12 = sh.ch[1].hid == sh.ch[2].hid;
            tf.use_he = any_he ? 1 : 0;
        }
    }
    int bad = 0;
    for (int r = 0; r < 3 && r < p.Ty; ++r) {
        double mn = INFINITY, mx = -INFINITY;
        for (int i = threadIdx.x; i < p.Tx; i += kPPThreads) {
            const double x = (double)tile[r * p.Tx + i];
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
// One CTA walks bands of `band_h` output rows of its tile: the input rows a band needs (band_h * Ty / new_h + 2) are
// evaluated once into shared memory (fp32 x 3), then every output pixel interpolates from shared memory.
// emit_only: bands are input rows, the evaluated rows are written to chain_out as they are (parity surface and the
// path of tiles whose bands do not fit shared memory, which then go through pp_resize_kernel).
struct FinalParams {
    PPParams p;
    void* out16;         // [B,Sh,Sw,4] bf16 / fp16
    float* out_f32;      // optional [B,3,Sh,Sw]
    int Sh, Sw, new_h, new_w, top, left;
    double scale_y, scale_x;   // 1 / (dst/src), as cv2.resize computes it
    int band_h, rows_cap, nbands;
    int f16, emit_only;
};

__device__ __forceinline__ void eval3(const TileFinal& tf, float xf, float* o) {
    const double x = (double)xf;
    const double v0 = in_zero_x(tf.cc[0], xf) ? 0.0 : eval_fast(tf.cc[0], tf.ch[0], tf.he, x);
    const double v1 = tf.same01 ? v0 : (in_zero_x(tf.cc[1], xf) ? 0.0 : eval_fast(tf.cc[1], tf.ch[1], tf.he, x));
    const double v2 = tf.same02 ? v0
                                : (tf.same12 ? v1 : (in_zero_x(tf.cc[2], xf) ? 0.0 : eval_fast(tf.cc[2], tf.ch[2], tf.he, x)));
    o[0] = (float)v0;
    o[1] = (float)v1;
    o[2] = (float)v2;
}

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

__global__ void __launch_bounds__(kFinThreads) pp_final_kernel(const __grid_constant__ FinalParams r) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    TileFinal& tf = *reinterpret_cast<TileFinal*>(fin_smem);
    int* xi0 = reinterpret_cast<int*>(fin_smem + sizeof(TileFinal));           // [Sw] source column of tap 0
    int* xi1 = xi0 + r.Sw;                                                      // [Sw] source column of tap 1 (-1: padding)
    float* xw = reinterpret_cast<float*>(xi1 + r.Sw);                           // [Sw]
    float* vals = reinterpret_cast<float*>(fin_smem + sizeof(TileFinal) + (((size_t)r.Sw * 12 + 15) & ~(size_t)15));
    const int b = blockIdx.x, tid = threadIdx.x;
    const PPParams& p = r.p;
    const int Ty = p.Ty, Tx = p.Tx;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&p.fin[b]);
        uint4* dst = reinterpret_cast<uint4*>(&tf);
        // the histogram-equalisation tables are the bulk of the state: skip them when no channel uses them
        const int n_head = (int)(offsetof(TileFinal, he) / 16), n_all = (int)(sizeof(TileFinal) / 16);
        const int n_tail0 = (int)((offsetof(TileFinal, he) + sizeof(HistEq)) / 16);
        for (int i = tid; i < n_head; i += kFinThreads) dst[i] = src[i];
        for (int i = n_tail0 + tid; i < n_all; i += kFinThreads) dst[i] = src[i];
        __syncthreads();
        if (tf.status == 0 && tf.use_he)
            for (int i = n_head + tid; i < n_tail0; i += kFinThreads) dst[i] = src[i];
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
    const bool ok = tf.status == 0;
    float* chain = p.chain_out ? p.chain_out + (long long)b * Ty * Tx * 3 : nullptr;
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
        // evaluate the input rows of the band once
        for (int i = tid; i < nrows * Tx; i += kFinThreads) {
            float o[3] = {0.f, 0.f, 0.f};
            if (ok) eval3(tf, load_pixel(p, b, yin0 * Tx + i), o);
            vals[3 * i + 0] = o[0];
            vals[3 * i + 1] = o[1];
            vals[3 * i + 2] = o[2];
        }
        __syncthreads();
        if (r.emit_only) {
            if (chain)
                for (int i = tid; i < nrows * Tx * 3; i += kFinThreads) chain[(long long)yin0 * Tx * 3 + i] = vals[i];
        } else {
            const int nout = (oy1 - oy0) * r.Sw;
            for (int o = tid; o < nout; o += kFinThreads) {
                const int dy = o / r.Sw, ox = o - dy * r.Sw, oy = oy0 + dy;
                float v[3] = {114.f, 114.f, 114.f};   // cv2.copyMakeBorder value
                const int ry = oy - r.top;
                const int x0 = xi0[ox];
                if (ry >= 0 && ry < r.new_h && x0 >= 0) {
                    int y0, y1;
                    float wy;
                    src_coord(ry, r.scale_y, Ty, y0, y1, wy);
                    const int x1 = xi1[ox];
                    const float wx = xw[ox];
                    const float* p00 = vals + ((y0 - yin0) * Tx + x0) * 3;
                    const float* p01 = vals + ((y0 - yin0) * Tx + x1) * 3;
                    const float* p10 = vals + ((y1 - yin0) * Tx + x0) * 3;
                    const float* p11 = vals + ((y1 - yin0) * Tx + x1) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float t0 = p00[c] * (1.f - wx) + p01[c] * wx;
                        const float t1 = p10[c] * (1.f - wx) + p11[c] * wx;
                        v[c] = t0 * (1.f - wy) + t1 * wy;
                    }
                }
                // predictor.preprocess: im[..., ::-1] (channel reversal), float32, /255
                const float m0 = v[2] / 255.f, m1 = v[1] / 255.f, m2 = v[0] / 255.f;
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
        __syncthreads();
    }
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
    // grid = (column blocks, output rows, tiles): no 64-bit div/mod per pixel (they were 40 % of this kernel's stalls)
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y, b = blockIdx.z;
    if (ox >= r.Sw) return;
    const long long idx = ((long long)b * r.Sh + oy) * r.Sw + ox;
    float v[3] = {114.f, 114.f, 114.f};  // cv2.copyMakeBorder value
    const int ry = oy - r.top, rx = ox - r.left;
    if (ry >= 0 && ry < r.new_h && rx >= 0 && rx < r.new_w) {
        // cv2.resize INTER_LINEAR: half-pixel centres, clamp at the borders
        double fy = ((double)ry + 0.5) * r.scale_y - 0.5;
        double fx = ((double)rx + 0.5) * r.scale_x - 0.5;
        int y0 = (int)floor(fy), x0 = (int)floor(fx);
        float wy = (float)(fy - (double)y0), wx = (float)(fx - (double)x0);
        if (y0 < 0) { y0 = 0; wy = 0.f; }
        if (y0 >= r.Ty - 1) { y0 = r.Ty - 1; wy = 0.f; }
        if (x0 < 0) { x0 = 0; wx = 0.f; }
        if (x0 >= r.Tx - 1) { x0 = r.Tx - 1; wx = 0.f; }
        const int y1 = min(y0 + 1, r.Ty - 1), x1 = min(x0 + 1, r.Tx - 1);
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

extern "C" int cy_sort_set_debug(void* dev_buf) {
    long long* p = (long long*)dev_buf;
    CY_CUDA_CHECK(cudaMemcpyToSymbol(cy::g_sort_dbg, &p, sizeof(p)));
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
static constexpr int kFinBandH = 8;            // output rows per band of the fused final kernel
static constexpr size_t kFinSmemMax = 100 * 1024;   // two CTAs per SM

static size_t fin_smem_bytes(int Sw, int rows, int Tx) {
    return sizeof(cy::TileFinal) + (((size_t)Sw * 12 + 15) & ~(size_t)15) + (size_t)rows * Tx * 12;
}
// rows of shared memory the fused bands need (0: does not fit -> chain_out + pp_resize_kernel path)
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

extern "C" size_t cy_preprocess_scratch_bytes(const cy_pp_config* cfg, int B, int Ty, int Tx) {
    const size_t N = (size_t)Ty * Tx;
    const int nbuf = 3 + ((cfg && cfg->enabled && cfg->subtract_bkg && cfg->use_box_mask_in_bkg) ? 1 : 0);
    // + the per-tile final maps + (tiles too large for the fused final kernel) the fp32 HWC chain image
    return (size_t)nbuf * align256((size_t)B * N * 4) + 2 * align256((size_t)B * 4) + align256((size_t)B * sizeof(cy::TileFinal)) +
           align256((size_t)B * N * 12) + 256;
}

extern "C" int cy_preprocess(const cy_pp_config* cfg, const void* img, long long row_stride, int big_endian,
                             const int32_t* tile_x0, const int32_t* tile_y0, int B, int Ty, int Tx, int imgsz,
                             float* chain_out, void* model_in, float* model_in_f32, int32_t* status, void* scratch,
                             uintptr_t stream) {
    using namespace cy;
    if (!cfg || !img || !tile_x0 || !tile_y0 || !status || !scratch)
        return set_error(CY_ERR_INVALID, "cy_preprocess: null argument");
    if (!chain_out && !model_in) return set_error(CY_ERR_INVALID, "cy_preprocess: no output requested");
    if (B <= 0 || Ty <= 0 || Tx <= 0) return set_error(CY_ERR_INVALID, "cy_preprocess: invalid shape");
    if ((long long)Ty * Tx >= (1ll << 30)) return set_error(CY_ERR_INVALID, "cy_preprocess: tile too large");
    if (B > 65535) return set_error(CY_ERR_INVALID, "cy_preprocess: at most 65535 tiles per call");
    if (cfg->enabled) {
        if (cfg->nchannels != 1 && cfg->nchannels != 3)
            return set_error(CY_ERR_INVALID, "nchannels must be 1 or 3 (the model takes 3-channel images)");
        if (cfg->chan3_preproc && cfg->nchannels != 3)
            return set_error(CY_ERR_INVALID, "chan3_preproc requires nchannels == 3 (scripts/run.py:253-256)");
    }
    cudaStream_t st = (cudaStream_t)stream;
    PPParams p;
    p.cfg = *cfg;
    p.img = (const uint32_t*)img;
    p.row_stride = row_stride;
    p.big_endian = big_endian;
    p.x0 = tile_x0;
    p.y0 = tile_y0;
    p.B = B; p.Ty = Ty; p.Tx = Tx;
    const size_t N = (size_t)Ty * Tx;
    const size_t buf = align256((size_t)B * N * 4);
    char* s = (char*)scratch;
    p.tilebuf = (float*)s; s += buf;
    p.keysA = (uint32_t*)s; s += buf;
    p.keysB = (uint32_t*)s; s += buf;
    const bool box = cfg->enabled && cfg->subtract_bkg && cfg->use_box_mask_in_bkg;
    p.keysC = nullptr;
    if (box) {
        p.keysC = (uint32_t*)s;
        s += buf;
    }
    p.nlive = (int*)s; s += align256((size_t)B * 4);
    p.nbox = (int*)s; s += align256((size_t)B * 4);
    p.fin = (TileFinal*)s; s += align256((size_t)B * sizeof(TileFinal));
    float* chain_scratch = (float*)s;
    p.chain_out = chain_out;
    p.status = status;
    static std::atomic<unsigned long long> attr_done{0};
    if (first_use_on_device(attr_done)) {
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared)));
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem)));
        CY_CUDA_CHECK(cudaFuncSetAttribute(pp_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemMax));
    }
    pp_sort_kernel<<<B, kSortThreads, sizeof(SortSmem), st>>>(p);
    pp_chain_kernel<<<B, kPPThreads, sizeof(Shared), st>>>(p);
    CY_CUDA_CHECK(cudaGetLastError());

    LbGeom g;
    int rc = lb_geometry(Ty, Tx, imgsz, &g);
    if (rc) return rc;
    FinalParams r;
    r.p = p;
    r.out16 = model_in;
    r.out_f32 = model_in_f32;
    r.Sh = g.Sh; r.Sw = g.Sw; r.new_h = g.new_h; r.new_w = g.new_w; r.top = g.top; r.left = g.left;
    r.scale_y = g.scale_y; r.scale_x = g.scale_x;
    r.f16 = cfg->out_f16 ? 1 : 0;
    const int sms = current_device_sms();
    const int cap = model_in ? fin_rows_cap(g, Ty, Tx) : 0;
    auto emit_chain = [&](float* dst) -> int {   // evaluated maps as an fp32 HWC image (parity output / resize input)
        FinalParams e = r;
        e.p.chain_out = dst;
        e.emit_only = 1;
        const int rows_fit = (int)((kFinSmemMax - sizeof(TileFinal) - 64) / ((size_t)Tx * 12));
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
    if (model_in) {
        if (cap > 0) {
            r.emit_only = 0;
            r.band_h = kFinBandH;
            r.rows_cap = cap;
            r.nbands = (g.Sh + kFinBandH - 1) / kFinBandH;
            int chunks = (2 * sms + B - 1) / B;
            chunks = chunks < 1 ? 1 : (chunks > r.nbands ? r.nbands : chunks);
            pp_final_kernel<<<dim3((unsigned)B, (unsigned)chunks), kFinThreads, fin_smem_bytes(g.Sw, cap, Tx), st>>>(r);
        } else {
            // bands of this tile shape do not fit shared memory: materialise the fp32 HWC image once, then resize it
            float* hwc = chain_out ? chain_out : chain_scratch;
            if (!chain_out && (rc = emit_chain(hwc))) return rc;
            rc = cy_letterbox_resize_fmt(hwc, B, Ty, Tx, imgsz, model_in, model_in_f32, r.f16, stream);
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
