// Detect-head decode, confidence filter, score sort, greedy NMS, box rescale and the per-tile IoU merge.
// Warp/CTA-level kernels, one CTA per tile; everything is batched over tiles.
//
// Replaces (reference call sites): ultralytics Detect._inference + ops.non_max_suppression +
// torchvision.ops.nms + ops.scale_boxes inside `model(...)` (caesar_yolo/evaluation.py:181-193; SURVEY App.
// A.6) and Analyzer.process_detections (caesar_yolo/evaluation.py:252-346) with utils.get_iou
// (caesar_yolo/utils.py:54-107) and Graph (caesar_yolo/graph.py:2-41).
#include "postprocess.h"
#include "common.h"
#include <math.h>

namespace cy {

// ------------------------------------------------------------------------------------------ helpers

__device__ __forceinline__ uint32_t float_orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Block-wide bitonic sort, descending, n a power of two; keys may live in shared or global memory.
__device__ void block_bitonic_sort_desc(unsigned long long* keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const unsigned long long a = keys[i], b = keys[p];
                const bool desc = ((i & k) == 0);
                if (desc ? (a < b) : (a > b)) {
                    keys[i] = b;
                    keys[p] = a;
                }
            }
            __syncthreads();
        }
    }
}

// torchvision nms_kernel_impl arithmetic (fp32, no contraction): suppress iff inter/(ai+aj-inter) > thr (double)
__device__ __forceinline__ bool nms_suppresses(const float4 a, float aa, const float4 b, float ab, double thr) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
    const float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
    // The IEEE division (a ~80-clock dependent chain, and the greedy passes are chains of these tests) is only needed
    // when the ratio is within 2e-5 of the threshold: outside that band the two products below decide with a margin
    // 100x wider than the rounding of the products and of the quotient, so the outcome equals the reference's
    // (double)(inter / uni) > thr bit for bit.  uni <= 0 (degenerate boxes: NaN / inf semantics) takes the division.
    const float t = (float)thr;
    if (uni > 0.f && t >= 1e-3f) {          // (a threshold near 0 would meet the underflow of the quotient)
        if (inter > __fmul_rn(uni, t * 1.00002f)) return true;
        if (inter < __fmul_rn(uni, t * 0.99998f)) return false;
    }
    const float ovr = __fdiv_rn(inter, uni);
    return (double)ovr > thr;
}
__device__ __forceinline__ float box_area(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

static constexpr int kNmsThreads = 512;  // == chunk size
static constexpr int kNmsWords = kNmsThreads / 32;

// Greedy NMS over K boxes already in descending score order.  boxes: global [K] float4.
// keep_pos: global scratch [>= min(K, max_keep)] receives positions (in sorted order) of kept boxes.
// Returns number kept (valid in all threads).  Uses static shared memory.
struct NoPrep {
    __device__ __forceinline__ void operator()(int) const {}
};
// prep(base): called by every thread before chunk [base, base + 512) is read — the fused path decodes the chunk's
// boxes there (thread t produces boxes[base + t], the element it reads itself), so only as many candidates are decoded
// as the greedy pass actually visits before max_keep boxes are kept.
template <class Prep>
__device__ int nms_sorted_block(const float4* boxes, int K, double thr, int max_keep, int* keep_pos, Prep prep) {
    __shared__ uint32_t s_mask[kNmsThreads][kNmsWords];
    __shared__ float4 s_kb[256];
    __shared__ float s_ka[256];
    __shared__ uint32_t s_alive[kNmsWords];
    __shared__ int s_nkeep;
    __shared__ float4 s_cb[kNmsThreads];
    __shared__ float s_ca[kNmsThreads];
    const int t = threadIdx.x;
    if (t == 0) s_nkeep = 0;
    __syncthreads();
    for (int base = 0; base < K; base += kNmsThreads) {
        const int nk0 = s_nkeep;
        if (nk0 >= max_keep) break;
        prep(base);
        const int idx = base + t;
        const bool have = idx < K;
        float4 bx = make_float4(0, 0, 0, 0);
        if (have) bx = boxes[idx];
        const float ar = box_area(bx);
        s_cb[t] = bx;
        s_ca[t] = ar;
        // phase A: test against boxes kept in earlier chunks
        bool alive = have;
        for (int kb = 0; kb < nk0; kb += 256) {
            const int nb = min(256, nk0 - kb);
            __syncthreads();
            if (t < nb) {
                const float4 q = boxes[keep_pos[kb + t]];
                s_kb[t] = q;
                s_ka[t] = box_area(q);
            }
            __syncthreads();
            if (alive) {
                for (int i = 0; i < nb; ++i)
                    if (nms_suppresses(s_kb[i], s_ka[i], bx, ar, thr)) {
                        alive = false;
                        break;
                    }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, alive);
        if ((t & 31) == 0) s_alive[t >> 5] = bal;
        __syncthreads();
        // phase B: intra-chunk suppression mask, row t = boxes j>t in this chunk that t would suppress
        {
            uint32_t* row = s_mask[t];
            for (int wd = 0; wd < kNmsWords; ++wd) {
                uint32_t bits = 0;
                if (alive && wd >= (t >> 5)) {
                    uint32_t cand = s_alive[wd];
                    if (wd == (t >> 5)) cand &= (t & 31) == 31 ? 0u : (0xffffffffu << ((t & 31) + 1));
                    while (cand) {
                        const int b = __ffs(cand) - 1;
                        cand &= cand - 1;
                        const int j = wd * 32 + b;
                        if (nms_suppresses(bx, ar, s_cb[j], s_ca[j], thr)) bits |= 1u << b;
                    }
                }
                row[wd] = bits;
            }
        }
        __syncthreads();
        // phase C: sequential resolve by warp 0; lane w owns word w of the removed set
        if (t < 32) {
            uint32_t remv = 0;
            uint32_t al = t < kNmsWords ? s_alive[t] : 0u;
            int nk = nk0;
            for (int wd = 0; wd < kNmsWords && nk < max_keep; ++wd) {
                // word wd of (alive & ~remv) evolves as we keep boxes inside it
                while (nk < max_keep) {
                    const uint32_t avail = __shfl_sync(0xffffffffu, al & ~remv, wd);
                    if (!avail) break;
                    const int b = __ffs(avail) - 1;
                    const int i = wd * 32 + b;
                    if (t == 0) keep_pos[nk] = base + i;
                    ++nk;
                    if (t < kNmsWords) remv |= s_mask[i][t];
                    if (t == wd) remv |= 1u << b;  // consume i itself
                }
            }
            if (t == 0) s_nkeep = nk;
        }
        __syncthreads();
    }
    return s_nkeep;
}

// ------------------------------------------------------------------------------------------ decode

struct HeadLevels {
    const float* p[3];
    int h[3], w[3];
    int a0[3];  // first anchor index of each level
    int A;
};

// Full decode (parity entry): pred [B, 4+nc, A] exactly like ultralytics Detect._inference.
__global__ void decode_pred_kernel(HeadLevels L, int B, int nc, float* __restrict__ pred) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * L.A) return;
    const int a = (int)(idx % L.A);
    const int b = (int)(idx / L.A);
    const int l = a >= L.a0[2] ? 2 : (a >= L.a0[1] ? 1 : 0);
    const int la = a - L.a0[l];
    const float stride = l == 0 ? 8.f : (l == 1 ? 16.f : 32.f);
    const float* rec = L.p[l] + ((long long)b * L.h[l] * L.w[l] + la) * kHeadC;
    const float ax = (float)(la % L.w[l]) + 0.5f, ay = (float)(la / L.w[l]) + 0.5f;
    float d[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[16], mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            v[k] = rec[s * 16 + k];
            mx = fmaxf(mx, v[k]);
        }
        float sum = 0.f, acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float e = expf(v[k] - mx);
            sum += e;
            acc += e * (float)k;
        }
        d[s] = acc / sum;
    }
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    float* o = pred + (long long)b * (4 + nc) * L.A + a;
    o[0 * (long long)L.A] = (x1 + x2) / 2.f * stride;
    o[1 * (long long)L.A] = (y1 + y2) / 2.f * stride;
    o[2 * (long long)L.A] = (x2 - x1) * stride;
    o[3 * (long long)L.A] = (y2 - y1) * stride;
    for (int c = 0; c < nc; ++c) o[(4 + c) * (long long)L.A] = 1.f / (1.f + expf(-rec[64 + c]));
}

// Stage 1 of the fused path: per anchor best class score; candidates (score > conf) are appended to the tile's key
// list (order fixed later by the sort; key = score | inverted anchor index | class).
__global__ void score_key_kernel(HeadLevels L, int B, int nc, float conf, unsigned long long* __restrict__ keys,
                                 int key_stride, int* __restrict__ cand_count) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * L.A) return;
    const int a = (int)(idx % L.A);
    const int b = (int)(idx / L.A);
    const int l = a >= L.a0[2] ? 2 : (a >= L.a0[1] ? 1 : 0);
    const int la = a - L.a0[l];
    const float* rec = L.p[l] + ((long long)b * L.h[l] * L.w[l] + la) * kHeadC + 64;
    float best = -1.f;
    int bc = 0;
    for (int c = 0; c < nc; ++c) {
        const float s = 1.f / (1.f + expf(-rec[c]));
        if (s > best) {  // torch max(): first index on ties
            best = s;
            bc = c;
        }
    }
    if (best > conf) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) |
                                       ((unsigned long long)(0xFFFFFu - (uint32_t)a) << 8) | (unsigned long long)bc;
        const int slot = atomicAdd(&cand_count[b], 1);
        keys[(long long)b * key_stride + slot] = key;
    }
}

// DFL decode of one candidate (key = score | inverted anchor | class) -> xywh centre form in letterboxed pixels.
__device__ __forceinline__ void decode_candidate(const HeadLevels& L, int b, unsigned long long key, float& cx, float& cy,
                                                 float& hw, float& hh, int& cls) {
    const int a = (int)(0xFFFFFu - (uint32_t)((key >> 8) & 0xFFFFFu));
    cls = (int)(key & 0xFFu);
    const int l = a >= L.a0[2] ? 2 : (a >= L.a0[1] ? 1 : 0);
    const int la = a - L.a0[l];
    const float stride = l == 0 ? 8.f : (l == 1 ? 16.f : 32.f);
    const float* rec = L.p[l] + ((long long)b * L.h[l] * L.w[l] + la) * kHeadC;
    const float ax = (float)(la % L.w[l]) + 0.5f, ay = (float)(la / L.w[l]) + 0.5f;
    float d[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[16], mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 r = *reinterpret_cast<const float4*>(rec + s * 16 + q * 4);
            v[4 * q] = r.x; v[4 * q + 1] = r.y; v[4 * q + 2] = r.z; v[4 * q + 3] = r.w;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) mx = fmaxf(mx, v[q]);
        float sum = 0.f, acc = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float e = expf(v[q] - mx);
            sum += e;
            acc += e * (float)q;
        }
        d[s] = acc / sum;
    }
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.f), stride);
    cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.f), stride);
    const float w = __fmul_rn(__fsub_rn(x2, x1), stride), h = __fmul_rn(__fsub_rn(y2, y1), stride);
    hw = __fdiv_rn(w, 2.f);
    hh = __fdiv_rn(h, 2.f);
}

static constexpr int kSelKeys = 4096;     // candidates sorted in shared memory
static constexpr int kSelBins = 2048;     // score histogram of the dense path (bins of the top 16 score bits)
static constexpr int kSelTarget = 2048;   // dense path: take at least this many of the best candidates first

// Stage 2: one CTA per tile.  Sort keys (score desc, anchor asc), NMS with class offsets over lazily decoded boxes,
// rescale to tile pixels, emit dets[max_det][6].
//  * up to 4096 candidates: the whole list is sorted in shared memory;
//  * more (dense tiles, BASELINE configs[4]): a histogram over the top 16 score bits finds the score bin above which
//    2048..4096 candidates lie; only those — exactly the head of the sorted list — are sorted (in shared memory) and
//    fed to the greedy pass, which stops at max_det kept boxes.  If it runs out of candidates first (rare: > 2048
//    candidates of which fewer than max_det survive), the full list is sorted in global memory and the pass repeated.
//  * boxes are decoded chunk by chunk inside the greedy pass (only the chunks it visits), and once more, without the
//    class offset, for the kept ones (subtracting the offset again would lose bits).
__global__ void __launch_bounds__(kNmsThreads) nms_tiles_kernel(HeadLevels L, int B, unsigned long long* keys,
                                                                int key_stride, const int* __restrict__ cand_count,
                                                                float4* boxes_scratch, int* keep_scratch,
                                                                unsigned long long* sel_keys, float conf, double iou_thr,
                                                                int max_det, int max_nms, float max_wh,
                                                                const LetterboxInfo* lb, float* __restrict__ dets,
                                                                int* __restrict__ ndets) {
    extern __shared__ __align__(16) unsigned long long nms_dyn[];
    unsigned long long* skeys = nms_dyn;                                   // [kSelKeys]
    int* hist = reinterpret_cast<int*>(nms_dyn + kSelKeys);                // [kSelBins]
    __shared__ int s_sel, s_thr_bin, s_total;
    const int b = blockIdx.x;
    unsigned long long* k = keys + (long long)b * key_stride;
    unsigned long long* sel = sel_keys + (long long)b * kSelKeys;
    const int ncand = cand_count[b];
    const int Kfull = min(ncand, max_nms);
    const uint32_t base16 = __float_as_uint(conf) >> 16;
    auto score_bin = [&](unsigned long long key) {
        const int v = (int)((uint32_t)(key >> 48)) - (int)base16;
        return min(max(v, 0), kSelBins - 1);
    };
    int n_sel;
    if (ncand <= kSelKeys) {
        n_sel = ncand;
        int n2 = 2;
        while (n2 < ncand) n2 <<= 1;
        for (int i = threadIdx.x; i < n2; i += blockDim.x) skeys[i] = i < ncand ? k[i] : 0ull;
        __syncthreads();
        block_bitonic_sort_desc(skeys, n2);
    } else {
        for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hist[i] = 0;
        if (threadIdx.x == 0) s_sel = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < ncand; i += blockDim.x) atomicAdd(&hist[score_bin(k[i])], 1);
        __syncthreads();
        if (threadIdx.x < 32) {   // highest bin T with count(bins >= T) >= kSelTarget (bin 0 if there is none)
            const int lane = threadIdx.x;
            int part = 0;         // lane owns bins [lane*64, lane*64+64)
            for (int q = 0; q < kSelBins / 32; ++q) part += hist[lane * (kSelBins / 32) + q];
            int above = 0;        // candidates in the lanes above mine
            for (int o = 0; o < 32; ++o) {
                const int v = __shfl_sync(0xffffffffu, part, o);
                if (o > lane) above += v;
            }
            const bool mine = above < kSelTarget && above + part >= kSelTarget;
            const uint32_t who = __ballot_sync(0xffffffffu, mine);
            if (who == 0u) {
                if (lane == 0) { s_thr_bin = 0; s_total = above + part; }
            } else if (mine) {
                int acc = above, T = lane * (kSelBins / 32);
                for (int q = kSelBins / 32 - 1; q >= 0; --q) {
                    acc += hist[lane * (kSelBins / 32) + q];
                    if (acc >= kSelTarget) { T = lane * (kSelBins / 32) + q; break; }
                }
                s_thr_bin = T;
                s_total = acc;
            }
        }
        __syncthreads();
        if (s_total <= kSelKeys) {
            const int T = s_thr_bin;
            for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
                const unsigned long long key = k[i];
                if (score_bin(key) >= T) skeys[atomicAdd(&s_sel, 1)] = key;
            }
            __syncthreads();
            n_sel = s_sel;
            int n2 = 2;
            while (n2 < n_sel) n2 <<= 1;
            for (int i = n_sel + threadIdx.x; i < n2; i += blockDim.x) skeys[i] = 0ull;
            __syncthreads();
            block_bitonic_sort_desc(skeys, n2);
        } else {
            n_sel = 0;   // one score bin holds thousands of candidates: straight to the full sort
        }
    }
    for (int i = threadIdx.x; i < n_sel; i += blockDim.x) sel[i] = skeys[i];
    __syncthreads();
    float4* bx = boxes_scratch + (long long)b * key_stride;
    int* kp = keep_scratch + (long long)b * max_det;
    const unsigned long long* ks = sel;
    auto prep = [&](int base) {
        const int i = base + (int)threadIdx.x;
        if (i < (ks == sel ? n_sel : Kfull)) {
            float cx, cy, hw, hh;
            int cls;
            decode_candidate(L, b, ks[i], cx, cy, hw, hh, cls);
            const float off = __fmul_rn((float)cls, max_wh);   // class offset as ultralytics does (agnostic = False)
            bx[i] = make_float4(__fadd_rn(__fsub_rn(cx, hw), off), __fadd_rn(__fsub_rn(cy, hh), off),
                                __fadd_rn(__fadd_rn(cx, hw), off), __fadd_rn(__fadd_rn(cy, hh), off));
        }
    };
    int nk = n_sel > 0 ? nms_sorted_block(bx, min(n_sel, max_nms), iou_thr, max_det, kp, prep) : 0;
    if (n_sel < Kfull && nk < max_det) {   // the head of the list was not enough: full sort, full pass
        int npow2 = 2;
        while (npow2 < ncand) npow2 <<= 1;
        __syncthreads();
        for (int i = ncand + threadIdx.x; i < npow2; i += blockDim.x) k[i] = 0ull;
        __syncthreads();
        block_bitonic_sort_desc(k, npow2);
        ks = k;
        nk = nms_sorted_block(bx, Kfull, iou_thr, max_det, kp, prep);
    }
    const LetterboxInfo li = lb[b];
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
        const unsigned long long key = ks[kp[i]];
        const float score = __uint_as_float((uint32_t)(key >> 32));
        float cx, cy, hw, hh;
        int cls;
        decode_candidate(L, b, key, cx, cy, hw, hh, cls);
        // scale_boxes: subtract pad, divide by gain, clip to the original tile
        float bx1 = __fdiv_rn(__fsub_rn(__fsub_rn(cx, hw), li.pad_x), li.gain);
        float by1 = __fdiv_rn(__fsub_rn(__fsub_rn(cy, hh), li.pad_y), li.gain);
        float bx2 = __fdiv_rn(__fsub_rn(__fadd_rn(cx, hw), li.pad_x), li.gain);
        float by2 = __fdiv_rn(__fsub_rn(__fadd_rn(cy, hh), li.pad_y), li.gain);
        bx1 = fminf(fmaxf(bx1, 0.f), (float)li.w0);
        bx2 = fminf(fmaxf(bx2, 0.f), (float)li.w0);
        by1 = fminf(fmaxf(by1, 0.f), (float)li.h0);
        by2 = fminf(fmaxf(by2, 0.f), (float)li.h0);
        float* o = dets + ((long long)b * max_det + i) * 6;
        o[0] = bx1; o[1] = by1; o[2] = bx2; o[3] = by2; o[4] = score; o[5] = (float)cls;
    }
    if (threadIdx.x == 0) ndets[b] = nk;
}

// Standalone batched NMS == torchvision.ops.nms per segment (parity entry cy_nms / cy_nms_batched).
__global__ void __launch_bounds__(kNmsThreads) nms_generic_kernel(const float* __restrict__ boxes,
                                                                  const float* __restrict__ scores,
                                                                  const int* __restrict__ counts, int stride_n,
                                                                  int npow2, unsigned long long* keys,
                                                                  float4* boxes_sorted, int* keep_pos, double thr,
                                                                  int max_keep, long long* __restrict__ keep,
                                                                  int* __restrict__ nkeep) {
    const int b = blockIdx.x;
    const int N = counts ? counts[b] : stride_n;
    unsigned long long* k = keys + (long long)b * npow2;
    const float* bsrc = boxes + (long long)b * stride_n * 4;
    const float* ssrc = scores + (long long)b * stride_n;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x)
        k[i] = i < N ? (((unsigned long long)float_orderable(ssrc[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i))
                     : 0ull;
    __syncthreads();
    block_bitonic_sort_desc(k, npow2);
    float4* bs = boxes_sorted + (long long)b * stride_n;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int src = (int)(0xFFFFFFFFu - (uint32_t)(k[i] & 0xFFFFFFFFull));
        bs[i] = make_float4(bsrc[src * 4 + 0], bsrc[src * 4 + 1], bsrc[src * 4 + 2], bsrc[src * 4 + 3]);
    }
    __syncthreads();
    int* kp = keep_pos + (long long)b * stride_n;
    const int nk = nms_sorted_block(bs, N, thr, max_keep, kp, NoPrep());
    for (int i = threadIdx.x; i < nk; i += blockDim.x)
        keep[(long long)b * stride_n + i] = (long long)(0xFFFFFFFFu - (uint32_t)(k[kp[i]] & 0xFFFFFFFFull));
    if (threadIdx.x == 0) nkeep[b] = nk;
}

// ------------------------------------------------------------------------------------------ per-tile merge

// utils.get_iou in fp32 (numpy>=2 scalar semantics).  Caller guarantees non-degenerate boxes.
__device__ __forceinline__ float ref_get_iou(const float4 a, const float4 b) {
    const float xl = fmaxf(a.x, b.x), yt = fmaxf(a.y, b.y), xr = fminf(a.z, b.z), yb = fminf(a.w, b.w);
    if (xr < xl || yb < yt) return 0.f;
    const float inter = __fmul_rn(__fsub_rn(xr, xl), __fsub_rn(yb, yt));
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

static constexpr int kMergeMaxN = 320;               // >= max_det (300)
static constexpr int kMergeWords = kMergeMaxN / 32;  // 10
static constexpr int kMergeThreads = 512;             // 16 warps: adjacency rows; warp 0 runs the sequential parts

// One CTA (512 threads) per tile.  dets [B, det_stride, 6]; keeps detections with !(score < thr_score), links
// pairs with iou >= hard or (same class and iou >= soft), connected components by recursive-DFS order, winner =
// first member in DFS preorder with strictly greatest score (score_best starts at 0).
// The three sequential parts of the reference (ordered score filter, DFS, ordered output) run on ONE WARP with the
// 32 lanes working on a step together: ballot-scan compaction, and a DFS whose visited set lives in registers (lane w
// owns word w of the bitmask) so that "next unvisited neighbour in ascending order" is one shared-memory load per lane +
// one ballot instead of a 10-word scan through local memory by a single thread.
__global__ void __launch_bounds__(kMergeThreads) merge_tile_kernel(const float* __restrict__ dets, const int* __restrict__ ndets,
                                                         int det_stride, float thr_score, float thr_soft,
                                                         float thr_hard, const int* __restrict__ pre_status,
                                                         int* __restrict__ keep_idx, int* __restrict__ nkeep,
                                                         int* __restrict__ status) {
    __shared__ float4 s_box[kMergeMaxN];
    __shared__ float s_score[kMergeMaxN];
    __shared__ int s_cls[kMergeMaxN];
    __shared__ int s_src[kMergeMaxN];
    __shared__ uint32_t s_adj[kMergeMaxN][kMergeWords];
    __shared__ short s_stack[kMergeMaxN];
    __shared__ int s_keep[kMergeMaxN];
    __shared__ int s_N, s_bad, s_nk;
    __shared__ float s_raw[kMergeMaxN * 6];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (pre_status && pre_status[b] != 0) {  // tile rejected upstream (predict returned -1): no detections
        if (threadIdx.x == 0) {
            nkeep[b] = 0;
            status[b] = pre_status[b];
        }
        return;
    }
    const int n_in = min(ndets[b], kMergeMaxN);
    {
        const float* G = dets + (long long)b * det_stride * 6;
        for (int i = threadIdx.x; i < n_in * 6; i += blockDim.x) s_raw[i] = G[i];
    }
    __syncthreads();
    const float* D = s_raw;
    if (warp == 0) {
        // score filter keeps order (evaluation.py:276-287): ballot scan, 32 entries per round
        int n = 0;
        bool bad = false;
        for (int base = 0; base < n_in; base += 32) {
            const int i = base + lane;
            float sc = 0.f;
            bool keep = false;
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_in) {
                sc = D[i * 6 + 4];
                keep = !(sc < thr_score);
                bx = make_float4(D[i * 6 + 0], D[i * 6 + 1], D[i * 6 + 2], D[i * 6 + 3]);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = n + __popc(m & ((1u << lane) - 1u));
                if (!(bx.x < bx.z) || !(bx.y < bx.w)) bad = true;  // get_iou would assert (utils.py:78-81)
                s_box[pos] = bx;
                s_score[pos] = sc;
                s_cls[pos] = (int)D[i * 6 + 5];
                s_src[pos] = i;
            }
            n += __popc(m);
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            s_N = n;
            s_bad = (bad && n >= 2) ? 1 : 0;
        }
    }
    __syncthreads();
    const int N = s_N;
    int* ko = keep_idx + (long long)b * det_stride;
    if (s_bad) {
        if (threadIdx.x == 0) {
            nkeep[b] = 0;
            status[b] = -2;
        }
        return;
    }
    // adjacency: one warp per row, lane = column inside a 32-column word, the word is the ballot of the link test.  The
    // full matrix is computed (the fp32 IoU is symmetric bit for bit: max / min / commutative add), so every word of every
    // row is written exactly once by its owner: no atomics, no zero fill, and the densest tile (300 boxes) costs 19 rows
    // x 10 words per warp instead of 600 sequential tests per thread.
    for (int i = warp; i < N; i += kMergeThreads / 32) {
        const float4 bi = s_box[i];
        const int ci = s_cls[i];
        for (int w = 0; w < kMergeWords; ++w) {
            const int j = w * 32 + lane;
            bool link = false;
            if (j < N && j != i) {
                const float iou = ref_get_iou(bi, s_box[j]);
                link = iou >= thr_hard || (ci == s_cls[j] && iou >= thr_soft);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, link);
            if (lane == 0) s_adj[i][w] = m;
        }
    }
    __syncthreads();
    if (warp == 0) {
        // iterative emulation of the recursive DFS (graph.py:9-23), adjacency ascending; all lanes run the same control
        // flow, lane w holds word w of the visited set
        uint32_t vis = 0u;
        int nk = 0;
        for (int v = 0; v < N; ++v) {
            const uint32_t vw = __shfl_sync(0xffffffffu, vis, v >> 5);
            if ((vw >> (v & 31)) & 1u) continue;
            float sbest = 0.f;
            int best = -1;
            int sp = 0;
            if (lane == 0) s_stack[0] = (short)v;
            sp = 1;
            if (lane == (v >> 5)) vis |= 1u << (v & 31);
            if (s_score[v] > sbest) {
                sbest = s_score[v];
                best = v;
            }
            int u = v;                                  // top of the stack
            while (sp > 0) {
                // next unvisited neighbour of u in ascending order
                const uint32_t m = lane < kMergeWords ? (s_adj[u][lane] & ~vis) : 0u;
                const uint32_t bal = __ballot_sync(0xffffffffu, m != 0u);
                if (bal == 0u) {
                    --sp;
                    __syncwarp();
                    if (sp > 0) u = s_stack[sp - 1];
                    continue;
                }
                const int wl = __ffs(bal) - 1;
                const uint32_t mm = __shfl_sync(0xffffffffu, m, wl);
                const int nxt = wl * 32 + __ffs(mm) - 1;
                if (lane == wl) vis |= 1u << (nxt & 31);
                const float sn = s_score[nxt];
                if (sn > sbest) {  // preorder visit
                    sbest = sn;
                    best = nxt;
                }
                if (lane == 0) s_stack[sp] = (short)nxt;
                ++sp;
                u = nxt;
            }
            // best == -1 only when every score in the component is <= 0; the reference then indexes [-1]
            if (lane == 0) s_keep[nk] = best >= 0 ? s_src[best] : s_src[N - 1];
            ++nk;
        }
        if (lane == 0) {
            s_nk = nk;
            nkeep[b] = nk;
            status[b] = 0;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s_nk; i += blockDim.x) ko[i] = s_keep[i];
}

// ------------------------------------------------------------------------------------------ host launchers

static HeadLevels make_levels(const float* h0, const float* h1, const float* h2, int Sh, int Sw) {
    HeadLevels L;
    L.p[0] = h0; L.p[1] = h1; L.p[2] = h2;
    int a = 0;
    for (int l = 0; l < 3; ++l) {
        const int s = 8 << l;
        L.h[l] = Sh / s;
        L.w[l] = Sw / s;
        L.a0[l] = a;
        a += L.h[l] * L.w[l];
    }
    L.A = a;
    return L;
}

int num_anchors(int Sh, int Sw) {
    return (Sh / 8) * (Sw / 8) + (Sh / 16) * (Sw / 16) + (Sh / 32) * (Sw / 32);
}

int decode_pred(const float* h0, const float* h1, const float* h2, int B, int Sh, int Sw, int nc, float* pred,
                cudaStream_t st) {
    HeadLevels L = make_levels(h0, h1, h2, Sh, Sw);
    const long long total = (long long)B * L.A;
    decode_pred_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L, B, nc, pred);
    return (int)cudaGetLastError();
}

static int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

size_t postprocess_scratch_bytes(int B, int Sh, int Sw, int max_det) {
    const int np2 = next_pow2(num_anchors(Sh, Sw));
    return (size_t)B * np2 * (sizeof(unsigned long long) + sizeof(float4)) + (size_t)B * max_det * sizeof(int) +
           (size_t)B * sizeof(int) + 256 + (size_t)B * kSelKeys * sizeof(unsigned long long) + 256;
}

int postprocess(const float* h0, const float* h1, const float* h2, int B, int Sh, int Sw, int nc, float conf,
                float iou, int max_det, const LetterboxInfo* lb, float* dets, int* ndets, void* scratch,
                cudaStream_t st) {
    HeadLevels L = make_levels(h0, h1, h2, Sh, Sw);
    if (L.A >= (1 << 20)) return -1;
    const int np2 = next_pow2(L.A);
    unsigned long long* keys = (unsigned long long*)scratch;
    float4* boxes = (float4*)(keys + (size_t)B * np2);
    int* keep = (int*)(boxes + (size_t)B * np2);
    int* cand = keep + (size_t)B * max_det;
    unsigned long long* sel = (unsigned long long*)(((uintptr_t)(cand + B) + 255) & ~(uintptr_t)255);
    const int dyn = kSelKeys * (int)sizeof(unsigned long long) + kSelBins * (int)sizeof(int);
    static std::atomic<unsigned long long> attr_done{0};
    if (first_use_on_device(attr_done)) {
        if (cudaFuncSetAttribute(nms_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) return -1;
    }
    const long long total = (long long)B * L.A;
    cudaMemsetAsync(cand, 0, (size_t)B * sizeof(int), st);
    score_key_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L, B, nc, conf, keys, np2, cand);
    nms_tiles_kernel<<<B, kNmsThreads, dyn, st>>>(L, B, keys, np2, cand, boxes, keep, sel, conf, (double)iou, max_det,
                                                  30000, 7680.f, lb, dets, ndets);
    return (int)cudaGetLastError();
}

size_t nms_scratch_bytes(int B, int N) {
    const int np2 = next_pow2(N < 2 ? 2 : N);
    return (size_t)B * np2 * sizeof(unsigned long long) + (size_t)B * N * (sizeof(float4) + sizeof(int)) + 256;
}

int nms_batched(const float* boxes, const float* scores, const int* counts, int B, int N, double thr, int max_keep,
                long long* keep, int* nkeep, void* scratch, cudaStream_t st) {
    const int np2 = next_pow2(N < 2 ? 2 : N);
    unsigned long long* keys = (unsigned long long*)scratch;
    float4* bs = (float4*)(keys + (size_t)B * np2);
    int* kp = (int*)(bs + (size_t)B * N);
    nms_generic_kernel<<<B, kNmsThreads, 0, st>>>(boxes, scores, counts, N, np2, keys, bs, kp, thr,
                                                  max_keep > 0 ? max_keep : N, keep, nkeep);
    return (int)cudaGetLastError();
}

int merge_tiles(const float* dets, const int* ndets, int B, int det_stride, float thr_score, float thr_soft,
                float thr_hard, const int* pre_status, int* keep_idx, int* nkeep, int* status, cudaStream_t st) {
    if (det_stride > kMergeMaxN) return -1;
    merge_tile_kernel<<<B, kMergeThreads, 0, st>>>(dets, ndets, det_stride, thr_score, thr_soft, thr_hard, pre_status, keep_idx,
                                         nkeep, status);
    return (int)cudaGetLastError();
}

}  // namespace cy
