// tcgen05 implicit-GEMM convolution kernel + host-side planning (tensor-map encoding).  See conv.cuh.
#include "conv.cuh"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace cy {

static constexpr int kConvThreads = 192;  // warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps2-5 = epilogue
static constexpr int kBlockM = 128;

template <int BLOCK_N>
__global__ void __launch_bounds__(kConvThreads) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t row_bytes = p.kc * 2;
    const uint32_t a_bytes = kBlockM * row_bytes;
    const uint32_t b_bytes_raw = BLOCK_N * row_bytes;
    const uint32_t b_bytes = (b_bytes_raw + 1023u) & ~1023u;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int stages = p.stages;

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* accum_bar = empty_bar + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

    // tile coordinates: blockIdx.x = N tile (fast, shares the A tile through L2), blockIdx.y = M tile
    const int n_tile = blockIdx.x;
    int mt = blockIdx.y;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    const int tn = mt / p.tiles_h;
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int k_iters = p.ntaps * p.cchunks;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const CUtensorMap* ma = &p.tmA[p.tap_map[tap]];
                const int cw = w0 + p.tap_dw[tap];
                const int ch = h0 + p.tap_dh[tap];
                for (int ck = 0; ck < p.cchunks; ++ck, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], a_bytes + b_bytes_raw);
                    tma_load_4d(sa, ma, &full_bar[s], ck * p.kc, cw, ch, n0);
                    tma_load_2d(sb, &p.tmB, &full_bar[s], tap * p.cin + ck * p.kc, n_tile * BLOCK_N);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
        const int ksteps = p.kc >> 4;
        for (int it = 0; it < k_iters; ++it) {
            const int s = it % stages;
            const uint32_t ph = (it / stages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t da = make_smem_desc(sa, row_bytes);
                const uint64_t db = make_smem_desc(sa + a_bytes, row_bytes);
                for (int k = 0; k < ksteps; ++k) {
                    // advancing 16 bf16 (=32 B) along K inside the swizzle atom = +2 in the 16-byte address field
                    umma_bf16(tmem_base, da + 2u * k, db + 2u * k, idesc, (it | k) != 0);
                }
                umma_commit(&empty_bar[s]);
                if (it == k_iters - 1) umma_commit(accum_bar);
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> bias + SiLU (+ residual) -> global NHWC
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int lw = row % p.bw;
        const int lh = (row / p.bw) % p.bh;
        const int ln = row / (p.bw * p.bh);
        const int ow = w0 + lw, oh = h0 + lh, on = n0 + ln;
        const bool valid = (ow < p.W) && (oh < p.H) && (on < p.B);
        const long long pix = ((long long)on * p.H + oh) * p.W + ow;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr int CW = BLOCK_N >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += CW) {
            __syncwarp();
            float x[CW];
            if constexpr (CW == 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            } else {
                uint32_t v[16];
                tmem_ld_32x16(taddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
            }
            const int col0 = n_tile * BLOCK_N + c0;
            if (!valid) continue;
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int g = 0; g < CW / 8; ++g) {
                const int col = col0 + g * 8;
                if (col + 8 > p.cout_store) break;
                const float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
                float y[8] = {x[g * 8 + 0] + ba.x, x[g * 8 + 1] + ba.y, x[g * 8 + 2] + ba.z, x[g * 8 + 3] + ba.w,
                              x[g * 8 + 4] + bb.x, x[g * 8 + 5] + bb.y, x[g * 8 + 6] + bb.z, x[g * 8 + 7] + bb.w};
                if (p.act) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = __fdividef(y[j], 1.0f + __expf(-y[j]));
                }
                if (p.res) {
                    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.res_cstride + p.res_coff + col));
                    const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(r2[j]);
                        y[2 * j] += f.x;
                        y[2 * j + 1] += f.y;
                    }
                }
                if (p.out_f32) {
                    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.out_cstride +
                                                          p.out_coff + col);
                    o[0] = make_float4(y[0], y[1], y[2], y[3]);
                    o[1] = make_float4(y[4], y[5], y[6], y[7]);
                } else {
                    uint4 o;
                    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_cstride +
                                              p.out_coff + col) = o;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ host side

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
    return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int conv_block_n(int cout) {
    static int force = -1;
    if (force < 0) {
        const char* e = getenv("CY_CONV_BLOCK_N");
        force = e ? atoi(e) : 0;
    }
    if (cout <= 16) return 16;
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    if (force == 256 && cout % 256 == 0) return 256;
    return 128;
}

static void choose_box(int B, int H, int W, int* bw, int* bh, int* bn) {
    double best = -1;
    for (int w = 1; w <= 128; w *= 2)
        for (int h = 1; w * h <= 128; h *= 2) {
            int n = 128 / (w * h);
            double cover = (double)((W + w - 1) / w * w) * ((H + h - 1) / h * h) * ((B + n - 1) / n * n);
            double util = (double)W * H * B / cover;
            // prefer wide boxes (longer contiguous runs) on ties, then tall
            double score = util + 1e-6 * w + 1e-9 * h;
            if (score > best) {
                best = score;
                *bw = w;
                *bh = h;
                *bn = n;
            }
        }
}

template <int BN>
static int launch_t(const ConvPlan& pl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    conv_igemm_kernel<BN><<<pl.grid, kConvThreads, pl.smem, st>>>(pl.kp);
    return (int)cudaGetLastError();
}

int conv_launch(const ConvPlan& pl, cudaStream_t st) {
    switch (pl.block_n) {
        case 16: return launch_t<16>(pl, st);
        case 32: return launch_t<32>(pl, st);
        case 64: return launch_t<64>(pl, st);
        case 128: return launch_t<128>(pl, st);
        case 256: return launch_t<256>(pl, st);
    }
    return -1;
}

int conv_make_plan(const ConvDesc& d, ConvPlan* plan, char* err, size_t errlen) {
#define FAIL(...)                          \
    do {                                   \
        snprintf(err, errlen, __VA_ARGS__); \
        return -1;                         \
    } while (0)
    auto enc = get_encode();
    if (!enc) FAIL("cuTensorMapEncodeTiled entry point not available");
    if (d.ksize != 1 && d.ksize != 3) FAIL("ksize must be 1 or 3");
    if (d.stride != 1 && d.stride != 2) FAIL("stride must be 1 or 2");
    if (d.ksize == 1 && d.stride != 1) FAIL("1x1 stride 2 unsupported");
    if (d.cin % 16) FAIL("cin must be a multiple of 16 (got %d)", d.cin);
    if (d.in_ctot % 8 || d.in_coff % 8) FAIL("input channel stride/offset must be multiples of 8");
    if (d.stride == 2 && ((d.Hin | d.Win) & 1)) FAIL("stride-2 conv needs even input extent");
    const int kc = d.cin % 64 == 0 ? 64 : (d.cin % 32 == 0 ? 32 : 16);
    const int row_bytes = kc * 2;
    const int bn_ = conv_block_n(d.cout);
    if (d.cout_pad % bn_) FAIL("cout_pad %d not a multiple of BLOCK_N %d", d.cout_pad, bn_);
    ConvKParams& kp = plan->kp;
    memset(&kp, 0, sizeof(kp));
    const int Hout = d.stride == 1 ? d.Hin : d.Hin / 2;
    const int Wout = d.stride == 1 ? d.Win : d.Win / 2;
    kp.B = d.B;
    kp.H = Hout;
    kp.W = Wout;
    choose_box(d.B, Hout, Wout, &kp.bw, &kp.bh, &kp.bn);
    kp.tiles_w = (Wout + kp.bw - 1) / kp.bw;
    kp.tiles_h = (Hout + kp.bh - 1) / kp.bh;
    kp.tiles_n = (d.B + kp.bn - 1) / kp.bn;
    kp.ntaps = d.ksize * d.ksize;
    kp.kc = kc;
    kp.cin = d.cin;
    kp.cchunks = d.cin / kc;
    kp.bias = d.bias;
    kp.out = d.out;
    kp.out_cstride = d.out_ctot;
    kp.out_coff = d.out_coff;
    kp.res = d.res;
    kp.res_cstride = d.res_ctot;
    kp.res_coff = d.res_coff;
    kp.cout_store = (d.cout + 7) / 8 * 8;
    kp.act = d.act;
    kp.out_f32 = d.out_f32;
    if ((d.out_ctot % 8) || (d.out_coff % 8)) FAIL("output channel stride/offset must be multiples of 8");
    if (kp.cout_store + d.out_coff > d.out_ctot) FAIL("output slice exceeds buffer channels");

    // ---- A tensor maps
    const CUtensorMapSwizzle sw = swizzle_for(row_bytes);
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)kp.bw, (cuuint32_t)kp.bh, (cuuint32_t)kp.bn};
    const char* base = reinterpret_cast<const char*>(d.in) + (size_t)d.in_coff * 2;
    const size_t pixb = (size_t)d.in_ctot * 2;
    if (d.stride == 1) {
        const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.Win, (cuuint64_t)d.Hin, (cuuint64_t)d.B};
        const cuuint64_t str[3] = {pixb, pixb * d.Win, pixb * d.Win * d.Hin};
        CUresult r = enc(&kp.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, str, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(A) failed: %d", (int)r);
        int t = 0;
        for (int kh = 0; kh < d.ksize; ++kh)
            for (int kw = 0; kw < d.ksize; ++kw, ++t) {
                kp.tap_map[t] = 0;
                kp.tap_dh[t] = (signed char)(kh - d.ksize / 2);
                kp.tap_dw[t] = (signed char)(kw - d.ksize / 2);
            }
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)(d.Win / 2), (cuuint64_t)(d.Hin / 2),
                                            (cuuint64_t)d.B};
                const cuuint64_t str[3] = {pixb * 2, pixb * d.Win * 2, pixb * d.Win * d.Hin};
                const char* b2 = base + ((size_t)ph * d.Win + pw) * pixb;
                CUresult r = enc(&kp.tmA[ph * 2 + pw], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)b2, dims, str, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(A s2) failed: %d", (int)r);
            }
        // input row ih = 2*oh + kh - 1:  kh=0 -> odd rows, coord oh-1; kh=1 -> even rows, coord oh; kh=2 -> odd, oh
        int t = 0;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw, ++t) {
                const int ph = (kh != 1), pw = (kw != 1);
                kp.tap_map[t] = (signed char)(ph * 2 + pw);
                kp.tap_dh[t] = (signed char)(kh == 0 ? -1 : 0);
                kp.tap_dw[t] = (signed char)(kw == 0 ? -1 : 0);
            }
    }
    // ---- B tensor map: weights [cout_pad, ntaps*cin], K fastest
    {
        const cuuint64_t dims[2] = {(cuuint64_t)kp.ntaps * d.cin, (cuuint64_t)d.cout_pad};
        const cuuint64_t str[1] = {(cuuint64_t)kp.ntaps * d.cin * 2};
        const cuuint32_t bbox[2] = {(cuuint32_t)kc, (cuuint32_t)bn_};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&kp.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d.w, dims, str, bbox, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    // ---- pipeline depth from the shared-memory budget: aim for 2 CTAs per SM
    const size_t a_bytes = (size_t)kBlockM * row_bytes;
    const size_t b_bytes = ((size_t)bn_ * row_bytes + 1023) & ~(size_t)1023;
    const size_t stage = a_bytes + b_bytes;
    const size_t budget = bn_ >= 256 ? 200 * 1024 : 100 * 1024;
    int stages = (int)(budget / stage);
    const int k_iters = kp.ntaps * kp.cchunks;
    if (stages > 8) stages = 8;
    if (stages > k_iters) stages = k_iters;
    if (stages < 1) stages = 1;
    kp.stages = stages;
    plan->smem = (size_t)stages * stage + (2 * stages + 1) * 8 + 16 + 1024;
    plan->block_n = bn_;
    plan->grid = dim3(d.cout_pad / bn_, kp.tiles_w * kp.tiles_h * kp.tiles_n, 1);
    plan->flops = 2.0 * d.B * Hout * Wout * (double)d.cout * kp.ntaps * d.cin;
    return 0;
#undef FAIL
}

}  // namespace cy
