// tcgen05 implicit-GEMM convolution kernel + host-side planning (tensor-map encoding).  See conv.cuh.
#include "conv.cuh"
#include "common.h"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace cy {

// warp roles: 0 = A-box TMA producer, 1 = TMEM alloc + MMA issuer, 2 = weight-tile TMA producer, 3..10 = epilogue
static constexpr int kNumEpiWarps = 8;
static constexpr int kFirstEpiWarp = 3;
static constexpr int kConvThreads = (kFirstEpiWarp + kNumEpiWarps) * 32;
static constexpr int kBlockM = 128;
static constexpr int kEpiBufBytes = 4096;   // 32 rows x 128 B staging buffer of one epilogue warp

struct HalfCoord {
    int w0, h0, n0;
};

__device__ __forceinline__ HalfCoord half_coord(const ConvKParams& p, int ht) {
    HalfCoord c;
    if (ht >= p.n_half_tiles) {  // padding half of the last unit: fully out of bounds -> zero box, masked stores
        c.w0 = 0;
        c.h0 = 0;
        c.n0 = p.B;
        return c;
    }
    const int tw = ht % p.tiles_w;
    ht /= p.tiles_w;
    const int th = ht % p.tiles_h;
    const int tn = ht / p.tiles_h;
    c.w0 = tw * p.bw;
    c.h0 = th * p.bh;
    c.n0 = tn * p.bn;
    return c;
}

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, uint32_t row_bytes, uint32_t sbo_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           (layout << 61);
}

// SiLU(y) = y * sigmoid(y) = h + h * tanh(h), h = y / 2: one MUFU op per element (tanh.approx.f32, abs. error <= 2^-11
// on tanh, i.e. below the bf16 rounding of the stored activation) instead of ex2 + rcp.
__device__ __forceinline__ float silu_f(float y) {
    const float h = 0.5f * y;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// 32 (or 16) accumulator columns of one row -> bias, SiLU, (+ residual read in place), 16-byte chunks into the
// swizzled staging row.  cc0 = first column inside the store group.
template <bool F32OUT>
__device__ __forceinline__ void epi_cols32(const uint32_t (&v)[32], const float4 (&bias4)[8], int act, bool has_res,
                                           uint8_t* sbuf, uint32_t my_row, uint32_t sw_mask, int cc0, int ncols, int f16) {
    // stage 1: bias (+ SiLU) on all 32 columns with the 32 MUFU ops issued back to back (their latency overlaps)
    float y[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        y[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bias4[j].x;
        y[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bias4[j].y;
        y[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bias4[j].z;
        y[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bias4[j].w;
    }
    if (act == 1) {
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            y[j] *= 0.5f;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t[j]) : "f"(y[j]));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = fmaf(y[j], t[j], y[j]);
    } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = __fdividef(y[j], 1.0f + __expf(-y[j]));
    }
    // stage 2: (+ residual read in place) and 16-byte chunks into the swizzled staging row
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
        if (g8 * 8 < ncols) {
            const int cc = cc0 + g8 * 8;
            if (!F32OUT) {
                uint32_t a = my_row + (uint32_t)(cc >> 3) * 16u;
                a ^= (a >> 3) & sw_mask;
                uint4* dst = reinterpret_cast<uint4*>(sbuf + a);
                if (has_res) {
                    const uint4 r = *dst;
                    const uint32_t* r2 = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_h2(r2[j], f16);
                        y[g8 * 8 + 2 * j] += f.x;
                        y[g8 * 8 + 2 * j + 1] += f.y;
                    }
                }
                uint4 o;
                uint32_t* o2 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) o2[j] = pack_h2(y[g8 * 8 + 2 * j], y[g8 * 8 + 2 * j + 1], f16);
                *dst = o;
            } else {
                uint32_t a0 = my_row + (uint32_t)(cc >> 2) * 16u;
                uint32_t a1 = a0 + 16u;
                a0 ^= (a0 >> 3) & sw_mask;
                a1 ^= (a1 >> 3) & sw_mask;
                *reinterpret_cast<float4*>(sbuf + a0) =
                    make_float4(y[g8 * 8 + 0], y[g8 * 8 + 1], y[g8 * 8 + 2], y[g8 * 8 + 3]);
                *reinterpret_cast<float4*>(sbuf + a1) =
                    make_float4(y[g8 * 8 + 4], y[g8 * 8 + 5], y[g8 * 8 + 6], y[g8 * 8 + 7]);
            }
        }
    }
}

template <int HALVES, int KSTEPS, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Programmatic dependent launch: all CTAs of a conv grid are resident at once (one per SM), so the next conv of the
    // stream may be scheduled onto an SM as soon as this grid's CTA there has exited; it runs its prologue (barrier
    // init, TMEM allocation, descriptor prefetch) during this grid's tail and blocks in
    // grid_dep_wait() below until this grid has completed.
    grid_dep_launch_dependents();
    const uint32_t row_bytes = p.kc * 2;
    const int a_stages = p.a_stages, b_stages = p.b_stages;
    // PAIR: two CTAs of a cluster (one TPC) work on one unit with tcgen05 cta_group::2: every MMA covers 128 rows
    // from each CTA (own A boxes) and block_n/2 weight rows from each CTA's shared memory, so each SM reads half of
    // B per flop.  The rank-0 CTA issues the MMAs; barriers "full" live in the leader, "empty" are multicast to both.
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int kCtasPerUnit = PAIR ? 2 : 1;

    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)a_stages * p.a_stage_bytes;
    uint8_t* epi_ring = b_ring + (size_t)b_stages * p.b_stage_bytes;   // [kNumEpiWarps][2][kEpiBufBytes]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(epi_ring + (size_t)kNumEpiWarps * 2 * kEpiBufBytes);
    uint64_t* a_empty = a_full + a_stages;
    uint64_t* b_full = a_empty + a_stages;
    uint64_t* b_empty = b_full + b_stages;
    uint64_t* acc_full = b_empty + b_stages;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* res_bar = acc_empty + 2;                       // [kNumEpiWarps][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kNumEpiWarps);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmO);
        for (int s = 0; s < a_stages; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < b_stages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kNumEpiWarps * kCtasPerUnit);
        }
        for (int s = 0; s < 2 * kNumEpiWarps; ++s) mbar_init(&res_bar[s], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (PAIR) {
            tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync();   // barrier inits + TMEM allocation of both CTAs visible before any remote signal
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int halves = HALVES;
    // everything past the prologue waits for the preceding grid (weights included: the parity entry cy_conv2d_nhwc may
    // be handed weights that a kernel just before it on the stream produced)
    grid_dep_wait();

    // The three single-issuer roles below run their loops with the WHOLE warp (all control flow warp-uniform) and
    // gate only the issuing instructions with elect.sync: the compiler then keeps descriptors / coordinates in uniform
    // registers and emits back-to-back UTCHMMA / UTMALDG without per-instruction R2UR waterfall loops.
    if (warp == 0) {
        // ------------------------------------------------------------ A-box producer
        int s = 0;
        uint32_t ph = 0;
        for (int u = worker; u < p.n_units; u += n_workers) {
            const int mu = u / p.n_tiles_n;
            const HalfCoord hc0 = half_coord(p, (mu * kCtasPerUnit + (int)rank) * HALVES);
            const HalfCoord hc1 = half_coord(p, (mu * kCtasPerUnit + (int)rank) * HALVES + 1);
            for (int ck = 0; ck < p.cchunks; ++ck) {
                for (int al = 0; al < p.n_aloads; ++al) {
                    const ConvALoad L = p.aload[al];
                    mbar_wait(&a_empty[s], ph ^ 1);
                    if (elect_one()) {
                        uint8_t* dst = a_ring + (size_t)s * p.a_stage_bytes;
                        if (rank == 0) mbar_arrive_expect_tx(&a_full[s], L.bytes * HALVES * kCtasPerUnit);
                        if (PAIR) {
                            const uint32_t bar = mapa_u32(smem_u32(&a_full[s]), 0);
                            tma_load_4d_pair(dst, &p.tmA[L.map], bar, ck * p.kc, hc0.w0 + L.dw, hc0.h0 + L.dh, hc0.n0);
                            if (HALVES == 2)
                                tma_load_4d_pair(dst + p.a_half_stride, &p.tmA[L.map], bar, ck * p.kc, hc1.w0 + L.dw,
                                                 hc1.h0 + L.dh, hc1.n0);
                        } else {
                            tma_load_4d(dst, &p.tmA[L.map], &a_full[s], ck * p.kc, hc0.w0 + L.dw, hc0.h0 + L.dh,
                                        hc0.n0);
                            if (HALVES == 2)
                                tma_load_4d(dst + p.a_half_stride, &p.tmA[L.map], &a_full[s], ck * p.kc,
                                            hc1.w0 + L.dw, hc1.h0 + L.dh, hc1.n0);
                        }
                    }
                    __syncwarp();
                    if (++s == a_stages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------ weight-tile producer
        int s = 0;
        uint32_t ph = 0;
        const int b_rows = p.block_n / kCtasPerUnit;   // weight rows this CTA stages per tile
        for (int u = worker; u < p.n_units; u += n_workers) {
            const int n_tile = u % p.n_tiles_n;
            for (int ck = 0; ck < p.cchunks; ++ck) {
                for (int al = 0; al < p.n_aloads; ++al) {
                    const ConvALoad L = p.aload[al];
                    for (int t = 0; t < L.ntaps; ++t) {
                        const int wtap = p.tap[L.tap0 + t].wtap;
                        mbar_wait(&b_empty[s], ph ^ 1);
                        if (elect_one()) {
                            if (rank == 0) mbar_arrive_expect_tx(&b_full[s], p.b_tile_bytes * kCtasPerUnit);
                            if (PAIR)
                                tma_load_2d_pair(b_ring + (size_t)s * p.b_stage_bytes, &p.tmB,
                                                 mapa_u32(smem_u32(&b_full[s]), 0), wtap * p.cin + ck * p.kc,
                                                 n_tile * p.block_n + (int)rank * b_rows);
                            else
                                tma_load_2d(b_ring + (size_t)s * p.b_stage_bytes, &p.tmB, &b_full[s],
                                            wtap * p.cin + ck * p.kc, n_tile * p.block_n);
                        }
                        __syncwarp();
                        if (++s == b_stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
      if (rank == 0) {
        // ------------------------------------------------------------ MMA issuer (one elected thread of the leader)
        const uint32_t idesc = make_idesc_bf16(kBlockM * kCtasPerUnit, p.block_n, p.f16);
        const uint32_t a_ring_addr = smem_u32(a_ring);
        const uint32_t b_ring_addr = smem_u32(b_ring);
        int sa = 0, sb = 0;
        uint32_t pha = 0, phb = 0;
        int i = 0;
        for (int u = worker; u < p.n_units; u += n_workers, ++i) {
            const int buf = p.acc_bufs == 2 ? (i & 1) : 0;
            const uint32_t use = p.acc_bufs == 2 ? (uint32_t)(i >> 1) : (uint32_t)i;
            unsigned long long* dbg =
                (p.dbg && i < p.dbg_units) ? p.dbg + ((size_t)worker * p.dbg_units + i) * 8 : nullptr;
            long long wait_a = 0, wait_b = 0;
            if (p.dbg_epi) dbg = nullptr;
            if (dbg && lane == 0) dbg[0] = clock64();
            mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
            tc_fence_after();
            if (dbg && lane == 0) dbg[1] = clock64();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * HALVES * p.acc_stride);
            uint32_t accum = 0;
            for (int ck = 0; ck < p.cchunks; ++ck) {
                for (int al = 0; al < p.n_aloads; ++al) {
                    const ConvALoad L = p.aload[al];
                    if (dbg) wait_a -= clock64();
                    mbar_wait(&a_full[sa], pha);
                    if (dbg) wait_a += clock64();
                    const uint32_t a_addr = a_ring_addr + (uint32_t)sa * p.a_stage_bytes;
                    for (int t = 0; t < L.ntaps; ++t) {
                        const uint32_t a_off = p.tap[L.tap0 + t].a_off;
                        if (dbg) wait_b -= clock64();
                        mbar_wait(&b_full[sb], phb);
                        if (dbg) wait_b += clock64();
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t db = make_desc_sbo(b_ring_addr + (uint32_t)sb * p.b_stage_bytes, row_bytes,
                                                              8u * row_bytes);
                            const uint64_t da0 = make_desc_sbo(a_addr + a_off, row_bytes, L.sbo);
#pragma unroll
                            for (int h = 0; h < HALVES; ++h) {
                                const uint64_t da = da0 + (uint64_t)((h * p.a_half_stride) >> 4);
                                const uint32_t td = tacc + (uint32_t)(h * p.acc_stride);
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) {
                                    // advancing 16 bf16 (=32 B) along K inside the swizzle row = +2 in the 16 B field
                                    if (PAIR) umma_bf16_pair(td, da + 2u * k, db + 2u * k, idesc, accum | (uint32_t)k);
                                    else umma_bf16(td, da + 2u * k, db + 2u * k, idesc, accum | (uint32_t)k);
                                }
                            }
                            if (PAIR) umma_commit_pair(&b_empty[sb]);
                            else umma_commit(&b_empty[sb]);
                        }
                        __syncwarp();
                        accum = 1;
                        if (++sb == b_stages) {
                            sb = 0;
                            phb ^= 1;
                        }
                    }
                    if (elect_one()) {
                        if (PAIR) umma_commit_pair(&a_empty[sa]);
                        else umma_commit(&a_empty[sa]);
                    }
                    __syncwarp();
                    if (++sa == a_stages) {
                        sa = 0;
                        pha ^= 1;
                    }
                }
            }
            if (elect_one()) {
                if (PAIR) umma_commit_pair(&acc_full[buf]);
                else umma_commit(&acc_full[buf]);
            }
            __syncwarp();
            if (dbg && lane == 0) {
                dbg[2] = clock64();
                dbg[3] = (unsigned long long)wait_b;
                dbg[4] = (unsigned long long)wait_a;
            }
        }
      }
    } else {
        // ------------------------------------------------------------ epilogue
        // TMEM -> registers -> bias + SiLU -> swizzled staging buffer (+ residual, TMA-loaded into the same buffer,
        // added in place) -> TMA store of the warp's 32-row sub box.  Every warp owns two 4 KB staging buffers and
        // issues its own loads/stores: no cross-warp synchronisation, out-of-image rows/channels are clipped by TMA.
        const int e = warp - kFirstEpiWarp;
        const int q = warp & 3;      // TMEM lane quarter this warp may access
        const int part = e >> 2;     // the two warps of a quarter split halves (halves == 2) or column groups
        uint8_t* stage_base = epi_ring + (size_t)e * 2 * kEpiBufBytes;
        uint64_t* rbar = res_bar + e * 2;
        const int r0 = q * 32;
        const int sub_w = r0 % p.bw, sub_h = (r0 / p.bw) % p.bh, sub_n = r0 / (p.bw * p.bh);
        const int gw = p.o_gw;
        const int ngroups = (p.block_n + gw - 1) / gw;
        int g_begin = 0, g_end = ngroups;
        if (halves == 1) {
            const int mid = (ngroups + 1) / 2;
            if (part == 0) g_end = mid;
            else g_begin = mid;
        }
        const int h = halves == 2 ? part : 0;
        const uint32_t my_row = (uint32_t)lane * p.o_row_bytes;
        int bufsel = 0;
        uint32_t rphase0 = 0, rphase1 = 0;
        int i = 0;
        for (int u = worker; u < p.n_units; u += n_workers, ++i) {
            const int buf = p.acc_bufs == 2 ? (i & 1) : 0;
            const uint32_t use = p.acc_bufs == 2 ? (uint32_t)(i >> 1) : (uint32_t)i;
            const int n_tile = u % p.n_tiles_n;
            const int mu = u / p.n_tiles_n;
            const HalfCoord hc = half_coord(p, (mu * kCtasPerUnit + (int)rank) * halves + h);
            const bool valid_half = hc.n0 < p.B;
            const int bw0 = hc.w0 + sub_w, bh0 = hc.h0 + sub_h, bn0 = hc.n0 + sub_n;
            unsigned long long* dbg = (p.dbg && e == 0 && lane == 0 && rank == 0 && i < p.dbg_units)
                                          ? p.dbg + ((size_t)worker * p.dbg_units + i) * 8
                                          : nullptr;
            if (dbg) dbg[5] = clock64();
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
            if (dbg) dbg[6] = clock64();
            const uint32_t taddr =
                tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * halves + h) * p.acc_stride);
            if (valid_half) {
#pragma unroll 1
                for (int g = g_begin; g < g_end; ++g) {
                    const int c0 = g * gw;
                    const int col0 = n_tile * p.block_n + c0;
                    const int b = bufsel;
                    bufsel ^= 1;
                    uint8_t* sbuf = stage_base + (size_t)b * kEpiBufBytes;
                    long long tq0 = 0, tq1 = 0, tq2 = 0, tq3 = 0;
                    if (dbg) tq0 = clock64();
                    if (lane == 0) bulk_wait_read<1>();  // the store that used this buffer two groups ago has read it
                    __syncwarp();
                    if (dbg) tq1 = clock64();
                    if (p.res != nullptr && lane == 0) {
                        mbar_arrive_expect_tx(&rbar[b], 32u * p.o_row_bytes);
                        tma_load_4d(sbuf, &p.tmR, &rbar[b], col0, bw0, bh0, bn0);
                    }
                    const bool has_res = p.res != nullptr;
                    // both 32-column TMEM reads of the group are issued before the wait (and the bias loads overlap
                    // them), then 64 columns of math run with full instruction-level parallelism
                    const bool two = gw > 32;
                    const int ncols0 = gw >= 32 ? 32 : 16;
                    uint32_t va[32], vb[32];
                    if (ncols0 == 32) {
                        tmem_ld_32x32(taddr + c0, va);
                    } else {
                        uint32_t v16[16];
                        tmem_ld_32x16(taddr + c0, v16);
#pragma unroll
                        for (int j = 0; j < 16; ++j) va[j] = v16[j];
#pragma unroll
                        for (int j = 16; j < 32; ++j) va[j] = 0;
                    }
                    if (two) tmem_ld_32x32(taddr + c0 + 32, vb);
                    float4 bias_a[8], bias_b[8];
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        bias_a[j] = (j * 4 < ncols0) ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        bias_b[j] = two ? __ldg(b4 + 8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    tmem_ld_wait();
                    if (dbg) tq2 = clock64();
                    if (has_res) {
                        mbar_wait(&rbar[b], b ? rphase1 : rphase0);
                        if (b) rphase1 ^= 1;
                        else rphase0 ^= 1;
                    }
                    if (p.o_esz == 2) {
                        epi_cols32<false>(va, bias_a, p.act, has_res, sbuf, my_row, p.o_sw_mask, 0, ncols0, p.f16);
                        if (two) epi_cols32<false>(vb, bias_b, p.act, has_res, sbuf, my_row, p.o_sw_mask, 32, 32, p.f16);
                    } else {
                        epi_cols32<true>(va, bias_a, p.act, false, sbuf, my_row, p.o_sw_mask, 0, ncols0, p.f16);
                    }
                    if (dbg) tq3 = clock64();
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&p.tmO, sbuf, col0, bw0, bh0, bn0);
                        bulk_commit();
                    }
                    if (dbg && p.dbg_epi) {
                        const long long tq4 = clock64();
                        dbg[0] += tq1 - tq0; dbg[1] += tq2 - tq1; dbg[2] += tq3 - tq2; dbg[3] += tq4 - tq3;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
                else mbar_arrive(&acc_empty[buf]);
            }
            if (dbg) dbg[7] = clock64();
        }
        if (lane == 0) bulk_wait_all();  // shared memory must outlive the stores that read it
        __syncwarp();
    }

    tc_fence_before();
    if (PAIR) cluster_sync();   // the peer's shared memory / TMEM stay valid until the leader's last MMA retired
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
        else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int num_sms() { return current_device_sms(); }   // per device (common.h)

static unsigned long long* g_dbg = nullptr;
static int g_dbg_units = 0;
void conv_set_debug(unsigned long long* dev_buf, int units_per_cta) {
    g_dbg = dev_buf;
    g_dbg_units = units_per_cta;
}

template <int HALVES, int KSTEPS, bool PAIR>
static int launch_t(const ConvPlan& pl, cudaStream_t st) {
    static std::atomic<unsigned long long> attr_done{0};
    auto kern = conv_igemm_kernel<HALVES, KSTEPS, PAIR>;
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
    }
    ConvKParams kp = pl.kp;
    if (g_dbg) {
        kp.dbg = g_dbg;
        kp.dbg_units = g_dbg_units & 0xffff;
        kp.dbg_epi = (g_dbg_units >> 30) & 1;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = pl.grid;
    cfg.blockDim = dim3(kConvThreads, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const int pdl = env_int("CY_CONV_PDL", 1);      // 0: plain stream order between conv launches
    cfg.numAttrs = pdl ? 2 : 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, kp);
}

int conv_launch(const ConvPlan& pl, cudaStream_t st) {
    const int key = pl.kp.pair * 100 + pl.kp.halves * 10 + (pl.kp.kc >> 4);
    switch (key) {
        case 11: return launch_t<1, 1, false>(pl, st);
        case 12: return launch_t<1, 2, false>(pl, st);
        case 14: return launch_t<1, 4, false>(pl, st);
        case 21: return launch_t<2, 1, false>(pl, st);
        case 22: return launch_t<2, 2, false>(pl, st);
        case 24: return launch_t<2, 4, false>(pl, st);
        case 114: return launch_t<1, 4, true>(pl, st);
        case 124: return launch_t<2, 4, true>(pl, st);
    }
    return (int)cudaErrorInvalidValue;
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
    int bn_ = conv_block_n(d.cout);
    if (d.cout_pad % bn_) FAIL("cout_pad %d not a multiple of BLOCK_N %d", d.cout_pad, bn_);
    ConvKParams& kp = plan->kp;
    memset(&kp, 0, sizeof(kp));
    const int Hout = d.stride == 1 ? d.Hin : d.Hin / 2;
    const int Wout = d.stride == 1 ? d.Win : d.Win / 2;
    kp.B = d.B;
    kp.H = Hout;
    kp.W = Wout;
    const int ntaps = d.ksize * d.ksize;

    // ---- mode: tap reuse for 3x3 stride-1 convs with 128-byte channel chunks on maps that 8 x 16 tiles cover well
    int mode = 0;
    // (kc = 32: 64-byte rows with SWIZZLE_64B take the same descriptor-offset reuse; CY_CONV_MODE32=0 disables it)
    if (d.ksize == 3 && d.stride == 1 && (kc == 64 || (kc == 32 && env_int("CY_CONV_MODE32", 1)))) {
        const double util = (double)Wout * Hout / ((double)((Wout + 7) / 8 * 8) * ((Hout + 15) / 16 * 16));
        if (util >= 0.8) mode = 2;
    }
    if (d.ksize == 3 && d.stride == 2 && kc == 64) {
        const double util = (double)Wout * Hout / ((double)((Wout + 7) / 8 * 8) * ((Hout + 15) / 16 * 16));
        if (util >= 0.8) mode = 3;
    }
    const int max_mode = env_int("CY_CONV_MODE", 3);
    if (mode > max_mode) mode = (mode == 3) ? 0 : max_mode;
    kp.mode = mode;
    int box_w, box_h, box_n;          // TMA box in pixels (mode 3: the largest of the four parity boxes)
    if (mode == 0) {
        choose_box(d.B, Hout, Wout, &kp.bw, &kp.bh, &kp.bn);
        box_w = kp.bw; box_h = kp.bh; box_n = kp.bn;
    } else if (mode == 3) {
        kp.bw = 8; kp.bh = 16; kp.bn = 1;
        box_w = 9; box_h = 17; box_n = 1;
    } else {
        kp.bw = 8; kp.bh = 16; kp.bn = 1;
        box_w = mode == 2 ? 10 : 8; box_h = 18; box_n = 1;
    }
    kp.tiles_w = (Wout + kp.bw - 1) / kp.bw;
    kp.tiles_h = (Hout + kp.bh - 1) / kp.bh;
    kp.tiles_n = (d.B + kp.bn - 1) / kp.bn;
    kp.n_half_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
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
    kp.f16 = d.f16 ? 1 : 0;
    kp.act = d.act ? (env_int("CY_CONV_SILU_EXACT", 0) ? 2 : 1) : 0;   // 1: tanh.approx form, 2: ex2 + rcp
    kp.out_f32 = d.out_f32;
    if ((d.out_ctot % 8) || (d.out_coff % 8)) FAIL("output channel stride/offset must be multiples of 8");
    if (kp.cout_store + d.out_coff > d.out_ctot) FAIL("output slice exceeds buffer channels");

    // ---- work decomposition: two half tiles per unit when that still fills the machine
    kp.block_n = bn_;
    kp.n_tiles_n = d.cout_pad / bn_;
    int halves = ((kp.n_half_tiles + 1) / 2) * kp.n_tiles_n >= num_sms() ? 2 : 1;
    const int force_halves = env_int("CY_CONV_HALVES", 0);
    if (force_halves == 1 || force_halves == 2) halves = force_halves;
    // CTA pairs: 128-byte channel chunks, a weight tile that splits into two halves of >= 32 rows, and enough units
    // to give every pair of SMs work
    // Measured per layer (profiles/): pairs win once the K loop is long enough to amortise the cluster hand-shakes
    // (K = taps * cin >= 768, or >= 512 without a residual operand: re-measured per layer at batch 296 — K = 576 / 640
    // layers gain 5-12 % unless they carry a residual (then -11 %), K <= 320 layers lose 9-35 %); short-K layers are
    // epilogue-bound and run better as independent CTAs.
    int pair = 0;
    const bool pair_ok = kc == 64 && bn_ >= 64;
    if (pair_ok && ntaps * d.cin >= (d.res ? 768 : env_int("CY_CONV_PAIR_MINK", 512)) &&
        ((kp.n_half_tiles + 2 * halves - 1) / (2 * halves)) * kp.n_tiles_n >= num_sms() / 2)
        pair = 1;
    const int force_pair = env_int("CY_CONV_PAIR", -1);   // -1 auto, 0 off, 2 force wherever possible
    if (force_pair == 0) pair = 0;
    if (force_pair == 2) pair = pair_ok ? 1 : 0;
    // Wide pair tiles: one 256 x 256 unit per CTA pair (cta_group::2, M = 256, N = 256; each CTA stages its own 128 A
    // rows and 128 of the 256 weight rows) instead of two 128-row halves x N = 128.  Operand rows fetched per K step and
    // CTA: 128 A + 128 B = 256 instead of 256 A + 64 B = 320, and per MMA and SM the shared-memory operand reads drop
    // from 6 KB / 64 clk to 8 KB / 128 clk; the two accumulator buffers are the full 512 TMEM columns.  Used where
    // every tap loads its own A box (1x1 convs, mode 0): with 3x3 halo reuse (mode 2/3) A is nearly free and the
    // weight rows dominate, so the two-half unit (64 weight rows per CTA) stays better there (measured per layer,
    // profiles/r01h_layers_wide_ab.txt).
    const int wide_env = env_int("CY_CONV_WIDE", 1);   // 0 off, 1 auto, 2 force wherever possible (tests)
    if (pair && bn_ == 128 && d.cout_pad % 256 == 0 && !d.out_f32 && wide_env &&
        (wide_env == 2 || (mode == 0 && ((kp.n_half_tiles + 1) / 2) * (d.cout_pad / 256) >= num_sms() / 2))) {
        bn_ = 256;
        halves = 1;
        kp.block_n = bn_;
        kp.n_tiles_n = d.cout_pad / bn_;
    }
    kp.pair = pair;
    kp.halves = halves;
    const int per_unit = halves * (pair ? 2 : 1);
    kp.n_units_m = (kp.n_half_tiles + per_unit - 1) / per_unit;
    kp.n_units = kp.n_units_m * kp.n_tiles_n;
    kp.acc_stride = bn_ < 32 ? 32 : bn_;
    kp.acc_bufs = 2 * halves * kp.acc_stride <= 512 ? 2 : 1;
    int cols = kp.acc_bufs * halves * kp.acc_stride;
    int pow2 = 32;
    while (pow2 < cols) pow2 *= 2;
    if (pow2 > 512) FAIL("accumulators do not fit TMEM");
    kp.tmem_cols = pow2;

    // ---- A-load / tap lists
    const uint32_t box_bytes_dflt = (uint32_t)(box_w * box_h * box_n * row_bytes);
    for (int t = 0; t < 9; ++t) {
        kp.aload[t].bytes = box_bytes_dflt;
        kp.aload[t].sbo = mode == 0 ? 8u * row_bytes : (uint32_t)box_w * row_bytes;
    }
    if (mode == 3) {
        // parity class (ph, pw): ph = 1 -> odd input rows, taps kh = 0 (row oh-1) and kh = 2 (row oh): box rows
        // [h0-1, h0+16); ph = 0 -> kh = 1 (row oh): box rows [h0, h0+16).  Same for columns.
        kp.n_aloads = 4;
        int t = 0;
        for (int ph = 1; ph >= 0; --ph)
            for (int pw = 1; pw >= 0; --pw) {
                const int al = (1 - ph) * 2 + (1 - pw);
                const int bwid = pw ? 9 : 8, bhgt = ph ? 17 : 16;
                kp.aload[al].map = (signed char)(ph * 2 + pw);
                kp.aload[al].dw = (signed char)(pw ? -1 : 0);
                kp.aload[al].dh = (signed char)(ph ? -1 : 0);
                kp.aload[al].tap0 = (signed char)t;
                kp.aload[al].bytes = (uint32_t)(bwid * bhgt * row_bytes);
                kp.aload[al].sbo = (uint32_t)(bwid * row_bytes);
                int nt = 0;
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < 3; ++kw) {
                        if ((kh != 1) != (ph == 1) || (kw != 1) != (pw == 1)) continue;
                        const int di = ph ? (kh == 0 ? 0 : 1) : 0, dj = pw ? (kw == 0 ? 0 : 1) : 0;
                        kp.tap[t].wtap = kh * 3 + kw;
                        kp.tap[t].a_off = (uint32_t)((di * bwid + dj) * row_bytes);
                        ++t;
                        ++nt;
                    }
                kp.aload[al].ntaps = (signed char)nt;
            }
    } else if (mode == 0) {
        kp.n_aloads = ntaps;
        for (int t = 0; t < ntaps; ++t) {
            kp.aload[t].ntaps = 1;
            kp.aload[t].tap0 = (signed char)t;
            kp.tap[t].wtap = t;
            kp.tap[t].a_off = 0;
        }
        if (d.stride == 1) {
            int t = 0;
            for (int kh = 0; kh < d.ksize; ++kh)
                for (int kw = 0; kw < d.ksize; ++kw, ++t) {
                    kp.aload[t].map = 0;
                    kp.aload[t].dh = (signed char)(kh - d.ksize / 2);
                    kp.aload[t].dw = (signed char)(kw - d.ksize / 2);
                }
        } else {
            // input row ih = 2*oh + kh - 1: kh=0 -> odd rows, coord oh-1; kh=1 -> even rows, coord oh; kh=2 -> odd, oh
            int t = 0;
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw, ++t) {
                    const int ph = (kh != 1), pw = (kw != 1);
                    kp.aload[t].map = (signed char)(ph * 2 + pw);
                    kp.aload[t].dh = (signed char)(kh == 0 ? -1 : 0);
                    kp.aload[t].dw = (signed char)(kw == 0 ? -1 : 0);
                }
        }
    } else if (mode == 1) {
        kp.n_aloads = 3;
        for (int kw = 0; kw < 3; ++kw) {
            kp.aload[kw].map = 0;
            kp.aload[kw].dw = (signed char)(kw - 1);
            kp.aload[kw].dh = -1;
            kp.aload[kw].ntaps = 3;
            kp.aload[kw].tap0 = (signed char)(3 * kw);
            for (int kh = 0; kh < 3; ++kh) {
                kp.tap[3 * kw + kh].wtap = kh * 3 + kw;
                kp.tap[3 * kw + kh].a_off = (uint32_t)(kh * box_w * row_bytes);
            }
        }
    } else {
        kp.n_aloads = 1;
        kp.aload[0].map = 0;
        kp.aload[0].dw = -1;
        kp.aload[0].dh = -1;
        kp.aload[0].ntaps = 9;
        kp.aload[0].tap0 = 0;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                kp.tap[kh * 3 + kw].wtap = kh * 3 + kw;
                kp.tap[kh * 3 + kw].a_off = (uint32_t)((kh * box_w + kw) * row_bytes);
            }
    }
    kp.a_half_stride = (box_bytes_dflt + 1023u) & ~1023u;
    kp.a_stage_bytes = kp.a_half_stride * halves;
    kp.b_tile_bytes = (uint32_t)((bn_ / (pair ? 2 : 1)) * row_bytes);   // per CTA
    kp.b_stage_bytes = (kp.b_tile_bytes + 1023u) & ~1023u;

    // ---- A tensor maps
    const CUtensorMapSwizzle sw = swizzle_for(row_bytes);
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
    const char* base = reinterpret_cast<const char*>(d.in) + (size_t)d.in_coff * 2;
    const size_t pixb = (size_t)d.in_ctot * 2;
    if (d.stride == 1) {
        const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.Win, (cuuint64_t)d.Hin, (cuuint64_t)d.B};
        const cuuint64_t str[3] = {pixb, pixb * d.Win, pixb * d.Win * d.Hin};
        CUresult r = enc(&kp.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, str, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)(d.Win / 2), (cuuint64_t)(d.Hin / 2),
                                            (cuuint64_t)d.B};
                const cuuint64_t str[3] = {pixb * 2, pixb * d.Win * 2, pixb * d.Win * d.Hin};
                const char* b2 = base + ((size_t)ph * d.Win + pw) * pixb;
                const cuuint32_t pbox[4] = {(cuuint32_t)kc, (cuuint32_t)(pw ? 9 : 8), (cuuint32_t)(ph ? 17 : 16), 1u};
                CUresult r = enc(&kp.tmA[ph * 2 + pw], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)b2, dims, str,
                                 mode == 3 ? pbox : box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(A s2) failed: %d", (int)r);
            }
    }
    // ---- B tensor map: weights [cout_pad, ntaps*cin], K fastest
    {
        const cuuint64_t dims[2] = {(cuuint64_t)ntaps * d.cin, (cuuint64_t)d.cout_pad};
        const cuuint64_t str[1] = {(cuuint64_t)ntaps * d.cin * 2};
        const cuuint32_t bbox[2] = {(cuuint32_t)kc, (cuuint32_t)(bn_ / (pair ? 2 : 1))};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&kp.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d.w, dims, str, bbox, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    // ---- output / residual tensor maps: one box = the 32 rows of an epilogue warp x o_gw channels (<= 128 bytes)
    {
        const int esz = d.out_f32 ? 4 : 2;
        int gw = 128 / esz;
        if (gw > bn_) gw = bn_;
        kp.o_gw = gw;
        kp.o_esz = esz;
        kp.o_row_bytes = (uint32_t)(gw * esz);
        kp.o_sw_mask = kp.o_row_bytes == 128 ? 0x70u : (kp.o_row_bytes == 64 ? 0x30u : 0x10u);
        if (d.res && d.out_f32) FAIL("residual with fp32 output unsupported");
        const int sw_ = kp.bw < 32 ? kp.bw : 32;
        const int sh_ = kp.bh < 32 / sw_ ? kp.bh : 32 / sw_;
        const int sn_ = 32 / (sw_ * sh_);
        const cuuint32_t obox[4] = {(cuuint32_t)gw, (cuuint32_t)sw_, (cuuint32_t)sh_, (cuuint32_t)sn_};
        const cuuint64_t odims[4] = {(cuuint64_t)kp.cout_store, (cuuint64_t)Wout, (cuuint64_t)Hout, (cuuint64_t)d.B};
        const size_t opix = (size_t)d.out_ctot * esz;
        const cuuint64_t ostr[3] = {opix, opix * Wout, opix * Wout * Hout};
        CUresult r = enc(&kp.tmO, d.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                         (char*)d.out + (size_t)d.out_coff * esz, odims, ostr, obox, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)kp.o_row_bytes),
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(out) failed: %d", (int)r);
        if (d.res) {
            if ((d.res_ctot % 8) || (d.res_coff % 8)) FAIL("residual channel stride/offset must be multiples of 8");
            const size_t rpix = (size_t)d.res_ctot * 2;
            const cuuint64_t rstr[3] = {rpix, rpix * Wout, rpix * Wout * Hout};
            r = enc(&kp.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (char*)d.res + (size_t)d.res_coff * 2, odims, rstr,
                    obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)kp.o_row_bytes),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) FAIL("cuTensorMapEncodeTiled(res) failed: %d", (int)r);
        }
    }
    // ---- ring depths from the shared-memory budget (one persistent CTA per SM)
    const size_t epi_bytes = (size_t)kNumEpiWarps * 2 * kEpiBufBytes;
    const size_t budget = 227 * 1024 - 1024 - 512 - epi_bytes;
    const int a_loads_per_unit = kp.cchunks * kp.n_aloads;
    const int b_loads_per_unit = kp.cchunks * ntaps;
    int a_stages, b_stages;
    if (mode == 0) {
        a_stages = (int)(budget / (kp.a_stage_bytes + kp.b_stage_bytes));
        if (a_stages > 8) a_stages = 8;
        b_stages = a_stages;
    } else {
        a_stages = mode == 3 ? 3 : 2;
        b_stages = (int)((budget - (size_t)a_stages * kp.a_stage_bytes) / kp.b_stage_bytes);
        if (b_stages > 12) b_stages = 12;
    }
    if (a_stages < 1 || b_stages < 1) FAIL("shared-memory budget too small for one stage");
    (void)a_loads_per_unit;
    (void)b_loads_per_unit;
    kp.a_stages = a_stages;
    kp.b_stages = b_stages;
    plan->smem = (size_t)a_stages * kp.a_stage_bytes + (size_t)b_stages * kp.b_stage_bytes + epi_bytes +
                 (size_t)(2 * a_stages + 2 * b_stages + 4 + 2 * kNumEpiWarps) * 8 + 16 + 1024;
    plan->block_n = bn_;
    int g = kp.n_units < num_sms() ? kp.n_units : num_sms();
    if (pair) {
        const int nclusters = kp.n_units < num_sms() / 2 ? kp.n_units : num_sms() / 2;
        g = 2 * nclusters;
    }
    plan->grid = dim3(g, 1, 1);
    plan->flops = 2.0 * d.B * Hout * Wout * (double)d.cout * ntaps * d.cin;
    plan->bytes = (double)d.B * d.Hin * d.Win * d.cin * 2.0 + (double)d.cout_pad * ntaps * d.cin * 2.0 +
                  (double)d.B * Hout * Wout * kp.cout_store * (d.out_f32 ? 4.0 : 2.0) +
                  (d.res ? (double)d.B * Hout * Wout * kp.cout_store * 2.0 : 0.0);
    return 0;
#undef FAIL
}

}  // namespace cy
