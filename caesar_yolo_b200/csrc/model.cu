// YOLOv8 detection model runtime for B200: graph builder, BN folding, weight packing, activation buffers,
// per-shape execution plans (tensor maps cached), and the non-GEMM layer kernels (stem im2col gather, max-pool,
// nearest upsample).  All dense contractions go through the tcgen05 implicit-GEMM kernel in conv.cu.
//
// Replaces the `model(image, ...)` call of the reference (caesar_yolo/evaluation.py:181-193), i.e. ultralytics
// DetectionModel.forward with fused Conv+BN+SiLU / C2f / SPPF / Upsample / Concat / Detect convs
// (yolov8.yaml).  Concats are zero-copy: producers store into channel slices.
#include "model.h"
#include "common.h"
#include <math.h>
#include <string.h>
#include <algorithm>

namespace cy {

int conv_block_n(int cout);

// ------------------------------------------------------------------------------------------ small kernels

// Stem (3x3 stride-2 pad-1 conv, Cin=3) runs on the tensor pipe as a 1x1 conv with K = 32 over an im2col tensor:
// one thread = one output pixel gathers its 3x3x3 patch (input NHWC with 4 channels, 4th ignored) into 32 bf16
// (k = (kh*3+kw)*3 + c, k >= 27 zero) = 64 contiguous bytes.  HBM-bound: reads 8 B/input pixel, writes 64 B/output pixel.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const __nv_bfloat16* __restrict__ in,  // [B,H,W,4]
                                                          __nv_bfloat16* __restrict__ col,       // [B,H/2,W/2,32]
                                                          int B, int H, int W) {
    const int Ho = H / 2, Wo = W / 2;
    const long long npix = (long long)B * Ho * Wo;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const int ow = (int)(pix % Wo);
    const int oh = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((long long)Wo * Ho));
    unsigned short v[32];
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = 2 * oh + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = 2 * ow + kw - 1;
            uint2 q = make_uint2(0u, 0u);
            if (ih >= 0 && ih < H && iw >= 0 && iw < W)
                q = __ldg(reinterpret_cast<const uint2*>(in + (((long long)b * H + ih) * W + iw) * 4));
            const int t = (kh * 3 + kw) * 3;
            v[t + 0] = (unsigned short)(q.x & 0xFFFFu);
            v[t + 1] = (unsigned short)(q.x >> 16);
            v[t + 2] = (unsigned short)(q.y & 0xFFFFu);
        }
    }
    uint4* o = reinterpret_cast<uint4*>(col + pix * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 w;
        w.x = (uint32_t)v[8 * j + 0] | ((uint32_t)v[8 * j + 1] << 16);
        w.y = (uint32_t)v[8 * j + 2] | ((uint32_t)v[8 * j + 3] << 16);
        w.z = (uint32_t)v[8 * j + 4] | ((uint32_t)v[8 * j + 5] << 16);
        w.w = (uint32_t)v[8 * j + 6] | ((uint32_t)v[8 * j + 7] << 16);
        o[j] = w;
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

// nn.Upsample(scale_factor=2, mode='nearest') into a channel slice; thread = (output pixel, 8-channel group).
__global__ void upsample2_kernel(const __nv_bfloat16* __restrict__ in, int in_ctot, int in_coff,
                                 __nv_bfloat16* __restrict__ out, int out_ctot, int out_coff, int B, int H, int W,
                                 int C) {  // H, W = input extent
    const int groups = C / 8;
    const int Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * Ho * Wo * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    long long pix = idx / groups;
    const int w = (int)(pix % Wo);
    const int h = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((long long)Wo * Ho));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(
        in + (((long long)b * H + (h >> 1)) * W + (w >> 1)) * in_ctot + in_coff + g * 8));
    *reinterpret_cast<uint4*>(out + pix * out_ctot + out_coff + g * 8) = v;
}

// ------------------------------------------------------------------------------------------ model

static int make_divisible(double x, int d) { return (int)(ceil(x / d) * d); }

int Model::init(const char* variant, int nc_) {
    double depth, width;
    int maxc;
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

int Model::add_conv(const std::string& p, int cin, int cout, int k, bool bn) {
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
    if (cin == 3) {  // stem: packed bf16 [cout_pad][32], k = (kh*3+kw)*3 + c, zero for k >= 27 (1x1 conv over im2col)
        const int bn_ = conv_block_n(cout);
        cw.cout_pad = (cout + bn_ - 1) / bn_ * bn_;
        std::vector<__nv_bfloat16> hw((size_t)cw.cout_pad * 32, __float2bfloat16(0.f));
        for (int o = 0; o < cout; ++o)
            for (int i = 0; i < 3; ++i)
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < 3; ++kw)
                        hw[(size_t)o * 32 + (kh * 3 + kw) * 3 + i] =
                            __float2bfloat16((*w)[((o * 3 + i) * 3 + kh) * 3 + kw] * scale[o]);
        CY_CUDA_CHECK(cudaMalloc(&cw.w, hw.size() * sizeof(__nv_bfloat16)));
        CY_CUDA_CHECK(cudaMemcpy(cw.w, hw.data(), hw.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
        cw.cin = 32;
        cw.k = 1;
    } else {
        const int bn_ = conv_block_n(cout);
        cw.cout_pad = (cout + bn_ - 1) / bn_ * bn_;
        const size_t K = (size_t)k * k * cin;
        std::vector<__nv_bfloat16> hw((size_t)cw.cout_pad * K, __float2bfloat16(0.f));
        for (int o = 0; o < cout; ++o)
            for (int i = 0; i < cin; ++i)
                for (int kh = 0; kh < k; ++kh)
                    for (int kw = 0; kw < k; ++kw)
                        hw[(size_t)o * K + (size_t)(kh * k + kw) * cin + i] =
                            __float2bfloat16((*w)[(((size_t)o * cin + i) * k + kh) * k + kw] * scale[o]);
        CY_CUDA_CHECK(cudaMalloc(&cw.w, hw.size() * sizeof(__nv_bfloat16)));
        CY_CUDA_CHECK(cudaMemcpy(cw.w, hw.data(), hw.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
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

int Model::finalize() {
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
    {   // model.0 stem = im2col gather + 1x1 conv (K = 32) on the tensor pipe
        Buf col = pb.alloc(H2, W2, 32);
        PB(!col.p);
        Op op;
        op.type = Op::IM2COL; op.name = "model.0.im2col";
        op.out = col;
        pl->ops.push_back(op);
        PB(pb.conv("model.0", col, 0, x0, 0, 1));
        const double f27 = 2.0 * B * H2 * W2 * c1 * 27;   // count the real 27 taps, not the zero padding
        pl->flops += f27 - pl->ops.back().conv.flops;
        pl->ops.back().conv.flops = f27;
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
    for (int i = 0; i < 3; ++i) {
        Op op;
        op.type = Op::MAXPOOL; op.name = "model.9.m";
        op.in = sp; op.in_off = i * (c5 / 2); op.out = sp; op.out_off = (i + 1) * (c5 / 2); op.C = c5 / 2;
        pl->ops.push_back(op);
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

int Model::launch_op(const Op& op, const void* in, int B, int Sh, int Sw, cudaStream_t st) {
    switch (op.type) {
        case Op::IM2COL: {
            const long long npix = (long long)B * (Sh / 2) * (Sw / 2);
            stem_im2col_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(
                (const __nv_bfloat16*)in, (__nv_bfloat16*)op.out.p, B, Sh, Sw);
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
        case Op::UPSAMPLE: {
            const long long total = (long long)B * op.in.H * 2 * op.in.W * 2 * (op.C / 8);
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
        if (flops) flops[i] = pl->ops[i].type == Op::CONV ? pl->ops[i].conv.flops : 0.0;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (nops) *nops = n;
    return CY_OK;
}

}  // namespace cy
