// YOLOv8 model runtime types (see model.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "conv.cuh"

namespace cy {

static constexpr int kHeadC = 80;  // per-anchor head record: 64 DFL logits + up to 16 class logits (fp32)

struct Buf {
    void* p = nullptr;
    int H = 0, W = 0, C = 0;
};

struct ConvW {
    void* w = nullptr;   // packed bf16 [cout_pad, k*k*cin]  (stem: mma.sync B fragments)
    float* b = nullptr;  // fp32 [cout_pad]
    int cin = 0, cout = 0, cout_pad = 0, k = 0;
};

struct Op {
    enum Type { STEM, CONV, MAXPOOL, SPPF_POOL, UPSAMPLE, DWCONV, ATTN } type;
    std::string name;
    ConvPlan conv;
    Buf in, out;
    int in_off = 0, out_off = 0, C = 0;
    // DWCONV: depthwise 3x3 (+ bias, optional SiLU, optional residual); logical channel c reads input channel
    // in_off + (c / gs) * gst + c % gs (gs = C, gst = 0 for a plain slice; the attention `pe` conv gathers v out of qkv)
    const float* dw_w = nullptr;   // fp32 [9][C], BN folded, bf16-rounded values
    const float* dw_b = nullptr;   // fp32 [C]
    int act = 0, gs = 0, gst = 0;
    Buf res;
    int res_off = 0;
    // ATTN: in = qkv buffer [B,H,W,nh*128] (per head 32 q | 32 k | 64 v), out = [B,H,W,nh*64] slice
    int nh = 0;
};

struct Plan {
    int B = 0, Sh = 0, Sw = 0;
    std::vector<Op> ops;
    std::vector<void*> allocs;
    Buf head[3];
    size_t bytes = 0;
    double flops = 0;
    ~Plan();
};

struct Model {
    char variant = 'n';
    int family = 8;   // 8: yolov8 (yolov8.yaml), 11: yolo11 (yolo11.yaml)
    int nc = 5;
    int f16 = 0;      // storage format of weights / activations: 0 bf16, 1 fp16 (set before finalize)
    // yolo11 widths / repeats (init11)
    int w64 = 0, w128 = 0, w256 = 0, w512 = 0, w1024 = 0, n11 = 1;
    bool c3k11 = false;
    int c1, c2, c3, c4, c5, n2, n4, n6, n8, nh, cb, cc;
    bool finalized = false;
    long long nparams = 0;
    std::map<std::string, std::vector<float>> raw;
    std::map<std::string, ConvW> convs;
    std::map<std::tuple<int, int, int>, Plan*> plans;

    int init(const char* variant, int nc);
    int set_tensor(const char* name, const float* data, long long numel);
    int finalize();
    int get_plan(int B, int Sh, int Sw, Plan** out);
    int build_plan(int B, int Sh, int Sw, Plan** out);
    int build_plan11(int B, int Sh, int Sw, Plan** out);
    int finalize11();
    int forward(const void* in, int B, int Sh, int Sw, cudaStream_t st, Plan** plan_out);
    int launch_op(const Op& op, const void* in, int B, int Sh, int Sw, cudaStream_t st);
    int profile(const void* in, int B, int Sh, int Sw, int cap, const char** names, float* ms, double* flops, int* nops,
                cudaStream_t st);
    ~Model();

   private:
    const std::vector<float>* get(const std::string& k) const;
    int add_conv(const std::string& p, int cin, int cout, int k, bool bn, int cin_pad = 0);
    int add_dwconv(const std::string& p, int c);
    int add_c3k2(const std::string& p, int cin, int cout, bool c3k, double e);
    int add_c2f(const std::string& p, int cin, int cout, int n);
};

}  // namespace cy
