// C-ABI entry points for the convolution layer primitive (see include/caesar_b200.h).
#include "../../include/caesar_b200.h"
#include "conv.cuh"
#include "common.h"
#include <string.h>

namespace cy {
int conv_block_n(int cout);
int stem_conv_run(const void* in, int B, int H, int W, const float* w_host, const float* bias_host, int cout, int act,
                  void* out, cudaStream_t st);
}

extern "C" int cy_conv_block_n(int cout) { return cy::conv_block_n(cout); }

extern "C" int cy_conv_set_debug(void* dev_buf, int units_per_cta) {
    cy::conv_set_debug((unsigned long long*)dev_buf, units_per_cta);
    return CY_OK;
}

extern "C" int cy_conv_plan_info(int B, int Hin, int Win, int cin, int cout, int ksize, int stride, int* info8) {
    cy::ConvDesc d;
    memset(&d, 0, sizeof(d));
    alignas(128) static char dummy[128];
    d.in = (const __nv_bfloat16*)dummy; d.in_ctot = cin; d.cin = cin; d.B = B; d.Hin = Hin; d.Win = Win;
    d.ksize = ksize; d.stride = stride; d.w = (const __nv_bfloat16*)dummy;
    const int bn = cy::conv_block_n(cout);
    d.cout = cout; d.cout_pad = (cout + bn - 1) / bn * bn; d.out = dummy; d.out_ctot = (cout + 7) / 8 * 8;
    cy::ConvPlan plan;
    char err[256];
    if (cy::conv_make_plan(d, &plan, err, sizeof(err)) != 0) return cy::set_error(CY_ERR_INVALID, "%s", err);
    info8[0] = plan.kp.mode + 10 * plan.kp.pair; info8[1] = plan.kp.halves; info8[2] = plan.kp.n_units; info8[3] = plan.kp.block_n;
    info8[4] = plan.kp.a_stages; info8[5] = plan.kp.b_stages; info8[6] = plan.kp.acc_bufs; info8[7] = (int)plan.grid.x;
    return CY_OK;
}

extern "C" int cy_conv2d_nhwc(const void* in, int B, int Hin, int Win, int in_ctot, int in_coff, int cin,
                              const void* w, const float* bias, int cout, int cout_pad, int ksize, int stride,
                              void* out, int out_ctot, int out_coff, int out_f32, const void* res, int res_ctot,
                              int res_coff, int act, uintptr_t stream) {
    cy::ConvDesc d;
    d.in = (const __nv_bfloat16*)in; d.in_ctot = in_ctot; d.in_coff = in_coff; d.cin = cin;
    d.B = B; d.Hin = Hin; d.Win = Win; d.ksize = ksize; d.stride = stride;
    d.w = (const __nv_bfloat16*)w; d.cout_pad = cout_pad; d.bias = bias; d.cout = cout;
    d.out = out; d.out_ctot = out_ctot; d.out_coff = out_coff; d.out_f32 = out_f32;
    d.res = (const __nv_bfloat16*)res; d.res_ctot = res_ctot; d.res_coff = res_coff; d.act = act;
    d.f16 = 0;
    cy::ConvPlan plan;
    char err[256];
    if (cy::conv_make_plan(d, &plan, err, sizeof(err)) != 0) return cy::set_error(CY_ERR_INVALID, "%s", err);
    int r = cy::conv_launch(plan, (cudaStream_t)stream);
    if (r != 0) return cy::set_error(CY_ERR_CUDA, "conv launch failed: %s", cudaGetErrorString((cudaError_t)r));
    return CY_OK;
}

extern "C" int cy_stem_conv_nhwc4(const void* in, int B, int H, int W, const float* w_host, const float* bias_host,
                                  int cout, int act, void* out, uintptr_t stream) {
    if (!in || !w_host || !bias_host || !out || B <= 0) return cy::set_error(CY_ERR_INVALID, "cy_stem_conv_nhwc4: null argument");
    return cy::stem_conv_run(in, B, H, W, w_host, bias_host, cout, act, out, (cudaStream_t)stream);
}
