"""Thin Python handles on the C-ABI stage kernels (torch used only for device memory + streams)."""
import torch

from . import _capi
from ._capi import c_int, c_float, check, cur_stream, lib, ptr


def conv_block_n(cout):
    return int(lib.cy_conv_block_n(c_int(cout)))


def pack_conv_weight(w_oihw, bias, device):
    """[Cout,Cin,kh,kw] float/bf16 (BN already folded) -> device bf16 [cout_pad, kh*kw*Cin] + fp32 bias[cout_pad]."""
    cout, cin, kh, kw = w_oihw.shape
    bn = conv_block_n(cout)
    cout_pad = (cout + bn - 1) // bn * bn
    wp = torch.zeros(cout_pad, kh * kw * cin, dtype=torch.bfloat16)
    wp[:cout] = w_oihw.permute(0, 2, 3, 1).reshape(cout, -1).to(torch.bfloat16)
    bp = torch.zeros(cout_pad, dtype=torch.float32)
    if bias is not None:
        bp[:cout] = bias.float()
    return wp.to(device), bp.to(device)


def conv2d_nhwc(x, in_coff, cin, wp, bp, cout, k, s, out, out_coff, act=True, res=None, res_coff=0):
    B, H, W, in_ctot = x.shape
    check(lib.cy_conv2d_nhwc(ptr(x), c_int(B), c_int(H), c_int(W), c_int(in_ctot), c_int(in_coff), c_int(cin),
                             ptr(wp), ptr(bp), c_int(cout), c_int(wp.shape[0]), c_int(k), c_int(s), ptr(out),
                             c_int(out.shape[-1]), c_int(out_coff), c_int(1 if out.dtype == torch.float32 else 0),
                             ptr(res), c_int(res.shape[-1] if res is not None else 0), c_int(res_coff),
                             c_int(1 if act else 0), cur_stream()))
