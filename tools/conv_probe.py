#!/usr/bin/env python
"""Times single conv layers of the YOLOv8l inventory through the C ABI and dumps the kernel's per-CTA timeline
(clock64 stamps of the MMA issuer and of epilogue warp 0) to see what bounds each shape.

usage: python tools/conv_probe.py [--timeline] [--shapes name,...]
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {
    # name: (B, H, W, cin, cout, k, s, res)
    'stem_1x1_320_32_64': (32, 320, 320, 32, 64, 1, 1, False),
    'm2.cv1_1x1_160_128_128': (32, 160, 160, 128, 128, 1, 1, False),
    'm2.m_3x3_160_64_64': (32, 160, 160, 64, 64, 3, 1, False),
    'm2.m_3x3_160_64_64_res': (32, 160, 160, 64, 64, 3, 1, True),
    'm4.m_3x3_80_128_128': (32, 80, 80, 128, 128, 3, 1, False),
    'm4.cv2_1x1_80_1024_256': (32, 80, 80, 1024, 256, 1, 1, False),
    'm6.m_3x3_40_256_256': (32, 40, 40, 256, 256, 3, 1, False),
    'm6.cv2_1x1_40_2048_512': (32, 40, 40, 2048, 512, 1, 1, False),
    'm8.m_3x3_20_256_256': (32, 20, 20, 256, 256, 3, 1, False),
    'm3_3x3s2_160_128_256': (32, 160, 160, 128, 256, 3, 2, False),
    'cv3.0.0_3x3_80_256_256': (32, 80, 80, 256, 256, 3, 1, False),
    'cv2.0.0_3x3_80_256_64': (32, 80, 80, 256, 64, 3, 1, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--timeline', action='store_true')
    ap.add_argument('--shapes', default='')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--epi', action='store_true', help='epilogue phase sums instead of the MMA timeline')
    a = ap.parse_args()
    import torch
    from caesar_yolo_b200 import ops
    from caesar_yolo_b200._capi import lib, check
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    names = [s for s in a.shapes.split(',') if s] or list(SHAPES)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in names:
        B, H, W, cin, cout, k, s, res = SHAPES[name]
        g = torch.Generator(device='cpu').manual_seed(0)
        x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, generator=g) * 0.1
        wp, bp = ops.pack_conv_weight(w, b, dev)
        Ho, Wo = H // s, W // s
        out = torch.zeros(B, Ho, Wo, cout, dtype=torch.bfloat16, device=dev)
        r = torch.randn(B, Ho, Wo, cout, generator=g).to(torch.bfloat16).to(dev) if res else None
        info = (ctypes.c_int * 8)()
        check(lib.cy_conv_plan_info(B, H, W, cin, cout, k, s, info))
        for _ in range(3):
            ops.conv2d_nhwc(x, 0, cin, wp, bp, cout, k, s, out, 0, act=True, res=r)
        ts = []
        for _ in range(a.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d_nhwc(x, 0, cin, wp, bp, cout, k, s, out, 0, act=True, res=r)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = 2.0 * B * Ho * Wo * cout * cin * k * k
        print("%-28s %8.4f ms %7.1f TFLOP/s  mode=%d halves=%d units=%d N=%d a_st=%d b_st=%d bufs=%d grid=%d"
              % ((name, ms, fl / ms / 1e9) + tuple(info)))
        if a.timeline:
            U = 64
            grid = info[7]
            buf = torch.zeros(grid * U * 8, dtype=torch.int64, device=dev)
            check(lib.cy_conv_set_debug(ctypes.c_void_p(buf.data_ptr()), U | ((1 << 30) if a.epi else 0)))
            ops.conv2d_nhwc(x, 0, cin, wp, bp, cout, k, s, out, 0, act=True, res=r)
            torch.cuda.synchronize()
            check(lib.cy_conv_set_debug(ctypes.c_void_p(0), 0))
            t = buf.cpu().numpy().reshape(grid, U, 8)
            if a.epi:
                nu0 = min((info[2] + (grid // 2 if (info[0] % 100) >= 10 else grid) - 1) // (grid // 2 if (info[0] % 100) >= 10 else grid), U)
                d = t[0, 1:nu0 - 1].astype(float)
                print("   CTA0 epilogue warp 0, clk per unit: wait staging buffer %.0f | TMEM load %.0f | math + st.shared %.0f | fence + TMA store %.0f | acc wait %.0f | total %.0f"
                      % (d[:, 0].mean(), d[:, 1].mean(), d[:, 2].mean(), d[:, 3].mean(), (d[:, 6] - d[:, 5]).mean(), (d[:, 7] - d[:, 6]).mean()))
                continue
            if (info[0] % 100) >= 10:            # CTA pairs: one timeline per cluster (worker = blockIdx.x >> 1)
                grid //= 2
            nu = [(info[2] - c + grid - 1) // grid for c in range(grid)]
            c = 0
            t0 = t[c, 0, 0]
            print("   CTA0: unit | mma wait-acc  start  end(issued) | waitB  waitA | epi wait-from  start  end   (clk rel. to start)")
            for i in range(min(nu[c], 8)):
                d = t[c, i]
                print("        %3d | %8d %8d %8d | %6d %6d | %8d %8d %8d" % (i, d[0] - t0, d[1] - t0, d[2] - t0, d[3], d[4], d[5] - t0, d[6] - t0, d[7] - t0))
            import numpy as np
            mm, ep, wb, wa, tot = [], [], [], [], []
            for c in range(grid):
                n = min(nu[c], U)
                if n < 1:
                    continue
                d = t[c, :n]
                mm.append((d[:, 2] - d[:, 1]).mean())
                ep.append((d[:, 7] - d[:, 6]).mean())
                wb.append(d[:, 3].mean())
                wa.append(d[:, 4].mean())
                tot.append((d[n - 1, 7] - d[0, 0]) / n)
            print("   mean over CTAs per unit: mma-issue span %.0f clk (waitB %.0f, waitA %.0f), epilogue %.0f clk, wall/unit %.0f clk"
                  % (np.mean(mm), np.mean(wb), np.mean(wa), np.mean(ep), np.mean(tot)))


if __name__ == '__main__':
    main()
