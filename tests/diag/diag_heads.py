"""Diagnostic: head-map error of the tcgen05 forward vs the oracle (bf16-emulated and fp32) on a real preprocessed
tile, plus per-stage device timing of one batch and host-side timing of run_image."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from caesar_yolo_b200 import ops, pipeline, synth, weights as W  # noqa: E402
from oracle import preprocessing as opp, yolo as oy  # noqa: E402

dev = torch.device('cuda:0')
flags = dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True, normalize_minmax=True,
             nchannels=3, norm_max=255.)
variant = sys.argv[1] if len(sys.argv) > 1 else 'n'
w = W.make_random_weights(variant, 5, seed=0, cls_bias=-16.0)
tile = synth.make_mosaic(512, 512, seed=3, nan_border_frac=0.0)
cube = np.stack([tile.astype(np.float64)] * 3, -1)
img = opp.DataPreprocessor(opp.build_stages(**flags))(cube)
x = oy.preprocess(img, 640)                       # [1,3,640,640] fp32
xin = torch.zeros(1, 640, 640, 4, dtype=torch.bfloat16)
xin[..., :3] = x[0].permute(1, 2, 0).to(torch.bfloat16)
dm = ops.DeviceModel(w)
heads = [h.cpu() for h in dm.forward_tensors(xin.to(dev))]
with torch.no_grad():
    he = oy.OracleYolo(w, emulate_bf16=True).forward_heads(x)
    hf = oy.OracleYolo(w, emulate_bf16=False).forward_heads(x)
for l in range(3):
    g = heads[l][..., :69].permute(0, 3, 1, 2)
    for name, sl in (('box', slice(0, 64)), ('cls', slice(64, 69))):
        a, e, f = g[:, sl], he[l][:, sl], hf[l][:, sl]
        rms = lambda t: float(t.pow(2).mean().sqrt())
        print("level %d %s: rms(f32)=%.4f  rms(ours-emu)=%.5f  rms(emu-f32)=%.5f  rms(ours-f32)=%.5f  max|ours-emu|=%.4f"
              % (l, name, rms(f), rms(a - e), rms(e - f), rms(a - f), float((a - e).abs().max())))

# ---- stage timing of one batch (device events) ----
B = 32
mos = synth.make_mosaic(512, 512 * B, seed=5, nan_border_frac=0.0)
raw = torch.from_numpy(mos.astype('>f4').view(np.int32).copy()).to(dev)
cfg = pipeline.make_pp_config(**flags)
x0 = (torch.arange(B, dtype=torch.int32) * 512).to(dev)
y0 = torch.zeros(B, dtype=torch.int32, device=dev)
wl = W.make_random_weights('l', 5, seed=0, cls_bias=-24.0)
dml = ops.DeviceModel(wl)
_, _, lb = ops.letterbox_shape(512, 512, 640)
lbd = ops.letterbox_array([lb] * B, dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(3):
    t = [ev()]
    chain, model_in, _, st = ops.preprocess(cfg, raw, 512 * B, True, x0, y0, 512, 512, 640)
    t.append(ev())
    hd = dml.forward(model_in)
    t.append(ev())
    dets, nd = ops.postprocess(hd, B, 640, 640, 5, 0.5, 0.5, lbd, dev)
    t.append(ev())
    keep, nk, ms = ops.merge_tile(dets, nd, 0.5, 0.3, 0.8, pre_status=st)
    t.append(ev())
    torch.cuda.synchronize()
    names = ['preprocess(sort+chain+resize)', 'forward v8l', 'decode+nms', 'merge_tile']
    if it == 2:
        for i, nme in enumerate(names):
            print("  %-32s %8.3f ms per batch of %d  (%.1f us/tile)" % (nme, t[i].elapsed_time(t[i + 1]), B,
                                                                       t[i].elapsed_time(t[i + 1]) * 1e3 / B))
        print("  ndets", nd[:8].tolist(), "status", st[:8].tolist())

# ---- host-side timing of run_image ----
n = 4096
img = synth.make_mosaic(n, n, seed=1234, nan_border_frac=0.0)
host = torch.empty((n, n), dtype=torch.int32, pin_memory=True)
host.numpy().view(np.uint32)[:] = img.view(np.uint32).byteswap()
tiles = ops.generate_tiles(0, n - 1, 0, n - 1, 512, 512, 1.0, 1.0)
eng = pipeline.Engine(dml, cfg, imgsz=640, score_thr=0.5, device=dev, batch_tiles=32)
import cProfile
import pstats
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    if it == 2:
        pr = cProfile.Profile()
        pr.enable()
    src, nrec = pipeline.run_image(eng, host, True, tiles)
    torch.cuda.synchronize()
    if it == 2:
        pr.disable()
    print("run_image 4096^2: %.1f ms, %d sources" % ((time.time() - t0) * 1e3, len(src)))
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)

# ---- single-conv precision: tcgen05 fp32-accumulate vs fp64 reference on identical bf16 operands ----
g = torch.Generator().manual_seed(1)
for (cin, cout, k) in ((256, 256, 3), (64, 64, 3), (512, 512, 1)):
    xx = torch.randn(2, 40, 40, cin, generator=g).to(torch.bfloat16)
    ww = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    bb = torch.zeros(cout)
    wp, bp = ops.pack_conv_weight(ww, bb, dev)
    out = torch.zeros(2, 40, 40, cout, dtype=torch.float32, device=dev)
    ops.conv2d_nhwc(xx.to(dev), 0, cin, wp, bp, cout, k, 1, out, 0, act=False)
    ref64 = torch.nn.functional.conv2d(xx.double().permute(0, 3, 1, 2), ww.double(), None, padding=k // 2).permute(0, 2, 3, 1)
    ref32 = torch.nn.functional.conv2d(xx.float().permute(0, 3, 1, 2), ww.float(), None, padding=k // 2).permute(0, 2, 3, 1)
    rms = float(ref64.pow(2).mean().sqrt())
    print("conv cin=%d k=%d: rel rms err tcgen05 vs fp64 %.3e | torch-CPU-fp32 vs fp64 %.3e" % (
        cin, k, float((out.cpu().double() - ref64).pow(2).mean().sqrt()) / rms,
        float((ref32.double() - ref64).pow(2).mean().sqrt()) / rms))
