#!/usr/bin/env python
"""Per-layer device timing of one batched forward (CUDA events around every op, Model::profile).

usage: python tools/profile_layers.py [--variant l] [--batch 32] [--imgsz 640] [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='l')
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--imgsz', type=int, default=640)
    ap.add_argument('--json', default=None)
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    import torch
    from caesar_yolo_b200 import ops, weights as W
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    w = W.make_random_weights(a.variant, 5, seed=0, cls_bias=-16.0)
    m = ops.DeviceModel(w)
    x = torch.rand((a.batch, a.imgsz, a.imgsz, 4), device=dev).to(m.dtype)
    best = None
    for _ in range(a.reps):
        prof = m.profile(x)
        if best is None:
            best = prof
        else:
            best = [(n, min(ms, b[1]), fl) for (n, ms, fl), b in zip(prof, best)]
    tot = sum(p[1] for p in best)
    totf = sum(p[2] for p in best)
    print("%-28s %9s %9s %7s" % ("op", "ms", "TFLOP/s", "share"))
    for n, ms, fl in best:
        print("%-28s %9.4f %9.1f %6.1f%%" % (n, ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0, 100 * ms / tot))
    print("total %.3f ms for batch %d = %.4f ms/tile, %.1f TFLOP/s overall" % (tot, a.batch, tot / a.batch, totf / (tot * 1e-3) / 1e12))
    if a.json:
        json.dump([dict(op=n, ms=ms, flops=fl) for n, ms, fl in best], open(a.json, 'w'))


if __name__ == '__main__':
    main()
