#!/usr/bin/env python
"""BASELINE.json configs[4]: batched imgsize-1024 tiles, scoreThr 0.05, dense-candidate NMS stress (~10k boxes/tile)
through Detect decode -> threshold -> per-class NMS -> max_det -> un-letterbox (a16 + a17) and the per-tile IoU merge
(a18).  Head maps are synthetic (seeded normal logits, class bias chosen so ~10k of the 21 504 anchors pass 0.05).

Prints one JSON line: tiles/s of cy_postprocess and cy_merge_tile (CUDA events, mean of --reps launches after warm-up),
candidates per tile, algorithmic bytes per tile (head maps read once + candidates + bitmask), the
resulting GB/s, and — for scale only — torchvision.ops.nms (the library call the reference's NMS ends in) on the host
for the candidate boxes of --cpu-tiles tiles.  Parity of this configuration is asserted in
tests/test_nms_merge_gpu.py::test_config5_dense_batch_matches_oracle, not here.

usage: python tools/nms_stress.py [--tiles 256] [--imgsz 1024] [--conf 0.05] [--bias -0.5] [--reps 10] [--cpu-tiles 4]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rand_heads(B, Sh, Sw, nc, seed, cls_bias, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    heads = []
    for s in (8, 16, 32):
        h = torch.zeros(B, Sh // s, Sw // s, 80, device=device)
        h[..., :64] = torch.randn(B, Sh // s, Sw // s, 64, generator=g, device=device) * 2.0
        h[..., 64:64 + nc] = torch.randn(B, Sh // s, Sw // s, nc, generator=g, device=device) * 1.5 + cls_bias
        heads.append(h)
    return heads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tiles', type=int, default=256)
    ap.add_argument('--imgsz', type=int, default=1024)
    ap.add_argument('--tile', type=int, default=512)
    ap.add_argument('--conf', type=float, default=0.05)
    ap.add_argument('--iou', type=float, default=0.5)
    ap.add_argument('--bias', type=float, default=-4.7, help='class-logit bias: -4.7 -> ~10k candidates, -0.5 -> all 21.5k')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--cpu-tiles', type=int, default=4)
    a = ap.parse_args()
    import torch
    from caesar_yolo_b200 import ops
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    B, S, nc = a.tiles, a.imgsz, 5
    heads = rand_heads(B, S, S, nc, 42, a.bias, dev)
    A = ops.num_anchors(S, S)
    _, _, lb = ops.letterbox_shape(a.tile, a.tile, S)
    lbd = ops.letterbox_array([lb] * B, dev)
    ncand = sum(int((torch.sigmoid(h[..., 64:64 + nc]).amax(-1) > a.conf).sum()) for h in heads) / float(B)

    need = int(ops.lib.cy_postprocess_scratch_bytes(B, S, S, ops.MAX_DET))
    scratch = torch.empty((need,), dtype=torch.uint8, device=dev)
    dets = torch.zeros((B, ops.MAX_DET, 6), dtype=torch.float32, device=dev)
    nd = torch.zeros((B,), dtype=torch.int32, device=dev)

    def post():
        return ops.postprocess(heads, B, S, S, nc, a.conf, a.iou, lbd, dev, scratch=scratch, dets=dets, ndets=nd)

    def merge():
        return ops.merge_tile(dets, nd, 0.05, 0.3, 0.8)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    ms_post = timed(post)
    ms_merge = timed(merge)
    keep_idx, nkeep, status = merge()
    torch.cuda.synchronize()

    # host scale: torchvision.ops.nms on the class-offset candidate boxes of the first tiles (decode on the GPU)
    import torchvision
    pred = ops.decode_pred([h[:a.cpu_tiles].contiguous() for h in heads], a.cpu_tiles, S, S, nc, dev).cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    t_cpu = 0.0
    for b in range(a.cpu_tiles):
        p = pred[b].T                                  # [A, 4 + nc]
        conf, cls = p[:, 4:].max(1)
        m = conf > a.conf
        xywh, conf, cls = p[m, :4], conf[m], cls[m].float()
        xyxy = torch.cat([xywh[:, :2] - xywh[:, 2:] / 2, xywh[:, :2] + xywh[:, 2:] / 2], 1)
        t0 = time.time()
        torchvision.ops.nms(xyxy + cls[:, None] * 7680.0, conf, a.iou)
        t_cpu += time.time() - t0
    cpu_nms_s = t_cpu / a.cpu_tiles
    words = (ncand + 63) // 64
    alg = 80 * A * 4 + ncand * 24 + 2 * ncand * words * 8 + ops.MAX_DET * 24
    line = {"workload": "configs[4]: %d tiles, imgsz %d (A = %d anchors), scoreThr %.2f, iou %.2f, nc 5" % (B, S, A, a.conf, a.iou),
            "candidates_per_tile": ncand, "dets_per_tile": float(nd.float().mean()), "kept_after_merge_per_tile": float(nkeep.float().mean()),
            "postprocess_ms": ms_post, "postprocess_tiles_per_s": B / (ms_post * 1e-3),
            "merge_tile_ms": ms_merge, "merge_tile_tiles_per_s": B / (ms_merge * 1e-3),
            "hbm_bytes_per_tile": 80 * A * 4 + ops.MAX_DET * 24, "hbm_GBps": (80 * A * 4 + ops.MAX_DET * 24) * B / (ms_post * 1e-3) / 1e9,
            "survey_formula_bytes_per_tile": alg, "survey_formula_GBps": alg * B / (ms_post * 1e-3) / 1e9,
            "note": "the K x K/64 suppression bitmask of the SURVEY formula never leaves shared memory here (512-box chunks), so hbm_* is the traffic the kernels really have to move",
            "cpu": {"nms_tiles_per_s": 1.0 / cpu_nms_s, "cores": os.cpu_count(),
                    "sample": "torchvision.ops.nms alone on the candidates of %d tiles" % a.cpu_tiles}}
    print(json.dumps(line))
    return 0


if __name__ == '__main__':
    sys.exit(main())
