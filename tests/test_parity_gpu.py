"""Parity of the BENCHMARKED configuration and of everything downstream of the head maps.

1. YOLOv8l at 640 x 640 with a batch of 148 tiles (the planner's real choices for the bench shape: CTA pairs, 256-wide
   tiles, stride-2 tap reuse): the head maps of sampled tiles against the CPU oracle in the same storage precision
   (kernel check) and in fp32 (quantisation check), rms bound per pyramid level.
2. Hybrid end-to-end: the head maps THIS path computes for a real tiled mosaic go through the ORACLE's decode -> NMS ->
   scale_boxes -> process_detections -> make_json_results -> find_sources_at_edge -> merge_edge_sources; the catalog
   must be identical to the one this path produces (coordinates, scores, classes, flags, order).  Together with (1)
   this splits the end-to-end criterion into "head maps within the storage-precision distance" and "everything after
   the head maps bit-exact".
"""
import numpy as np
import pytest
import torch

import helpers
from oracle import yolo as oy

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PP = dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True, normalize_minmax=True,
          nchannels=3, norm_max=255.)


def _preprocessed_tiles(mosaic, tiles, pp_cfg, imgsz=640):
    from caesar_yolo_b200 import ops
    dev = torch.device(DEV)
    img = torch.from_numpy(np.ascontiguousarray(mosaic, dtype=np.float32)).to(dev)
    x0 = torch.from_numpy(tiles['xmin'].astype(np.int32)).to(dev)
    y0 = torch.from_numpy(tiles['ymin'].astype(np.int32)).to(dev)
    Ty = int(tiles['ymax'][0] - tiles['ymin'][0])
    Tx = int(tiles['xmax'][0] - tiles['xmin'][0])
    _, model_in, _, status = ops.preprocess(pp_cfg, img, mosaic.shape[1], False, x0, y0, Ty, Tx, imgsz)
    return model_in, status


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_forward_v8l_bench_shape_b148(precision):
    """Bench plan (yolov8l, 640^2, large batch): sampled tiles vs the same-precision and fp32 oracles."""
    from caesar_yolo_b200 import ops, pipeline, synth, weights as W
    w = W.make_random_weights('l', 5, seed=0, cls_bias=-24.0)
    dm = ops.DeviceModel(w, precision=precision)
    mosaic = synth.make_mosaic(2048, 2048, seed=77, nan_border_frac=0.0)
    tiles = ops.generate_tiles(0, 2047, 0, 2047, 512, 512, 1.0, 1.0)
    m16, status = _preprocessed_tiles(mosaic, tiles, pipeline.make_pp_config(out_f16=(precision == 'fp16'), **PP))
    assert int((status != 0).sum()) == 0
    B = 148
    reps = [torch.roll(m16[k % 16], shifts=(17 * (k // 16), 29 * (k // 16)), dims=(0, 1)) for k in range(B)]
    x = torch.stack(reps).contiguous()
    assert x.dtype == ops.storage_dtype(precision)
    info = dm.plan_summary(B, 640, 640)
    assert info['pair_launches'] > 0 and info['mode3_launches'] > 0, info     # the plan really is the bench plan
    heads = dm.forward_tensors(x)
    torch.cuda.synchronize()
    sample = [0, 61, 147]
    xs = x[sample, :, :, :3].float().cpu().permute(0, 3, 1, 2).contiguous()   # the exact model input, NCHW fp32
    with torch.no_grad():
        he = oy.OracleYolo(w, emulate_bf16=precision).forward_heads(xs)
        hf = oy.OracleYolo(w, emulate_bf16=False).forward_heads(xs)
    # storage-precision noise relative to the map rms: bf16 keeps 8 significand bits, fp16 11
    bound_emu = {"bf16": 0.02, "fp16": 0.004}[precision]
    bound_f32 = {"bf16": 0.03, "fp16": 0.006}[precision]
    for l in range(3):
        got = heads[l][sample].cpu()[..., :69].permute(0, 3, 1, 2)
        rms = hf[l].pow(2).mean().sqrt().item()
        e_emu = (got - he[l]).pow(2).mean().sqrt().item() / rms
        e_f32 = (got - hf[l]).pow(2).mean().sqrt().item() / rms
        q = (he[l] - hf[l]).pow(2).mean().sqrt().item() / rms
        print("%s level %d: rms(ours-emu)/rms %.5f  rms(ours-fp32)/rms %.5f  rms(emu-fp32)/rms %.5f"
              % (precision, l, e_emu, e_f32, q))
        assert e_emu < bound_emu, (l, e_emu, q)
        assert e_f32 < max(bound_f32, 1.5 * q), (l, e_f32, q)


def _oracle_tile_dets(heads, b, nc, conf, iou, lb_shape, tile_shape):
    """Oracle decode -> non_max_suppression -> scale_boxes on OUR head maps of tile b -> [N,6] float32."""
    net = oy.OracleYolo.__new__(oy.OracleYolo)
    net.nc = nc
    hb = [h[b:b + 1, :, :, :64 + nc].permute(0, 3, 1, 2).contiguous() for h in heads]
    pred = net.decode(hb)[0]
    det = oy.nms_single(pred, conf, iou)
    if det.shape[0]:
        det[:, :4] = oy.scale_boxes(lb_shape, det[:, :4], tile_shape)
    return det.numpy().astype(np.float32)


@pytest.mark.parametrize("step,variant,bias", [(1.0, 'n', -12.0), (0.5, 'n', -12.0), (1.0, 'l', -24.0)])
def test_hybrid_our_heads_through_oracle_catalog_identical(step, variant, bias):
    from caesar_yolo_b200 import catalog, ops, pipeline, synth, weights as W
    conf, iou, soft, hard = 0.5, 0.5, 0.3, 0.8
    ny, nx = (1536, 2048) if step == 1.0 else (1024, 1536)
    mosaic = synth.make_mosaic(ny, nx, seed=31, nan_border_frac=0.0)
    mosaic[-90:, :] = np.nan
    mosaic[:, -40:] = np.nan
    tiles = ops.generate_tiles(0, nx - 1, 0, ny - 1, 512, 512, step, step)
    w = W.make_random_weights(variant, 5, seed=0, cls_bias=bias)
    pp = pipeline.make_pp_config(**PP)
    eng = pipeline.Engine(w, pp, imgsz=640, score_thr=conf, iou_thr=iou, thr_soft=soft, thr_hard=hard, device=DEV)
    raw = torch.from_numpy(np.nan_to_num(mosaic, nan=np.nan).astype('>f4').view(np.int32).copy())
    src, nrec = pipeline.run_image(eng, raw, True, tiles)
    ours = catalog.sources_to_dicts(src, eng.names)

    # our head maps, tile shape by tile shape (same grouping / batch sizes as the engine used)
    per_tile = [np.zeros((0, 6), np.float32) for _ in range(len(tiles))]
    w_ = tiles['xmax'] - tiles['xmin']
    h_ = tiles['ymax'] - tiles['ymin']
    for (Ty, Tx) in sorted(set(zip(h_.tolist(), w_.tolist())), reverse=True):
        ids = np.nonzero((h_ == Ty) & (w_ == Tx))[0]
        model_in, status = _preprocessed_tiles(mosaic, tiles[ids], eng.pp_cfg)
        heads = [h.cpu() for h in eng.model.forward_tensors(model_in)]
        st = status.cpu().numpy()
        Sh, Sw, _ = ops.letterbox_shape(Ty, Tx, 640)
        for k, tid in enumerate(ids):
            if st[k] != 0:
                continue                                   # predict() returned -1 for this tile: no detections
            d = _oracle_tile_dets(heads, k, 5, conf, iou, (Sh, Sw), (Ty, Tx))
            if len(d):
                keep, _ = helpers.oracle_merge_tile(d, conf, soft, hard)
                per_tile[tid] = d[keep]
    want, _ = helpers.oracle_catalog([tuple(int(v) for v in t) for t in tiles], per_tile, conf, soft, hard)
    assert len(want) >= 15
    assert len(ours) == len(want), (len(ours), len(want))
    for a, b in zip(ours, want):
        assert a['name'] == b['name']
        assert (a['x1'], a['y1'], a['x2'], a['y2']) == (b['x1'], b['y1'], b['x2'], b['y2']), (a, b)
        assert a['class_id'] == b['class_id'] and a['class_name'] == b['class_name']
        # the class sigmoid of the decode kernel and torch.sigmoid agree to 1 ulp, not bit for bit
        assert abs(np.float32(a['score']) - np.float32(b['score'])) <= 2e-7 * max(1.0, abs(b['score'])), (a, b)
        assert bool(a['edge']) == bool(b['edge']) and bool(a['merged']) == bool(b['merged'])


def test_exchange_slots_roundtrip_two_ranks():
    """Engine.pack_send / unpack_recv (the device side of the all-gather exchange) with two emulated ranks on one GPU:
    the catalog assembled from the two gathered slots equals the single-rank catalog, also when a slot overflows."""
    from caesar_yolo_b200 import ops, pipeline, synth, weights as W
    mosaic = synth.make_mosaic(1024, 1536, seed=5, nan_border_frac=0.0)
    tiles = ops.generate_tiles(0, 1535, 0, 1023, 512, 512, 0.5, 0.5)
    w = W.make_random_weights('n', 5, seed=0, cls_bias=-12.0)
    eng = pipeline.Engine(w, pipeline.make_pp_config(**PP), imgsz=640, score_thr=0.5, device=DEV)
    dev_img = torch.from_numpy(mosaic).to(DEV)
    eng.begin(tiles)
    eng.process_tiles(dev_img, 1536, False, 0, 0, np.arange(len(tiles), dtype=np.int32))
    ref, n_ref = eng.exchange_and_merge(1)
    parts = pipeline.split_tile_rows(tiles, 2)
    for cap in (None, 7):                                   # None: default capacity; 7: forces the overflow redo path
        slots, counts = [], []
        for r in range(2):
            eng.begin(tiles)
            eng.process_tiles(dev_img, 1536, False, 0, 0, np.arange(parts[r][0], parts[r][1], dtype=np.int32))
            packed = eng._compact_local()
            c = eng._exchange_cap(2) if cap is None else cap
            slots.append(eng.pack_send(packed, c).clone())
            counts.append(int(eng._buf['total'][0].item()))
        recv = torch.cat(slots)
        allp, n, cmax = eng.unpack_recv(recv, 2, c)
        assert cmax == max(counts)
        if cap is None:
            assert n == n_ref == sum(counts)
            got = eng.global_merge(allp[:n * 32], n)
            assert got.tobytes() == ref.tobytes()
        else:
            assert cmax > c and n == min(counts[0], c) + min(counts[1], c)   # overflow is detected, nothing out of range
