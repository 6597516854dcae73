#!/usr/bin/env python
"""bench.py — headline benchmark of the caesar-yolo hot path on B200 (contract: see the task prompt).

Workload (BASELINE.json configs[2]): synthetic 16384 x 16384 float32 radio mosaic (FITS payload byte order), 512 x 512
tiles at step 1.0 -> 1024 tiles, random-init YOLOv8l (nc=5, seeded), imgsz 640, full preprocessing chain
(subtract_bkg, clip_data, zscale_stretch, chan3_preproc, normalize_minmax), FITS payload -> merged catalog.
One "step" = one full pass mosaic -> catalog.  N GPUs split the SAME mosaic into contiguous bands of tile rows
(strong scaling), exchange 32-byte detection records with one NCCL all-gather and merge on every rank.

  value : tiles/s with the mosaic band already resident in HBM (device timed, max over ranks)
  e2e   : the same through the host-facing call with the mosaic in pinned HOST memory: H2D of the payload and D2H of
          the catalog inside the timed region
  --impl reference : the CPU oracle (restatement of the reference's --devices=cpu path) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ['NCCL_DEBUG'] = 'WARN'   # keep warnings; the banner it prints is kept off stdout below

# The contract is ONE JSON line on stdout.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there at WARN/INFO), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
# private duplicate of the original stdout.
_JSON_OUT = None


def claim_stdout():
    """Called once by main(): keeps a private handle on the real stdout and sends everything else to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

PP_FLAGS = dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True, normalize_minmax=True,
                nchannels=3, norm_max=255.)
SCORE_THR, IOU_THR, SOFT, HARD = 0.5, 0.5, 0.3, 0.8
CLS_BIAS = {'n': -16.0, 'l': -24.0, '11n': -20.0, '11l': -24.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mosaic', type=int, default=16384)
    ap.add_argument('--tile', type=int, default=512)
    ap.add_argument('--step', type=float, default=1.0)
    ap.add_argument('--variant', default='l')
    ap.add_argument('--imgsz', type=int, default=640)
    ap.add_argument('--batch', type=int, default=296)
    ap.add_argument('--cpu-tiles', type=int, default=16, help='tiles in the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for k, nme in enumerate(names):
                    if r[4 + k].lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), d.get('bf16_tflops', 1590.0), 'measured'
    return 6650.0, 1400.0, 1590.0, 'fallback'


def make_mosaic_pinned(args):
    """Synthetic mosaic in raw FITS payload order (big-endian float32) in pinned host memory."""
    import torch
    from caesar_yolo_b200 import synth
    n = args.mosaic
    seed = 1234 if n <= 16384 else 5678
    img = synth.make_mosaic(n, n, seed=seed, nan_border_frac=0.0)
    b = int(0.02 * n)
    img[-b:, :] = np.nan           # 2% NaN strips (bottom/right: NaN top rows make the reference reject those tiles)
    img[:, -b:] = np.nan
    host = torch.empty((n, n), dtype=torch.int32, pin_memory=True)
    hv = host.numpy().view(np.uint32)
    rows = max(1, (1 << 24) // n)
    for y0 in range(0, n, rows):
        hv[y0:y0 + rows] = img[y0:y0 + rows].view(np.uint32).byteswap()
    return img, host


def cpu_reference(args, img_native, ntiles, threads):
    """The oracle (restated reference --devices=cpu path) on the first `ntiles` tiles of the mosaic, as a FITS file ->
    catalog run of oracle.SFinder.run_parallel.  Returns (tiles/s, seconds, tiles)."""
    import tempfile
    import torch
    from caesar_yolo_b200 import synth, weights as W
    from oracle import inference as oinf, preprocessing as opp, yolo as oy
    torch.set_num_threads(threads)
    side = int(np.ceil(np.sqrt(ntiles)))
    sub = img_native[:side * args.tile, :side * args.tile]
    tmp = tempfile.mkdtemp(prefix='cybench_')
    path = os.path.join(tmp, 'sub.fits')
    synth.write_fits(path, sub)
    w = W.make_random_weights(args.variant, 5, seed=0, cls_bias=CLS_BIAS.get(args.variant, -16.0))
    dp = opp.DataPreprocessor(opp.build_stages(**PP_FLAGS))
    cfg = dict(img_size=args.imgsz, preprocess_fcn=dp, image_path=path, image_xmin=-1, image_xmax=-1, image_ymin=-1,
               image_ymax=-1, mpi=None, split_image_in_tiles=True, tile_xsize=args.tile, tile_ysize=args.tile,
               tile_xstep=args.step, tile_ystep=args.step, max_ntasks_per_worker=1 << 30, devices=['cpu'],
               iou_thr=IOU_THR, merge_overlap_iou_thr_soft=SOFT, merge_overlap_iou_thr_hard=HARD, score_thr=SCORE_THR,
               save_catalog=True, outdir=tmp)
    model = oy.OracleModel(w)
    sf = oinf.SFinder(model, cfg)
    t0 = time.time()
    sf.run_parallel()
    dt = time.time() - t0
    nt = len(sf.tasks_per_worker[0])
    return nt / dt, dt, nt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from caesar_yolo_b200 import synth
    threads = os.cpu_count() or 1
    n = args.tile * int(np.ceil(np.sqrt(args.cpu_tiles)))
    img = synth.make_mosaic(n, n, seed=1234, nan_border_frac=0.0)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, nt = cpu_reference(args, img, args.cpu_tiles, threads)
        if i >= args.warmup:
            vals.append((v, dt, nt))
        if sum(x[1] for x in vals) > 240:
            break
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals])) * 1e3
    nt = vals[0][2]
    sample = "%d tiles (%dx%d sub-mosaic of the same generator), oracle SFinder.run_parallel FITS->catalog, torch CPU fp32 batch 1 per tile, logging silenced" % (nt, n, n)
    line = {"impl": "reference", "metric": "tiles/s FITS->catalog", "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mpix_per_s": v * args.tile * args.tile / 1e6}
    emit(line)
    return 0


def workload_config(args, n, T=None):
    if T is None:
        from caesar_yolo_b200 import ops
        T = len(ops.generate_tiles(0, args.mosaic - 1, 0, args.mosaic - 1, args.tile, args.tile, args.step, args.step))
    return {"workload": "synthetic %dx%d f32 mosaic (FITS byte order), %dx%d tiles step %.1f (%d tiles), %s nc=5 "
                        "random-init, imgsz %d, full preprocessing chain, FITS payload -> merged catalog"
                        % (args.mosaic, args.mosaic, args.tile, args.tile, args.step, T,
                           ("YOLO" + args.variant) if args.variant.startswith('11') else ("YOLOv8" + args.variant), args.imgsz),
            "tiles": T, "tile_batch": args.batch, "parallelism": "tile-row bands x%d + NCCL all-gather of records" % n,
            "l2": "inputs larger than L2 (mosaic band >= 134 MB, activations > 1 GB per batch)",
            "score_thr": SCORE_THR, "iou_thr": IOU_THR}


def main():
    args = parse()
    claim_stdout()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda:%d' % local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from caesar_yolo_b200 import ops, pipeline, weights as W
    hbm_peak, tf_sust, tf_burst, peak_kind = peaks()

    img_native, host = make_mosaic_pinned(args)
    tiles = ops.generate_tiles(0, args.mosaic - 1, 0, args.mosaic - 1, args.tile, args.tile, args.step, args.step)
    T = len(tiles)
    w = W.make_random_weights(args.variant, 5, seed=0, cls_bias=CLS_BIAS.get(args.variant, -16.0))
    eng = pipeline.Engine(w, pipeline.make_pp_config(**PP_FLAGS), imgsz=args.imgsz, score_thr=SCORE_THR,
                          iou_thr=IOU_THR, thr_soft=SOFT, thr_hard=HARD, device=dev, batch_tiles=args.batch)
    a, b = pipeline.split_tile_rows(tiles, world)[rank]
    y0b, y1b = int(tiles['ymin'][a:b].min()), int(tiles['ymax'][a:b].max())
    band_dev = host[y0b:y1b].to(dev)          # resident copy for the device-timed `value`
    my_ids = np.arange(a, b, dtype=np.int32)

    def step_resident():
        eng.begin(tiles)
        eng.process_tiles(band_dev, args.mosaic, True, 0, y0b, my_ids)
        packed, n = eng.finish()
        if world > 1:
            packed, n = pipeline.allgather_records(packed, n, world)
        return eng.global_merge(packed, n), n

    def step_e2e():
        return pipeline.run_image(eng, host, True, tiles, rank=rank, world=world, on_rank0_only=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        out = None
        for _ in range(k):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.time() - t0
        barrier()
        ms = max(e0.elapsed_time(e1), wall * 1e3 * 0.0)  # device time on the launching stream
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ms_dev, wall_ms, (src, nrec) = timed(step_resident, args.steps)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(max(2, args.warmup - 1)):
        step_e2e()
    ms_e2e, wall_e2e, (src2, _) = timed(step_e2e, args.steps)

    # one extra resident step with per-stage CUDA events (outside the timed region)
    eng.stage_events = []
    step_resident()
    stage_ms = {k: round(v, 3) for k, v in eng.stage_times_ms().items()}
    eng.stage_events = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_step = ms_dev / args.steps
    value = T / (ms_step * 1e-3)
    e2e_ms = max(ms_e2e, wall_e2e) / args.steps     # host-facing call: wall clock covers the host side as well
    e2e_val = T / (e2e_ms * 1e-3)
    mpix = T * args.tile * args.tile / 1e6
    line = {"metric": "tiles/s FITS->catalog", "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world, T),
            "mpix_per_s": mpix / (ms_step * 1e-3), "sources": int(len(src)), "records": int(nrec),
            "e2e": {"value": e2e_val, "unit": "tiles/s", "h2d_bytes_per_step": int(args.mosaic) * int(args.mosaic) * 4,
                    "d2h_bytes_per_step": int(len(src2)) * 32 * world, "ms_per_step": e2e_ms,
                    "mpix_per_s": mpix / (e2e_ms * 1e-3)},
            "gpu_launches": int(launches), "clocks": clocks, "stage_ms_per_step": stage_ms}

    # ---- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv), measured live with CUDA events on the
    # launching stream: per-op events around one batch forward (Model::profile), summed over the conv launches.
    if not args.no_profile:
        Sh, Sw, _ = ops.letterbox_shape(args.tile, args.tile, args.imgsz)
        x = torch.rand((args.batch, Sh, Sw, 4), device=dev).to(torch.bfloat16)
        prof = eng.model.profile(x)
        prof = eng.model.profile(x)
        # model.0 is the fused mma.sync stem kernel (HBM-bound), every other op with flops is a conv_igemm_kernel launch
        conv = [(n_, ms, fl) for (n_, ms, fl) in prof if fl > 0 and n_ != 'model.0']
        conv_ms = sum(p[1] for p in conv)
        conv_fl = sum(p[2] for p in conv)
        tot_ms = sum(p[1] for p in prof)
        ach = conv_fl / (conv_ms * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05 implicit GEMM, %d launches per batch of %d tiles)" % (len(conv), args.batch),
                            "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s", "frac": ach / tf_sust,
                            "peak_kind": "%s sustained bf16 cuBLAS (kernel timed inside a long step)" % peak_kind,
                            "traffic": None, "avg_launch_us": conv_ms * 1e3 / len(conv),
                            "algorithmic_flops_per_launch": conv_fl / len(conv),
                            "conv_share_of_forward": conv_ms / tot_ms,
                            "forward_ms_per_tile": tot_ms / args.batch}
        # dram traffic per launch from the committed ncu capture (profiles/conv_traffic.json), scaled to this batch;
        # algorithmic bytes per launch from the plan (input + weights + output + residual, each once)
        try:
            if args.variant != 'l' or args.imgsz != 640 or args.tile != 512:
                raise ValueError('the committed ncu capture is of the yolov8l / 512 / 640 workload')
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'conv_traffic.json')))
            line["roofline"]["traffic"] = (tj["dram_read_bytes"] + tj["dram_write_bytes"]) / tj["conv_launches"] * (
                args.batch / float(tj["tiles"]))
            line["roofline"]["traffic_source"] = tj["source"]
            line["roofline"]["tensor_pipe_active_pct_ncu"] = tj.get("tensor_pipe_active_pct_time_weighted")
        except Exception:
            pass
        ab, nconv = eng.model.conv_bytes(args.batch, Sh, Sw)
        line["roofline"]["algorithmic_bytes_per_launch"] = ab / max(nconv, 1)
        slow = sorted(prof, key=lambda p: -p[1])[:8]
        line["roofline"]["top_ops"] = [{"op": n_, "ms": round(ms, 4), "tflops": round(fl / (ms * 1e-3) / 1e12, 1) if ms > 0 else 0} for n_, ms, fl in slow]

    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, nt = cpu_reference(args, img_native, args.cpu_tiles, threads)
        line["cpu_baseline"] = {"value": v, "unit": "tiles/s", "cores": threads, "kind": "port",
                                "sample": "%d tiles of the same mosaic (top-left sub-mosaic), oracle SFinder.run_parallel FITS->catalog in %.1f s, torch CPU fp32 batch 1, logging silenced" % (nt, dt)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
