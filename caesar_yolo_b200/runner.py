"""Thin Python handle on the whole-path C entry (cy_ctx_* / cy_run_mosaic, include/caesar_b200.h): FITS file -> merged
catalog with the orchestration (tile grouping, file read + upload overlap, stage launches, exchange, global merge) done
in C++.  The Python engine (pipeline.Engine / run_image) drives the same stage entry points from Python; both give
identical catalogs (tests/test_runner_gpu.py).  Replaces SFinder.run_parallel (caesar_yolo/inference.py:578-658)."""
import ctypes

import numpy as np

from . import ops
from ._capi import ALLGATHER_FN, CaesarB200Error, PPChain, RunConfig, c_int, c_void_p, check, lib


class MosaicRunner(object):
    def __init__(self, model, pp_chain=None, imgsz=640, score_thr=0.7, iou_thr=0.5, thr_soft=0.3, thr_hard=0.8,
                 tile=(512, 512), step=(1.0, 1.0), region=(-1, -1, -1, -1), batch_tiles=0, read_threads=0, rank=0,
                 world=1):
        """model: ops.DeviceModel; pp_chain: PPChain / PPConfig / None; tile = (tile_x, tile_y), (0, 0): no tiling."""
        self.model = model
        cfg = RunConfig()
        cfg.imgsz = int(imgsz)
        cfg.score_thr, cfg.iou_thr = float(score_thr), float(iou_thr)
        cfg.thr_soft, cfg.thr_hard = float(thr_soft), float(thr_hard)
        cfg.tile_x, cfg.tile_y = int(tile[0]), int(tile[1])
        cfg.step_x, cfg.step_y = float(step[0]), float(step[1])
        cfg.xmin, cfg.xmax, cfg.ymin, cfg.ymax = [int(v) for v in region]
        cfg.batch_tiles, cfg.read_threads = int(batch_tiles), int(read_threads)
        cfg.rank, cfg.world = int(rank), int(world)
        self.cfg = cfg
        if pp_chain is not None and not isinstance(pp_chain, PPChain):
            pp_chain = ops.chain_from_config(pp_chain)
        self._chain = pp_chain
        h = c_void_p(0)
        check(lib.cy_ctx_create(model._h, ctypes.byref(pp_chain) if pp_chain is not None else None, ctypes.byref(cfg),
                                ctypes.byref(h)))
        self._h = h
        self._cb = None

    def set_allgather(self, fn):
        """fn(send_ptr, recv_ptr, bytes_per_rank, stream) -> 0 on success (device pointers as ints)."""
        self._cb = ALLGATHER_FN(lambda user, s, r, n, st: int(fn(s, r, n, st)))
        check(lib.cy_ctx_set_allgather(self._h, self._cb, None))

    def set_allgather_torch(self):
        """Registers torch.distributed's all-gather (NCCL) as the exchange of cy_run_mosaic: the callback stages the
        slot through two torch tensors on the context's own stream (wrapped as an ExternalStream, so the copies and the
        collective are ordered with the C side's work without a host synchronisation)."""
        import torch
        import torch.distributed as dist
        from ._capi import c_uptr
        state = {}

        def gather(send_ptr, recv_ptr, nbytes, stream):
            try:
                world = dist.get_world_size()
                if state.get('n') != nbytes:
                    dev = torch.device('cuda', torch.cuda.current_device())
                    state['send'] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    state['recv'] = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
                    state['n'] = nbytes
                ext = torch.cuda.ExternalStream(stream)
                with torch.cuda.stream(ext):
                    check(lib.cy_memcpy_d2d(c_void_p(state['send'].data_ptr()), c_void_p(send_ptr),
                                            ctypes.c_size_t(nbytes), c_uptr(stream)))
                    dist.all_gather_into_tensor(state['recv'], state['send'])
                    check(lib.cy_memcpy_d2d(c_void_p(recv_ptr), c_void_p(state['recv'].data_ptr()),
                                            ctypes.c_size_t(nbytes * world), c_uptr(stream)))
                return 0
            except Exception:
                return 1
        self.set_allgather(gather)

    def run(self, fits_path, capacity=1 << 20):
        """-> (sources structured array (ops.SRC_DTYPE), number of records)."""
        out = np.zeros(capacity, dtype=ops.SRC_DTYPE)
        ns, nr = c_int(0), c_int(0)
        check(lib.cy_run_mosaic(self._h, str(fits_path).encode(), out.ctypes.data_as(c_void_p), c_int(capacity),
                                ctypes.byref(ns), ctypes.byref(nr)))
        return out[:ns.value].copy(), nr.value

    def run_local(self, fits_path):
        n = c_int(0)
        check(lib.cy_run_local(self._h, str(fits_path).encode(), ctypes.byref(n)))
        return n.value

    def pack_slot(self, cap):
        p = c_void_p(0)
        check(lib.cy_ctx_pack_slot(self._h, c_int(cap), ctypes.byref(p)))
        return p.value

    def unpack_slots(self, slots_ptr, world, cap):
        p, n, cmax = c_void_p(0), c_int(0), c_int(0)
        check(lib.cy_ctx_unpack_slots(self._h, c_void_p(slots_ptr), c_int(world), c_int(cap), ctypes.byref(p),
                                      ctypes.byref(n), ctypes.byref(cmax)))
        return p.value, n.value, cmax.value

    def merge(self, recs_ptr, n, capacity=1 << 20):
        out = np.zeros(capacity, dtype=ops.SRC_DTYPE)
        ns = c_int(0)
        check(lib.cy_run_merge(self._h, c_void_p(recs_ptr), c_int(n), out.ctypes.data_as(c_void_p), c_int(capacity),
                               ctypes.byref(ns)))
        return out[:ns.value].copy()

    def info(self):
        a = (ctypes.c_double * 8)()
        check(lib.cy_ctx_info(self._h, a))
        return dict(tiles=int(a[0]), tiles_processed=int(a[1]), bytes_uploaded=int(a[2]), first=int(a[3]), last=int(a[4]),
                    records=int(a[5]))

    def close(self):
        if self._h:
            lib.cy_ctx_destroy(self._h)
            self._h = c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
