"""Host-side driver of the B200 hot path: FITS mosaic -> tiles -> preprocessing -> YOLOv8 forward -> decode/NMS ->
per-tile merge -> records -> (all-gather) -> cross-tile merge -> catalog.  Python only orchestrates; every stage is a
CUDA kernel behind the C ABI (ops.py)."""
from ._capi import PPConfig


def make_pp_config(enabled=True, subtract_bkg=False, sigma_bkg=3.0, use_box_mask_in_bkg=False, bkg_box_mask_fract=0.7,
                   bkg_chid=-1, clip_shift_data=False, sigma_clip=1.0, clip_chid=-1, clip_data=False,
                   sigma_clip_low=10.0, sigma_clip_up=10.0, nchannels=1, zscale_stretch=False,
                   zscale_contrasts=(0.25, 0.25, 0.25), chan3_preproc=False, sigma_clip_baseline=0.0,
                   normalize_minmax=False, norm_min=0.0, norm_max=1.0):
    """cy_pp_config from the run.py option names and defaults (scripts/run.py:80-107, stage order :272-293)."""
    c = PPConfig()
    c.enabled = 1 if enabled else 0
    c.subtract_bkg = int(bool(subtract_bkg))
    c.sigma_bkg = float(sigma_bkg)
    c.use_box_mask_in_bkg = int(bool(use_box_mask_in_bkg))
    c.bkg_box_mask_fract = float(bkg_box_mask_fract)
    c.bkg_chid = int(bkg_chid)
    c.clip_shift_data = int(bool(clip_shift_data))
    c.sigma_clip = float(sigma_clip)
    c.clip_chid = int(clip_chid)
    c.clip_data = int(bool(clip_data))
    c.sigma_clip_low = float(sigma_clip_low)
    c.sigma_clip_up = float(sigma_clip_up)
    c.nchannels = int(nchannels)
    c.zscale_stretch = int(bool(zscale_stretch))
    zc = list(zscale_contrasts) + [0.25] * 3
    for i in range(3):
        c.zscale_contrasts[i] = float(zc[i])
    c.chan3_preproc = int(bool(chan3_preproc))
    c.sigma_clip_baseline = float(sigma_clip_baseline)
    c.normalize_minmax = int(bool(normalize_minmax))
    c.norm_min = float(norm_min)
    c.norm_max = float(norm_max)
    return c


# ======================================================================================================= engine

import time

import numpy as np
import torch

from . import ops
from ._capi import CaesarB200Error, Letterbox


class Engine(object):
    """One GPU's share of the hot path.  Holds the device model and reusable buffers; `process_tiles` runs
    preprocessing -> forward -> decode/NMS -> per-tile merge -> records for a list of tiles cut out of an image that
    is already in device memory; `finish` compacts the records; `global_merge` builds the catalog.

    Replaces TileTask.find_sources + Analyzer.predict (caesar_yolo/inference.py:173-275, evaluation.py:128-245) for
    all tiles of a rank at once (the reference runs them one by one, batch 1)."""

    def __init__(self, weights, pp_cfg, imgsz=640, score_thr=0.7, iou_thr=0.5, thr_soft=0.3, thr_hard=0.8,
                 device=None, batch_tiles=296, pp_tiles=296):
        if not torch.cuda.is_available():
            raise CaesarB200Error("no CUDA device: the B200 path has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        ops.check(ops.lib.cy_device_check())
        self.model = weights if isinstance(weights, ops.DeviceModel) else ops.DeviceModel(weights)
        self.nc = self.model.nc
        self.names = self.model.names
        self.pp_cfg = pp_cfg if pp_cfg is not None else make_pp_config(enabled=False)
        self.imgsz = int(imgsz)
        self.score_thr, self.iou_thr = float(score_thr), float(iou_thr)
        self.thr_soft, self.thr_hard = float(thr_soft), float(thr_hard)
        self.batch_tiles = int(batch_tiles)
        self.pp_tiles = max(int(pp_tiles), self.batch_tiles)
        self._buf = {}
        self.launches = 0      # kernels launched by this engine (bench.py's gpu_launches)
        self.stage_events = None   # set to [] to collect (stage, start_event, end_event) triples (bench breakdown)
        self._fwd_launches = {}
        # per-tile debug outputs (--save_tile_catalog / --save_tile_region / --save_tile_img, inference.py:218-229):
        self.collect_tile_status = False   # keep the per-tile status (0 ok, <0 rejected) in self.tile_status
        self.tile_img_sink = None          # callable(tile_id, ndarray[Ty,Tx] float32): channel 0 after the chain

    def _mark(self):
        if self.stage_events is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _stage(self, name, e0):
        if e0 is not None:
            self.stage_events.append((name, e0, self._mark()))

    def stage_times_ms(self):
        """Sum of CUDA-event durations per stage since stage_events was last reset (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.stage_events or []:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    # ---- buffers -----------------------------------------------------------------------------------------
    def _get(self, key, shape, dtype):
        t = self._buf.get(key)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty((n,), dtype=dtype, device=self.device)
            self._buf[key] = t
        return t[:n].view(*shape)

    def begin(self, tiles):
        """tiles: structured array (ops.TILE_DTYPE) of ALL tiles of the mosaic (global ids = row-major index)."""
        self.tiles = np.ascontiguousarray(tiles)
        self.T = len(tiles)
        self.tiles_dev = ops.to_device_bytes(self.tiles, self.device)
        self.rec_slots = self._get('rec_slots', (self.T * ops.MAX_DET * 32,), torch.uint8)
        self.nrec = self._get('nrec', (self.T,), torch.int32)
        self.nrec.zero_()
        self.launches += 1
        self.tile_status = None
        if self.collect_tile_status:
            self.tile_status = torch.full((self.T,), -3, dtype=torch.int32, device=self.device)  # -3: not processed

    def process_tiles(self, img_dev, row_stride, big_endian, origin_x, origin_y, tile_ids, ready=None):
        """img_dev covers mosaic rows/cols starting at (origin_x, origin_y); tile_ids: global ids of the tiles to
        process (all must lie inside img_dev).  Groups by tile shape (edge tiles are smaller).
        Preprocessing runs in groups of `pp_tiles` tiles (one CTA per tile: a multiple of the SM count keeps all 148
        SMs busy); the conv stack and the detect/merge kernels run in batches of `batch_tiles`.
        ready(y): optional callback invoked before a group is launched, y = last mosaic row (exclusive) the group
        reads (run_image uses it to wait for / issue the upload of those rows)."""
        tile_ids = np.asarray(tile_ids, dtype=np.int32)
        if tile_ids.size == 0:
            return
        t = self.tiles[tile_ids]
        w = t['xmax'] - t['xmin']
        h = t['ymax'] - t['ymin']
        shapes = sorted(set(zip(h.tolist(), w.tolist())), reverse=True)
        for (Ty, Tx) in shapes:
            ids = tile_ids[(h == Ty) & (w == Tx)]
            for s in range(0, len(ids), self.pp_tiles):
                g = ids[s:s + self.pp_tiles]
                if ready is not None:
                    ready(int(self.tiles['ymax'][g].max()))
                self._run_group(img_dev, row_stride, big_endian, origin_x, origin_y, g, Ty, Tx)

    def _run_group(self, img_dev, row_stride, big_endian, ox, oy, ids, Ty, Tx):
        G = len(ids)
        dev = self.device
        tl = self.tiles[ids]
        meta = np.concatenate([(tl['xmin'] - ox).astype(np.int32), (tl['ymin'] - oy).astype(np.int32),
                               ids.astype(np.int32)])
        meta_dev = torch.from_numpy(meta).to(dev, non_blocking=True)
        x0, y0, ids_dev = meta_dev[:G], meta_dev[G:2 * G], meta_dev[2 * G:]
        Sh, Sw, lb = ops.letterbox_shape(Ty, Tx, self.imgsz)
        chain = self._get('chain', (G, Ty, Tx, 3), torch.float32)
        model_in = self._get('model_in', (G, Sh, Sw, 4), torch.bfloat16)
        status = self._get('pp_status', (G,), torch.int32)
        need = int(ops.lib.cy_preprocess_scratch_bytes(ops.ctypes.byref(self.pp_cfg), G, Ty, Tx))
        scratch = self._get('pp_scratch', (need,), torch.uint8)
        e = self._mark()
        ops.preprocess(self.pp_cfg, img_dev, row_stride, big_endian, x0, y0, Ty, Tx, self.imgsz, scratch=scratch,
                       chain_out=chain, model_in=model_in, status=status)
        self._stage('preprocess', e)
        self.launches += 3
        if self.tile_img_sink is not None:   # Analyzer.write_fits: channel 0 of the preprocessed image, accepted tiles
            ok = status.cpu().numpy()
            ch0 = chain[:, :, :, 0].cpu().numpy()
            for k in range(G):
                if ok[k] == 0:
                    self.tile_img_sink(int(ids[k]), ch0[k])
        for s in range(0, G, self.batch_tiles):
            e = min(G, s + self.batch_tiles)
            self._run_batch(model_in[s:e], status[s:e], ids_dev[s:e], Ty, Tx, Sh, Sw, lb)

    def _run_batch(self, model_in, status, ids_dev, Ty, Tx, Sh, Sw, lb):
        B = model_in.shape[0]
        dev = self.device
        key = (B, Ty, Tx)
        lbd = self._buf.get(('lb',) + key)
        if lbd is None:
            lbd = ops.letterbox_array([lb] * B, dev)
            self._buf[('lb',) + key] = lbd
        e = self._mark()
        heads = self.model.forward(model_in)
        self._stage('forward', e)
        e = self._mark()
        need = int(ops.lib.cy_postprocess_scratch_bytes(B, Sh, Sw, ops.MAX_DET))
        pscr = self._get('post_scratch', (need,), torch.uint8)
        dets = self._get('dets', (B, ops.MAX_DET, 6), torch.float32)
        ndets = self._get('ndets', (B,), torch.int32)
        ops.postprocess(heads, B, Sh, Sw, self.nc, self.score_thr, self.iou_thr, lbd, dev, scratch=pscr, dets=dets,
                        ndets=ndets)
        self._stage('decode_nms', e)
        e = self._mark()
        keep = self._get('keep', (B, ops.MAX_DET), torch.int32)
        nkeep = self._get('nkeep', (B,), torch.int32)
        mstat = self._get('mstat', (B,), torch.int32)
        ops.merge_tile(dets, ndets, self.score_thr, self.thr_soft, self.thr_hard, keep_idx=keep, nkeep=nkeep,
                       status=mstat, pre_status=status)
        ops.make_records(dets, keep, nkeep, mstat, self.tiles_dev, ids_dev, self.rec_slots, self.nrec)
        self._stage('merge_tile_records', e)
        if self.tile_status is not None:
            self.tile_status.index_copy_(0, ids_dev.long(), mstat)
        fl = self._fwd_launches.get((B, Sh, Sw))
        if fl is None:
            fl = int(self.model.info(B, Sh, Sw)['launches'])
            self._fwd_launches[(B, Sh, Sw)] = fl
        # forward + memset/score/nms (3) + merge + records (2)
        self.launches += fl + 3 + 2

    def finish(self):
        """Compacts the per-tile record slots -> (packed uint8 tensor of n cy_det_record, n) in tile-id order."""
        total = self._get('total', (1,), torch.int32)
        packed = self._get('packed', (self.T * ops.MAX_DET * 32,), torch.uint8)
        ops.compact_records(self.rec_slots, self.nrec, self.T, ops.MAX_DET, packed, total)
        self.launches += 4
        n = int(total.item())
        return packed[:n * 32], n

    def global_merge(self, packed, n, nb_off=None, nb_idx=None):
        """find_sources_at_edge + merge_edge_sources -> numpy structured array (ops.SRC_DTYPE)."""
        if nb_off is None:
            nb_off, nb_idx = ops.tile_neighbors(self.tiles)
        key = ('nb', self.T, int(nb_off[-1]))
        if key not in self._buf:
            self._buf[key] = (torch.from_numpy(nb_off).to(self.device),
                              torch.from_numpy(nb_idx if len(nb_idx) else np.zeros(1, np.int32)).to(self.device))
        off_dev, idx_dev = self._buf[key]
        out = self._get('sources', (max(n, 1) * 32,), torch.uint8)
        self.launches += 33
        e = self._mark()
        res = ops.merge_global(packed, n, self.tiles_dev, self.T, off_dev, idx_dev, out=out)
        self._stage('merge_global', e)
        return res


def split_tile_rows(tiles, nparts):
    """Partition of the row-major tile grid into `nparts` contiguous bands of whole tile rows, balanced by tile count.
    Returns list of (first_tile_id, last_tile_id_exclusive).  (Replaces the reference's round-robin tile->rank map,
    inference.py:1008-1029; the gathered list stays in tile-id order for any rank count.)"""
    T = len(tiles)
    ymins = tiles['ymin']
    row_starts = [0] + [i for i in range(1, T) if ymins[i] != ymins[i - 1]] + [T]
    nrows = len(row_starts) - 1
    parts = []
    for r in range(nparts):
        a = row_starts[(nrows * r) // nparts]
        b = row_starts[(nrows * (r + 1)) // nparts]
        parts.append((a, b))
    return parts


def allgather_records(packed, n, world_size):
    """Exchange step of the path (replaces SFinder.gather_task_data_from_workers, inference.py:936-984): all-gather of
    the fixed 32-byte detection records (counts first, then records padded to the max count — NCCL has no
    all-gather-v).  Ranks own ascending contiguous tile-id ranges, so concatenating in rank order keeps the list in
    tile-id order, i.e. the reference's nproc=1 order, for any world size.  Works on CUDA tensors over NCCL and on
    CPU tensors over gloo (tests)."""
    import torch.distributed as dist
    if world_size == 1:
        return packed, n
    dev = packed.device
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    counts = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world_size)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    send = torch.zeros((mx * 32,), dtype=torch.uint8, device=dev)
    send[:n * 32] = packed[:n * 32]
    if dev.type == 'cuda':
        recv = torch.empty((world_size * mx * 32,), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send)
        parts = [recv[r * mx * 32: r * mx * 32 + counts[r] * 32] for r in range(world_size)]
    else:
        bufs = [torch.empty((mx * 32,), dtype=torch.uint8) for _ in range(world_size)]
        dist.all_gather(bufs, send)
        parts = [bufs[r][:counts[r] * 32] for r in range(world_size)]
    return torch.cat(parts), sum(counts)


def run_image(engine, img_host, big_endian, tiles, rank=0, world=1, piece_bytes=1 << 26, on_rank0_only=True,
              on_local_records=None):
    """FITS payload in HOST memory -> catalog.  img_host: 2-D array [ny,nx] of 4-byte pixels (numpy array / memmap, or
    a pinned torch tensor for zero-copy staging); big_endian: raw FITS byte order.  The rows of this rank's tiles
    (contiguous band of tile rows) are uploaded in pieces of ~piece_bytes on a copy stream; every tile group waits
    only for the rows it needs, so all but the first group's upload overlaps the compute of earlier groups.
    on_local_records(packed, n, first_tile_id, last_tile_id_excl): called with this rank's records before the exchange
    (per-tile output files).  Returns (sources structured array or None on ranks != 0, n_records_total)."""
    engine.begin(tiles)
    a, b = split_tile_rows(tiles, world)[rank]
    ids = np.arange(a, b, dtype=np.int32)
    is_torch = isinstance(img_host, torch.Tensor)
    ny, nx = (img_host.shape[0], img_host.shape[1])
    if len(ids):
        Y0, Y1 = int(tiles['ymin'][ids].min()), int(tiles['ymax'][ids].max())
        copy_stream = engine._buf.get('copy_stream')
        if copy_stream is None:
            copy_stream = torch.cuda.Stream(device=engine.device)
            engine._buf['copy_stream'] = copy_stream
        compute = torch.cuda.current_stream()
        band = engine._get('band', (Y1 - Y0, nx), torch.int32)   # reused across calls: no allocation per image
        copy_stream.wait_stream(compute)                          # earlier work may still read the band buffer
        rows_per_piece = max(1, int(piece_bytes) // (nx * 4))
        state = {'row': Y0, 'events': []}                         # events[i] = (last_row_excl, event)

        def upload_piece():
            r0 = state['row']
            r1 = min(Y1, r0 + rows_per_piece)
            k = len(state['events'])
            with torch.cuda.stream(copy_stream):
                if is_torch:
                    src = img_host[r0:r1]
                else:
                    # two reusable pinned staging buffers (file/memmap rows -> pinned -> HBM): the host copy of piece
                    # k+1 overlaps the DMA of piece k; a buffer is rewritten only after its last DMA completed
                    nwords = (r1 - r0) * nx
                    slot = ('pin', k & 1)
                    pin, pin_ev = engine._buf.get(slot, (None, None))
                    if pin is None or pin.numel() < nwords:
                        pin = torch.empty((rows_per_piece * nx,), dtype=torch.int32, pin_memory=True)
                        pin_ev = None
                    if pin_ev is not None:
                        pin_ev.synchronize()
                    src = pin[:nwords].view(r1 - r0, nx)
                    np.copyto(src.numpy().view(np.uint8).reshape(r1 - r0, nx * 4),
                              np.ascontiguousarray(img_host[r0:r1]).view(np.uint8).reshape(r1 - r0, nx * 4))
                band[r0 - Y0:r1 - Y0].copy_(src.view(torch.int32) if src.dtype != torch.int32 else src,
                                            non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                if not is_torch:
                    engine._buf[slot] = (pin, ev)
            state['events'].append((r1, ev))
            state['row'] = r1

        def ready(y_needed):
            """Called before a tile group is launched: rows [Y0, y_needed) must be in HBM."""
            while state['row'] < min(y_needed, Y1):
                upload_piece()
            for r1, ev in state['events']:
                if r1 >= min(y_needed, Y1):
                    compute.wait_event(ev)
                    break

        engine.process_tiles(band, nx, big_endian, 0, Y0, ids, ready=ready)
    packed, n = engine.finish()
    if on_local_records is not None:
        on_local_records(packed, n, a, b)
    if world > 1:
        packed, n = allgather_records(packed, n, world)
    if rank == 0 or not on_rank0_only:
        return engine.global_merge(packed, n), n
    return None, n
