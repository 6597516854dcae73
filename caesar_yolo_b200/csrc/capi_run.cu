// Whole-path C entry: FITS file -> merged catalog without any host-language orchestration (include/caesar_b200.h,
// "whole path").  Replaces SFinder.run_parallel / TileTask.find_sources (caesar_yolo/inference.py:578-658,173-275),
// the per-tile file read (utils.py:340-418), gather_task_data_from_workers (:936-984) and find_sources_at_edge +
// merge_edge_sources (:663-931) for one GPU's share of the tiles.
//
// One context per GPU (one process per GPU).  cy_run_local:
//   tile grid (cy_generate_tiles) -> this rank's contiguous band of tile rows -> the band's rows are read from the file
//   with pread by a few threads into two pinned staging buffers and uploaded piece by piece on a copy stream; every
//   tile group (<= batch_tiles tiles of one shape; a band smaller than one group is split in two so the second upload
//   overlaps the first group's compute) waits only for the rows it needs -> cy_preprocess_chain -> cy_model_forward ->
//   cy_postprocess -> cy_merge_tile -> cy_make_records -> device compaction in tile-id order.
// Exchange: cy_ctx_pack_slot gives the fixed-capacity all-gather slot [count | records]; the host moves the slots with
// ncclAllGather (or anything else) -- or registers a callback -- and cy_ctx_unpack_slots rebuilds the tile-id-ordered
// list.  cy_run_merge: edge flags + cross-tile merge -> sources on the host.  cy_run_mosaic strings the three together.
#include "../../include/caesar_b200.h"
#include "common.h"
#include "model.h"

#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <thread>
#include <vector>

using namespace cy;

namespace {

constexpr int kMaxDet = 300;   // ultralytics non_max_suppression max_det (ops.MAX_DET)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = (n + 255) & ~(size_t)255;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return set_error(CY_ERR_NOMEM, "cy_run: cudaMalloc of %zu bytes failed", want);
        }
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct FitsInfo {
    long long offset = 0;   // byte offset of the payload
    int nx = 0, ny = 0;
};

struct RunCtx {
    Model* model = nullptr;
    cy_pp_chain chain;
    cy_run_config cfg;
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    bool pin_used[2] = {false, false};
    void* pin[2] = {nullptr, nullptr};
    size_t pin_cap = 0;
    std::vector<cudaEvent_t> piece_ev;
    cy_allgather_fn gather = nullptr;
    void* gather_user = nullptr;
    // geometry of the current run
    std::vector<cy_tile> tiles;
    int T = 0, first = 0, last = 0;   // this rank owns tiles [first, last)
    int n_local = 0;                  // records of this rank after cy_run_local
    long long launches = 0;
    double stats[8] = {0};
    DevBuf band, meta, model_in, pp_scratch, status, post_scratch, dets, ndets, keep, nkeep, mstat, lb, tiles_dev,
        rec_slots, nrec, packed, total, compact_scratch, nb_off, nb_idx, sources, nout, xsend, xrecv, xall, xtotal,
        xcounts, xscratch;
    std::string nb_key;
};

int parse_fits_header(int fd, FitsInfo* fi) {
    // Primary HDU of a FITS file: 80-byte cards in 2880-byte blocks up to END (utils.read_fits, utils.py:150-246: the
    // reference takes plane [0, 0] of a 4-D cube and rejects everything that is not 2-D or 4-D).
    char block[2880];
    long long off = 0;
    int bitpix = 0, naxis = -1;
    long long ax[8] = {0};
    double bscale = 1.0, bzero = 0.0;
    bool end = false;
    while (!end) {
        if (pread(fd, block, sizeof(block), off) != (ssize_t)sizeof(block))
            return set_error(CY_ERR_INVALID, "cy_run: truncated FITS header");
        if (off == 0 && strncmp(block, "SIMPLE  =", 9) != 0) return set_error(CY_ERR_INVALID, "cy_run: not a FITS file");
        for (int c = 0; c < 36 && !end; ++c) {
            char card[81];
            memcpy(card, block + 80 * c, 80);
            card[80] = 0;
            if (strncmp(card, "END     ", 8) == 0) {
                end = true;
                break;
            }
            if (card[8] != '=') continue;
            char key[9];
            memcpy(key, card, 8);
            key[8] = 0;
            const char* v = card + 10;
            if (!strcmp(key, "BITPIX  ")) bitpix = atoi(v);
            else if (!strcmp(key, "NAXIS   ")) naxis = atoi(v);
            else if (!strncmp(key, "NAXIS", 5) && key[5] >= '1' && key[5] <= '8' && key[6] == ' ') ax[key[5] - '1'] = atoll(v);
            else if (!strcmp(key, "BSCALE  ")) bscale = atof(v);
            else if (!strcmp(key, "BZERO   ")) bzero = atof(v);
        }
        off += 2880;
        if (off > (1ll << 24)) return set_error(CY_ERR_INVALID, "cy_run: FITS header without END");
    }
    if (naxis != 2 && naxis != 4)   // utils.read_fits / read_fits_crop (utils.py:207-216,378-386): 2-D images and 4-D cubes only
        return set_error(CY_ERR_INVALID, "cy_run: invalid/unsupported number of channels (NAXIS=%d)", naxis);
    if (bitpix != -32 || bscale != 1.0 || bzero != 0.0)
        return set_error(CY_ERR_INVALID,
                         "cy_run: only unscaled BITPIX=-32 payloads are read in place (BITPIX=%d BSCALE=%g BZERO=%g): "
                         "convert on the host and use cy_run_payload", bitpix, bscale, bzero);
    if (ax[0] <= 0 || ax[1] <= 0 || ax[0] > 0x7fffffff || ax[1] > 0x7fffffff)
        return set_error(CY_ERR_INVALID, "cy_run: bad NAXIS1/NAXIS2");
    fi->offset = off;
    fi->nx = (int)ax[0];
    fi->ny = (int)ax[1];
    return CY_OK;
}

// pread of [offset, offset + n) into dst, split over `nt` threads
int read_range(int fd, long long offset, char* dst, size_t n, int nt) {
    if (n < (8u << 20)) nt = 1;
    nt = std::max(1, std::min(nt, 32));
    size_t step = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
    std::vector<int> rc(nt, 0);
    auto one = [&](int i) {
        const size_t a = (size_t)i * step, b = std::min(n, a + step);
        size_t done = a;
        while (done < b) {
            const ssize_t got = pread(fd, dst + done, b - done, offset + (long long)done);
            if (got <= 0) {
                rc[i] = -1;
                return;
            }
            done += (size_t)got;
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i)
        if ((size_t)i * step < n) th.emplace_back(one, i);
    one(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < nt; ++i)
        if (rc[i]) return set_error(CY_ERR_INVALID, "cy_run: short read from the FITS payload");
    return CY_OK;
}

// split_tile_rows (pipeline.py): contiguous bands of whole tile rows, balanced by row count
void split_rows(const std::vector<cy_tile>& tiles, int world, int rank, int* a, int* b) {
    const int T = (int)tiles.size();
    std::vector<int> row_starts;
    row_starts.push_back(0);
    for (int i = 1; i < T; ++i)
        if (tiles[i].ymin != tiles[i - 1].ymin) row_starts.push_back(i);
    row_starts.push_back(T);
    const long long nrows = (long long)row_starts.size() - 1;
    *a = row_starts[(size_t)((nrows * rank) / world)];
    *b = row_starts[(size_t)((nrows * (rank + 1)) / world)];
}

struct RowSource {      // where the payload rows come from
    int fd = -1;
    long long offset = 0;
    const char* host = nullptr;   // payload in host memory (pageable or pinned)
    int host_pinned = 0;
    int nx = 0;
};

int run_batch(RunCtx* c, const char* model_in, const int32_t* status, const int32_t* ids_dev, int B, int Ty, int Tx,
              int Sh, int Sw, const cy_letterbox& lb) {
    cudaStream_t st = c->compute;
    int rc;
    {   // cy_letterbox[B]: the same geometry for every tile of the batch
        std::vector<cy_letterbox> h((size_t)B, lb);
        if ((rc = c->lb.ensure(sizeof(cy_letterbox) * (size_t)B))) return rc;
        CY_CUDA_CHECK(cudaMemcpyAsync(c->lb.p, h.data(), sizeof(cy_letterbox) * (size_t)B, cudaMemcpyHostToDevice, st));
        CY_CUDA_CHECK(cudaStreamSynchronize(st));   // h goes out of scope; tiny and once per batch
    }
    const float* heads[3];
    if ((rc = cy_model_forward(c->model, model_in, B, Sh, Sw, heads, (uintptr_t)st))) return rc;
    if ((rc = c->post_scratch.ensure(cy_postprocess_scratch_bytes(B, Sh, Sw, kMaxDet)))) return rc;
    if ((rc = c->dets.ensure((size_t)B * kMaxDet * 6 * 4))) return rc;
    if ((rc = c->ndets.ensure((size_t)B * 4))) return rc;
    if ((rc = c->keep.ensure((size_t)B * kMaxDet * 4))) return rc;
    if ((rc = c->nkeep.ensure((size_t)B * 4))) return rc;
    if ((rc = c->mstat.ensure((size_t)B * 4))) return rc;
    if ((rc = cy_postprocess(heads[0], heads[1], heads[2], B, Sh, Sw, c->model->nc, c->cfg.score_thr, c->cfg.iou_thr,
                             kMaxDet, (const cy_letterbox*)c->lb.p, (float*)c->dets.p, (int32_t*)c->ndets.p,
                             c->post_scratch.p, (uintptr_t)st)))
        return rc;
    if ((rc = cy_merge_tile((const float*)c->dets.p, (const int32_t*)c->ndets.p, B, kMaxDet, c->cfg.score_thr,
                            c->cfg.thr_soft, c->cfg.thr_hard, status, (int32_t*)c->keep.p, (int32_t*)c->nkeep.p,
                            (int32_t*)c->mstat.p, (uintptr_t)st)))
        return rc;
    if ((rc = cy_make_records((const float*)c->dets.p, (const int32_t*)c->keep.p, (const int32_t*)c->nkeep.p,
                              (const int32_t*)c->mstat.p, kMaxDet, (const cy_tile*)c->tiles_dev.p, ids_dev, B,
                              (cy_det_record*)c->rec_slots.p, (int32_t*)c->nrec.p, (uintptr_t)st)))
        return rc;
    (void)Ty;
    (void)Tx;
    return CY_OK;
}

int run_group(RunCtx* c, long long row_stride, int big_endian, int ox, int oy, const std::vector<int>& ids, int Ty,
              int Tx) {
    const int G = (int)ids.size();
    cudaStream_t st = c->compute;
    int rc;
    std::vector<int32_t> meta((size_t)3 * G);
    for (int k = 0; k < G; ++k) {
        meta[k] = c->tiles[ids[k]].xmin - ox;
        meta[G + k] = c->tiles[ids[k]].ymin - oy;
        meta[2 * G + k] = ids[k];
    }
    if ((rc = c->meta.ensure(meta.size() * 4))) return rc;
    CY_CUDA_CHECK(cudaMemcpyAsync(c->meta.p, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, st));
    CY_CUDA_CHECK(cudaStreamSynchronize(st));
    const int32_t* x0 = (const int32_t*)c->meta.p;
    const int32_t* y0 = x0 + G;
    const int32_t* ids_dev = x0 + 2 * G;
    int Sh = 0, Sw = 0;
    cy_letterbox lb;
    if ((rc = cy_letterbox_shape(Ty, Tx, c->cfg.imgsz, &Sh, &Sw, &lb))) return rc;
    if ((rc = c->model_in.ensure((size_t)G * Sh * Sw * 4 * 2))) return rc;
    if ((rc = c->status.ensure((size_t)G * 4))) return rc;
    if ((rc = c->pp_scratch.ensure(cy_preprocess_chain_scratch_bytes(&c->chain, G, Ty, Tx)))) return rc;
    if ((rc = cy_preprocess_chain(&c->chain, c->band.p, row_stride, big_endian, x0, y0, G, Ty, Tx, c->cfg.imgsz, nullptr,
                                  c->model_in.p, nullptr, (int32_t*)c->status.p, c->pp_scratch.p, (uintptr_t)st)))
        return rc;
    const int bt = c->cfg.batch_tiles;
    for (int s = 0; s < G; s += bt) {
        const int B = std::min(bt, G - s);
        if ((rc = run_batch(c, (const char*)c->model_in.p + (size_t)s * Sh * Sw * 4 * 2, (const int32_t*)c->status.p + s,
                            ids_dev + s, B, Ty, Tx, Sh, Sw, lb)))
            return rc;
    }
    c->stats[1] += G;
    return CY_OK;
}

int compact_local(RunCtx* c) {
    int rc;
    if ((rc = c->total.ensure(32))) return rc;
    if ((rc = c->packed.ensure((size_t)c->T * kMaxDet * 32))) return rc;
    if ((rc = c->compact_scratch.ensure(cy_compact_scratch_bytes(c->T)))) return rc;
    return cy_compact_records((const cy_det_record*)c->rec_slots.p, (const int32_t*)c->nrec.p, c->T, kMaxDet,
                              (cy_det_record*)c->packed.p, (int32_t*)c->total.p, c->compact_scratch.p,
                              (uintptr_t)c->compute);
}

int run_local(RunCtx* c, const RowSource& src, int ny, int nx, int big_endian) {
    const cy_run_config& f = c->cfg;
    int rc;
    CY_CUDA_CHECK(cudaSetDevice(c->device));
    // ---- tile grid (utils.generate_tiles) over the requested region
    const int xmin = f.xmin >= 0 ? f.xmin : 0, xmax = f.xmax >= 0 ? f.xmax : nx - 1;
    const int ymin = f.ymin >= 0 ? f.ymin : 0, ymax = f.ymax >= 0 ? f.ymax : ny - 1;
    if (xmax >= nx || ymax >= ny || xmin > xmax || ymin > ymax) return set_error(CY_ERR_INVALID, "cy_run: region outside the image");
    c->tiles.clear();
    if (f.tile_x > 0 && f.tile_y > 0) {
        int nt = 0;
        if ((rc = cy_generate_tiles(xmin, xmax, ymin, ymax, f.tile_x, f.tile_y, f.step_x, f.step_y, nullptr, 0, &nt))) return rc;
        c->tiles.resize((size_t)nt);
        if ((rc = cy_generate_tiles(xmin, xmax, ymin, ymax, f.tile_x, f.tile_y, f.step_x, f.step_y, c->tiles.data(), nt, &nt)))
            return rc;
    } else {
        c->tiles.push_back(cy_tile{xmin, xmax + 1, ymin, ymax + 1});   // serial path: the whole region is one tile
    }
    c->T = (int)c->tiles.size();
    if (c->T == 0) return set_error(CY_ERR_INVALID, "cy_run: empty tile grid");
    split_rows(c->tiles, f.world, f.rank, &c->first, &c->last);
    if ((rc = c->tiles_dev.ensure(sizeof(cy_tile) * (size_t)c->T))) return rc;
    CY_CUDA_CHECK(cudaMemcpyAsync(c->tiles_dev.p, c->tiles.data(), sizeof(cy_tile) * (size_t)c->T, cudaMemcpyHostToDevice,
                                  c->compute));
    if ((rc = c->rec_slots.ensure((size_t)c->T * kMaxDet * 32))) return rc;
    if ((rc = c->nrec.ensure((size_t)c->T * 4))) return rc;
    CY_CUDA_CHECK(cudaMemsetAsync(c->nrec.p, 0, (size_t)c->T * 4, c->compute));
    c->stats[0] = c->T;
    c->stats[1] = 0;
    if (c->last > c->first) {
        int Y0 = 0x7fffffff, Y1 = 0;
        for (int i = c->first; i < c->last; ++i) {
            Y0 = std::min(Y0, c->tiles[i].ymin);
            Y1 = std::max(Y1, c->tiles[i].ymax);
        }
        const size_t row_bytes = (size_t)nx * 4;
        if ((rc = c->band.ensure((size_t)(Y1 - Y0) * row_bytes))) return rc;
        const int rpp = (int)std::max<size_t>(1, ((size_t)64 << 20) / row_bytes);   // rows per upload piece (~64 MB)
        if (!src.host_pinned) {
            const size_t need = (size_t)rpp * row_bytes;
            if (need > c->pin_cap) {
                for (int k = 0; k < 2; ++k) {
                    if (c->pin[k]) cudaFreeHost(c->pin[k]);
                    c->pin[k] = nullptr;
                    CY_CUDA_CHECK(cudaMallocHost(&c->pin[k], need));
                    c->pin_used[k] = false;
                }
                c->pin_cap = need;
            }
        }
        // earlier work on the compute stream may still read the band buffer
        cudaEvent_t ev0;
        CY_CUDA_CHECK(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
        CY_CUDA_CHECK(cudaEventRecord(ev0, c->compute));
        CY_CUDA_CHECK(cudaStreamWaitEvent(c->copy, ev0, 0));
        cudaEventDestroy(ev0);
        int row = Y0, npieces = 0;
        std::vector<int> piece_end;
        auto upload_piece = [&](int y_stop) -> int {
            const int r0 = row;
            int r1 = std::min(Y1, r0 + rpp);
            if (r0 < y_stop && y_stop < r1) r1 = y_stop;   // a piece ends where the waiting group's rows end
            const size_t nb = (size_t)(r1 - r0) * row_bytes;
            const char* hsrc;
            if (src.host_pinned) {
                hsrc = src.host + (size_t)r0 * row_bytes;
            } else {
                const int k = npieces & 1;
                if (c->pin_used[k]) CY_CUDA_CHECK(cudaEventSynchronize(c->pin_ev[k]));   // previous DMA out of this buffer
                if (src.fd >= 0) {
                    int r = read_range(src.fd, src.offset + (long long)r0 * (long long)row_bytes, (char*)c->pin[k], nb,
                                       c->cfg.read_threads);
                    if (r) return r;
                } else {
                    memcpy(c->pin[k], src.host + (size_t)r0 * row_bytes, nb);
                }
                hsrc = (const char*)c->pin[k];
            }
            CY_CUDA_CHECK(cudaMemcpyAsync((char*)c->band.p + (size_t)(r0 - Y0) * row_bytes, hsrc, nb, cudaMemcpyHostToDevice,
                                          c->copy));
            if (!src.host_pinned) {
                CY_CUDA_CHECK(cudaEventRecord(c->pin_ev[npieces & 1], c->copy));
                c->pin_used[npieces & 1] = true;
            }
            if ((int)c->piece_ev.size() <= npieces) {
                cudaEvent_t e;
                CY_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                c->piece_ev.push_back(e);
            }
            CY_CUDA_CHECK(cudaEventRecord(c->piece_ev[(size_t)npieces], c->copy));
            piece_end.push_back(r1);
            row = r1;
            ++npieces;
            c->stats[2] += (double)nb;
            return CY_OK;
        };
        auto ready = [&](int y_needed) -> int {     // rows [Y0, y_needed) must be in HBM before the group runs
            const int want = std::min(y_needed, Y1);
            while (row < want) {
                int r = upload_piece(want);
                if (r) return r;
            }
            for (int i = 0; i < npieces; ++i)
                if (piece_end[(size_t)i] >= want) {
                    CY_CUDA_CHECK(cudaStreamWaitEvent(c->compute, c->piece_ev[(size_t)i], 0));
                    break;
                }
            return CY_OK;
        };
        // groups by tile shape (edge tiles are smaller), larger shapes first, ids ascending inside a shape
        std::map<std::pair<int, int>, std::vector<int>, std::greater<std::pair<int, int>>> shapes;
        for (int i = c->first; i < c->last; ++i)
            shapes[{c->tiles[i].ymax - c->tiles[i].ymin, c->tiles[i].xmax - c->tiles[i].xmin}].push_back(i);
        const int bt = c->cfg.batch_tiles;
        for (auto& kv : shapes) {
            const std::vector<int>& ids = kv.second;
            const int n = (int)ids.size();
            // a small LEADING group (the first whole tile rows with >= 32 tiles) starts computing as soon as its rows
            // are in HBM; the upload of everything else overlaps it (pipeline.lead_group_size)
            int lead = 0;
            for (int k = 0; k < n;) {
                const int y = c->tiles[ids[k]].ymin;
                while (k < n && c->tiles[ids[k]].ymin == y) ++k;
                if (k >= 32) {
                    lead = k;
                    break;
                }
            }
            std::vector<int> starts;
            if (lead > 0 && 2 * lead <= n) {
                starts.push_back(0);
                for (int s = lead; s < n; s += bt) starts.push_back(s);
            } else {
                for (int s = 0; s < n; s += bt) starts.push_back(s);
            }
            for (size_t gi = 0; gi < starts.size(); ++gi) {
                const int s = starts[gi], e = gi + 1 < starts.size() ? starts[gi + 1] : n;
                std::vector<int> g(ids.begin() + s, ids.begin() + e);
                int ylast = 0;
                for (int id : g) ylast = std::max(ylast, c->tiles[id].ymax);
                if ((rc = ready(ylast))) return rc;
                if ((rc = run_group(c, nx, big_endian, 0, Y0, g, kv.first.first, kv.first.second))) return rc;
            }
        }
    }
    if ((rc = compact_local(c))) return rc;
    int32_t n = 0;
    CY_CUDA_CHECK(cudaMemcpyAsync(&n, c->total.p, 4, cudaMemcpyDeviceToHost, c->compute));
    CY_CUDA_CHECK(cudaStreamSynchronize(c->compute));
    c->n_local = n;
    return CY_OK;
}

int ensure_neighbors(RunCtx* c) {
    std::string key((const char*)c->tiles.data(), sizeof(cy_tile) * c->tiles.size());
    if (key == c->nb_key) return CY_OK;
    int rc, total = 0;
    std::vector<int> off((size_t)c->T + 1);
    if ((rc = cy_tile_neighbors(c->tiles.data(), c->T, off.data(), nullptr, 0, &total))) return rc;
    std::vector<int> idx((size_t)std::max(total, 1));
    if ((rc = cy_tile_neighbors(c->tiles.data(), c->T, off.data(), idx.data(), total, &total))) return rc;
    if ((rc = c->nb_off.ensure(off.size() * 4))) return rc;
    if ((rc = c->nb_idx.ensure(idx.size() * 4))) return rc;
    CY_CUDA_CHECK(cudaMemcpyAsync(c->nb_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, c->compute));
    CY_CUDA_CHECK(cudaMemcpyAsync(c->nb_idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, c->compute));
    CY_CUDA_CHECK(cudaStreamSynchronize(c->compute));
    c->nb_key.swap(key);
    return CY_OK;
}

}  // namespace

extern "C" int cy_ctx_create(void* model, const cy_pp_chain* chain, const cy_run_config* cfg, void** ctx_host) {
    if (!model || !cfg || !ctx_host) return set_error(CY_ERR_INVALID, "cy_ctx_create: null argument");
    int rc = cy_device_check();
    if (rc) return rc;
    Model* m = (Model*)model;
    if (!m->finalized) return set_error(CY_ERR_STATE, "cy_ctx_create: the model is not finalized");
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return set_error(CY_ERR_INVALID, "cy_ctx_create: bad rank / world");
    if (cfg->imgsz < 32 || cfg->imgsz % 32) return set_error(CY_ERR_INVALID, "cy_ctx_create: imgsz must be a positive multiple of 32");
    RunCtx* c = new RunCtx();
    c->model = m;
    c->cfg = *cfg;
    if (c->cfg.batch_tiles <= 0) c->cfg.batch_tiles = 296;
    if (c->cfg.read_threads <= 0) c->cfg.read_threads = 8;
    memset(&c->chain, 0, sizeof(c->chain));
    if (chain) c->chain = *chain;
    if ((rc = cy_pp_chain_validate(&c->chain))) {
        delete c;
        return rc;
    }
    c->chain.out_f16 = m->f16 ? 1 : 0;   // the model input format follows the model's storage format
    cudaGetDevice(&c->device);
    if (cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->pin_ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->pin_ev[1], cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return set_error(CY_ERR_CUDA, "cy_ctx_create: stream / event creation failed");
    }
    *ctx_host = c;
    return CY_OK;
}

extern "C" int cy_ctx_set_allgather(void* ctx, cy_allgather_fn fn, void* user) {
    if (!ctx) return set_error(CY_ERR_INVALID, "null context");
    ((RunCtx*)ctx)->gather = fn;
    ((RunCtx*)ctx)->gather_user = user;
    return CY_OK;
}

extern "C" int cy_ctx_destroy(void* ctx) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c) return CY_OK;
    cudaSetDevice(c->device);
    if (c->compute) cudaStreamSynchronize(c->compute);
    if (c->copy) cudaStreamSynchronize(c->copy);
    DevBuf* bufs[] = {&c->band, &c->meta, &c->model_in, &c->pp_scratch, &c->status, &c->post_scratch, &c->dets, &c->ndets,
                      &c->keep, &c->nkeep, &c->mstat, &c->lb, &c->tiles_dev, &c->rec_slots, &c->nrec, &c->packed,
                      &c->total, &c->compact_scratch, &c->nb_off, &c->nb_idx, &c->sources, &c->nout, &c->xsend,
                      &c->xrecv, &c->xall, &c->xtotal, &c->xcounts, &c->xscratch};
    for (DevBuf* b : bufs) b->release();
    for (int k = 0; k < 2; ++k) {
        if (c->pin[k]) cudaFreeHost(c->pin[k]);
        if (c->pin_ev[k]) cudaEventDestroy(c->pin_ev[k]);
    }
    for (cudaEvent_t e : c->piece_ev) cudaEventDestroy(e);
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->copy) cudaStreamDestroy(c->copy);
    delete c;
    return CY_OK;
}

extern "C" int cy_run_local(void* ctx, const char* fits_path, int* nrecords_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !fits_path) return set_error(CY_ERR_INVALID, "cy_run_local: null argument");
    const int fd = open(fits_path, O_RDONLY);
    if (fd < 0) return set_error(CY_ERR_INVALID, "cy_run_local: cannot open %s", fits_path);
    FitsInfo fi;
    int rc = parse_fits_header(fd, &fi);
    if (!rc) {
        struct stat sb;
        if (fstat(fd, &sb) == 0 && (long long)sb.st_size < fi.offset + (long long)fi.nx * fi.ny * 4)
            rc = set_error(CY_ERR_INVALID, "cy_run_local: FITS payload shorter than NAXIS1 x NAXIS2");
    }
    if (!rc) {
        RowSource src;
        src.fd = fd;
        src.offset = fi.offset;
        src.nx = fi.nx;
        rc = run_local(c, src, fi.ny, fi.nx, 1);
    }
    close(fd);
    if (!rc && nrecords_host) *nrecords_host = c->n_local;
    return rc;
}

extern "C" int cy_run_local_payload(void* ctx, const void* payload_host, int ny, int nx, int big_endian, int pinned,
                                    int* nrecords_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !payload_host || ny <= 0 || nx <= 0) return set_error(CY_ERR_INVALID, "cy_run_local_payload: bad argument");
    RowSource src;
    src.host = (const char*)payload_host;
    src.host_pinned = pinned ? 1 : 0;
    src.nx = nx;
    int rc = run_local(c, src, ny, nx, big_endian ? 1 : 0);
    if (!rc && nrecords_host) *nrecords_host = c->n_local;
    return rc;
}

extern "C" int cy_ctx_records(void* ctx, const cy_det_record** recs_dev, int* n_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !recs_dev || !n_host) return set_error(CY_ERR_INVALID, "cy_ctx_records: null argument");
    *recs_dev = (const cy_det_record*)c->packed.p;
    *n_host = c->n_local;
    return CY_OK;
}

extern "C" int cy_ctx_pack_slot(void* ctx, int cap, const void** slot_dev) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !slot_dev || cap < 0) return set_error(CY_ERR_INVALID, "cy_ctx_pack_slot: bad argument");
    if (!c->packed.p) return set_error(CY_ERR_STATE, "cy_ctx_pack_slot: no records (run cy_run_local first)");
    int rc;
    if ((rc = c->xsend.ensure(((size_t)cap + 1) * 32))) return rc;
    CY_CUDA_CHECK(cudaMemsetAsync(c->xsend.p, 0, 32, c->compute));
    CY_CUDA_CHECK(cudaMemcpyAsync(c->xsend.p, c->total.p, 4, cudaMemcpyDeviceToDevice, c->compute));
    const size_t nb = (size_t)std::min(cap, c->n_local) * 32;
    if (nb) CY_CUDA_CHECK(cudaMemcpyAsync((char*)c->xsend.p + 32, c->packed.p, nb, cudaMemcpyDeviceToDevice, c->compute));
    CY_CUDA_CHECK(cudaStreamSynchronize(c->compute));
    *slot_dev = c->xsend.p;
    return CY_OK;
}

extern "C" int cy_ctx_unpack_slots(void* ctx, const void* slots_dev, int world, int cap, const cy_det_record** recs_dev,
                                   int* n_host, int* max_count_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !slots_dev || world < 1 || cap < 0 || !recs_dev || !n_host)
        return set_error(CY_ERR_INVALID, "cy_ctx_unpack_slots: bad argument");
    int rc;
    std::vector<int32_t> counts((size_t)world);
    CY_CUDA_CHECK(cudaMemcpy2DAsync(counts.data(), 4, slots_dev, ((size_t)cap + 1) * 32, 4, (size_t)world,
                                    cudaMemcpyDeviceToHost, c->compute));
    CY_CUDA_CHECK(cudaStreamSynchronize(c->compute));
    int cmax = 0;
    long long tot = 0;
    std::vector<int32_t> clamped(counts);
    for (int r = 0; r < world; ++r) {
        if (counts[(size_t)r] < 0) return set_error(CY_ERR_INVALID, "cy_ctx_unpack_slots: negative record count in slot %d", r);
        cmax = std::max(cmax, counts[(size_t)r]);
        clamped[(size_t)r] = std::min(counts[(size_t)r], cap);
        tot += clamped[(size_t)r];
    }
    if (max_count_host) *max_count_host = cmax;
    if ((rc = c->xcounts.ensure((size_t)world * 4))) return rc;
    if ((rc = c->xall.ensure((size_t)std::max<long long>(1, (long long)world * cap) * 32))) return rc;
    if ((rc = c->xtotal.ensure(32))) return rc;
    if ((rc = c->xscratch.ensure(cy_compact_scratch_bytes(world)))) return rc;
    CY_CUDA_CHECK(cudaMemcpyAsync(c->xcounts.p, clamped.data(), (size_t)world * 4, cudaMemcpyHostToDevice, c->compute));
    if ((rc = cy_compact_records((const cy_det_record*)((const char*)slots_dev + 32), (const int32_t*)c->xcounts.p, world,
                                 cap + 1, (cy_det_record*)c->xall.p, (int32_t*)c->xtotal.p, c->xscratch.p,
                                 (uintptr_t)c->compute)))
        return rc;
    CY_CUDA_CHECK(cudaStreamSynchronize(c->compute));
    *recs_dev = (const cy_det_record*)c->xall.p;
    *n_host = (int)tot;
    return CY_OK;
}

extern "C" int cy_run_merge(void* ctx, const cy_det_record* recs_dev, int n, cy_source* sources_host, int capacity,
                            int* nsources_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || (n > 0 && !recs_dev) || !nsources_host || n < 0) return set_error(CY_ERR_INVALID, "cy_run_merge: bad argument");
    if (c->T == 0) return set_error(CY_ERR_STATE, "cy_run_merge: no tile grid (run cy_run_local first)");
    int rc;
    if ((rc = ensure_neighbors(c))) return rc;
    if ((rc = c->sources.ensure((size_t)std::max(n, 1) * sizeof(cy_source)))) return rc;
    if ((rc = c->nout.ensure(8))) return rc;
    CY_CUDA_CHECK(cudaMemsetAsync(c->nout.p, 0, 8, c->compute));
    if ((rc = cy_merge_global((cy_det_record*)recs_dev, n, (const cy_tile*)c->tiles_dev.p, c->T, (const int32_t*)c->nb_off.p,
                              (const int32_t*)c->nb_idx.p, (cy_source*)c->sources.p, (int64_t*)c->nout.p,
                              (uintptr_t)c->compute)))
        return rc;
    long long k = 0;
    CY_CUDA_CHECK(cudaMemcpyAsync(&k, c->nout.p, 8, cudaMemcpyDeviceToHost, c->compute));
    CY_CUDA_CHECK(cudaStreamSynchronize(c->compute));
    *nsources_host = (int)k;
    if (sources_host) {
        if (k > capacity) return set_error(CY_ERR_INVALID, "cy_run_merge: %lld sources do not fit capacity %d", k, capacity);
        if (k) CY_CUDA_CHECK(cudaMemcpy(sources_host, c->sources.p, (size_t)k * sizeof(cy_source), cudaMemcpyDeviceToHost));
    }
    return CY_OK;
}

static int exchange_and_merge(RunCtx* c, cy_source* sources_host, int capacity, int* nsources_host, int* nrecords_host) {
    int rc;
    const cy_det_record* recs = (const cy_det_record*)c->packed.p;
    int n = c->n_local;
    if (c->cfg.world > 1) {
        if (!c->gather)
            return set_error(CY_ERR_STATE, "cy_run_mosaic: world > 1 needs an all-gather (cy_ctx_set_allgather), or drive "
                                           "cy_run_local / cy_ctx_pack_slot / cy_ctx_unpack_slots / cy_run_merge yourself");
        int cap = 48 * ((c->T + c->cfg.world - 1) / c->cfg.world) + 1024;   // pipeline.Engine._exchange_cap
        for (;;) {
            const void* slot = nullptr;
            if ((rc = cy_ctx_pack_slot(c, cap, &slot))) return rc;
            const size_t per = ((size_t)cap + 1) * 32;
            if ((rc = c->xrecv.ensure(per * (size_t)c->cfg.world))) return rc;
            if (c->gather(c->gather_user, slot, c->xrecv.p, per, (uintptr_t)c->compute))
                return set_error(CY_ERR_STATE, "cy_run_mosaic: the all-gather callback failed");
            int cmax = 0;
            if ((rc = cy_ctx_unpack_slots(c, c->xrecv.p, c->cfg.world, cap, &recs, &n, &cmax))) return rc;
            if (cmax <= cap) break;
            cap = (int)(cmax * 1.25) + 1;   // a rank overflowed its slot: grow and redo (every rank sees the same cmax)
        }
    }
    if (nrecords_host) *nrecords_host = n;
    return cy_run_merge(c, recs, n, sources_host, capacity, nsources_host);
}

extern "C" int cy_run_mosaic(void* ctx, const char* fits_path, cy_source* sources_host, int capacity, int* nsources_host,
                             int* nrecords_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !nsources_host) return set_error(CY_ERR_INVALID, "cy_run_mosaic: null argument");
    int rc = cy_run_local(ctx, fits_path, nullptr);
    if (rc) return rc;
    return exchange_and_merge(c, sources_host, capacity, nsources_host, nrecords_host);
}

extern "C" int cy_run_payload(void* ctx, const void* payload_host, int ny, int nx, int big_endian, int pinned,
                              cy_source* sources_host, int capacity, int* nsources_host, int* nrecords_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !nsources_host) return set_error(CY_ERR_INVALID, "cy_run_payload: null argument");
    int rc = cy_run_local_payload(ctx, payload_host, ny, nx, big_endian, pinned, nullptr);
    if (rc) return rc;
    return exchange_and_merge(c, sources_host, capacity, nsources_host, nrecords_host);
}

extern "C" int cy_ctx_info(void* ctx, double* info_host) {
    RunCtx* c = (RunCtx*)ctx;
    if (!c || !info_host) return set_error(CY_ERR_INVALID, "cy_ctx_info: null argument");
    info_host[0] = c->T;           // tiles of the grid
    info_host[1] = c->stats[1];    // tiles this rank processed in the last run
    info_host[2] = c->stats[2];    // payload bytes uploaded so far
    info_host[3] = c->first;
    info_host[4] = c->last;
    info_host[5] = c->n_local;
    info_host[6] = info_host[7] = 0;
    return CY_OK;
}
