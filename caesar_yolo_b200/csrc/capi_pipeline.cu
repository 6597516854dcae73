// C-ABI entry points for the model, detect/NMS and merge stages (see include/caesar_b200.h).
#include "../../include/caesar_b200.h"
#include "common.h"
#include "merge_global.h"
#include "model.h"
#include "postprocess.h"
#include <vector>

using namespace cy;

#define CY_LAUNCH_CHECK(expr, what)                                                                      \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r > 0)                                                                                      \
            return set_error(CY_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString((cudaError_t)_r));   \
        if (_r < 0) return set_error(_r == -3 ? CY_ERR_NOMEM : CY_ERR_INVALID, "%s: invalid arguments or out of memory", what); \
    } while (0)

extern "C" int cy_model_create(const char* variant, int nc, void** model_host) {
    if (!variant || !model_host) return set_error(CY_ERR_INVALID, "null argument");
    Model* m = new Model();
    int r = m->init(variant, nc);
    if (r) {
        delete m;
        return r;
    }
    *model_host = m;
    return CY_OK;
}
extern "C" int cy_model_set_tensor(void* model, const char* name, const float* data_host, long long numel) {
    if (!model || !name || !data_host) return set_error(CY_ERR_INVALID, "null argument");
    return ((Model*)model)->set_tensor(name, data_host, numel);
}
extern "C" int cy_model_set_precision(void* model, int fp16) {
    if (!model) return set_error(CY_ERR_INVALID, "null model");
    Model* m = (Model*)model;
    if (m->finalized) return set_error(CY_ERR_STATE, "cy_model_set_precision must be called before cy_model_finalize");
    m->f16 = fp16 ? 1 : 0;
    return CY_OK;
}
extern "C" int cy_model_plan_summary(void* model, int B, int Sh, int Sw, int* info_host) {
    if (!model || !info_host) return set_error(CY_ERR_INVALID, "null argument");
    Plan* pl = nullptr;
    int r = ((Model*)model)->get_plan(B, Sh, Sw, &pl);
    if (r) return r;
    for (int i = 0; i < 8; ++i) info_host[i] = 0;
    for (const Op& op : pl->ops)
        if (op.type == Op::CONV) {
            const ConvKParams& kp = op.conv.kp;
            info_host[0] += 1;
            info_host[1] += kp.pair ? 1 : 0;
            if (kp.mode >= 0 && kp.mode <= 3) info_host[2 + kp.mode] += 1;
            info_host[6] += kp.block_n > 128 ? 1 : 0;
            info_host[7] += kp.halves == 2 ? 1 : 0;
        }
    return CY_OK;
}
extern "C" int cy_model_finalize(void* model) {
    if (!model) return set_error(CY_ERR_INVALID, "null model");
    return ((Model*)model)->finalize();
}
extern "C" int cy_model_forward(void* model, const void* in, int B, int Sh, int Sw, const float** heads_host,
                                uintptr_t stream) {
    if (!model || !in) return set_error(CY_ERR_INVALID, "null argument");
    Plan* pl = nullptr;
    int r = ((Model*)model)->forward(in, B, Sh, Sw, (cudaStream_t)stream, &pl);
    if (r) return r;
    if (heads_host)
        for (int l = 0; l < 3; ++l) heads_host[l] = (const float*)pl->head[l].p;
    return CY_OK;
}
extern "C" int cy_model_info(void* model, int B, int Sh, int Sw, double* info_host) {
    if (!model || !info_host) return set_error(CY_ERR_INVALID, "null argument");
    Model* m = (Model*)model;
    Plan* pl = nullptr;
    int r = m->get_plan(B, Sh, Sw, &pl);
    if (r) return r;
    int nconv = 0;
    for (const Op& op : pl->ops) nconv += (op.type == Op::CONV || op.type == Op::STEM);
    info_host[0] = (double)m->nparams;
    info_host[1] = pl->flops;
    info_host[2] = (double)pl->ops.size();
    info_host[3] = (double)pl->bytes;
    info_host[4] = m->c3;
    info_host[5] = m->c4;
    info_host[6] = m->c5;
    info_host[7] = nconv;
    return CY_OK;
}
extern "C" int cy_model_profile(void* model, const void* in, int B, int Sh, int Sw, int cap, const char** names_host,
                                float* ms_host, double* flops_host, int* nops_host, uintptr_t stream) {
    if (!model || !in) return set_error(CY_ERR_INVALID, "null argument");
    return ((Model*)model)->profile(in, B, Sh, Sw, cap, names_host, ms_host, flops_host, nops_host,
                                    (cudaStream_t)stream);
}
extern "C" int cy_model_conv_bytes(void* model, int B, int Sh, int Sw, double* bytes_host, int* nconv_host) {
    if (!model || !bytes_host) return set_error(CY_ERR_INVALID, "null argument");
    Plan* pl = nullptr;
    int r = ((Model*)model)->get_plan(B, Sh, Sw, &pl);
    if (r) return r;
    double tot = 0;
    int n = 0;
    for (const Op& op : pl->ops)
        if (op.type == Op::CONV) {
            tot += op.conv.bytes;
            ++n;
        }
    *bytes_host = tot;
    if (nconv_host) *nconv_host = n;
    return CY_OK;
}
extern "C" int cy_model_destroy(void* model) {
    delete (Model*)model;
    return CY_OK;
}

extern "C" int cy_num_anchors(int Sh, int Sw) { return num_anchors(Sh, Sw); }

extern "C" int cy_decode_pred(const float* h0, const float* h1, const float* h2, int B, int Sh, int Sw, int nc,
                              float* pred, uintptr_t stream) {
    if (nc < 1 || nc > kHeadC - 64) return set_error(CY_ERR_INVALID, "nc out of range");
    CY_LAUNCH_CHECK(decode_pred(h0, h1, h2, B, Sh, Sw, nc, pred, (cudaStream_t)stream), "decode_pred");
    return CY_OK;
}
extern "C" size_t cy_postprocess_scratch_bytes(int B, int Sh, int Sw, int max_det) {
    return postprocess_scratch_bytes(B, Sh, Sw, max_det);
}
extern "C" int cy_postprocess(const float* h0, const float* h1, const float* h2, int B, int Sh, int Sw, int nc,
                              float conf, float iou, int max_det, const cy_letterbox* lb, float* dets, int32_t* ndets,
                              void* scratch, uintptr_t stream) {
    if (nc < 1 || nc > kHeadC - 64) return set_error(CY_ERR_INVALID, "nc out of range");
    if (max_det < 1 || max_det > 320) return set_error(CY_ERR_INVALID, "max_det must be in [1,320]");
    CY_LAUNCH_CHECK(postprocess(h0, h1, h2, B, Sh, Sw, nc, conf, iou, max_det, lb, dets, ndets, scratch,
                                (cudaStream_t)stream),
                    "postprocess");
    return CY_OK;
}
extern "C" size_t cy_nms_scratch_bytes(int B, int N) { return nms_scratch_bytes(B, N); }
extern "C" int cy_nms_batched(const float* boxes, const float* scores, const int32_t* counts, int B, int N,
                              double iou_thr, int max_keep, int64_t* keep, int32_t* nkeep, void* scratch,
                              uintptr_t stream) {
    if (B <= 0 || N <= 0) return set_error(CY_ERR_INVALID, "B and N must be positive");
    CY_LAUNCH_CHECK(nms_batched(boxes, scores, counts, B, N, iou_thr, max_keep, (long long*)keep, nkeep, scratch,
                                (cudaStream_t)stream),
                    "nms_batched");
    return CY_OK;
}
extern "C" int cy_merge_tile(const float* dets, const int32_t* ndets, int B, int det_stride, float thr_score,
                             float thr_soft, float thr_hard, const int32_t* pre_status, int32_t* keep_idx,
                             int32_t* nkeep, int32_t* status, uintptr_t stream) {
    CY_LAUNCH_CHECK(merge_tiles(dets, ndets, B, det_stride, thr_score, thr_soft, thr_hard, pre_status, keep_idx, nkeep, status,
                                (cudaStream_t)stream),
                    "merge_tile");
    return CY_OK;
}
extern "C" int cy_make_records(const float* dets, const int32_t* keep_idx, const int32_t* nkeep, const int32_t* status,
                               int det_stride, const cy_tile* tiles, const int32_t* tile_ids, int B,
                               cy_det_record* recs, int32_t* nrec, uintptr_t stream) {
    CY_LAUNCH_CHECK(make_records(dets, keep_idx, nkeep, status, det_stride, tiles, tile_ids, B, recs, nrec,
                                 (cudaStream_t)stream),
                    "make_records");
    return CY_OK;
}
extern "C" size_t cy_compact_scratch_bytes(int T) { return compact_scratch_bytes(T); }
extern "C" int cy_compact_records(const cy_det_record* slots, const int32_t* counts, int T, int slot_stride,
                                  cy_det_record* out, int32_t* total, void* scratch, uintptr_t stream) {
    CY_LAUNCH_CHECK(compact_records(slots, counts, T, slot_stride, out, total, scratch, (cudaStream_t)stream),
                    "compact_records");
    return CY_OK;
}
extern "C" int cy_merge_global(cy_det_record* recs, int n, const cy_tile* tiles, int T, const int32_t* nb_off,
                               const int32_t* nb_idx, cy_source* out, int64_t* nout, uintptr_t stream) {
    CY_LAUNCH_CHECK(merge_global(recs, n, tiles, T, nb_off, nb_idx, out, (long long*)nout, (cudaStream_t)stream),
                    "merge_global");
    return CY_OK;
}
