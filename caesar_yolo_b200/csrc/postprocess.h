#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "model.h"
#include "../../include/caesar_b200.h"

namespace cy {

typedef cy_letterbox LetterboxInfo;

int num_anchors(int Sh, int Sw);
int decode_pred(const float* h0, const float* h1, const float* h2, int B, int Sh, int Sw, int nc, float* pred,
                cudaStream_t st);
size_t postprocess_scratch_bytes(int B, int Sh, int Sw, int max_det);
int postprocess(const float* h0, const float* h1, const float* h2, int B, int Sh, int Sw, int nc, float conf,
                float iou, int max_det, const LetterboxInfo* lb, float* dets, int* ndets, void* scratch,
                cudaStream_t st);
size_t nms_scratch_bytes(int B, int N);
int nms_batched(const float* boxes, const float* scores, const int* counts, int B, int N, double thr, int max_keep,
                long long* keep, int* nkeep, void* scratch, cudaStream_t st);
int merge_tiles(const float* dets, const int* ndets, int B, int det_stride, float thr_score, float thr_soft,
                float thr_hard, const int* pre_status, int* keep_idx, int* nkeep, int* status, cudaStream_t st);

}  // namespace cy
