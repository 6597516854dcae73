#pragma once
#include <cuda_runtime.h>
#include "../../include/caesar_b200.h"

namespace cy {
int make_records(const float* dets, const int* keep_idx, const int* nkeep, const int* status, int det_stride,
                 const cy_tile* tiles, const int* tile_ids, int B, cy_det_record* recs, int* nrec, cudaStream_t st);
size_t compact_scratch_bytes(int T);
int compact_records(const cy_det_record* slots, const int* counts, int T, int slot_stride, cy_det_record* out,
                    int* total, void* scratch, cudaStream_t st);
int merge_global(cy_det_record* recs, int n, const cy_tile* tiles, int T, const int* nb_off, const int* nb_idx,
                 cy_source* out, long long* nout, cudaStream_t st);
}  // namespace cy
