// TEMPORARY stub (replaced by the real chain kernels in the next commit).
#include "common.h"
extern "C" size_t cy_preprocess_scratch_bytes(const cy_pp_config*, int, int, int) { return 256; }
extern "C" int cy_preprocess(const cy_pp_config*, const void*, long long, int, const int32_t*, const int32_t*, int, int,
                             int, int, float*, void*, float*, int32_t*, void*, uintptr_t) {
    return cy::set_error(CY_ERR_STATE, "cy_preprocess not built yet");
}
