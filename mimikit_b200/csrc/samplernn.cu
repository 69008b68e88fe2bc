// SampleRNN generation (sm_100a) — placeholder entry points until the persistent kernel lands.
#include "common.cuh"
#include "../../include/mmk_b200.h"

extern "C" int mmk_samplernn_create(const mmk_samplernn_desc*, int, mmk_samplernn_t*) {
    MMK_FAIL("mmk_samplernn_create: not implemented yet");
}
extern "C" int mmk_samplernn_destroy(mmk_samplernn_t) { return 0; }
extern "C" int mmk_samplernn_launch_info(mmk_samplernn_t, mmk_launch_info*) {
    MMK_FAIL("mmk_samplernn_launch_info: not implemented yet");
}
extern "C" int mmk_samplernn_run(mmk_samplernn_t, int64_t*, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                                 int64_t, int, int, const float*, int, const float*, int64_t, int64_t, float*,
                                 int64_t*, unsigned long long*, void*) {
    MMK_FAIL("mmk_samplernn_run: not implemented yet");
}
extern "C" int mmk_samplernn_generate(mmk_samplernn_t, int64_t*, int, int64_t, int64_t, int64_t, const float*, int,
                                      const float*, float*, unsigned long long*, void*) {
    MMK_FAIL("mmk_samplernn_generate: not implemented yet");
}
