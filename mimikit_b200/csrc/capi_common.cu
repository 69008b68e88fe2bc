// Error plumbing and device queries of the C ABI.
#include "common.cuh"
#include "../../include/mmk_b200.h"

namespace mmk {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const char* file, int line, const std::string& msg) {
    const char* base = file;
    for (const char* p = file; *p; ++p) if (*p == '/') base = p + 1;
    g_last_error = std::string(base) + ":" + std::to_string(line) + ": " + msg;
    return 1;
}
}  // namespace mmk

extern "C" int mmk_abi_version(void) { return MMK_ABI_VERSION; }
extern "C" const char* mmk_last_error(void) { return mmk::g_last_error.c_str(); }

extern "C" int mmk_get_device_info(mmk_device_info* out) {
    MMK_CHECK(out != nullptr, "mmk_get_device_info: null out");
    int dev = 0;
    MMK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    MMK_CUDA(cudaGetDeviceProperties(&p, dev));
    out->sm_count = p.multiProcessorCount;
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->max_smem_optin = (int)p.sharedMemPerBlockOptin;
    out->l2_bytes = p.l2CacheSize;
    return 0;
}
