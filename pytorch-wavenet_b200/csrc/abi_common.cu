// abi_common.cu -- version, error string, device info.
#include "common.cuh"
#include <cstring>

namespace wn {
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace wn

extern "C" int wn_version(void) { return WN_ABI_VERSION; }

extern "C" const char* wn_last_error_string(void) { return wn::err_buf(); }

extern "C" int wn_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin, int* l2_bytes) {
    int dev = 0;
    WN_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    WN_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = (int)p.sharedMemPerBlockOptin;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    return 0;
}
