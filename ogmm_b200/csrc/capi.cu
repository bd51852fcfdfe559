// C-ABI housekeeping: version, thread-local last error, device info.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ogmm {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return OGMM_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return OGMM_ECUDA;
}

}  // namespace ogmm

extern "C" __attribute__((visibility("default"))) int ogmm_version(void) { return OGMM_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) const char* ogmm_last_error(void) { return ogmm::g_last_error; }

extern "C" __attribute__((visibility("default"))) int ogmm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    int st = ogmm::cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (st != OGMM_OK) return st;
    int v = 0;
    if (sm_count) {
        st = ogmm::cuda_status(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
        if (st != OGMM_OK) return st;
        *sm_count = v;
    }
    if (cc_major) {
        st = ogmm::cuda_status(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev), "cudaDeviceGetAttribute");
        if (st != OGMM_OK) return st;
        *cc_major = v;
    }
    if (cc_minor) {
        st = ogmm::cuda_status(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev), "cudaDeviceGetAttribute");
        if (st != OGMM_OK) return st;
        *cc_minor = v;
    }
    return OGMM_OK;
}
