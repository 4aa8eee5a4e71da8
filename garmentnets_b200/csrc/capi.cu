// Library-level entry points: version, last error, device query.
#include "common.cuh"
#include <string.h>

namespace gnb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
    return n;
}

}  // namespace gnb

extern "C" {

int32_t gnb_version(void) { return 100; /* 0.1.0 */ }

const char* gnb_last_error(void) { return gnb::g_err; }

int32_t gnb_device_sm_count(void) {
    int dev = 0, n = 0;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

}  // extern "C"
