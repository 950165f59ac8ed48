// Error reporting + device probing for libhm_b200.so
#include <stdarg.h>
#include <string.h>
#include "hm_common.cuh"

static thread_local char g_err[512] = "ok";

void hm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hm_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] > 0) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
    return n;
}

extern "C" const char* hm_last_error(void) { return g_err; }

extern "C" int hm_device_sm_count(void) {
    int dev = 0, n = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        hm_set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        hm_set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    return n;
}

extern "C" int hm_version(void) { return 100; }
