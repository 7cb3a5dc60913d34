// Error reporting, version and device probe of libcopo_b200.so.
#include <stdarg.h>
#include <stdio.h>
#include "b2c_internal.h"

static thread_local char g_err[512] = "";

int b2c_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" {
const char* b2c_last_error(void) { return g_err; }
int b2c_version(void) { return 100; }
int b2c_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return major == 10 ? 1 : 0;
}
}
