// Error reporting, version and device queries of the C-ABI.
#include <cstdarg>
#include <cstring>
#include "common.cuh"

static thread_local char g_err[512] = "";

void pn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int pn_sm_count_cached() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

extern "C" const char *pn_last_error(void) { return g_err; }
extern "C" int pn_version(void) { return 100; }
extern "C" int pn_device_sm_count(int *sm_count) {
    PN_REQUIRE(sm_count, "null pointer");
    int dev = 0;
    PN_CUDA(cudaGetDevice(&dev));
    PN_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return PN_OK;
}
