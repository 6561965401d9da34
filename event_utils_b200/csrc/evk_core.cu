// evk_core.cu -- error reporting, version and device checks of libevk.so.
#include <stdarg.h>
#include <stdio.h>

#include "evk_common.cuh"

namespace evk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return EVK_E_CUDA;
}

int num_sms()
{
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
            cached_sms = sms;
            cached_dev = dev;
        }
    }
    return cached_sms;
}

}  // namespace evk

extern "C" {

int evk_version(void) { return EVK_VERSION; }

const char *evk_last_error(void) { return evk::g_err; }

int evk_device_check(void)
{
    int dev = 0, major = 0;
    EVK_CUDA(cudaGetDevice(&dev));
    EVK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        evk::set_error("libevk is built for sm_100a only; device %d has compute capability %d.x", dev, major);
        return EVK_E_DEVICE;
    }
    return EVK_OK;
}

}  // extern "C"
