// ABI housekeeping: version, per-thread error string, device check.
#include <stdarg.h>

#include "common.cuh"

namespace wm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

}  // namespace wm

extern "C" int wm_abi_version(void) { return WM_ABI_VERSION; }

extern "C" const char *wm_last_error(void) { return wm::g_err; }

extern "C" int wm_device_check(void)
{
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        wm::set_error("no CUDA device is available; wavemamba_b200 has no CPU fallback");
        return WM_ENODEVICE;
    }
    if (major != 10) {
        wm::set_error("device compute capability %d.x is not sm_100-class; this library is built "
                      "for sm_100a only", major);
        return WM_ENODEVICE;
    }
    return WM_OK;
}
