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

// One 32-bit word per device in which the mbarrier pipelines (conv3x3_tc5.cu, pw_dw_tc5.cu) record a
// wait that timed out; zero-initialised, allocated on first use, never freed.
unsigned int *pipeline_err_word()
{
    static thread_local int cached_dev = -1;
    static thread_local unsigned int *cached = nullptr;
    static unsigned int *per_dev[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (dev == cached_dev) return cached;
    if (per_dev[dev] == nullptr) {
        unsigned int *p = nullptr;
        if (cudaMalloc(&p, 16) != cudaSuccess || cudaMemset(p, 0, 16) != cudaSuccess) return nullptr;
        per_dev[dev] = p;
    }
    cached_dev = dev;
    cached = per_dev[dev];
    return cached;
}

}  // namespace wm

/* Developer aid: synchronises the device and returns (and clears) the pipeline error word:
 * 0 = every mbarrier wait of the tcgen05 kernels completed; otherwise
 * 0x80000000 | role << 24 | barrier << 16 | iteration of the first wait that timed out. */
extern "C" int wm_debug_pipeline_error(unsigned int *out)
{
    if (out == nullptr) return WM_EINVAL;
    unsigned int *w = wm::pipeline_err_word();
    if (w == nullptr) { wm::set_error("wm_debug_pipeline_error: no device"); return WM_ENODEVICE; }
    WM_CUDA_OK(cudaDeviceSynchronize());
    WM_CUDA_OK(cudaMemcpy(out, w, 4, cudaMemcpyDeviceToHost));
    WM_CUDA_OK(cudaMemset(w, 0, 4));
    return WM_OK;
}

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
