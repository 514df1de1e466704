// Shared helpers for the wavemamba_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/wavemamba_b200.h"

namespace wm {

// Thread-local message for wm_last_error().
void set_error(const char *fmt, ...);

#define WM_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            wm::set_error(__VA_ARGS__); \
            return WM_EINVAL;          \
        }                              \
    } while (0)

#define WM_CUDA_OK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            wm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                          __FILE__, __LINE__);                                        \
            return WM_ECUDA;                                                          \
        }                                                                             \
    } while (0)

#define WM_LAUNCH_OK(what)                                                            \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            wm::set_error("launch of %s failed: %s", what, cudaGetErrorString(_e));   \
            return WM_ECUDA;                                                          \
        }                                                                             \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

// Number of SMs of the current device (cached per process; B200 = 148).
int sm_count();
// Device word in which the mbarrier pipelines record a timed-out wait (abi.cu).
unsigned int *pipeline_err_word();

// ---- device-side load/store helpers -------------------------------------------------
// Streaming 128-bit load: read-only path, do not allocate in L1 (data is touched once).
__device__ __forceinline__ float4 ld_stream4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// Streaming 128-bit store (evict-first in L2: consumers re-read it much later or never).
__device__ __forceinline__ void st_stream4(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
// 256-bit forms (sm_100: LDG.E.256 / STG.E.256; 32-byte aligned): half the LSU instructions of a
// streaming kernel whose warps otherwise stall on the load/store queue (lg_throttle).
struct float8 { float4 lo, hi; };
__device__ __forceinline__ float8 ld_stream8(const float *p)
{
    float8 v;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y),
                   "=f"(v.hi.z), "=f"(v.hi.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream8(float *p, float4 a, float4 b)
{
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y),
                 "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace wm
