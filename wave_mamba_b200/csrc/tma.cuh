// Tensor-map (TMA) helpers shared by the kernels that stage NCHW halo tiles with cp.async.bulk.tensor:
// the driver entry point for cuTensorMapEncodeTiled (no -lcuda: resolved through the runtime), a 4-D map
// over a (B, C, h, w) fp32 tensor, and the device-side box load.
#pragma once
#include <cuda.h>

#include "tc5_common.cuh"

namespace wm {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// NCHW fp32 tensor (B, C, h, w) as a 4-D tensor map {w, h, C, B} with a box_w x box_h x box_c x 1 box; elements
// outside the tensor read as zero.  Needs a 16-byte aligned base and w % 4 == 0 (strides in 16-byte units).
inline bool make_tmap_nchw(CUtensorMap *tm, const float *x, int64_t B, int64_t C, int64_t h, int64_t w,
                           uint32_t box_w, uint32_t box_h, uint32_t box_c)
{
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)C * h * w * 4};
    const cuuint32_t box[4] = {box_w, box_h, box_c, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS;
}

// The innermost start coordinate c0 must be a multiple of 4 elements (16 bytes): measured with
// tools/probes/tma_probe.cu, anything else raises an illegal-instruction fault.  dst: 128-byte aligned.
__device__ __forceinline__ void load_box(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                         uint32_t mbar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(mbar)
        : "memory");
}

}  // namespace tma
}  // namespace wm
