// The LFSSBlock.ffn tail  y = residual * res_scale + W (gelu(x[:, :32]) * x[:, 32:]) + b  (reference
// wavemamba_arch.py:227-230, :526; x is the (B,64,h,w) output of conv1+conv2) as a persistent,
// warp-specialised, TMA-fed pipeline (one CTA of 256 threads per SM).  A tile is 128 consecutive pixels of
// one image; its input box (128 px x 64 channels, 32 KB) arrives by cp.async.bulk.tensor three tiles
// ahead.  The register-staged kernel in pixelwise.cu is bound by global-load latency at 24 % occupancy:
// 0.285 -> 0.218 ms at the 4K level-1 size (57 -> 74 % of the HBM peak).
//   warps 0-3  thread = pixel: gelu(a) * b on the packed FP32 pipe -> v[tile & 1][32][128]
//   warps 4-7  thread = (4 adjacent pixels, 8 of the 32 outputs): per input channel one LDS.128 of pixels and
//              two of weights feed 16 FFMA2; residual quads straight from global memory, fetched before
//              the FMAs; 16-byte stores
// Needs hw % 4 == 0 and 16-byte aligned tensors; wm_pw_fwd falls back to the register-staged kernels
// otherwise.  (Measured, not kept: the plain 32 -> 32 form with per-image weights -- the CMTAttention tail --
// through the same pipeline without the first warp group: 0.193 ms against 0.156 ms of the two-pixel
// register-staged kernel, which already runs at 78 % of the peak.)
#include "tma.cuh"

namespace wm {
namespace pwt {

using namespace wm::tc5;

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx_ftz(float v)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Exact (erf) GELU of two values; see spatial32.cu (erf(t) = 1 - 2^(-t P(t)), degree-7 fit, |error| of the
// GELU <= 5e-7).
__device__ __forceinline__ f32x2 gelu2(f32x2 v)
{
    float x0, x1;
    unpack2(v, x0, x1);
    const float t0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.0f);
    const float t1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.0f);
    const f32x2 t = pack2(t0, t1);
    f32x2 p = ffma2(pack2(4.535860352916643e-05f, 4.535860352916643e-05f), t,
                    pack2(-0.00044550769962370396f, -0.00044550769962370396f));
    p = ffma2(p, t, pack2(0.001489441841840744f, 0.001489441841840744f));
    p = ffma2(p, t, pack2(0.0007746291812509298f, 0.0007746291812509298f));
    p = ffma2(p, t, pack2(-0.02825368009507656f, -0.02825368009507656f));
    p = ffma2(p, t, pack2(0.1484816074371338f, 0.1484816074371338f));
    p = ffma2(p, t, pack2(0.9184163808822632f, 0.9184163808822632f));
    p = ffma2(p, t, pack2(1.6279085874557495f, 1.6279085874557495f));
    float g0, g1;
    unpack2(fmul2(p, t), g0, g1);
    const float r0 = copysignf(1.0f - ex2_approx_ftz(-g0), x0), r1 = copysignf(1.0f - ex2_approx_ftz(-g1), x1);
    const float h0 = 0.5f * x0, h1 = 0.5f * x1;
    return pack2(fmaf(h0, r0, h0), fmaf(h1, r1, h1));
}

constexpr int kTP = 128;                      // pixels per tile
constexpr int kC = 32;
constexpr int kStages = 3;
constexpr int kThreadsA = 128, kThreadsB = 128, kThreads = kThreadsA + kThreadsB;

struct G {
    static constexpr int kRows = 64;                                   // channels of the input box
    static constexpr uint32_t kStageBytes = kRows * kTP * 4;           // 32 KB
    static constexpr uint32_t kVBytes = 2 * kC * kTP * 4;              // v[2][32][128]
    static constexpr size_t kSmem = kStages * kStageBytes + kVBytes + sizeof(float) * (kC * kC + 2 * kC) + kStages * 8;
};

__global__ void __launch_bounds__(kThreads, 1)
pw32_gate_tma_kernel(const __grid_constant__ CUtensorMap xmap, const float *__restrict__ w,
                     const float *__restrict__ bias, const float *__restrict__ residual,
                     const float *__restrict__ res_scale, float *__restrict__ y, int64_t hw, int tiles_per_img,
                     int total_tiles)
{
    constexpr int kBarVFull = 1, kBarVEmpty = 3, kBarA = 5;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stage = reinterpret_cast<float *>(smem_raw);                                   // [3][rows][128]
    float *vbuf = reinterpret_cast<float *>(smem_raw + kStages * G::kStageBytes);         // [2][32][128] (GATE)
    float *wt = reinterpret_cast<float *>(smem_raw + kStages * G::kStageBytes + G::kVBytes);   // [32 ci][32 co]
    float *pb = wt + kC * kC, *rs = pb + kC;
    const uint32_t bar0 = smem_u32(rs + kC);
    const int tid = threadIdx.x;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto issue = [&](int j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int b = tile / tiles_per_img, px0 = (tile - b * tiles_per_img) * kTP;
        const int sg = j % kStages;
        const uint32_t bar = bar0 + 8u * (uint32_t)sg;
        mbar_expect_tx(bar, G::kStageBytes);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(smem_u32(stage) + (uint32_t)sg * G::kStageBytes), "l"(reinterpret_cast<uint64_t>(&xmap)), "r"(px0),
            "r"(b * G::kRows), "r"(bar)
            : "memory");
    };
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(bar0 + 8u * (uint32_t)i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < kStages; ++i)
            if (my_tiles > i) issue(i);
    }
    if (tid < kC) {
        pb[tid] = bias ? __ldg(bias + tid) : 0.0f;
        rs[tid] = res_scale ? __ldg(res_scale + tid) : 1.0f;
    }
    for (int i = tid; i < kC * kC; i += kThreads) {
        const int co = i / kC, ci = i - co * kC;
        wt[ci * kC + co] = __ldg(w + i);
    }
    __syncthreads();

    if (tid < kThreadsA) {
        // =========================== gelu(a) * b ================================================
        const int px = tid;                                // lane = pixel: conflict-free rows of 128 floats
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int sg = j % kStages, buf = j & 1;
            mbar_wait(bar0 + 8u * (uint32_t)sg, (uint32_t)(j / kStages) & 1u);
            const float *st = stage + sg * (G::kStageBytes / 4) + px;
            float g[kC];
#pragma unroll
            for (int i = 0; i < kC; i += 2) {
                float g0, g1;
                unpack2(gelu2(pack2(st[i * kTP], st[(i + 1) * kTP])), g0, g1);
                g[i] = g0 * st[(kC + i) * kTP];
                g[i + 1] = g1 * st[(kC + i + 1) * kTP];
            }
            bar_sync(kBarA, kThreadsA);                    // every thread of the group has left the stage
            if (tid == 0 && j + kStages < my_tiles) issue(j + kStages);
            if (j >= 2) bar_sync(kBarVEmpty + buf, kThreads);     // the 1x1 warps are done with v[buf]
            float *vp = vbuf + buf * (kC * kTP) + px;
#pragma unroll
            for (int i = 0; i < kC; ++i) vp[i * kTP] = g[i];
            bar_arrive(kBarVFull + buf, kThreads);         // v[buf] is complete
        }
    } else {
        // =========================== 1x1 + bias + residual =====================================
        const int t = tid - kThreadsA;
        const int q = t & 3, pq = t >> 2;                  // output quarter (8 outputs), pixel quad 0..31
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            const int tile = blockIdx.x + j * gridDim.x;
            const int b = tile / tiles_per_img;
            const int64_t p = (int64_t)(tile - b * tiles_per_img) * kTP + 4 * pq;
            const bool inside = p < hw;                    // hw % 4 == 0: the four pixels are inside or outside together
            const int64_t o = ((int64_t)b * kC + q * 8) * hw + p;
            float4 rv[8];
            if (residual != nullptr && inside) {
#pragma unroll
                for (int i = 0; i < 8; ++i) rv[i] = __ldg(reinterpret_cast<const float4 *>(residual + o + (int64_t)i * hw));
            }
            bar_sync(kBarVFull + buf, kThreads);           // v[buf] is complete
            const float *src = vbuf + buf * (kC * kTP) + 4 * pq;
            f32x2 acc[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[pp][i] = reinterpret_cast<const f32x2 *>(pb)[q * 4 + i];
#pragma unroll 4
            for (int ci = 0; ci < kC; ++ci) {
                const float4 xv = *reinterpret_cast<const float4 *>(src + ci * kTP);
                const f32x2 x2[4] = {pack2(xv.x, xv.x), pack2(xv.y, xv.y), pack2(xv.z, xv.z), pack2(xv.w, xv.w)};
                const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(wt + ci * kC + q * 8);
                const ulonglong2 wa = wr[0], wb2 = wr[1];
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    acc[pp][0] = ffma2(x2[pp], wa.x, acc[pp][0]);
                    acc[pp][1] = ffma2(x2[pp], wa.y, acc[pp][1]);
                    acc[pp][2] = ffma2(x2[pp], wb2.x, acc[pp][2]);
                    acc[pp][3] = ffma2(x2[pp], wb2.y, acc[pp][3]);
                }
            }
            if (j + 2 < my_tiles) bar_arrive(kBarVEmpty + buf, kThreads);     // v[buf] may be overwritten
            if (inside) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v[2][4];
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) unpack2(acc[pp][i], v[0][pp], v[1][pp]);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int co = 2 * i + k;
                        float4 r = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                        if (residual != nullptr) {
                            const float sc = rs[q * 8 + co];
                            const float4 u = rv[co];
                            r = make_float4(fmaf(u.x, sc, r.x), fmaf(u.y, sc, r.y), fmaf(u.z, sc, r.z), fmaf(u.w, sc, r.w));
                        }
                        *reinterpret_cast<float4 *>(y + o + (int64_t)co * hw) = r;
                    }
                }
            }
        }
    }
}

// Returns WM_OK when the pipeline ran, 1 when its preconditions do not hold, or an error code.
int forward_gate(const float *x, int64_t x_bstride, const float *w, const float *bias, const float *residual,
                 const float *res_scale, float *y, int64_t B, int64_t hw, cudaStream_t s)
{
    if (hw % 4 != 0 || hw < kTP || (x_bstride != 0 && x_bstride != 64 * hw)) return 1;
    if (!aligned16(x) || !aligned16(y) || (residual && !aligned16(residual))) return 1;
    const int64_t rows = B * G::kRows;
    const int64_t tiles_per_img = (hw + kTP - 1) / kTP, total = tiles_per_img * B;
    if (total >= ((int64_t)1 << 31) || rows >= ((int64_t)1 << 31)) return 1;
    tma::EncodeTiledFn enc = tma::encode_fn();
    if (enc == nullptr) return 1;
    CUtensorMap xmap;
    const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)hw * 4};
    const cuuint32_t box[2] = {kTP, (cuuint32_t)G::kRows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    WM_CUDA_OK(cudaFuncSetAttribute(pw32_gate_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem));
    const int grid = total < sm_count() ? (int)total : sm_count();
    pw32_gate_tma_kernel<<<grid, kThreads, G::kSmem, s>>>(xmap, w, bias, residual, res_scale, y, hw, (int)tiles_per_img,
                                                          (int)total);
    WM_LAUNCH_OK("pw gate 32->32 (TMA)");
    return WM_OK;
}

}  // namespace pwt
}  // namespace wm
