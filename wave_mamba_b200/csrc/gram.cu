// 32x32 Gram matrix + squared row norms of two (32, hw) channel stacks, one pass over HBM.
//
// Serves the two global reductions of HFEBlock (reference wavemamba_arch.py):
//   * Matching: torch.cdist(x, perception) (:664) is sqrt(|x_i|^2 + |p_j|^2 - 2 x_i.p_j) in its
//     default mm mode; the argmin over j (:624) only needs G = X P^T and the two norm vectors.
//   * CMTAttention: normalize(q) @ normalize(k)^T (:787-790) = (q k^T) / (|q| |k|^T).
// cuBLAS handles this shape (M=N=32, K = hw up to 2 M) with split-K SGEMMs at ~10 % of the HBM
// roofline; here every CTA streams a pixel range, keeps a 2x2 register block per thread, sums
// 128-pixel tiles in fp32 and carries tile sums in fp64; per-CTA partials are reduced in a fixed
// order by a second tiny kernel => deterministic, and closer to the fp64 truth than SGEMM.
// Algorithmic bytes: 2 * 32 * hw * 4 per batch item.
#include "common.cuh"

namespace wm {
namespace gram {

constexpr int kC = 32;
constexpr int kTile = 128;           // pixels per smem tile
constexpr int kRS = 132;             // smem row stride (33 quads: conflict-free LDS.128 rows)
constexpr int kThreads = 256;
constexpr int kOut = kC * kC + 2 * kC;   // 1088 values per batch item: G | |X|^2 | |Y|^2

__device__ __forceinline__ float sq4(const float4 v, float acc)
{
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    return acc;
}

__global__ void __launch_bounds__(kThreads)
gram_partial_kernel(const float *__restrict__ x, int64_t x_bstride, const float *__restrict__ y,
                    int64_t y_bstride, double *__restrict__ partial, int64_t hw, int chunk,
                    int nchunks, int vec)
{
    __shared__ __align__(16) float xs[kC * kRS];
    __shared__ __align__(16) float ysm[kC * kRS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = tid >> 4, tj = tid & 15;     // rows {ti, ti+16} of X, rows {tj, tj+16} of Y
    const int64_t b = blockIdx.y;
    const float *xb = x + b * x_bstride;
    const float *yb = y + b * y_bstride;
    const int64_t p_begin = (int64_t)blockIdx.x * chunk;
    int64_t p_end = p_begin + chunk;
    if (p_end > hw) p_end = hw;

    double g00 = 0, g01 = 0, g10 = 0, g11 = 0;
    // staging role: warp w loads rows w, w+8, w+16, w+24 (lane = quad); it also owns their norms
    double nx[4] = {0, 0, 0, 0}, ny[4] = {0, 0, 0, 0};

    for (int64_t p0 = p_begin; p0 < p_end; p0 += kTile) {
        const bool full = vec && p0 + kTile <= p_end;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = warp + 8 * r;
            float4 a, c;
            if (full) {
                a = ld_stream4(xb + (int64_t)row * hw + p0 + 4 * lane);
                c = ld_stream4(yb + (int64_t)row * hw + p0 + 4 * lane);
            } else {
                float av[4], cv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t p = p0 + 4 * lane + e;
                    const bool ok = p < p_end;
                    av[e] = ok ? __ldg(xb + (int64_t)row * hw + p) : 0.0f;
                    cv[e] = ok ? __ldg(yb + (int64_t)row * hw + p) : 0.0f;
                }
                a = make_float4(av[0], av[1], av[2], av[3]);
                c = make_float4(cv[0], cv[1], cv[2], cv[3]);
            }
            *reinterpret_cast<float4 *>(xs + row * kRS + 4 * lane) = a;
            *reinterpret_cast<float4 *>(ysm + row * kRS + 4 * lane) = c;
            nx[r] += (double)sq4(a, 0.0f);
            ny[r] += (double)sq4(c, 0.0f);
        }
        __syncthreads();
        float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
        const float4 *xr0 = reinterpret_cast<const float4 *>(xs + ti * kRS);
        const float4 *xr1 = reinterpret_cast<const float4 *>(xs + (ti + 16) * kRS);
        const float4 *yr0 = reinterpret_cast<const float4 *>(ysm + tj * kRS);
        const float4 *yr1 = reinterpret_cast<const float4 *>(ysm + (tj + 16) * kRS);
#pragma unroll 4
        for (int q = 0; q < kTile / 4; ++q) {
            const float4 u0 = xr0[q], u1 = xr1[q], v0 = yr0[q], v1 = yr1[q];
            a00 = fmaf(u0.x, v0.x, a00); a00 = fmaf(u0.y, v0.y, a00);
            a00 = fmaf(u0.z, v0.z, a00); a00 = fmaf(u0.w, v0.w, a00);
            a01 = fmaf(u0.x, v1.x, a01); a01 = fmaf(u0.y, v1.y, a01);
            a01 = fmaf(u0.z, v1.z, a01); a01 = fmaf(u0.w, v1.w, a01);
            a10 = fmaf(u1.x, v0.x, a10); a10 = fmaf(u1.y, v0.y, a10);
            a10 = fmaf(u1.z, v0.z, a10); a10 = fmaf(u1.w, v0.w, a10);
            a11 = fmaf(u1.x, v1.x, a11); a11 = fmaf(u1.y, v1.y, a11);
            a11 = fmaf(u1.z, v1.z, a11); a11 = fmaf(u1.w, v1.w, a11);
        }
        g00 += (double)a00; g01 += (double)a01; g10 += (double)a10; g11 += (double)a11;
        __syncthreads();
    }
    double *out = partial + ((int64_t)b * nchunks + blockIdx.x) * kOut;
    out[ti * kC + tj] = g00;
    out[ti * kC + tj + 16] = g01;
    out[(ti + 16) * kC + tj] = g10;
    out[(ti + 16) * kC + tj + 16] = g11;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double sx = nx[r], sy = ny[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
            sy += __shfl_xor_sync(0xffffffffu, sy, off);
        }
        if (lane == 0) {
            out[kC * kC + warp + 8 * r] = sx;
            out[kC * kC + kC + warp + 8 * r] = sy;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
gram_reduce_kernel(const double *__restrict__ partial, float *__restrict__ out, int nchunks)
{
    const int64_t b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= kOut) return;
    const double *p = partial + b * nchunks * kOut + i;
    double acc = 0.0;
    for (int c = 0; c < nchunks; ++c) acc += p[(int64_t)c * kOut];
    out[b * kOut + i] = (float)acc;
}

inline void plan(int64_t hw, int &chunk, int &nchunks)
{
    // about two CTAs per SM, chunk a multiple of the tile
    const int64_t target = (int64_t)sm_count() * 2;
    int64_t c = (hw + target - 1) / target;
    c = (c + kTile - 1) / kTile * kTile;
    if (c < kTile) c = kTile;
    chunk = (int)c;
    nchunks = (int)((hw + c - 1) / c);
}

}  // namespace gram
}  // namespace wm

extern "C" size_t wm_gram32_workspace_bytes(int64_t B, int64_t hw)
{
    if (B <= 0 || hw <= 0) return 0;
    int chunk, nchunks;
    wm::gram::plan(hw, chunk, nchunks);
    return (size_t)B * nchunks * wm::gram::kOut * sizeof(double);
}

extern "C" int wm_gram32_fwd(const float *x, int64_t x_bstride, const float *y, int64_t y_bstride,
                             float *out, void *workspace, size_t workspace_bytes, int64_t B,
                             int64_t hw, wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::gram;
    WM_REQUIRE(B >= 0 && hw >= 0 && B <= 65535, "wm_gram32_fwd: bad sizes");
    if (B == 0) return WM_OK;
    WM_REQUIRE(x && y && out, "wm_gram32_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (hw == 0) {
        WM_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)B * kOut * sizeof(float), s));
        return WM_OK;
    }
    WM_REQUIRE(x_bstride >= kC * hw && y_bstride >= kC * hw, "wm_gram32_fwd: batch stride too small");
    int chunk, nchunks;
    plan(hw, chunk, nchunks);
    const size_t need = (size_t)B * nchunks * kOut * sizeof(double);
    WM_REQUIRE(workspace && workspace_bytes >= need, "wm_gram32_fwd: workspace too small (%zu < %zu)",
               workspace_bytes, need);
    WM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "wm_gram32_fwd: workspace alignment");
    const int vec = (hw % 4 == 0 && aligned16(x) && aligned16(y) && x_bstride % 4 == 0 &&
                     y_bstride % 4 == 0) ? 1 : 0;
    dim3 grid(nchunks, (unsigned)B);
    gram_partial_kernel<<<grid, kThreads, 0, s>>>(x, x_bstride, y, y_bstride,
                                                  static_cast<double *>(workspace), hw, chunk,
                                                  nchunks, vec);
    WM_LAUNCH_OK("gram partial");
    dim3 rgrid((kOut + kThreads - 1) / kThreads, (unsigned)B);
    gram_reduce_kernel<<<rgrid, kThreads, 0, s>>>(static_cast<const double *>(workspace), out, nchunks);
    WM_LAUNCH_OK("gram reduce");
    return WM_OK;
}
