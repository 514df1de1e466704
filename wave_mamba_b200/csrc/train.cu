// Training-only kernels (reference training step basicsr/models/femasr_model.py:157-185):
// the pieces of the backward pass that are not a forward kernel run with transposed weights.
//
//   wm_dw3x3_fwd          depthwise 3x3 (zero pad 1), any channel count; flip=1 applies the 180-degree
//                         rotated taps = the data gradient of the same convolution
//   wm_dw3x3_wgrad        tap and bias gradients of a depthwise 3x3 (fixed-order fp64 reduction)
//   wm_layernorm2d_bwd    LayerNorm over channels (NCHW): dx, and the affine gradients through a
//                         per-pixel (mean, rstd) scratch + a fixed-order fp64 reduction
//
// Everything else in the backward reuses forward kernels: 1x1 / dense 3x3 data gradients are the same
// kernels with transposed (and rotated) weights, their weight gradients are Gram matrices over the
// pixels (wm_gram32_fwd), the Haar pair is its own adjoint, the SS2D core has wm_ss2d_core_bwd.
// These are simple streaming kernels: the training step is a functional path, not a tuned one.
#include "common.cuh"

namespace wm {
namespace train {

constexpr int kThreads = 256;
constexpr int kSplit = 16;     // pixel-range splits of the per-channel reductions

__global__ void __launch_bounds__(kThreads)
dw3x3_kernel(const float *__restrict__ x, const float *__restrict__ wgt, const float *__restrict__ bias,
             float *__restrict__ y, int64_t planes, int C, int h, int w, int flip)
{
    const int64_t hw = (int64_t)h * w;
    const int64_t total = planes * hw;
    for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * kThreads) {
        const int64_t pl = idx / hw;
        const int64_t rem = idx - pl * hw;
        const int i = (int)(rem / w), j = (int)(rem - (int64_t)i * w);
        const int c = (int)(pl % C);
        const float *xp = x + pl * hw;
        float acc = bias ? __ldg(bias + c) : 0.0f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int ii = i + t / 3 - 1, jj = j + t % 3 - 1;
            if (ii >= 0 && ii < h && jj >= 0 && jj < w)
                acc = fmaf(__ldg(wgt + c * 9 + (flip ? 8 - t : t)), __ldg(xp + (int64_t)ii * w + jj), acc);
        }
        y[idx] = acc;
    }
}

// block-wide sum of a double (fixed tree order), result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kThreads / 32; ++i) t += red[i];
    return t;
}

// part[c][split][10]: 9 tap sums + the bias sum of channel c over this split's pixels (all images)
__global__ void __launch_bounds__(kThreads)
dw3x3_wgrad_part_kernel(const float *__restrict__ dy, const float *__restrict__ x, double *__restrict__ part,
                        int B, int C, int h, int w)
{
    __shared__ double red[kThreads / 32];
    const int c = blockIdx.x, sp = blockIdx.y;
    const int64_t hw = (int64_t)h * w;
    const int64_t n = (int64_t)B * hw;
    const int64_t per = (n + kSplit - 1) / kSplit;
    const int64_t lo = sp * per, hi = lo + per < n ? lo + per : n;
    double acc[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t] = 0.0;
    for (int64_t q = lo + threadIdx.x; q < hi; q += kThreads) {
        const int64_t b = q / hw, rem = q - b * hw;
        const int i = (int)(rem / w), j = (int)(rem - (int64_t)i * w);
        const float g = __ldg(dy + (b * C + c) * hw + rem);
        const float *xp = x + (b * C + c) * hw;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int ii = i + t / 3 - 1, jj = j + t % 3 - 1;
            if (ii >= 0 && ii < h && jj >= 0 && jj < w) acc[t] += (double)g * (double)__ldg(xp + (int64_t)ii * w + jj);
        }
        acc[9] += (double)g;
    }
#pragma unroll 1
    for (int t = 0; t < 10; ++t) {
        const double v = block_sum(acc[t], red);
        if (threadIdx.x == 0) part[((int64_t)c * kSplit + sp) * 10 + t] = v;
    }
}

__global__ void dw3x3_wgrad_final_kernel(const double *__restrict__ part, float *__restrict__ dwgt,
                                         float *__restrict__ dbias, int C)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * 10) return;
    const int c = i / 10, t = i - c * 10;
    double v = 0.0;
    for (int sp = 0; sp < kSplit; ++sp) v += part[((int64_t)c * kSplit + sp) * 10 + t];
    if (t < 9) dwgt[c * 9 + t] = (float)v;
    else if (dbias) dbias[c] = (float)v;
}

// dx of y = w * (x - mu) * rstd + b over the channels of each pixel; stats[p] = (mu, rstd)
template <int C>
__global__ void __launch_bounds__(kThreads)
ln2d_bwd_dx_kernel(const float *__restrict__ x, const float *__restrict__ ln_w, const float *__restrict__ dy,
                   float eps, float *__restrict__ dx, float2 *__restrict__ stats, int64_t hw)
{
    const int64_t b = blockIdx.y;
    const float *xb = x + b * C * hw, *gb = dy + b * C * hw;
    float *ob = dx + b * C * hw;
    for (int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x; p < hw; p += (int64_t)gridDim.x * kThreads) {
        float v[C];
        float mu = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) { v[c] = __ldg(xb + c * hw + p); mu += v[c]; }
        mu *= (1.0f / C);
        float var = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) { v[c] -= mu; var = fmaf(v[c], v[c], var); }
        var *= (1.0f / C);
        const float rstd = 1.0f / sqrtf(var + eps);
        float m1 = 0.0f, m2 = 0.0f;     // mean(g), mean(g * xhat),  g = dy * w
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float g = __ldg(gb + c * hw + p) * __ldg(ln_w + c);
            v[c] *= rstd;               // xhat
            m1 += g;
            m2 = fmaf(g, v[c], m2);
        }
        m1 *= (1.0f / C);
        m2 *= (1.0f / C);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float g = __ldg(gb + c * hw + p) * __ldg(ln_w + c);
            ob[c * hw + p] = rstd * (g - m1 - v[c] * m2);
        }
        stats[b * hw + p] = make_float2(mu, rstd);
    }
}

// part[c][split][2]: sum dy * xhat, sum dy of channel c
__global__ void __launch_bounds__(kThreads)
ln2d_bwd_param_part_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                           const float2 *__restrict__ stats, double *__restrict__ part, int B, int C, int64_t hw)
{
    __shared__ double red[kThreads / 32];
    const int c = blockIdx.x, sp = blockIdx.y;
    const int64_t n = (int64_t)B * hw;
    const int64_t per = (n + kSplit - 1) / kSplit;
    const int64_t lo = sp * per, hi = lo + per < n ? lo + per : n;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t q = lo + threadIdx.x; q < hi; q += kThreads) {
        const int64_t b = q / hw, rem = q - b * hw;
        const float g = __ldg(dy + (b * C + c) * hw + rem);
        const float2 st = stats[q];
        a0 += (double)g * (double)((__ldg(x + (b * C + c) * hw + rem) - st.x) * st.y);
        a1 += (double)g;
    }
    const double s0 = block_sum(a0, red);
    const double s1 = block_sum(a1, red);
    if (threadIdx.x == 0) {
        part[((int64_t)c * kSplit + sp) * 2 + 0] = s0;
        part[((int64_t)c * kSplit + sp) * 2 + 1] = s1;
    }
}

__global__ void ln2d_bwd_param_final_kernel(const double *__restrict__ part, float *__restrict__ dw,
                                            float *__restrict__ db, int C)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double v0 = 0.0, v1 = 0.0;
    for (int sp = 0; sp < kSplit; ++sp) {
        v0 += part[((int64_t)c * kSplit + sp) * 2 + 0];
        v1 += part[((int64_t)c * kSplit + sp) * 2 + 1];
    }
    dw[c] = (float)v0;
    db[c] = (float)v1;
}

inline int flat_grid(int64_t n)
{
    const int64_t want = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace train
}  // namespace wm

using namespace wm;
using namespace wm::train;

extern "C" int wm_dw3x3_fwd(const float *x, const float *wgt, const float *bias, float *y, int64_t B, int64_t C,
                            int64_t h, int64_t w, int flip, wm_stream_t stream)
{
    WM_REQUIRE(B >= 0 && C > 0 && h >= 0 && w >= 0 && h < (1 << 24) && w < (1 << 24), "wm_dw3x3_fwd: bad sizes");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(x && wgt && y, "wm_dw3x3_fwd: null pointer");
    dw3x3_kernel<<<flat_grid(B * C * h * w), kThreads, 0, (cudaStream_t)stream>>>(x, wgt, bias, y, B * C, (int)C,
                                                                                    (int)h, (int)w, flip ? 1 : 0);
    WM_LAUNCH_OK("dw3x3");
    return WM_OK;
}

extern "C" size_t wm_train_workspace_bytes(int64_t B, int64_t C, int64_t h, int64_t w)
{
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return 0;
    // per-channel partials (10 doubles x kSplit) + per-pixel LayerNorm statistics
    return (size_t)C * kSplit * 10 * sizeof(double) + (size_t)B * h * w * sizeof(float2) + 256;
}

extern "C" int wm_dw3x3_wgrad(const float *dy, const float *x, float *dwgt, float *dbias, void *workspace,
                              size_t workspace_bytes, int64_t B, int64_t C, int64_t h, int64_t w,
                              wm_stream_t stream)
{
    WM_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && C <= 65535, "wm_dw3x3_wgrad: bad sizes");
    WM_REQUIRE(dy && x && dwgt && workspace, "wm_dw3x3_wgrad: null pointer");
    WM_REQUIRE(workspace_bytes >= (size_t)C * kSplit * 10 * sizeof(double) && aligned16(workspace),
               "wm_dw3x3_wgrad: workspace too small or unaligned");
    double *part = static_cast<double *>(workspace);
    cudaStream_t s = (cudaStream_t)stream;
    dw3x3_wgrad_part_kernel<<<dim3((unsigned)C, kSplit), kThreads, 0, s>>>(dy, x, part, (int)B, (int)C, (int)h, (int)w);
    WM_LAUNCH_OK("dw3x3 wgrad");
    dw3x3_wgrad_final_kernel<<<(int)((C * 10 + 255) / 256), 256, 0, s>>>(part, dwgt, dbias, (int)C);
    WM_LAUNCH_OK("dw3x3 wgrad final");
    return WM_OK;
}

extern "C" int wm_layernorm2d_bwd(const float *x, const float *ln_w, const float *dy, float eps, float *dx,
                                  float *dln_w, float *dln_b, void *workspace, size_t workspace_bytes,
                                  int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream)
{
    WM_REQUIRE(B > 0 && h > 0 && w > 0 && B <= 65535, "wm_layernorm2d_bwd: bad sizes");
    WM_REQUIRE(C == 32 || C == 64, "wm_layernorm2d_bwd: C=%lld unsupported (32 or 64)", (long long)C);
    WM_REQUIRE(x && ln_w && dy && dx && dln_w && dln_b && workspace, "wm_layernorm2d_bwd: null pointer");
    const int64_t hw = h * w;
    const size_t part_bytes = (size_t)C * kSplit * 10 * sizeof(double);
    WM_REQUIRE(workspace_bytes >= part_bytes + (size_t)B * hw * sizeof(float2) && aligned16(workspace),
               "wm_layernorm2d_bwd: workspace too small or unaligned");
    double *part = static_cast<double *>(workspace);
    float2 *stats = reinterpret_cast<float2 *>(static_cast<char *>(workspace) + part_bytes);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t want = (hw + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)B);
    if (C == 32) ln2d_bwd_dx_kernel<32><<<grid, kThreads, 0, s>>>(x, ln_w, dy, eps, dx, stats, hw);
    else ln2d_bwd_dx_kernel<64><<<grid, kThreads, 0, s>>>(x, ln_w, dy, eps, dx, stats, hw);
    WM_LAUNCH_OK("layernorm2d bwd dx");
    ln2d_bwd_param_part_kernel<<<dim3((unsigned)C, kSplit), kThreads, 0, s>>>(x, dy, stats, part, (int)B, (int)C, hw);
    WM_LAUNCH_OK("layernorm2d bwd params");
    ln2d_bwd_param_final_kernel<<<1, 64, 0, s>>>(part, dln_w, dln_b, (int)C);
    WM_LAUNCH_OK("layernorm2d bwd params final");
    return WM_OK;
}
