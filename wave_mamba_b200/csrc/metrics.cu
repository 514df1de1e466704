// PSNR and SSIM on the Y channel, on the device, for sm_100a -- the two numbers the inference loop prints
// per image (inference_wavemamba.py:117-118 -> comput_psnr_ssim.py calculate_psnr :387-438 and
// calculate_ssim :596-668 with their defaults: uint8 BGR HWC images, crop_border = 1, test_y_channel).
// On the host they cost five cv2.filter2D passes over float64 4K planes per image (seconds); here the
// enhanced uint8 image never has to leave the device for them.
//
//   Y  (to_y_channel :374-385 -> bgr2ycbcr(y_only) :210-238, restated operation by operation):
//        f = float32(byte) / 255   (fp32 division);  y = f_b*24.966 + f_g*128.553 + f_r*65.481 + 16  (fp64, np.dot)
//        Y = float32(y / 255) * 255                  (fp32; _convert_output_type_range + "* 255.")
//   PSNR = 20 log10(255 / sqrt(mean((Y1 - Y2)^2))), +inf when the mean is 0.  (Y1 - Y2)^2 is formed in fp32 as
//        numpy does; the mean is accumulated in fp64 (numpy: fp32 pairwise -- the reference's own result
//        carries ~1e-7 relative noise, this one does not).
//   SSIM (_ssim_cly :559-592): 11x11 Gaussian window (sigma 1.5, cv2.getGaussianKernel) correlated with
//        Y1, Y2, Y1^2, Y2^2, Y1*Y2 in fp64, BORDER_REPLICATE, C1 = 6.5025, C2 = 58.5225, mean over the map.
//        The window is an outer product, so it is applied as a row pass and a column pass.
//
// One kernel: a CTA owns a 32x32 tile of the cropped image; the 42x42 Y values it needs are computed from
// the bytes into shared memory (coordinates clamped = replicate border), row pass -> five fp64 maps in
// shared memory, column pass -> the SSIM value per pixel.  Per-CTA fp64 partials, summed in a fixed order
// by a second kernel: bit-reproducible run to run.
#include <math.h>

#include "common.cuh"

namespace wm {
namespace metrics {

constexpr int kT = 32;                 // tile edge
constexpr int kTaps = 11, kHalo = kTaps / 2;
constexpr int kIn = kT + 2 * kHalo;    // 42
constexpr int kPitch = kIn + 1;
constexpr int kThreads = 256;
constexpr size_t kSmem = sizeof(double) * 5 * kIn * kT + sizeof(float) * 2 * kIn * kPitch;

struct Window { double w[kTaps]; };

__device__ __forceinline__ float y_of(const uint8_t *p)
{
    const float b = __fdiv_rn((float)p[0], 255.0f), g = __fdiv_rn((float)p[1], 255.0f),
                r = __fdiv_rn((float)p[2], 255.0f);
    const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)b, 24.966), __dmul_rn((double)g, 128.553)),
                                         __dmul_rn((double)r, 65.481)), 16.0);
    return __fmul_rn((float)__ddiv_rn(y, 255.0), 255.0f);
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
    // fixed tree: lanes by shuffle, then the 8 warp sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(kThreads)
psnr_ssim_tile_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int H, int W, int crop,
                      const Window win, double *__restrict__ partial)
{
    extern __shared__ double sm[];
    __shared__ double red[kThreads / 32];
    double *hs = sm;                                        // [5][kIn][kT]: row-filtered Y1, Y2, Y1^2, Y2^2, Y1*Y2
    float *ya = reinterpret_cast<float *>(hs + 5 * kIn * kT);   // [kIn][kPitch]
    float *yb = ya + kIn * kPitch;
    const int h = H - 2 * crop, w = W - 2 * crop;
    const int tx0 = blockIdx.x * kT, ty0 = blockIdx.y * kT;
    const int64_t img = blockIdx.z;
    const uint8_t *pa = a + img * H * W * 3, *pb = b + img * H * W * 3;
    const int tid = threadIdx.x;

    double sq = 0.0;
    for (int i = tid; i < kIn * kIn; i += kThreads) {
        const int r = i / kIn, c = i - r * kIn;
        const int y = ty0 + r - kHalo, x = tx0 + c - kHalo;
        const int yc = min(max(y, 0), h - 1), xc = min(max(x, 0), w - 1);      // BORDER_REPLICATE
        const int64_t off = ((int64_t)(yc + crop) * W + xc + crop) * 3;
        const float fa = y_of(pa + off), fb = y_of(pb + off);
        ya[r * kPitch + c] = fa;
        yb[r * kPitch + c] = fb;
        if (r >= kHalo && r < kHalo + kT && c >= kHalo && c < kHalo + kT && y < h && x < w) {
            const float d = __fsub_rn(fa, fb);
            sq += (double)__fmul_rn(d, d);
        }
    }
    __syncthreads();
    for (int i = tid; i < kIn * kT; i += kThreads) {
        const int r = i / kT, c = i - r * kT;
        double s1 = 0.0, s2 = 0.0, s11 = 0.0, s22 = 0.0, s12 = 0.0;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
            const double u = (double)ya[r * kPitch + c + k], v = (double)yb[r * kPitch + c + k], wk = win.w[k];
            s1 += wk * u;
            s2 += wk * v;
            s11 += wk * (u * u);
            s22 += wk * (v * v);
            s12 += wk * (u * v);
        }
        hs[(0 * kIn + r) * kT + c] = s1;
        hs[(1 * kIn + r) * kT + c] = s2;
        hs[(2 * kIn + r) * kT + c] = s11;
        hs[(3 * kIn + r) * kT + c] = s22;
        hs[(4 * kIn + r) * kT + c] = s12;
    }
    __syncthreads();
    const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
    double ss = 0.0;
    for (int i = tid; i < kT * kT; i += kThreads) {
        const int r = i / kT, c = i - r * kT;
        if (ty0 + r < h && tx0 + c < w) {
            double m1 = 0.0, m2 = 0.0, e11 = 0.0, e22 = 0.0, e12 = 0.0;
#pragma unroll
            for (int k = 0; k < kTaps; ++k) {
                const double wk = win.w[k];
                m1 += wk * hs[(0 * kIn + r + k) * kT + c];
                m2 += wk * hs[(1 * kIn + r + k) * kT + c];
                e11 += wk * hs[(2 * kIn + r + k) * kT + c];
                e22 += wk * hs[(3 * kIn + r + k) * kT + c];
                e12 += wk * hs[(4 * kIn + r + k) * kT + c];
            }
            const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
            const double v1 = e11 - m11, v2 = e22 - m22, v12 = e12 - m12;
            ss += ((2.0 * m12 + C1) * (2.0 * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
        }
    }
    const double tsq = block_sum(sq, red);
    const double tss = block_sum(ss, red);
    if (tid == 0) {
        const int64_t nct = (int64_t)gridDim.x * gridDim.y;
        double *o = partial + (img * nct + (int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
        o[0] = tsq;
        o[1] = tss;
    }
}

// one CTA per image: thread t sums partials t, t + 256, ... in order, then the fixed tree
__global__ void __launch_bounds__(kThreads)
psnr_ssim_final_kernel(const double *__restrict__ partial, int64_t nct, double count, double *__restrict__ out)
{
    __shared__ double red[kThreads / 32];
    const int64_t img = blockIdx.x;
    const double *p = partial + img * nct * 2;
    double sq = 0.0, ss = 0.0;
    for (int64_t i = threadIdx.x; i < nct; i += kThreads) {
        sq += p[2 * i];
        ss += p[2 * i + 1];
    }
    const double tsq = block_sum(sq, red);
    const double tss = block_sum(ss, red);
    if (threadIdx.x == 0) {
        const double mse = tsq / count;
        out[img * 2 + 0] = mse == 0.0 ? INFINITY : 20.0 * log10(255.0 / sqrt(mse));
        out[img * 2 + 1] = tss / count;
    }
}

inline int64_t tiles(int64_t n) { return (n + kT - 1) / kT; }

}  // namespace metrics
}  // namespace wm

extern "C" size_t wm_psnr_ssim_y_workspace_bytes(int64_t B, int64_t H, int64_t W, int crop_border)
{
    using namespace wm::metrics;
    const int64_t h = H - 2 * (int64_t)crop_border, w = W - 2 * (int64_t)crop_border;
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)(B * tiles(h) * tiles(w) * 2) * sizeof(double);
}

extern "C" int wm_psnr_ssim_y_u8(const uint8_t *img1, const uint8_t *img2, double *out, void *workspace,
                                 size_t workspace_bytes, int64_t B, int64_t H, int64_t W, int crop_border,
                                 wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::metrics;
    WM_REQUIRE(B >= 0 && B <= 65535 && H >= 0 && W >= 0, "wm_psnr_ssim_y_u8: bad sizes");
    WM_REQUIRE(crop_border >= 0, "wm_psnr_ssim_y_u8: crop_border must be >= 0");
    if (B == 0) return WM_OK;
    const int64_t h = H - 2 * (int64_t)crop_border, w = W - 2 * (int64_t)crop_border;
    WM_REQUIRE(h >= 1 && w >= 1, "wm_psnr_ssim_y_u8: the cropped image is empty (H=%lld W=%lld crop=%d)",
               (long long)H, (long long)W, crop_border);
    WM_REQUIRE(H * W < (int64_t)1 << 31, "wm_psnr_ssim_y_u8: image too large");
    WM_REQUIRE(img1 && img2 && out && workspace, "wm_psnr_ssim_y_u8: null pointer");
    const size_t need = wm_psnr_ssim_y_workspace_bytes(B, H, W, crop_border);
    WM_REQUIRE(workspace_bytes >= need, "wm_psnr_ssim_y_u8: workspace %zu < %zu bytes", workspace_bytes, need);
    WM_REQUIRE(tiles(h) <= 65535, "wm_psnr_ssim_y_u8: image too tall");

    // cv2.getGaussianKernel(11, 1.5): exp(-(i - 5)^2 / (2 sigma^2)) scaled by the reciprocal of the sum
    Window win;
    double sum = 0.0;
    for (int i = 0; i < kTaps; ++i) {
        const double x = i - kHalo;
        win.w[i] = exp(-0.5 / (1.5 * 1.5) * x * x);
        sum += win.w[i];
    }
    const double inv = 1.0 / sum;
    for (int i = 0; i < kTaps; ++i) win.w[i] *= inv;

    cudaStream_t s = (cudaStream_t)stream;
    WM_CUDA_OK(cudaFuncSetAttribute(psnr_ssim_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    dim3 grid((unsigned)tiles(w), (unsigned)tiles(h), (unsigned)B);
    psnr_ssim_tile_kernel<<<grid, kThreads, kSmem, s>>>(img1, img2, (int)H, (int)W, crop_border, win,
                                                       static_cast<double *>(workspace));
    WM_LAUNCH_OK("psnr_ssim tile");
    psnr_ssim_final_kernel<<<(unsigned)B, kThreads, 0, s>>>(static_cast<const double *>(workspace),
                                                           tiles(h) * tiles(w), (double)h * (double)w, out);
    WM_LAUNCH_OK("psnr_ssim final");
    return WM_OK;
}
