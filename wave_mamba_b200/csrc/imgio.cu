// Image I/O edges of the inference loop for sm_100a -- the only per-image host<->device traffic
// becomes the uint8 image itself (3 bytes per pixel each way instead of 12).
//
//   wm_img_u8_to_f32_fwd   cv2 image (B,H,W,3) uint8 BGR  ->  (B,3,Hp,Wp) float32 RGB in [0,1],
//       = img2tensor (basicsr/utils/img_util.py:9-33: BGR->RGB, HWC->CHW, float) followed by "/ 255."
//       and check_image_size (inference_wavemamba.py:28-36,103-106: reflect pad at the bottom/right
//       up to a multiple of the window).  true fp32 division => bit-exact with the reference.
//   wm_img_f32_to_u8_fwd   (B,3,Hs,Ws) float32 RGB  ->  crop [:h,:w], clamp to [0,1], * 255, round
//       half to even, uint8, RGB->BGR, CHW->HWC = the crop at inference_wavemamba.py:112 + tensor2img
//       (img_util.py:36-98).  Bit-exact as well.
// Thread = four consecutive pixels of a row: three 4-byte words of packed BGR <-> one float4 per plane.
#include "common.cuh"

namespace wm {
namespace imgio {

constexpr int kThreads = 256;

__device__ __forceinline__ int reflect(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }

__global__ void __launch_bounds__(kThreads)
u8_to_f32_kernel(const uint8_t *__restrict__ img, float *__restrict__ out, int H, int W, int Hp, int Wp,
                 int recip)
{
    // "/ 255." is an IEEE division on the CPU; torch's CUDA kernel for tensor / python-scalar
    // multiplies by the rounded reciprocal instead (1 ulp apart for some bytes).  Both are offered.
    const float inv255 = __fdiv_rn(1.0f, 255.0f);
    auto scale = [&](uint8_t v) { return recip ? __fmul_rn((float)v, inv255) : __fdiv_rn((float)v, 255.0f); };
    const int b = blockIdx.z, y = blockIdx.y;
    const int x0 = (blockIdx.x * kThreads + threadIdx.x) * 4;
    if (x0 >= Wp) return;
    const int sy = reflect(y, H);
    const uint8_t *row = img + ((int64_t)b * H + sy) * W * 3;
    float r[4], g[4], bl[4];
    if (x0 + 4 <= W && (W % 4) == 0) {
        // 12 bytes = 4 BGR pixels; the row start is 4-byte aligned because W % 4 == 0
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + (int64_t)x0 * 3);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const uint8_t by[12] = {(uint8_t)(w0), (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                                (uint8_t)(w1), (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                                (uint8_t)(w2), (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bl[i] = scale(by[3 * i + 0]);
            g[i] = scale(by[3 * i + 1]);
            r[i] = scale(by[3 * i + 2]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = x0 + i;
            const int sx = reflect(x < Wp ? x : Wp - 1, W);
            const uint8_t *p = row + (int64_t)sx * 3;
            bl[i] = scale(__ldg(p + 0));
            g[i] = scale(__ldg(p + 1));
            r[i] = scale(__ldg(p + 2));
        }
    }
    const int64_t plane = (int64_t)Hp * Wp;
    float *o = out + (int64_t)b * 3 * plane + (int64_t)y * Wp + x0;
    if (x0 + 4 <= Wp && (Wp % 4) == 0) {
        *reinterpret_cast<float4 *>(o) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4 *>(o + plane) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4 *>(o + 2 * plane) = make_float4(bl[0], bl[1], bl[2], bl[3]);
    } else {
        for (int i = 0; i < 4 && x0 + i < Wp; ++i) {
            o[i] = r[i]; o[plane + i] = g[i]; o[2 * plane + i] = bl[i];
        }
    }
}

__device__ __forceinline__ uint32_t to_u8(float v)
{
    v = fminf(fmaxf(v, 0.0f), 1.0f);                       // clamp_(0, 1)
    return (uint32_t)__float2int_rn(__fmul_rn(v, 255.0f));  // (x * 255.0).round(): half to even
}

__global__ void __launch_bounds__(kThreads)
f32_to_u8_kernel(const float *__restrict__ x, uint8_t *__restrict__ img, int h, int w, int Hs, int Ws)
{
    const int b = blockIdx.z, y = blockIdx.y;
    const int x0 = (blockIdx.x * kThreads + threadIdx.x) * 4;
    if (x0 >= w) return;
    const int64_t plane = (int64_t)Hs * Ws;
    const float *p = x + (int64_t)b * 3 * plane + (int64_t)y * Ws + x0;
    uint8_t *row = img + (((int64_t)b * h + y) * w + x0) * 3;
    if (x0 + 4 <= w && (w % 4) == 0 && (Ws % 4) == 0) {
        const float4 r = ld_stream4(p), g = ld_stream4(p + plane), bl = ld_stream4(p + 2 * plane);
        const uint32_t c[12] = {to_u8(bl.x), to_u8(g.x), to_u8(r.x), to_u8(bl.y), to_u8(g.y), to_u8(r.y),
                                to_u8(bl.z), to_u8(g.z), to_u8(r.z), to_u8(bl.w), to_u8(g.w), to_u8(r.w)};
        uint32_t *o = reinterpret_cast<uint32_t *>(row);
        o[0] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
        o[1] = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
        o[2] = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
    } else {
        for (int i = 0; i < 4 && x0 + i < w; ++i) {
            row[3 * i + 0] = (uint8_t)to_u8(p[2 * plane + i]);
            row[3 * i + 1] = (uint8_t)to_u8(p[plane + i]);
            row[3 * i + 2] = (uint8_t)to_u8(p[i]);
        }
    }
}

}  // namespace imgio
}  // namespace wm

extern "C" int wm_img_u8_to_f32_fwd(const uint8_t *img, float *out, int64_t B, int64_t H, int64_t W,
                                    int64_t Hp, int64_t Wp, int reciprocal, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(B >= 0 && H >= 0 && W >= 0 && B <= 65535 && Hp <= 65535, "wm_img_u8_to_f32_fwd: bad sizes");
    WM_REQUIRE(Hp >= H && Wp >= W, "wm_img_u8_to_f32_fwd: padded size smaller than the image");
    if (B == 0 || H == 0 || W == 0) return WM_OK;
    WM_REQUIRE(Hp - H < H && Wp - W < W, "wm_img_u8_to_f32_fwd: reflect padding must be smaller than the image");
    WM_REQUIRE(img && out, "wm_img_u8_to_f32_fwd: null pointer");
    WM_REQUIRE((reinterpret_cast<uintptr_t>(img) & 3u) == 0 && aligned16(out),
               "wm_img_u8_to_f32_fwd: img must be 4-byte and out 16-byte aligned");
    dim3 grid((unsigned)((Wp + 4 * imgio::kThreads - 1) / (4 * imgio::kThreads)), (unsigned)Hp, (unsigned)B);
    imgio::u8_to_f32_kernel<<<grid, imgio::kThreads, 0, (cudaStream_t)stream>>>(img, out, (int)H, (int)W,
                                                                                  (int)Hp, (int)Wp,
                                                                                  reciprocal ? 1 : 0);
    WM_LAUNCH_OK("img u8->f32");
    return WM_OK;
}

extern "C" int wm_img_f32_to_u8_fwd(const float *x, uint8_t *img, int64_t B, int64_t h, int64_t w,
                                    int64_t Hs, int64_t Ws, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(B >= 0 && h >= 0 && w >= 0 && B <= 65535 && h <= 65535, "wm_img_f32_to_u8_fwd: bad sizes");
    WM_REQUIRE(Hs >= h && Ws >= w, "wm_img_f32_to_u8_fwd: crop larger than the source");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(x && img, "wm_img_f32_to_u8_fwd: null pointer");
    WM_REQUIRE((reinterpret_cast<uintptr_t>(img) & 3u) == 0 && aligned16(x),
               "wm_img_f32_to_u8_fwd: img must be 4-byte and x 16-byte aligned");
    dim3 grid((unsigned)((w + 4 * imgio::kThreads - 1) / (4 * imgio::kThreads)), (unsigned)h, (unsigned)B);
    imgio::f32_to_u8_kernel<<<grid, imgio::kThreads, 0, (cudaStream_t)stream>>>(x, img, (int)h, (int)w,
                                                                                  (int)Hs, (int)Ws);
    WM_LAUNCH_OK("img f32->u8");
    return WM_OK;
}
