// SKFF band fusion and the pixel-unshuffle side inputs for sm_100a -- HBM-bound streaming kernels.
//
// SKFF (reference wavemamba_arch.py:923-959), three bands HL / LH / HH of one DWT level:
//   U = (f0 + f1) + f2 ; S = mean_hw(U) ; Z = PReLU(W_du S) ; a_k = softmax_k(W_k Z) ;
//   V = (f0 a_0 + f1 a_1) + f2 a_2
// as two passes over the bands instead of the reference's cat + sum + pool + 3 mul + 2 add:
//   skff_pool_kernel   reads the three bands once, per-CTA partial sums of U (fp32 per thread,
//                      fp64 across threads; fixed order => deterministic)
//   skff_apply_kernel  every CTA re-derives the 3 x 32 attention weights of its image from the
//                      partial sums (32->4->96 MACs, negligible), then streams V
// Algorithmic bytes: pool 3 planes read, apply 3 read + 1 written (per B*C*h*w*4).
//
// ps_down (reference :1014-1025, :1043-1045): PixelUnshuffle(r) followed by a 1x1 conv
// (3 r^2 -> 32, bias).  One kernel reads the r x r x 3 patch of an output pixel straight from the
// image and writes the 32 outputs: the unshuffled copy is never materialised.
#include "common.cuh"

namespace wm {
namespace skff {

constexpr int kThreads = 256;
constexpr int kC = 32;      // channels per band
constexpr int kMid = 4;     // max(32 / 8, 4)

// partial[(b*C + c) * nblk + blk] = sum over this CTA's slice of plane (b,c) of (f0 + f1) + f2
__global__ void __launch_bounds__(kThreads)
skff_pool_kernel(const float *__restrict__ f0, const float *__restrict__ f1,
                 const float *__restrict__ f2, double *__restrict__ partial, int64_t hw, int vec)
{
    const int nblk = gridDim.x;
    const int64_t plane = blockIdx.y;
    const float *p0 = f0 + plane * hw, *p1 = f1 + plane * hw, *p2 = f2 + plane * hw;
    float acc = 0.0f;
    if (vec == 2) {      // 256-bit loads (hw % 8 == 0, 32-byte aligned planes): half the LSU instructions
        const int64_t n8 = hw >> 3;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4;
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (int64_t)nblk * kThreads) {
            const float8 a = ld_stream8(p0 + 8 * i), b = ld_stream8(p1 + 8 * i), c = ld_stream8(p2 + 8 * i);
            a4.x += (a.lo.x + b.lo.x) + c.lo.x; a4.y += (a.lo.y + b.lo.y) + c.lo.y;
            a4.z += (a.lo.z + b.lo.z) + c.lo.z; a4.w += (a.lo.w + b.lo.w) + c.lo.w;
            b4.x += (a.hi.x + b.hi.x) + c.hi.x; b4.y += (a.hi.y + b.hi.y) + c.hi.y;
            b4.z += (a.hi.z + b.hi.z) + c.hi.z; b4.w += (a.hi.w + b.hi.w) + c.hi.w;
        }
        acc = ((a4.x + a4.y) + (a4.z + a4.w)) + ((b4.x + b4.y) + (b4.z + b4.w));
    } else if (vec) {
        const int64_t n4 = hw >> 2;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)nblk * kThreads) {
            const float4 a = ld_stream4(p0 + 4 * i), b = ld_stream4(p1 + 4 * i), c = ld_stream4(p2 + 4 * i);
            a4.x += (a.x + b.x) + c.x; a4.y += (a.y + b.y) + c.y;
            a4.z += (a.z + b.z) + c.z; a4.w += (a.w + b.w) + c.w;
        }
        acc = (a4.x + a4.y) + (a4.z + a4.w);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < hw; i += (int64_t)nblk * kThreads)
            acc += (p0[i] + p1[i]) + p2[i];
    }
    double d = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    __shared__ double ws[kThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kThreads / 32; ++i) t += ws[i];
        partial[plane * nblk + blockIdx.x] = t;
    }
}

// grid (nblk_apply, C, B): CTA = a slice of plane (b, c)
__global__ void __launch_bounds__(kThreads)
skff_apply_kernel(const float *__restrict__ f0, const float *__restrict__ f1,
                  const float *__restrict__ f2, const double *__restrict__ partial, int nblk_pool,
                  const float *__restrict__ w_du, const float *__restrict__ prelu,
                  const float *__restrict__ w_fc0, const float *__restrict__ w_fc1,
                  const float *__restrict__ w_fc2, float *__restrict__ out, int64_t hw, int vec)
{
    __shared__ float s_mean[kC];
    __shared__ float s_z[kMid];
    __shared__ float s_att[3];
    const int c = blockIdx.y, b = blockIdx.z;
    if (threadIdx.x < kC) {
        const double *pp = partial + ((int64_t)b * kC + threadIdx.x) * nblk_pool;
        double t = 0.0;
        for (int i = 0; i < nblk_pool; ++i) t += pp[i];
        s_mean[threadIdx.x] = (float)(t / (double)hw);        // AdaptiveAvgPool2d(1)  (:948)
    }
    __syncthreads();
    if (threadIdx.x < kMid) {                                  // conv_du: 1x1 32->4, PReLU  (:949)
        float z = 0.0f;
        for (int i = 0; i < kC; ++i) z = fmaf(__ldg(w_du + threadIdx.x * kC + i), s_mean[i], z);
        const float slope = __ldg(prelu);
        s_z[threadIdx.x] = z >= 0.0f ? z : slope * z;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                    // fcs + softmax over the bands (:951-955)
        float a[3];
        const float *wf[3] = {w_fc0, w_fc1, w_fc2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < kMid; ++j) v = fmaf(__ldg(wf[k] + c * kMid + j), s_z[j], v);
            a[k] = v;
        }
        const float m = fmaxf(a[0], fmaxf(a[1], a[2]));
        const float e0 = expf(a[0] - m), e1 = expf(a[1] - m), e2 = expf(a[2] - m);
        const float inv = 1.0f / ((e0 + e1) + e2);
        s_att[0] = e0 * inv; s_att[1] = e1 * inv; s_att[2] = e2 * inv;
    }
    __syncthreads();
    const float a0 = s_att[0], a1 = s_att[1], a2 = s_att[2];
    const int64_t plane = (int64_t)b * kC + c;
    const float *p0 = f0 + plane * hw, *p1 = f1 + plane * hw, *p2 = f2 + plane * hw;
    float *po = out + plane * hw;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    // V = (f0 a0 + f1 a1) + f2 a2 with separate multiplies and adds, as torch.sum(x * a, dim=1)  (:957)
    auto mix = [&](const float4 &x, const float4 &y, const float4 &z) {
        float4 r;
        r.x = __fadd_rn(__fadd_rn(__fmul_rn(x.x, a0), __fmul_rn(y.x, a1)), __fmul_rn(z.x, a2));
        r.y = __fadd_rn(__fadd_rn(__fmul_rn(x.y, a0), __fmul_rn(y.y, a1)), __fmul_rn(z.y, a2));
        r.z = __fadd_rn(__fadd_rn(__fmul_rn(x.z, a0), __fmul_rn(y.z, a1)), __fmul_rn(z.z, a2));
        r.w = __fadd_rn(__fadd_rn(__fmul_rn(x.w, a0), __fmul_rn(y.w, a1)), __fmul_rn(z.w, a2));
        return r;
    };
    if (vec == 2) {
        const int64_t n8 = hw >> 3;
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n8; i += stride) {
            const float8 x = ld_stream8(p0 + 8 * i), y = ld_stream8(p1 + 8 * i), z = ld_stream8(p2 + 8 * i);
            const float4 rl = mix(x.lo, y.lo, z.lo), rh = mix(x.hi, y.hi, z.hi);
            // plain (not evict-first) store: the next kernel reads this map
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(po + 8 * i), "f"(rl.x), "f"(rl.y),
                         "f"(rl.z), "f"(rl.w), "f"(rh.x), "f"(rh.y), "f"(rh.z), "f"(rh.w)
                         : "memory");
        }
    } else if (vec) {
        const int64_t n4 = hw >> 2;
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
            const float4 x = ld_stream4(p0 + 4 * i), y = ld_stream4(p1 + 4 * i), z = ld_stream4(p2 + 4 * i);
            float4 r;
            r.x = __fadd_rn(__fadd_rn(__fmul_rn(x.x, a0), __fmul_rn(y.x, a1)), __fmul_rn(z.x, a2));
            r.y = __fadd_rn(__fadd_rn(__fmul_rn(x.y, a0), __fmul_rn(y.y, a1)), __fmul_rn(z.y, a2));
            r.z = __fadd_rn(__fadd_rn(__fmul_rn(x.z, a0), __fmul_rn(y.z, a1)), __fmul_rn(z.z, a2));
            r.w = __fadd_rn(__fadd_rn(__fmul_rn(x.w, a0), __fmul_rn(y.w, a1)), __fmul_rn(z.w, a2));
            *reinterpret_cast<float4 *>(po + 4 * i) = r;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < hw; i += stride)
            po[i] = __fadd_rn(__fadd_rn(__fmul_rn(p0[i], a0), __fmul_rn(p1[i], a1)), __fmul_rn(p2[i], a2));
    }
}

// ---------------------------------------------------------------------------------------------
// PixelUnshuffle(R) + 1x1 conv 3R^2 -> 32 (+bias).  Persistent CTAs walk tiles of 256 output pixels of
// one output row; a step is G input row segments (ci, dy) of the tile, 256*R floats each (G = 2, 4, 6 for
// R = 8, 4, 2: 384-512 FMAs per thread between barriers), loaded coalesced into registers one step ahead
// (across tile boundaries) and parked in shared memory as [segment][dx quad][pixel],
// so the compute threads read them conflict-free.  thread = 4 pixels (pg, pg+64, ...) x 8 output
// channels: every broadcast weight quad feeds 16 FMAs (a broadcast LDS.128 returns 512 bytes to the warp,
// DESIGN.md 4.3), the input is read once for all 32 outputs, and the weights are transposed once per CTA.
// Round 2's first form (thread = pixel x 8 channels, one weight quad per FMA quad, x read by four channel
// groups) sat at 0.10-0.12 ms per call at 4K whatever R: L1 tag-stage bound.
// ---------------------------------------------------------------------------------------------
constexpr int kPsTile = 256;                      // output pixels per tile

// packed FMA (two fp32 FMAs per issue slot; a 3-register FFMA issues every other cycle)
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

template <int R> struct PsCfg {
    static constexpr int G = R == 8 ? 2 : R == 4 ? 4 : 6;                 // row segments per step
    static constexpr int QN = R >= 4 ? R / 4 : 1;                         // float4 per pixel and row (R = 2: half a float4)
    static constexpr int QPITCH = kPsTile + 4;                            // float4 pitch of a dx quad: the two quads of R = 8 land in different banks
    static constexpr int SEG = R >= 4 ? QN * QPITCH : kPsTile / 2;        // float4 per parked segment
    static constexpr int BUF = G * SEG;                                   // float4 per staging buffer
    static constexpr size_t kSmem = sizeof(float) * (3 * R * R * 32 + 32) + sizeof(float4) * 2 * BUF;
};

template <int R>
__global__ void __launch_bounds__(kThreads, 2)
ps_down_kernel(const float *__restrict__ x, const float *__restrict__ wgt,
               const float *__restrict__ bias, float *__restrict__ y, int B, int H, int W)
{
    // unshuffled channel index = ci * R*R + dy * R + dx   (torch.nn.PixelUnshuffle)
    static_assert(kThreads == 256, "thread mapping");
    constexpr int CIN = 3 * R * R;
    using C = PsCfg<R>;
    constexpr int G = C::G, QN = C::QN, QPITCH = C::QPITCH, SEG = C::SEG, BUF = C::BUF;
    constexpr int STEPS = 3 * R / G;              // steps per tile
    static_assert(STEPS * G == 3 * R, "G divides 3R");
    constexpr int F4 = kPsTile * R / 4;           // float4 per row segment
    constexpr int NLD = G * F4 / kThreads;        // float4 per thread and step: 4, 4, 3
    static_assert(NLD * kThreads == G * F4, "whole loads");
    extern __shared__ float4 ps_smem[];
    float *ws = reinterpret_cast<float *>(ps_smem);           // [CIN][32] transposed weights, then [32] bias
    float4 *stage = ps_smem + (CIN * 32 + 32) / 4;            // [2][BUF]
    {
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        for (int k = wp; k < CIN; k += kThreads / 32) ws[k * 32 + lane] = __ldg(wgt + lane * CIN + k);
        if (threadIdx.x < 32) ws[CIN * 32 + threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.0f;
    }
    const int h = H / R, w = W / R;
    const int tpr = (w + kPsTile - 1) / kPsTile;  // tiles per output row
    const int64_t hw = (int64_t)h * w, HW = (int64_t)H * W;
    const int64_t ntiles = (int64_t)B * h * tpr;
    const int pg = threadIdx.x & 63, cog = threadIdx.x >> 6;  // a warp = 32 pixel groups of ONE channel group

    float4 ld[NLD];
    // the loads of step `st` of tile `t` (zeros past the row end / past the last tile)
    auto fetch = [&](int64_t t, int st) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) ld[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= ntiles) return;
        const int tx = (int)(t % tpr);
        const int64_t rowi = t / tpr;
        const int oy = (int)(rowi % h), b = (int)(rowi / h);
        const int64_t col0 = (int64_t)tx * kPsTile * R;       // first input column of the segments
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const int kk = threadIdx.x + i * kThreads;
            const int sg = st * G + kk / F4, k = kk % F4;     // segment (ci, dy), float4 inside it
            const int ci = sg / R, dy = sg - ci * R;
            const float *row = x + ((int64_t)b * 3 + ci) * HW + (int64_t)(oy * R + dy) * W + col0;
            if (col0 + 4 * k < W) ld[i] = ld_stream4(row + 4 * k);
        }
    };
    auto park = [&](int buf) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const int kk = threadIdx.x + i * kThreads;
            const int g = kk / F4, k = kk % F4;
            if (R >= 4) stage[buf * BUF + g * SEG + (k % QN) * QPITCH + k / QN] = ld[i];
            else stage[buf * BUF + g * SEG + k] = ld[i];      // R = 2: the row as it is, a float2 per pixel
        }
    };

    int64_t tile = blockIdx.x;
    fetch(tile, 0);
    park(0);
    __syncthreads();
    int buf = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        f32x2 acc[4][4];                            // [pixel][channel pair]
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                acc[p][j] = pack2(ws[CIN * 32 + cog * 8 + 2 * j], ws[CIN * 32 + cog * 8 + 2 * j + 1]);
#pragma unroll 1
        for (int st = 0; st < STEPS; ++st) {
            // next segment in flight during this one's FMAs
            if (st + 1 < STEPS) fetch(tile, st + 1);
            else fetch(tile + gridDim.x, 0);
#pragma unroll
            for (int g = 0; g < G; ++g) {
            const float4 *sb = stage + buf * BUF + g * SEG;
            const float *wrow = ws + (st * G + g) * R * 32 + cog * 8;
#pragma unroll
            for (int q = 0; q < QN; ++q) {
                float v[4][4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (R >= 4) {
                        const float4 t = sb[q * QPITCH + pg + 64 * p];
                        v[p][0] = t.x; v[p][1] = t.y; v[p][2] = t.z; v[p][3] = t.w;
                    } else {
                        const float2 t = reinterpret_cast<const float2 *>(sb)[pg + 64 * p];
                        v[p][0] = t.x; v[p][1] = t.y; v[p][2] = 0.f; v[p][3] = 0.f;
                    }
                }
#pragma unroll
                for (int d = 0; d < (R >= 4 ? 4 : 2); ++d) {
                    const ulonglong2 w0 = *reinterpret_cast<const ulonglong2 *>(wrow + (4 * q + d) * 32);
                    const ulonglong2 w1 = *reinterpret_cast<const ulonglong2 *>(wrow + (4 * q + d) * 32 + 4);
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const f32x2 t = pack2(v[p][d], v[p][d]);
                        acc[p][0] = ffma2(t, w0.x, acc[p][0]); acc[p][1] = ffma2(t, w0.y, acc[p][1]);
                        acc[p][2] = ffma2(t, w1.x, acc[p][2]); acc[p][3] = ffma2(t, w1.y, acc[p][3]);
                    }
                }
            }
            }
            park(buf ^ 1);          // read by the step before this one: every thread is past that barrier
            __syncthreads();
            buf ^= 1;
        }
        const int tx = (int)(tile % tpr);
        const int64_t rowi = tile / tpr;
        const int oy = (int)(rowi % h), b = (int)(rowi / h);
        float *yo = y + ((int64_t)b * 32 + cog * 8) * hw + (int64_t)oy * w + tx * kPsTile;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int px = pg + 64 * p;
            if (tx * kPsTile + px < w)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float lo, hi;
                    unpack2(acc[p][j], lo, hi);
                    yo[(int64_t)(2 * j) * hw + px] = lo;
                    yo[(int64_t)(2 * j + 1) * hw + px] = hi;
                }
        }
    }
}

template <int R>
int launch_ps_down(const float *x, const float *wgt, const float *bias, float *y, int64_t B,
                   int64_t H, int64_t W, cudaStream_t s)
{
    const size_t smem = PsCfg<R>::kSmem;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        WM_CUDA_OK(cudaGetDevice(&dev));
        WM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    WM_CUDA_OK(cudaFuncSetAttribute(ps_down_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = B * (H / R) * ((W / R + kPsTile - 1) / kPsTile);
    const int grid = (int)(tiles < 2 * sms ? tiles : 2 * sms);
    ps_down_kernel<R><<<grid, kThreads, smem, s>>>(x, wgt, bias, y, (int)B, (int)H, (int)W);
    WM_LAUNCH_OK("ps_down");
    return WM_OK;
}

int pool_blocks(int64_t hw)
{
    // enough CTAs per plane to fill the machine at B*C = 32 planes, never more than the data needs
    int64_t want = (hw / 4 + kThreads * 4 - 1) / (kThreads * 4);
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return (int)want;
}

static int band_vec(const float *f0, const float *f1, const float *f2, const float *out, int64_t hw)
{
    int vec = (hw % 4 == 0 && aligned16(f0) && aligned16(f1) && aligned16(f2) && (!out || aligned16(out))) ? 1 : 0;
    if (vec && hw % 8 == 0 && aligned32(f0) && aligned32(f1) && aligned32(f2) && (!out || aligned32(out))) vec = 2;
    return vec;
}

// pool pass alone (used by wm_dwt_haar_pool_fwd when the fused kernel's preconditions do not hold)
int launch_pool(const float *f0, const float *f1, const float *f2, double *partial, int64_t planes, int64_t hw,
                cudaStream_t s)
{
    skff_pool_kernel<<<dim3(pool_blocks(hw), (unsigned)planes), kThreads, 0, s>>>(f0, f1, f2, partial, hw,
                                                                                band_vec(f0, f1, f2, nullptr, hw));
    WM_LAUNCH_OK("skff pool");
    return WM_OK;
}

}  // namespace skff
}  // namespace wm

extern "C" size_t wm_skff_workspace_bytes(int64_t B, int64_t h, int64_t w)
{
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)B * wm::skff::kC * wm::skff::pool_blocks(h * w) * sizeof(double);
}

static int skff_run(bool pooled, const float *f0, const float *f1, const float *f2, const float *w_du,
                    const float *prelu_weight, const float *w_fc0, const float *w_fc1,
                    const float *w_fc2, float *out, void *workspace, size_t workspace_bytes,
                    int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::skff;
    WM_REQUIRE(B >= 0 && h >= 0 && w >= 0 && B <= 65535, "wm_skff_fwd: bad sizes");
    WM_REQUIRE(C == kC, "wm_skff_fwd: C=%lld unsupported (32)", (long long)C);
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(f0 && f1 && f2 && w_du && prelu_weight && w_fc0 && w_fc1 && w_fc2 && out,
               "wm_skff_fwd: null pointer");
    const int64_t hw = h * w;
    const int nblk = pool_blocks(hw);
    WM_REQUIRE(workspace && workspace_bytes >= (size_t)B * kC * nblk * sizeof(double),
               "wm_skff_fwd: workspace too small");
    WM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "wm_skff_fwd: workspace must be 8-byte aligned");
    const int vec = band_vec(f0, f1, f2, out, hw);
    cudaStream_t s = (cudaStream_t)stream;
    double *partial = static_cast<double *>(workspace);
    if (!pooled) {
        skff_pool_kernel<<<dim3(nblk, (unsigned)(B * kC)), kThreads, 0, s>>>(f0, f1, f2, partial, hw, vec);
        WM_LAUNCH_OK("skff pool");
    }
    skff_apply_kernel<<<dim3(nblk, kC, (unsigned)B), kThreads, 0, s>>>(f0, f1, f2, partial, nblk, w_du,
                                                                        prelu_weight, w_fc0, w_fc1,
                                                                        w_fc2, out, hw, vec);
    WM_LAUNCH_OK("skff apply");
    return WM_OK;
}

extern "C" int wm_skff_fwd(const float *f0, const float *f1, const float *f2, const float *w_du,
                           const float *prelu_weight, const float *w_fc0, const float *w_fc1,
                           const float *w_fc2, float *out, void *workspace, size_t workspace_bytes,
                           int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream)
{
    return skff_run(false, f0, f1, f2, w_du, prelu_weight, w_fc0, w_fc1, w_fc2, out, workspace, workspace_bytes, B, C,
                    h, w, stream);
}

extern "C" int wm_skff_apply_fwd(const float *f0, const float *f1, const float *f2, const float *w_du,
                                 const float *prelu_weight, const float *w_fc0, const float *w_fc1,
                                 const float *w_fc2, float *out, const void *pool_partials, size_t workspace_bytes,
                                 int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream)
{
    return skff_run(true, f0, f1, f2, w_du, prelu_weight, w_fc0, w_fc1, w_fc2, out, const_cast<void *>(pool_partials),
                    workspace_bytes, B, C, h, w, stream);
}

extern "C" int wm_ps_down_fwd(const float *x, const float *weight, const float *bias, float *y,
                              int64_t B, int64_t H, int64_t W, int r, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(B >= 0 && H >= 0 && W >= 0 && B <= 65535, "wm_ps_down_fwd: bad sizes");
    WM_REQUIRE(r == 2 || r == 4 || r == 8, "wm_ps_down_fwd: r=%d unsupported (2, 4, 8)", r);
    WM_REQUIRE(H % r == 0 && W % r == 0, "wm_ps_down_fwd: H and W must be multiples of r");
    if (B == 0 || H == 0 || W == 0) return WM_OK;
    WM_REQUIRE(x && weight && y, "wm_ps_down_fwd: null pointer");
    WM_REQUIRE(aligned16(x) && W % 4 == 0, "wm_ps_down_fwd: x must be 16-byte aligned with W %% 4 == 0");
    cudaStream_t s = (cudaStream_t)stream;
    if (r == 2) return skff::launch_ps_down<2>(x, weight, bias, y, B, H, W, s);
    if (r == 4) return skff::launch_ps_down<4>(x, weight, bias, y, B, H, W, s);
    return skff::launch_ps_down<8>(x, weight, bias, y, B, H, W, s);
}
