// Three small spatial kernels on the packed FP32 pipe (Blackwell FFMA2: two fp32 FMAs per lane and
// instruction; a three-register scalar FFMA issues every other cycle, so the scalar versions of these
// kernels were bound by the FMA pipe at 25-30 % of the HBM roofline):
//   wm_dw_act_pw_fwd     y = residual? + pw1x1( gelu?( dw3x3(x) ) ), C = 32   (FeedForward.project_out,
//                        reference wavemamba_arch.py:739-742,851)
//   wm_stem_conv3x3_fwd  UNet.conv_01, 3 -> 32                               (:1026,1048)
//   wm_head_conv3x3_fwd  UNet.last, 32 -> 3, + the global residual            (:1039,1061)
// NCHW float32; halo tiles in shared memory [channel][row][col]; lanes run along the image row.  After
// the packing these kernels are bound by the RETURN PATH of their shared-memory loads (128 B/clk into
// the register file: a broadcast LDS.128 of weights still delivers 512 bytes to the warp), so a loaded
// weight quad is reused for four pixels where the registers allow it.
#include <stdlib.h>

#include "tma.cuh"

namespace wm {
namespace sp32 {

typedef unsigned long long f32x2;      // packed fp32 pair
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
__device__ __forceinline__ float ex2_approx(float v)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// Exact (erf) GELU of two values: erf(t) = 1 - 2^(-t P(t)), t = min(|x| / sqrt 2, 4), P the degree-7
// weighted-minimax fit of -log2(erfc(t)) / t on [0, 4] (absolute error of erf <= 1.1e-7 in fp32, of the
// GELU <= 5e-7 at |x| ~ 4.4 and relative 1e-7 near 0; beyond t = 4 erf is 1 to fp32).  One MUFU and ~12
// FMA-pipe instructions per value, half of them packed; erff costs ~30.
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
    const float r0 = copysignf(1.0f - ex2_approx(-g0), x0), r1 = copysignf(1.0f - ex2_approx(-g1), x1);
    const float h0 = 0.5f * x0, h1 = 0.5f * x1;
    return pack2(fmaf(h0, r0, h0), fmaf(h1, r1, h1));
}

constexpr int kTH = 8;                 // tile rows
constexpr int kHH = kTH + 2;

// Halo tile of image b at (ty0-1, tx0-1), TW+2 columns, into xs[c * CS + row * PITCH + col]; zeros outside
// the image.  4-byte cp.async with zero fill (every load of the tile in flight at once); warp w takes
// channels w, w+nwarps, ...; lanes run along the row.  Row offsets, row validity and the column
// predicates are the same for every channel and computed once (the per-element index arithmetic is
// otherwise a large share of these kernels' instructions).
template <int TW, int PITCH, int CS>
__device__ __forceinline__ void load_halo(const float *__restrict__ x, float *xs, int64_t b, int C, int h,
                                          int w, int ty0, int tx0, int nthreads)
{
    constexpr int NCH = (TW + 2 + 31) / 32;            // 32-column chunks of a halo row
    const int64_t hw = (int64_t)h * w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
    const uint32_t xs_base = (uint32_t)__cvta_generic_to_shared(xs);
    const float *xb = x + (int64_t)b * C * hw;
    int64_t roff[kHH];
    uint32_t okrows = 0u;
#pragma unroll
    for (int py = 0; py < kHH; ++py) {
        const int gy = ty0 - 1 + py;
        const bool oky = gy >= 0 && gy < h;
        roff[py] = oky ? (int64_t)gy * w : 0;
        okrows |= oky ? (1u << py) : 0u;
    }
    bool okx[NCH], inrow[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int pc = 32 * k + lane, gx = tx0 - 1 + pc;
        inrow[k] = pc < TW + 2;
        okx[k] = inrow[k] && gx >= 0 && gx < w;
    }
#pragma unroll 1
    for (int c = warp; c < C; c += nwarps) {
        const float *plane = xb + (int64_t)c * hw + (tx0 - 1 + lane);
        const uint32_t dst_c = xs_base + (uint32_t)(c * CS + lane) * 4u;
#pragma unroll
        for (int py = 0; py < kHH; ++py) {
            const bool oky = (okrows >> py) & 1u;
            const float *row = plane + roff[py];
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if (k < NCH - 1 || inrow[k]) {
                    const bool ok = oky && okx[k];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_c + (uint32_t)(py * PITCH + 32 * k) * 4u),
                                 "l"(ok ? row + 32 * k : xb), "r"(ok ? 4u : 0u)
                                 : "memory");
                }
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void halo_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// y = residual? + pw1x1( act( dw3x3(x) ) ),  C = 32, 8x32-pixel tiles, 256 threads
// ---------------------------------------------------------------------------------------------
constexpr int kDwTW = 32, kDwPitch = 36, kDwCS = kHH * kDwPitch;   // 360 floats per channel
constexpr int kDwThreads = 256, kDwPix = kTH * kDwTW;
constexpr size_t kDwSmem = sizeof(float) * (32 * kDwCS + 32 * kDwPix + 32 * 32 + 32 + 32 * 12);

template <bool GELU>
__global__ void __launch_bounds__(kDwThreads, 2)
dw_act_pw_kernel(const float *__restrict__ x, const float *__restrict__ dw_w, const float *__restrict__ dw_b,
                 const float *__restrict__ pw_w, const float *__restrict__ pw_b,
                 const float *__restrict__ residual, float *__restrict__ y, int h, int w)
{
    constexpr int C = 32;
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                  // [32][10][36] halo tile
    float *ds = xs + C * kDwCS;        // [32][256] after dw + act
    float *wt = ds + C * kDwPix;       // [32 ci][32 co] transposed 1x1 weights
    float *pb = wt + C * C;            // [32]
    float *dwk = pb + C;               // [32][12]: 9 taps, bias, 2 pad

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * kDwTW, ty0 = blockIdx.y * kTH;
    const int64_t b = blockIdx.z;
    const int64_t hw = (int64_t)h * w;

    for (int i = tid; i < C * C; i += kDwThreads) {
        const int co = i / C, ci = i - co * C;
        wt[ci * C + co] = __ldg(pw_w + i);
    }
    if (tid < C) pb[tid] = __ldg(pw_b + tid);
    for (int i = tid; i < C * 12; i += kDwThreads) {
        const int c = i / 12, t = i - c * 12;
        dwk[i] = t < 9 ? __ldg(dw_w + c * 9 + t) : (t == 9 ? __ldg(dw_b + c) : 0.0f);
    }
    load_halo<kDwTW, kDwPitch, kDwCS>(x, xs, b, C, h, w, ty0, tx0, kDwThreads);
    halo_wait();
    __syncthreads();

    // depthwise 3x3 (+ GELU): thread = (channel, 4 adjacent columns), all 8 rows with a sliding window of
    // three input rows held as packed pairs; per input row one LDS.128 + one LDS.64
    {
        const int c = tid >> 3, j4 = (tid & 7) * 4;
        const float4 *kp = reinterpret_cast<const float4 *>(dwk + c * 12);
        const float4 ka = kp[0], kb = kp[1], kc = kp[2];
        const f32x2 k2[9] = {pack2(ka.x, ka.x), pack2(ka.y, ka.y), pack2(ka.z, ka.z), pack2(ka.w, ka.w),
                             pack2(kb.x, kb.x), pack2(kb.y, kb.y), pack2(kb.z, kb.z), pack2(kb.w, kb.w),
                             pack2(kc.x, kc.x)};
        const f32x2 bias2 = pack2(kc.y, kc.y);
        const float *pc = xs + c * kDwCS + j4;
        // packed pairs of one input row: P0=(v0,v1) P1=(v2,v3) P2=(v4,v5) Q0=(v1,v2) Q1=(v3,v4)
        f32x2 P[3][3], Q[3][2];
        auto load_row = [&](int r, int slot) {
            const float4 a4 = *reinterpret_cast<const float4 *>(pc + r * kDwPitch);
            const float2 b2 = *reinterpret_cast<const float2 *>(pc + r * kDwPitch + 4);
            P[slot][0] = pack2(a4.x, a4.y); P[slot][1] = pack2(a4.z, a4.w); P[slot][2] = pack2(b2.x, b2.y);
            Q[slot][0] = pack2(a4.y, a4.z); Q[slot][1] = pack2(a4.w, b2.x);
        };
        load_row(0, 0);
        load_row(1, 1);
        float *dp = ds + c * kDwPix + j4;
#pragma unroll
        for (int row = 0; row < kTH; ++row) {
            load_row(row + 2, (row + 2) % 3);
            f32x2 oa = bias2, ob = bias2;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int sl = (row + dy) % 3;
                oa = ffma2(k2[3 * dy + 0], P[sl][0], oa); ob = ffma2(k2[3 * dy + 0], P[sl][1], ob);
                oa = ffma2(k2[3 * dy + 1], Q[sl][0], oa); ob = ffma2(k2[3 * dy + 1], Q[sl][1], ob);
                oa = ffma2(k2[3 * dy + 2], P[sl][1], oa); ob = ffma2(k2[3 * dy + 2], P[sl][2], ob);
            }
            if (GELU) { oa = gelu2(oa); ob = gelu2(ob); }
            float o0, o1, o2, o3;
            unpack2(oa, o0, o1);
            unpack2(ob, o2, o3);
            *reinterpret_cast<float4 *>(dp + row * kDwTW) = make_float4(o0, o1, o2, o3);
        }
    }
    __syncthreads();

    // 1x1: thread = interior pixel, 32 outputs as 16 packed pairs (co, co+1); the weight row of an input
    // channel comes in as 8 broadcast LDS.128 = 16 ready-made pairs
    f32x2 acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = reinterpret_cast<const f32x2 *>(pb)[j];
#pragma unroll 4
    for (int ci = 0; ci < C; ++ci) {
        const float xv = ds[ci * kDwPix + tid];
        const f32x2 xv2 = pack2(xv, xv);
        const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(wt + ci * C);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 wv = wr[j];
            acc[2 * j] = ffma2(xv2, wv.x, acc[2 * j]);
            acc[2 * j + 1] = ffma2(xv2, wv.y, acc[2 * j + 1]);
        }
    }
    const int gy = ty0 + (tid >> 5), gx = tx0 + (tid & 31);
    if (gy < h && gx < w) {
        const int64_t o = (int64_t)b * C * hw + (int64_t)gy * w + gx;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v0, v1;
            unpack2(acc[j], v0, v1);
            if (residual != nullptr) {
                v0 += __ldg(residual + o + (2 * j) * hw);
                v1 += __ldg(residual + o + (2 * j + 1) * hw);
            }
            y[o + (2 * j) * hw] = v0;
            y[o + (2 * j + 1) * hw] = v1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same op as a persistent, warp-specialised, TMA-fed pipeline (one CTA of 512 threads per SM).  The
// cp.async form above is bound by exposed load latency (ncu: 44 % long-scoreboard stalls, no pipe above
// 30 %) and runs its two phases -- bound by different resources -- one after the other.  Here:
//   thread 0     requests the halo box of the tile after next as soon as the depthwise warps have consumed
//                a buffer (cp.async.bulk.tensor 40 x 10 x 32 channels at (tx0-4, ty0-1): halo column j is
//                box column j+3 -- the innermost start coordinate must be a multiple of 16 bytes; zero
//                fill outside the image), two buffers, one mbarrier each
//   warps 0-7    depthwise 3x3 + GELU: thread = (channel, 4 columns), 8 rows with a sliding window, packed
//                FFMA2 -> ds[tile & 1][32][256]
//   warps 8-15   1x1 of the PREVIOUS tile: thread = (4 adjacent pixels, 8 of the 32 outputs); per input
//                channel two LDS.128 of weights and one of pixels feed 16 FFMA2 (the return path of the
//                shared-memory loads bounds this phase); the residual quads are fetched from global
//                memory before the FMAs; 16-byte stores
// The two groups hand the ds buffers over with named barriers (bar.arrive / bar.sync).  Needs w % 4 == 0
// and 16-byte aligned tensors.
// ---------------------------------------------------------------------------------------------
constexpr int kTmBoxW = 40, kTmCS = kHH * kTmBoxW;          // 400 floats per channel
constexpr int kTmThreads = 512, kTmGroup = 256;
constexpr uint32_t kTmXBytes = 32 * kTmCS * 4;
constexpr size_t kTmSmem = 2 * kTmXBytes + sizeof(float) * (2 * 32 * kDwPix + 32 * 32 + 32 + 32 * 12) + 2 * 8;

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <bool GELU, bool RES>
__global__ void __launch_bounds__(kTmThreads, 1)
dw_act_pw_tma_kernel(const __grid_constant__ CUtensorMap xmap, const float *__restrict__ dw_w,
                     const float *__restrict__ dw_b, const float *__restrict__ pw_w,
                     const float *__restrict__ pw_b, const float *__restrict__ residual, float *__restrict__ y,
                     int h, int w, int tiles_x, int tiles_y, int total_tiles)
{
    using namespace wm::tc5;
    constexpr int C = 32;
    constexpr int kBarFull = 1, kBarEmpty = 3, kBarDw = 5;                 // named barriers: +buffer index
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);                       // [2][32][10][40]
    float *ds = reinterpret_cast<float *>(smem_raw + 2 * kTmXBytes);       // [2][32][256] after dw + act
    float *wt = ds + 2 * C * kDwPix;                                       // [32 ci][32 co]
    float *pb = wt + C * C;
    float *dwk = pb + C;                                                   // [32][12]
    const uint32_t bar0 = smem_u32(dwk + C * 12);                          // full[2] of the halo buffers
    const int tid = threadIdx.x;
    const int64_t hw = (int64_t)h * w;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto issue = [&](int j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const uint32_t bar = bar0 + 8u * (uint32_t)(j & 1);
        mbar_expect_tx(bar, kTmXBytes);
        tma::load_box(smem_u32(xs) + (uint32_t)(j & 1) * kTmXBytes, &xmap, txi * kDwTW - 4, tyi * kTH - 1, 0, b, bar);
    };
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (my_tiles > 0) issue(0);
        if (my_tiles > 1) issue(1);
    }
    for (int i = tid; i < C * C; i += kTmThreads) {
        const int co = i / C, ci = i - co * C;
        wt[ci * C + co] = __ldg(pw_w + i);
    }
    if (tid < C) pb[tid] = __ldg(pw_b + tid);
    for (int i = tid; i < C * 12; i += kTmThreads) {
        const int c = i / 12, t = i - c * 12;
        dwk[i] = t < 9 ? __ldg(dw_w + c * 9 + t) : (t == 9 ? __ldg(dw_b + c) : 0.0f);
    }
    __syncthreads();

    if (tid < kTmGroup) {
        // =========================== depthwise 3x3 (+ GELU) ====================================
        const int dc = tid >> 3, j4 = (tid & 7) * 4;
        f32x2 k2[9];
        f32x2 bias2;
        {
            const float4 *kp = reinterpret_cast<const float4 *>(dwk + dc * 12);
            const float4 ka = kp[0], kb = kp[1], kc = kp[2];
            k2[0] = pack2(ka.x, ka.x); k2[1] = pack2(ka.y, ka.y); k2[2] = pack2(ka.z, ka.z);
            k2[3] = pack2(ka.w, ka.w); k2[4] = pack2(kb.x, kb.x); k2[5] = pack2(kb.y, kb.y);
            k2[6] = pack2(kb.z, kb.z); k2[7] = pack2(kb.w, kb.w); k2[8] = pack2(kc.x, kc.x);
            bias2 = pack2(kc.y, kc.y);
        }
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)(j >> 1) & 1u);
            if (j >= 2) bar_sync(kBarEmpty + buf, kTmThreads);          // the 1x1 warps are done with ds[buf]
            // halo columns j4 .. j4+5 = box columns j4+3 .. j4+8: a scalar, an aligned LDS.128, a scalar
            const float *pc = xs + buf * (C * kTmCS) + dc * kTmCS + j4 + 3;
            f32x2 P[3][3], Q[3][2];
            auto load_row = [&](int r, int slot) {
                const float v0 = pc[r * kTmBoxW];
                const float4 a4 = *reinterpret_cast<const float4 *>(pc + r * kTmBoxW + 1);
                const float v5 = pc[r * kTmBoxW + 5];
                P[slot][0] = pack2(v0, a4.x); P[slot][1] = pack2(a4.y, a4.z); P[slot][2] = pack2(a4.w, v5);
                Q[slot][0] = pack2(a4.x, a4.y); Q[slot][1] = pack2(a4.z, a4.w);
            };
            load_row(0, 0);
            load_row(1, 1);
            float *dp = ds + buf * (C * kDwPix) + dc * kDwPix + j4;
#pragma unroll
            for (int row = 0; row < kTH; ++row) {
                load_row(row + 2, (row + 2) % 3);
                f32x2 oa = bias2, ob = bias2;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int sl = (row + dy) % 3;
                    oa = ffma2(k2[3 * dy + 0], P[sl][0], oa); ob = ffma2(k2[3 * dy + 0], P[sl][1], ob);
                    oa = ffma2(k2[3 * dy + 1], Q[sl][0], oa); ob = ffma2(k2[3 * dy + 1], Q[sl][1], ob);
                    oa = ffma2(k2[3 * dy + 2], P[sl][1], oa); ob = ffma2(k2[3 * dy + 2], P[sl][2], ob);
                }
                if (GELU) { oa = gelu2(oa); ob = gelu2(ob); }
                float o0, o1, o2, o3;
                unpack2(oa, o0, o1);
                unpack2(ob, o2, o3);
                *reinterpret_cast<float4 *>(dp + row * kDwTW) = make_float4(o0, o1, o2, o3);
            }
            bar_arrive(kBarFull + buf, kTmThreads);                     // ds[buf] is complete
            bar_sync(kBarDw, kTmGroup);                                 // every depthwise thread has left xs[buf]
            if (tid == 0 && j + 2 < my_tiles) issue(j + 2);
        }
    } else {
        // =========================== 1x1 + residual ============================================
        const int t = tid - kTmGroup;
        const int q = t & 3, pq = t >> 2;                 // output quarter, pixel quad 0..63
        f32x2 pbq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pbq[i] = reinterpret_cast<const f32x2 *>(pb)[q * 4 + i];
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            const int tile = blockIdx.x + j * gridDim.x;
            const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
            const int gy = tyi * kTH + (pq >> 3), gx = txi * kDwTW + (pq & 7) * 4;
            const bool inside = gy < h && gx < w;         // w % 4 == 0: the four pixels are inside or outside together
            const int64_t o = (int64_t)b * C * hw + (int64_t)gy * w + gx + (int64_t)(q * 8) * hw;
            float4 rv[8];
            if (RES && inside) {
#pragma unroll
                for (int i = 0; i < 8; ++i) rv[i] = __ldg(reinterpret_cast<const float4 *>(residual + o + (int64_t)i * hw));
            }
            bar_sync(kBarFull + buf, kTmThreads);                       // ds[buf] is complete
            f32x2 acc[4][4];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[p][i] = pbq[i];
            const float *dsb = ds + buf * (C * kDwPix) + pq * 4;
#pragma unroll 4
            for (int ci = 0; ci < C; ++ci) {
                const float4 xv = *reinterpret_cast<const float4 *>(dsb + ci * kDwPix);
                const f32x2 x2[4] = {pack2(xv.x, xv.x), pack2(xv.y, xv.y), pack2(xv.z, xv.z), pack2(xv.w, xv.w)};
                const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(wt + ci * C + q * 8);
                const ulonglong2 wa = wr[0], wb = wr[1];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[p][0] = ffma2(x2[p], wa.x, acc[p][0]);
                    acc[p][1] = ffma2(x2[p], wa.y, acc[p][1]);
                    acc[p][2] = ffma2(x2[p], wb.x, acc[p][2]);
                    acc[p][3] = ffma2(x2[p], wb.y, acc[p][3]);
                }
            }
            if (j + 2 < my_tiles) bar_arrive(kBarEmpty + buf, kTmThreads);   // ds[buf] may be overwritten
            if (inside) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v[2][4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) unpack2(acc[p][i], v[0][p], v[1][p]);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        float4 r = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                        if (RES) {
                            const float4 u = rv[2 * i + k];
                            r.x += u.x; r.y += u.y; r.z += u.z; r.w += u.w;
                        }
                        *reinterpret_cast<float4 *>(y + o + (int64_t)(2 * i + k) * hw) = r;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stem: out[co] = bias[co] + sum_{ci,tap} w[co][ci][tap] * in[ci][tap]   (CIN = 3, COUT = 32)
// 8x64-pixel tiles, 256 threads, thread = (4 adjacent pixels, 16 of the 32 outputs) as packed (co, co+1)
// pairs: a tap's 4 broadcast LDS.128 of weights feed 32 FFMA2 (the load return path is the bound, see
// dw_act_pw above); per input channel and row one LDS.128 + one LDS.64 bring the six pixels.
// ---------------------------------------------------------------------------------------------
constexpr int kStTW = 64, kStPitch = 68, kStCS = kHH * kStPitch;
constexpr int kStThreads = 256;

__global__ void __launch_bounds__(kStThreads, 2)
stem_conv3x3_kernel(const float *__restrict__ x, const float *__restrict__ wgt,
                    const float *__restrict__ bias, float *__restrict__ y, int h, int w)
{
    constexpr int CIN = 3, COUT = 32;
    __shared__ __align__(16) float xs[CIN * kStCS];
    __shared__ __align__(16) float wt[CIN * 9 * COUT];   // [ci*9+tap][co]
    __shared__ __align__(16) float bs[COUT];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * kStTW, ty0 = blockIdx.y * kTH;
    const int64_t b = blockIdx.z;
    const int64_t hw = (int64_t)h * w;
    for (int i = tid; i < COUT * CIN * 9; i += kStThreads) {
        const int co = i / (CIN * 9), r = i - co * (CIN * 9);
        wt[r * COUT + co] = __ldg(wgt + i);
    }
    if (tid < COUT) bs[tid] = bias ? __ldg(bias + tid) : 0.0f;
    load_halo<kStTW, kStPitch, kStCS>(x, xs, b, CIN, h, w, ty0, tx0, kStThreads);
    halo_wait();
    __syncthreads();
    const int half = tid & 1, pq = tid >> 1;               // output half, pixel quad 0..127
    const int row = pq >> 4, j4 = (pq & 15) * 4;           // pixels (row, j4 .. j4+3)
    f32x2 acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const f32x2 bj = reinterpret_cast<const f32x2 *>(bs)[half * 8 + j];
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p][j] = bj;
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const float *xr = xs + ci * kStCS + (row + dy) * kStPitch + j4;   // halo columns j4 .. j4+5
            const float4 va = *reinterpret_cast<const float4 *>(xr);
            const float2 vb = *reinterpret_cast<const float2 *>(xr + 4);
            const float v[6] = {va.x, va.y, va.z, va.w, vb.x, vb.y};
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(wt + (ci * 9 + dy * 3 + dx) * COUT + half * 16);
                const ulonglong2 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const f32x2 u = pack2(v[p + dx], v[p + dx]);
                    acc[p][0] = ffma2(u, w0.x, acc[p][0]); acc[p][1] = ffma2(u, w0.y, acc[p][1]);
                    acc[p][2] = ffma2(u, w1.x, acc[p][2]); acc[p][3] = ffma2(u, w1.y, acc[p][3]);
                    acc[p][4] = ffma2(u, w2.x, acc[p][4]); acc[p][5] = ffma2(u, w2.y, acc[p][5]);
                    acc[p][6] = ffma2(u, w3.x, acc[p][6]); acc[p][7] = ffma2(u, w3.y, acc[p][7]);
                }
            }
        }
    const int gy = ty0 + row, gx = tx0 + j4;
    if (gy < h && gx < w) {
        const int64_t o = b * COUT * hw + (int64_t)gy * w + gx;
        const bool vec = gx + 3 < w && (w & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v[2][4];
#pragma unroll
            for (int p = 0; p < 4; ++p) unpack2(acc[p][j], v[0][p], v[1][p]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float *oc = y + o + (int64_t)(half * 16 + 2 * j + k) * hw;
                if (vec) {
                    *reinterpret_cast<float4 *>(oc) = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                } else {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if (gx + p < w) oc[p] = v[k][p];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Head: out[co] = residual[co] + bias[co] + sum_{ci,tap} w[co][ci][tap] * in[ci][tap]  (CIN = 32, COUT = 3)
// A CTA takes TWO horizontally adjacent 8x32 tiles; a thread computes the same (row, col) of both, so a
// tap costs one broadcast LDS.128 (w0, w1, w2, -), two LDS.32 and six FFMA.
// (Measured alternatives, not kept, 0.80 ms at 3840x2160 for this form: the two pixels as one packed
// pair and three FFMA2 per tap -- the (w, w) and (u, v) register pairs cost more issue slots than the
// packing saves, 1.00 ms; 4 adjacent pixels per thread on 8x64 tiles with 128 threads -- fewer load
// bytes per FMA but only 8 warps per SM and the halo load exposed, 0.87 ms.)
// ---------------------------------------------------------------------------------------------
constexpr int kHdTW = 32, kHdPitch = 34, kHdCS = 360;
constexpr int kHdThreads = 256;
constexpr size_t kHdSmem = sizeof(float) * 2 * 32 * kHdCS + sizeof(float4) * 32 * 9;

__global__ void __launch_bounds__(kHdThreads)
head_conv3x3_kernel(const float *__restrict__ x, const float *__restrict__ wgt,
                    const float *__restrict__ bias, const float *__restrict__ residual,
                    float *__restrict__ y, int h, int w)
{
    constexpr int CIN = 32, COUT = 3;
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                                           // [32][360] halo of the left tile
    float *xs2 = xs + CIN * kHdCS;                              // right tile
    float4 *wt = reinterpret_cast<float4 *>(xs2 + CIN * kHdCS); // [ci*9+tap] -> (w0, w1, w2, 0)
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * (2 * kHdTW), ty0 = blockIdx.y * kTH;
    const int64_t b = blockIdx.z;
    const int64_t hw = (int64_t)h * w;
    for (int i = tid; i < CIN * 9; i += kHdThreads)
        wt[i] = make_float4(__ldg(wgt + i), __ldg(wgt + CIN * 9 + i), __ldg(wgt + 2 * CIN * 9 + i), 0.0f);
    load_halo<kHdTW, kHdPitch, kHdCS>(x, xs, b, CIN, h, w, ty0, tx0, kHdThreads);
    load_halo<kHdTW, kHdPitch, kHdCS>(x, xs2, b, CIN, h, w, ty0, tx0 + kHdTW, kHdThreads);   // zeros past the edge
    halo_wait();
    __syncthreads();
    const int col = tid & 31, row = tid >> 5;
    const float b0 = bias ? __ldg(bias + 0) : 0.0f, b1 = bias ? __ldg(bias + 1) : 0.0f,
                b2 = bias ? __ldg(bias + 2) : 0.0f;
    float a0[2] = {b0, b0}, a1[2] = {b1, b1}, a2[2] = {b2, b2};          // [left tile, right tile]
#pragma unroll 4
    for (int ci = 0; ci < CIN; ++ci) {
        const float *xc = xs + ci * kHdCS + row * kHdPitch + col;
        const float *xd = xs2 + ci * kHdCS + row * kHdPitch + col;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 wv = wt[ci * 9 + t];
            const float u = xc[(t / 3) * kHdPitch + t % 3], v = xd[(t / 3) * kHdPitch + t % 3];
            a0[0] = fmaf(u, wv.x, a0[0]); a1[0] = fmaf(u, wv.y, a1[0]); a2[0] = fmaf(u, wv.z, a2[0]);
            a0[1] = fmaf(v, wv.x, a0[1]); a1[1] = fmaf(v, wv.y, a1[1]); a2[1] = fmaf(v, wv.z, a2[1]);
        }
    }
    const float r[3][2] = {{a0[0], a0[1]}, {a1[0], a1[1]}, {a2[0], a2[1]}};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int gy = ty0 + row, gx = tx0 + k * kHdTW + col;
        if (gy < h && gx < w) {
            const int64_t o = b * COUT * hw + (int64_t)gy * w + gx;
            float r0 = r[0][k], r1 = r[1][k], r2 = r[2][k];
            if (residual) { r0 += __ldg(residual + o); r1 += __ldg(residual + o + hw); r2 += __ldg(residual + o + 2 * hw); }
            y[o] = r0; y[o + hw] = r1; y[o + 2 * hw] = r2;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Head as a persistent, TMA-fed pipeline (one CTA of 256 threads per SM).  The form above is bound by
// the return path of its shared-memory loads (6 cycles of it per tap and 64 pixels); here a thread owns
// 2 rows x 4 adjacent pixels, so per input channel 4 rows x (LDS.64 + LDS.128) of pixels and 7 LDS.128
// of weights feed 108 FFMA2 (pixel pairs x (w, w) pairs): ~60 cycles of return path per channel and 256
// pixels instead of ~216.  A tile is 32 rows x 64 columns, streamed in four 8-channel stages (box 68 x 34
// x 8 = 74 KB, two stages in flight, requested two stages ahead), accumulators carried across the stages.
// Tiles start ONE column left of a multiple of 64 (pixel columns 64t-1 .. 64t+62): the six input columns
// a thread needs then start at an even box column (box origin 64t-4, the 16-byte rule of the TMA start
// coordinate) and come in as one LDS.64 + one aligned LDS.128.  Needs w % 4 == 0, 16-byte aligned x.
// ---------------------------------------------------------------------------------------------
constexpr int kHtRows = 32, kHtBoxW = 68, kHtBoxH = kHtRows + 2, kHtCh = 8;
constexpr int kHtThreads = 256;
constexpr uint32_t kHtStageBytes = kHtBoxW * kHtBoxH * kHtCh * 4;        // 73,984
constexpr size_t kHtSmem = 2 * kHtStageBytes + sizeof(float) * 32 * 28 + 2 * 8;
static_assert(kHtStageBytes % 128 == 0, "TMA destinations must stay 128-byte aligned");

__global__ void __launch_bounds__(kHtThreads, 1)
head_conv3x3_tma_kernel(const __grid_constant__ CUtensorMap xmap, const float *__restrict__ wgt,
                        const float *__restrict__ bias, const float *__restrict__ residual,
                        float *__restrict__ y, int h, int w, int tiles_x, int tiles_y, int total_tiles)
{
    using namespace wm::tc5;
    constexpr int CIN = 32, COUT = 3, NST = CIN / kHtCh;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);                         // [2][8][34][68]
    float *wt = reinterpret_cast<float *>(smem_raw + 2 * kHtStageBytes);     // [ci][tap*3 + co], 28 per ci
    const uint32_t bar0 = smem_u32(wt + CIN * 28);
    const int tid = threadIdx.x;
    const int64_t hw = (int64_t)h * w;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int my_units = my_tiles * NST;

    auto issue = [&](int u) {
        const int tile = blockIdx.x + (u / NST) * gridDim.x, st = u % NST;
        const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const uint32_t bar = bar0 + 8u * (uint32_t)(u & 1);
        mbar_expect_tx(bar, kHtStageBytes);
        tma::load_box(smem_u32(xs) + (uint32_t)(u & 1) * kHtStageBytes, &xmap, txi * 64 - 4, tyi * kHtRows - 1,
                      st * kHtCh, b, bar);
    };
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (my_units > 0) issue(0);
        if (my_units > 1) issue(1);
    }
    for (int i = tid; i < CIN * 28; i += kHtThreads) {
        const int ci = i / 28, r = i - ci * 28;          // r = tap * 3 + co
        wt[i] = r < 27 ? __ldg(wgt + ((r % 3) * CIN + ci) * 9 + r / 3) : 0.0f;
    }
    __syncthreads();

    const int rp = tid >> 4, k = tid & 15;               // row pair, column group: pixels 4k-1 .. 4k+2 of the tile
    const float b0 = bias ? __ldg(bias + 0) : 0.0f, b1 = bias ? __ldg(bias + 1) : 0.0f,
                b2 = bias ? __ldg(bias + 2) : 0.0f;
    f32x2 acc[2][2][3];                                   // [row][pixel pair][co]
    int u = 0;
#pragma unroll 1
    for (int j = 0; j < my_tiles; ++j) {
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                acc[o][pp][0] = pack2(b0, b0); acc[o][pp][1] = pack2(b1, b1); acc[o][pp][2] = pack2(b2, b2);
            }
#pragma unroll 1
        for (int st = 0; st < NST; ++st, ++u) {
            mbar_wait(bar0 + 8u * (uint32_t)(u & 1), (uint32_t)(u >> 1) & 1u);
            const float *xb = xs + (u & 1) * (kHtStageBytes / 4) + (2 * rp) * kHtBoxW + 4 * k + 2;
#pragma unroll 2
            for (int c = 0; c < kHtCh; ++c) {
                float wv[28];
                {
                    const float4 *wp = reinterpret_cast<const float4 *>(wt + (st * kHtCh + c) * 28);
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        const float4 t = wp[i];
                        wv[4 * i] = t.x; wv[4 * i + 1] = t.y; wv[4 * i + 2] = t.z; wv[4 * i + 3] = t.w;
                    }
                }
                // four input rows, six columns each (box columns 4k+2 .. 4k+7), as packed neighbour pairs
                f32x2 P[4][3], Q[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float *xr = xb + c * (kHtBoxW * kHtBoxH) + r * kHtBoxW;
                    const float2 a2 = *reinterpret_cast<const float2 *>(xr);
                    const float4 a4 = *reinterpret_cast<const float4 *>(xr + 2);
                    P[r][0] = pack2(a2.x, a2.y); P[r][1] = pack2(a4.x, a4.y); P[r][2] = pack2(a4.z, a4.w);
                    Q[r][0] = pack2(a2.y, a4.x); Q[r][1] = pack2(a4.y, a4.z);
                }
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int t = dy * 3 + dx;
                        const f32x2 w0 = pack2(wv[3 * t], wv[3 * t]), w1 = pack2(wv[3 * t + 1], wv[3 * t + 1]),
                                    w2 = pack2(wv[3 * t + 2], wv[3 * t + 2]);
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const f32x2 A0 = dx == 0 ? P[o + dy][0] : (dx == 1 ? Q[o + dy][0] : P[o + dy][1]);
                            const f32x2 A1 = dx == 0 ? P[o + dy][1] : (dx == 1 ? Q[o + dy][1] : P[o + dy][2]);
                            acc[o][0][0] = ffma2(A0, w0, acc[o][0][0]); acc[o][1][0] = ffma2(A1, w0, acc[o][1][0]);
                            acc[o][0][1] = ffma2(A0, w1, acc[o][0][1]); acc[o][1][1] = ffma2(A1, w1, acc[o][1][1]);
                            acc[o][0][2] = ffma2(A0, w2, acc[o][0][2]); acc[o][1][2] = ffma2(A1, w2, acc[o][1][2]);
                        }
                    }
            }
            __syncthreads();                              // every thread is done with this stage buffer
            if (tid == 0 && u + 2 < my_units) issue(u + 2);
        }
        const int tile = blockIdx.x + j * gridDim.x;
        const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int gx0 = txi * 64 + 4 * k - 1;             // pixels gx0 .. gx0+3; gx0+1 is a multiple of 4
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int gy = tyi * kHtRows + 2 * rp + o;
            if (gy >= h) continue;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float v[4];
                unpack2(acc[o][0][co], v[0], v[1]);
                unpack2(acc[o][1][co], v[2], v[3]);
                const int64_t base = ((int64_t)b * COUT + co) * hw + (int64_t)gy * w + gx0;
                // w % 4 == 0: pixels gx0+1, gx0+2 are inside together and 8-byte aligned
                if (gx0 >= 0 && gx0 < w) y[base] = v[0] + (residual ? __ldg(residual + base) : 0.0f);
                if (gx0 + 1 < w) {
                    float2 r = make_float2(v[1], v[2]);
                    if (residual) {
                        const float2 t = __ldg(reinterpret_cast<const float2 *>(residual + base + 1));
                        r.x += t.x; r.y += t.y;
                    }
                    *reinterpret_cast<float2 *>(y + base + 1) = r;
                }
                if (gx0 + 3 < w) y[base + 3] = v[3] + (residual ? __ldg(residual + base + 3) : 0.0f);
            }
        }
    }
}

inline bool dims_ok(int64_t B, int64_t h, int64_t w)
{
    return B >= 0 && B <= 65535 && h >= 0 && w >= 0 && h < (1 << 24) && w < (1 << 24) &&
           (h + kTH - 1) / kTH <= 65535;
}

}  // namespace sp32
}  // namespace wm

using namespace wm;
using namespace wm::sp32;

extern "C" int wm_dw_act_pw_fwd(const float *x, const float *dw_w, const float *dw_b,
                                const float *pw_w, const float *pw_b, int act,
                                const float *residual, float *y, int64_t B, int64_t C, int64_t h,
                                int64_t w, wm_stream_t stream)
{
    WM_REQUIRE(x && dw_w && dw_b && pw_w && pw_b && y, "wm_dw_act_pw_fwd: null pointer");
    WM_REQUIRE(dims_ok(B, h, w), "wm_dw_act_pw_fwd: bad sizes");
    WM_REQUIRE(C == 32, "wm_dw_act_pw_fwd: C=%lld unsupported (32)", (long long)C);
    WM_REQUIRE(act == 0 || act == 1, "wm_dw_act_pw_fwd: act must be 0 or 1");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int tiles_x = (int)((w + kDwTW - 1) / kDwTW), tiles_y = (int)((h + kTH - 1) / kTH);
    const int64_t total = (int64_t)tiles_x * tiles_y * B;
    // persistent TMA pipeline when its preconditions hold (WM_DW_ACT_PW_LEGACY=1: developer A/B switch)
    static const bool legacy = getenv("WM_DW_ACT_PW_LEGACY") != nullptr;
    CUtensorMap xmap;
    if (!legacy && w % 4 == 0 && aligned16(x) && aligned16(y) && (residual == nullptr || aligned16(residual)) &&
        total < ((int64_t)1 << 31) && tma::make_tmap_nchw(&xmap, x, B, 32, h, w, kTmBoxW, kHH, 32)) {
        const int grid = total < sm_count() ? (int)total : sm_count();
#define WM_DWPW_TMA(G, R)                                                                                   \
    do {                                                                                                    \
        WM_CUDA_OK(cudaFuncSetAttribute(dw_act_pw_tma_kernel<G, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)kTmSmem));                                                     \
        dw_act_pw_tma_kernel<G, R><<<grid, kTmThreads, kTmSmem, s>>>(xmap, dw_w, dw_b, pw_w, pw_b, residual, y, (int)h, \
                                                                     (int)w, tiles_x, tiles_y, (int)total); \
    } while (0)
        if (act == 1 && residual) WM_DWPW_TMA(true, true);
        else if (act == 1) WM_DWPW_TMA(true, false);
        else if (residual) WM_DWPW_TMA(false, true);
        else WM_DWPW_TMA(false, false);
#undef WM_DWPW_TMA
        WM_LAUNCH_OK("dw_act_pw (TMA)");
        return WM_OK;
    }
    dim3 grid((unsigned)tiles_x, (unsigned)tiles_y, (unsigned)B);
    if (act == 1) {
        WM_CUDA_OK(cudaFuncSetAttribute(dw_act_pw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDwSmem));
        dw_act_pw_kernel<true><<<grid, kDwThreads, kDwSmem, s>>>(x, dw_w, dw_b, pw_w, pw_b, residual, y, (int)h, (int)w);
    } else {
        WM_CUDA_OK(cudaFuncSetAttribute(dw_act_pw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDwSmem));
        dw_act_pw_kernel<false><<<grid, kDwThreads, kDwSmem, s>>>(x, dw_w, dw_b, pw_w, pw_b, residual, y, (int)h, (int)w);
    }
    WM_LAUNCH_OK("dw_act_pw");
    return WM_OK;
}

extern "C" int wm_stem_conv3x3_fwd(const float *x, const float *w3x3, const float *bias, float *y,
                                   int64_t B, int64_t h, int64_t w, wm_stream_t stream)
{
    WM_REQUIRE(dims_ok(B, h, w), "wm_stem_conv3x3_fwd: bad sizes");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(x && w3x3 && y, "wm_stem_conv3x3_fwd: null pointer");
    dim3 grid((unsigned)((w + kStTW - 1) / kStTW), (unsigned)((h + kTH - 1) / kTH), (unsigned)B);
    stem_conv3x3_kernel<<<grid, kStThreads, 0, (cudaStream_t)stream>>>(x, w3x3, bias, y, (int)h, (int)w);
    WM_LAUNCH_OK("stem conv3x3");
    return WM_OK;
}

extern "C" int wm_head_conv3x3_fwd(const float *x, const float *w3x3, const float *bias,
                                   const float *residual, float *y, int64_t B, int64_t h, int64_t w,
                                   wm_stream_t stream)
{
    WM_REQUIRE(dims_ok(B, h, w), "wm_head_conv3x3_fwd: bad sizes");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(x && w3x3 && y, "wm_head_conv3x3_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    // persistent TMA pipeline when its preconditions hold (WM_HEAD_LEGACY=1: developer A/B switch)
    static const bool legacy = getenv("WM_HEAD_LEGACY") != nullptr;
    CUtensorMap xmap;
    const int ttx = (int)((w + 1 + 63) / 64), tty = (int)((h + kHtRows - 1) / kHtRows);   // tiles start at column 64t-1
    const int64_t total = (int64_t)ttx * tty * B;
    if (!legacy && w % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(y) & 7u) == 0 &&
        (residual == nullptr || (reinterpret_cast<uintptr_t>(residual) & 7u) == 0) && total < ((int64_t)1 << 31) &&
        tma::make_tmap_nchw(&xmap, x, B, 32, h, w, kHtBoxW, kHtBoxH, kHtCh)) {
        WM_CUDA_OK(cudaFuncSetAttribute(head_conv3x3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHtSmem));
        const int grid = total < sm_count() ? (int)total : sm_count();
        head_conv3x3_tma_kernel<<<grid, kHtThreads, kHtSmem, s>>>(xmap, w3x3, bias, residual, y, (int)h, (int)w, ttx,
                                                                  tty, (int)total);
        WM_LAUNCH_OK("head conv3x3 (TMA)");
        return WM_OK;
    }
    WM_CUDA_OK(cudaFuncSetAttribute(head_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHdSmem));
    dim3 grid((unsigned)((w + 2 * kHdTW - 1) / (2 * kHdTW)), (unsigned)((h + kTH - 1) / kTH), (unsigned)B);
    head_conv3x3_kernel<<<grid, kHdThreads, kHdSmem, s>>>(x, w3x3, bias, residual, y,
                                                                             (int)h, (int)w);
    WM_LAUNCH_OK("head conv3x3");
    return WM_OK;
}
