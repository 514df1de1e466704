// SS2D tail of an LFSSBlock as a persistent, warp-specialised, TMA-fed pipeline (one CTA of 192 threads
// per SM):   out = x * skip_scale + out_proj( LayerNorm_64( ((y0 + y2) + y1) + y3 ) * silu(z) )
// (reference wavemamba_arch.py:490-494, 525; the four direction planes of wm_ss2d_dirs_fwd, zs of
// wm_lfss_z_fwd).  The op moves 384 channel planes per call and is pointwise in the pixel, so a tile is 64
// consecutive pixels of one image: five TMA boxes (64 pixels x 64 channels of each direction plane and
// of z, 80 KB) per tile and stage, two stages, requested two tiles ahead.  The register-staged form in
// pixelwise.cu is bound by exposed load latency (ncu: 61 % long-scoreboard stalls at 24 % occupancy).
//   thread 0     requests the boxes of the tile after next once the LayerNorm warps have left a stage
//   warps 0-3    lane pair = pixel (32 channels each): direction sum, LayerNorm statistics exchanged with
//                one shuffle, affine, * z  ->  v[tile & 1][64 channels][64 pixels]
//   warps 4-5    out_proj of the PREVIOUS tile: thread = (4 adjacent pixels, 8 of the 32 outputs), per
//                input channel one LDS.128 of pixels and two of weights feed 16 FFMA2; x (the residual
//                input) comes straight from global memory, fetched before the FMAs; 16-byte stores
// The two groups hand the v buffers over with named barriers.  Needs hw % 4 == 0, all four planes and
// 16-byte aligned tensors; wm_lfss_out_fwd falls back to the register-staged kernels otherwise.
#include "mma_frag.cuh"
#include "tma.cuh"

namespace wm {
namespace lfss {

using namespace wm::tc5;
using namespace wm::frag;

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
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

constexpr int kTP = 64;                       // pixels per tile
constexpr int kC = 64, kCout = 32;
constexpr int kThreadsA = 128, kThreadsB = 64, kThreads = kThreadsA + kThreadsB;
constexpr uint32_t kBoxBytes = kC * kTP * 4;  // 16 KB
constexpr uint32_t kStageBytes = 5 * kBoxBytes;
constexpr size_t kSmem = 2 * kStageBytes + 2 * kBoxBytes + sizeof(float) * (kC * kCout + 2 * kC + kCout) + 2 * 8;

struct Maps {
    CUtensorMap p[4];    // direction planes in summation order
    CUtensorMap z;
};

__global__ void __launch_bounds__(kThreads, 1)
lfss_out_tma_kernel(const __grid_constant__ Maps maps, const float *__restrict__ on_w,
                    const float *__restrict__ on_b, float eps, const float *__restrict__ w_out,
                    const float *__restrict__ x, const float *__restrict__ skip_scale, float *__restrict__ out,
                    int64_t hw, int tiles_per_img, int total_tiles)
{
    constexpr int kBarFull = 1, kBarEmpty = 3, kBarA = 5;                  // named barriers: +buffer index
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stage = reinterpret_cast<float *>(smem_raw);                    // [2][5][64 ch][64 px]
    float *vbuf = reinterpret_cast<float *>(smem_raw + 2 * kStageBytes);   // [2][64 ch][64 px]
    float *wt = vbuf + 2 * kC * kTP;                                       // [64 ci][32 co]
    float *lw = wt + kC * kCout, *lb = lw + kC, *rs = lb + kC;
    const uint32_t bar0 = smem_u32(rs + kCout);
    const int tid = threadIdx.x;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto issue = [&](int j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int b = tile / tiles_per_img, px0 = (tile - b * tiles_per_img) * kTP;
        const uint32_t bar = bar0 + 8u * (uint32_t)(j & 1);
        const uint32_t dst = smem_u32(stage) + (uint32_t)(j & 1) * kStageBytes;
        mbar_expect_tx(bar, kStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const CUtensorMap *m = &maps.p[k];
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                    "r"(dst + (uint32_t)k * kBoxBytes), "l"(reinterpret_cast<uint64_t>(m)), "r"(px0), "r"(b * kC), "r"(bar)
                : "memory");
        }
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(dst + 4u * kBoxBytes), "l"(reinterpret_cast<uint64_t>(&maps.z)), "r"(px0), "r"(b * kC), "r"(bar)
            : "memory");
    };
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (my_tiles > 0) issue(0);
        if (my_tiles > 1) issue(1);
    }
    for (int i = tid; i < kC * kCout; i += kThreads) {
        const int co = i / kC, ci = i - co * kC;
        wt[ci * kCout + co] = __ldg(w_out + i);
    }
    for (int i = tid; i < kC; i += kThreads) { lw[i] = __ldg(on_w + i); lb[i] = __ldg(on_b + i); }
    if (tid < kCout) rs[tid] = __ldg(skip_scale + tid);
    __syncthreads();

    if (tid < kThreadsA) {
        // =========================== direction sum, LayerNorm, * z ==============================
        const int side = tid & 1, px = tid >> 1, c0 = side * 32;
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)(j >> 1) & 1u);
            const float *st = stage + buf * (kStageBytes / 4) + c0 * kTP + px;
            float xv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                // the caller's order: ((y + ya) + yb) + yc  (= ((y0 + y2) + y1) + y3 of the reference)
                float v = st[i * kTP];
                v += st[kC * kTP + i * kTP];
                v += st[2 * kC * kTP + i * kTP];
                v += st[3 * kC * kTP + i * kTP];
                xv[i] = v;
            }
            float mu = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) mu += xv[i];
            mu += __shfl_xor_sync(0xffffffffu, mu, 1);
            mu *= (1.0f / kC);
            float var = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float d = xv[i] - mu; var = fmaf(d, d, var); }
            var += __shfl_xor_sync(0xffffffffu, var, 1);
            var *= (1.0f / kC);
            const float rstd = 1.0f / sqrtf(var + eps);
            if (j >= 2) bar_sync(kBarEmpty + buf, kThreads);             // the out_proj warps are done with v[buf]
            float *vp = vbuf + buf * (kC * kTP) + c0 * kTP + px;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                vp[i * kTP] = fmaf((xv[i] - mu) * rstd, lw[c0 + i], lb[c0 + i]) * st[4 * kC * kTP + i * kTP];
            bar_arrive(kBarFull + buf, kThreads);                        // v[buf] is complete
            bar_sync(kBarA, kThreadsA);                                  // every thread of the group has left the stage
            if (tid == 0 && j + 2 < my_tiles) issue(j + 2);
        }
    } else {
        // =========================== out_proj + skip ===========================================
        const int t = tid - kThreadsA;
        const int q = t & 3, pq = t >> 2;                 // output quarter (8 outputs), pixel quad 0..15
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            const int tile = blockIdx.x + j * gridDim.x;
            const int b = tile / tiles_per_img;
            const int64_t p = (int64_t)(tile - b * tiles_per_img) * kTP + 4 * pq;
            const bool inside = p < hw;                   // hw % 4 == 0: the four pixels are inside or outside together
            const int64_t o = ((int64_t)b * kCout + q * 8) * hw + p;
            float4 xr[8];
            if (inside) {
#pragma unroll
                for (int i = 0; i < 8; ++i) xr[i] = __ldg(reinterpret_cast<const float4 *>(x + o + (int64_t)i * hw));
            }
            bar_sync(kBarFull + buf, kThreads);                          // v[buf] is complete
            f32x2 acc[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[pp][i] = pack2(0.0f, 0.0f);
            const float *vb = vbuf + buf * (kC * kTP) + 4 * pq;
#pragma unroll 4
            for (int ci = 0; ci < kC; ++ci) {
                const float4 xv = *reinterpret_cast<const float4 *>(vb + ci * kTP);
                const f32x2 x2[4] = {pack2(xv.x, xv.x), pack2(xv.y, xv.y), pack2(xv.z, xv.z), pack2(xv.w, xv.w)};
                const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(wt + ci * kCout + q * 8);
                const ulonglong2 wa = wr[0], wb = wr[1];
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    acc[pp][0] = ffma2(x2[pp], wa.x, acc[pp][0]);
                    acc[pp][1] = ffma2(x2[pp], wa.y, acc[pp][1]);
                    acc[pp][2] = ffma2(x2[pp], wb.x, acc[pp][2]);
                    acc[pp][3] = ffma2(x2[pp], wb.y, acc[pp][3]);
                }
            }
            if (j + 2 < my_tiles) bar_arrive(kBarEmpty + buf, kThreads);  // v[buf] may be overwritten
            if (inside) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v[2][4];
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) unpack2(acc[pp][i], v[0][pp], v[1][pp]);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int co = 2 * i + k;
                        const float sc = rs[q * 8 + co];
                        const float4 u = xr[co];
                        // x * skip_scale + out_proj(...)  (reference :525), one FMA per element as before
                        const float4 r = make_float4(fmaf(u.x, sc, v[k][0]), fmaf(u.y, sc, v[k][1]),
                                                     fmaf(u.z, sc, v[k][2]), fmaf(u.w, sc, v[k][3]));
                        *reinterpret_cast<float4 *>(out + o + (int64_t)co * hw) = r;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same tail with the gate computed in the kernel: zs = silu(in_proj[64:] . LayerNorm_32(x)) (reference
// ln_1 :524, in_proj / chunk :483-484, F.silu(z) :493) instead of read from a tensor that wm_lfss_z_fwd
// wrote -- one kernel and 224 channel planes of traffic less per LFSSBlock (x is needed here anyway, as
// the residual input).  Twelve warps in four roles, a 4-deep software pipeline over the tiles:
//   thread 0     TMA: four direction-plane boxes + the x box (72 KB) per tile, two stages
//   warps 8-11   Z: LayerNorm_32 of the x tile (one thread per pixel) -> xn; 32 -> 64 projection on the
//                tensor cores (a warp per 16 pixels), SiLU -> zbuf
//   warps 0-3    A: direction sum + LayerNorm_64 (statistics need no z; lane = pixel, the two 32-channel
//                halves of a pixel sit in different warps and exchange partial sums through shared memory),
//                then * zbuf -> v[tile & 1]
//   warps 4-7    B: out_proj of v on the tensor cores (a warp per 16 pixels) + x * skip_scale
// The two 1x1 GEMMs run as mma.sync m16n8k8 TF32 + one bf16 m16n8k16 MMA per k-step for the two 3xTF32
// correction terms (as the SS2D projection, ss2d.cu): with FFMA2 and weights from shared memory this kernel
// was bound by the return path of its shared-memory loads (~5 k of 5.8 k cycles per tile against 3.5 k of
// HBM time); the fragments need a fifth of those bytes.  v, zbuf and xn have padded row pitches
// (72 / 68 / 72 floats) so that the fragment loads and the accumulator stores are bank-conflict free.
// Named barriers (id: who arrives -> who waits): 1,2 vFull[buf] A->B; 3,4 vEmpty[buf] B->A; 5 zFull Z->A;
// 6 zEmpty A->Z; 7,8 stageFree[buf] Z->A (thread 0 of A then re-issues the stage); 9 inside Z; 10 inside A.
// ---------------------------------------------------------------------------------------------
constexpr int kCx = 32;
constexpr int kThreadsB2 = 128, kThreadsZ = 128, kThreadsF = kThreadsA + kThreadsB2 + kThreadsZ;   // 384
constexpr int kPV = 72, kPZ = 68, kPX = 72;        // row pitches of v, zbuf, xn
constexpr uint32_t kXBoxBytes = kCx * kTP * 4;                                    // 8 KB
constexpr uint32_t kStageBytesF = 4 * kBoxBytes + kXBoxBytes;                     // 72 KB
constexpr size_t kSmemF = 2 * kStageBytesF +
                          sizeof(float) * (2 * kC * kPV /* v */ + kC * kPZ /* zbuf */ + kCx * kPX /* xn */ +
                                           kC * kCout + kCx * kC + 2 * kC + kCout + 2 * kCx + 4 * kTP) + 2 * 8;
static_assert(kSmemF <= 232448, "shared memory budget");

struct MapsF {
    CUtensorMap p[4];    // direction planes in summation order
    CUtensorMap x;
};

__device__ __forceinline__ void tma_box2d(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

__global__ void __launch_bounds__(kThreadsF, 1)
lfss_tail_tma_kernel(const __grid_constant__ MapsF maps, const float *__restrict__ ln1_w,
                     const float *__restrict__ ln1_b, float ln1_eps, const float *__restrict__ w_z,
                     const float *__restrict__ on_w, const float *__restrict__ on_b, float eps,
                     const float *__restrict__ w_out, const float *__restrict__ x,
                     const float *__restrict__ skip_scale, float *__restrict__ out, int64_t hw,
                     int tiles_per_img, int total_tiles)
{
    constexpr int kBarVFull = 1, kBarVEmpty = 3, kBarZFull = 5, kBarZEmpty = 6, kBarStage = 7, kBarZ = 9, kBarA = 10;
    constexpr int kCntAB = kThreadsA + kThreadsB2, kCntAZ = kThreadsA + kThreadsZ;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stage = reinterpret_cast<float *>(smem_raw);                     // [2]{[4][64 ch][64 px], [32 ch][64 px]}
    float *vbuf = reinterpret_cast<float *>(smem_raw + 2 * kStageBytesF);   // [2][64 ch][kPV]
    float *zbuf = vbuf + 2 * kC * kPV;                                      // [64 ch][kPZ]
    float *xn = zbuf + kC * kPZ;                                            // [32 ch][kPX]
    float *wt = xn + kCx * kPX;        // out_proj B fragments: float2 [8 k-steps][4 n-tiles][32 lanes]
    float *wz = wt + kC * kCout;       // in_proj (z half) B fragments: float2 [4 k-steps][8 n-tiles][32 lanes]
    float *lw = wz + kCx * kC, *lb = lw + kC, *rs = lb + kC;
    float *l1w = rs + kCout, *l1b = l1w + kCx;
    float *sbuf = l1b + kCx;                                                // [mean | var][side][64 px] partial sums
    const uint32_t bar0 = smem_u32(sbuf + 4 * kTP);
    const int tid = threadIdx.x;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto issue = [&](int j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int b = tile / tiles_per_img, px0 = (tile - b * tiles_per_img) * kTP;
        const uint32_t bar = bar0 + 8u * (uint32_t)(j & 1);
        const uint32_t dst = smem_u32(stage) + (uint32_t)(j & 1) * kStageBytesF;
        mbar_expect_tx(bar, kStageBytesF);
#pragma unroll
        for (int k = 0; k < 4; ++k) tma_box2d(dst + (uint32_t)k * kBoxBytes, &maps.p[k], px0, b * kC, bar);
        tma_box2d(dst + 4u * kBoxBytes, &maps.x, px0, b * kCx, bar);
    };
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (my_tiles > 0) issue(0);
        if (my_tiles > 1) issue(1);
    }
    // B fragments of m16n8k8: lane (g = lane / 4, t = lane % 4) holds W[k = 8 ks + t][n = 8 nt + g] and k + 4
    for (int i = tid; i < 8 * 4 * 32; i += kThreadsF) {          // out_proj: k = ci (64), n = co (32); w_out (32, 64)
        const int ln = i & 31, nt = (i >> 5) & 3, ks = i >> 7;
        const int n = 8 * nt + (ln >> 2), k = 8 * ks + (ln & 3);
        reinterpret_cast<float2 *>(wt)[i] = make_float2(__ldg(w_out + n * kC + k), __ldg(w_out + n * kC + k + 4));
    }
    for (int i = tid; i < 4 * 8 * 32; i += kThreadsF) {          // in_proj z half: k = ci (32), n = co (64); w_z (64, 32)
        const int ln = i & 31, nt = (i >> 5) & 7, ks = i >> 8;
        const int n = 8 * nt + (ln >> 2), k = 8 * ks + (ln & 3);
        reinterpret_cast<float2 *>(wz)[i] = make_float2(__ldg(w_z + n * kCx + k), __ldg(w_z + n * kCx + k + 4));
    }
    for (int i = tid; i < kC; i += kThreadsF) { lw[i] = __ldg(on_w + i); lb[i] = __ldg(on_b + i); }
    if (tid < kCout) rs[tid] = __ldg(skip_scale + tid);
    if (tid < kCx) { l1w[tid] = __ldg(ln1_w + tid); l1b[tid] = __ldg(ln1_b + tid); }
    __syncthreads();

    if (tid < kThreadsA) {
        // =========================== A: direction sum, LayerNorm_64, * z =========================
        // lane = pixel (32 distinct shared-memory banks per access: every [channel][64 px] row starts at bank
        // 0, so two lanes on the same pixel would always collide), warps 0,1 own channels 0-31, warps 2,3
        // channels 32-63; the two halves of a pixel exchange their LayerNorm partial sums through sbuf
        const int side = tid >> 6, px = tid & 63, c0 = side * 32;
        float *smu = sbuf, *svar = sbuf + 2 * kTP;
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)(j >> 1) & 1u);
            const float *st = stage + buf * (kStageBytesF / 4) + c0 * kTP + px;
            float xv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float v = st[i * kTP];
                v += st[kC * kTP + i * kTP];
                v += st[2 * kC * kTP + i * kTP];
                v += st[3 * kC * kTP + i * kTP];
                xv[i] = v;
            }
            bar_sync(kBarStage + buf, kCntAZ);          // A and Z have both left the stage
            if (tid == 0 && j + 2 < my_tiles) issue(j + 2);
            float mu = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) mu += xv[i];
            smu[side * kTP + px] = mu;
            bar_sync(kBarA, kThreadsA);
            mu = (mu + smu[(side ^ 1) * kTP + px]) * (1.0f / kC);
            float var = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float d = xv[i] - mu; var = fmaf(d, d, var); }
            svar[side * kTP + px] = var;
            bar_sync(kBarA, kThreadsA);
            var = (var + svar[(side ^ 1) * kTP + px]) * (1.0f / kC);
            const float rstd = 1.0f / sqrtf(var + eps);
            bar_sync(kBarZFull, kCntAZ);                // zbuf holds this tile's gate
            if (j >= 2) bar_sync(kBarVEmpty + buf, kCntAB);   // B is done with v[buf]
            float *vp = vbuf + buf * (kC * kPV) + c0 * kPV + px;
            const float *zp = zbuf + c0 * kPZ + px;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                vp[i * kPV] = fmaf((xv[i] - mu) * rstd, lw[c0 + i], lb[c0 + i]) * zp[i * kPZ];
            bar_arrive(kBarVFull + buf, kCntAB);        // v[buf] is complete
            if (j + 1 < my_tiles) bar_arrive(kBarZEmpty, kCntAZ);   // zbuf may be overwritten
        }
    } else if (tid < kThreadsA + kThreadsB2) {
        // =========================== B: out_proj + skip ========================================
        const int t = tid - kThreadsA;
        const int lane = t & 31, m0 = (t >> 5) * 16;      // this warp's 16 pixels
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            const int tile = blockIdx.x + j * gridDim.x;
            const int b = tile / tiles_per_img;
            const int64_t p0 = (int64_t)(tile - b * tiles_per_img) * kTP + m0 + g;   // accumulator rows g and g + 8
            const bool in0 = p0 < hw, in1 = p0 + 8 < hw;
            // accumulator fragment (n-tile nt, i): pixel p0 + 8 (i / 2), output 8 nt + 2 t4 + (i & 1)
            const int64_t o = ((int64_t)b * kCout + 2 * t4) * hw + p0;
            float xr[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool ok = (i & 2) ? in1 : in0;
                    xr[nt][i] = ok ? __ldg(x + o + (int64_t)(8 * nt + (i & 1)) * hw + ((i & 2) ? 8 : 0)) : 0.0f;
                }
            bar_sync(kBarVFull + buf, kCntAB);
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
            gemm_frag<8, 4, kPV>(acc, vbuf + buf * (kC * kPV) + m0, reinterpret_cast<const float2 *>(wt), lane);
            if (j + 2 < my_tiles) bar_arrive(kBarVEmpty + buf, kCntAB);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool ok = (i & 2) ? in1 : in0;
                    const int co = 8 * nt + 2 * t4 + (i & 1);
                    if (ok)
                        out[o + (int64_t)(8 * nt + (i & 1)) * hw + ((i & 2) ? 8 : 0)] = fmaf(xr[nt][i], rs[co], acc[nt][i]);
                }
        }
    } else {
        // =========================== Z: LayerNorm_32(x), 32 -> 64, SiLU ==========================
        const int t = tid - kThreadsA - kThreadsB2;       // 0..127
        const int lane = t & 31, m0 = (t >> 5) * 16;      // this warp's 16 pixels in the projection
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll 1
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)(j >> 1) & 1u);
            if (t < kTP) {
                const float *xs = stage + buf * (kStageBytesF / 4) + 4 * kC * kTP + t;   // x box, pixel t
                float xv[kCx];
                float mu = 0.0f;
#pragma unroll
                for (int i = 0; i < kCx; ++i) { xv[i] = xs[i * kTP]; mu += xv[i]; }
                mu *= (1.0f / kCx);
                float var = 0.0f;
#pragma unroll
                for (int i = 0; i < kCx; ++i) { const float d = xv[i] - mu; var = fmaf(d, d, var); }
                var *= (1.0f / kCx);
                const float rstd = 1.0f / sqrtf(var + ln1_eps);
#pragma unroll
                for (int i = 0; i < kCx; ++i) xn[i * kPX + t] = fmaf((xv[i] - mu) * rstd, l1w[i], l1b[i]);
            }
            bar_arrive(kBarStage + buf, kCntAZ);          // done with the stage (x box)
            bar_sync(kBarZ, kThreadsZ);                   // xn is complete
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
            gemm_frag<4, 8, kPX>(acc, xn + m0, reinterpret_cast<const float2 *>(wz), lane);
            bar_sync(kBarZ, kThreadsZ);                   // every thread has left xn (the next tile rewrites it)
            // SiLU before the hand-over wait: off the A <-> Z cycle
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = __fdividef(acc[nt][i], 1.0f + __expf(-acc[nt][i]));
            if (j >= 1) bar_sync(kBarZEmpty, kCntAZ);     // A is done with the previous tile's gate
            // accumulator fragment (nt, i): pixel m0 + g + 8 (i / 2), channel 8 nt + 2 t4 + (i & 1)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    zbuf[(8 * nt + 2 * t4 + (i & 1)) * kPZ + m0 + g + ((i & 2) ? 8 : 0)] = acc[nt][i];
            bar_arrive(kBarZFull, kCntAZ);                // zbuf holds this tile's gate
        }
    }
}

// 2-D map over a (rows, hw) fp32 matrix with a 64-pixel x 64-row box
static bool make_map2d(CUtensorMap *tm, const float *base, int64_t rows, int64_t hw)
{
    tma::EncodeTiledFn enc = tma::encode_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)hw * 4};
    const cuuint32_t box[2] = {kTP, kC};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns WM_OK when the pipeline ran, 1 when its preconditions do not hold, or an error code.
int forward(const float *y, const float *ya, const float *yb, const float *yc, const float *zs,
            const float *on_w, const float *on_b, float eps, const float *w_out, const float *x,
            const float *skip_scale, float *out, int64_t B, int64_t hw, cudaStream_t s)
{
    if (!ya || !yb || !yc || hw % 4 != 0 || hw < kTP) return 1;
    if (!aligned16(y) || !aligned16(ya) || !aligned16(yb) || !aligned16(yc) || !aligned16(zs) || !aligned16(x) ||
        !aligned16(out))
        return 1;
    const int64_t tiles_per_img = (hw + kTP - 1) / kTP, total = tiles_per_img * B;
    if (total >= ((int64_t)1 << 31) || B * kC >= ((int64_t)1 << 31)) return 1;
    Maps maps;
    const float *planes[4] = {y, ya, yb, yc};
    for (int k = 0; k < 4; ++k)
        if (!make_map2d(&maps.p[k], planes[k], B * kC, hw)) return 1;
    if (!make_map2d(&maps.z, zs, B * kC, hw)) return 1;
    WM_CUDA_OK(cudaFuncSetAttribute(lfss_out_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    const int grid = total < sm_count() ? (int)total : sm_count();
    lfss_out_tma_kernel<<<grid, kThreads, kSmem, s>>>(maps, on_w, on_b, eps, w_out, x, skip_scale, out, hw,
                                                      (int)tiles_per_img, (int)total);
    WM_LAUNCH_OK("lfss out (TMA)");
    return WM_OK;
}

static bool make_map2d_rows(CUtensorMap *tm, const float *base, int64_t rows, int64_t hw, uint32_t box_rows)
{
    tma::EncodeTiledFn enc = tma::encode_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)hw * 4};
    const cuuint32_t box[2] = {kTP, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The tail with the gate computed in the kernel.  Returns WM_OK, 1 (preconditions do not hold), or an error.
int forward_tail(const float *y, const float *ya, const float *yb, const float *yc, const float *x,
                 const float *ln1_w, const float *ln1_b, float ln1_eps, const float *w_z, const float *on_w,
                 const float *on_b, float eps, const float *w_out, const float *skip_scale, float *out,
                 int64_t B, int64_t hw, cudaStream_t s)
{
    if (hw % 4 != 0 || hw < kTP) return 1;
    if (!aligned16(y) || !aligned16(ya) || !aligned16(yb) || !aligned16(yc) || !aligned16(x) || !aligned16(out))
        return 1;
    const int64_t tiles_per_img = (hw + kTP - 1) / kTP, total = tiles_per_img * B;
    if (total >= ((int64_t)1 << 31) || B * kC >= ((int64_t)1 << 31)) return 1;
    MapsF maps;
    const float *planes[4] = {y, ya, yb, yc};
    for (int k = 0; k < 4; ++k)
        if (!make_map2d_rows(&maps.p[k], planes[k], B * kC, hw, kC)) return 1;
    if (!make_map2d_rows(&maps.x, x, B * kCx, hw, kCx)) return 1;
    WM_CUDA_OK(cudaFuncSetAttribute(lfss_tail_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemF));
    const int grid = total < sm_count() ? (int)total : sm_count();
    lfss_tail_tma_kernel<<<grid, kThreadsF, kSmemF, s>>>(maps, ln1_w, ln1_b, ln1_eps, w_z, on_w, on_b, eps, w_out, x,
                                                         skip_scale, out, hw, (int)tiles_per_img, (int)total);
    WM_LAUNCH_OK("lfss tail (TMA)");
    return WM_OK;
}

}  // namespace lfss
}  // namespace wm
