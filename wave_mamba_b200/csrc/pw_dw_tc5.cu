// y = act( dw3x3( pw1x1( LayerNorm_c?(x) ) ) ),  Cin = 32, Cout in {32, 64, 96}: the fused
// pointwise + depthwise groups of HFEBlock / LFSSBlock (reference wavemamba_arch.py:483-487
// in_proj+conv2d+SiLU, :226 ffn conv1+conv2, :729-732 project_in, :762-764 qkv+qkv_dwconv), as a
// persistent, warp-specialised sm_100a pipeline (one CTA per SM):
//
//   TMA        one cp.async.bulk.tensor (4-D box 40 x 10 x 32ch, zero fill outside the image) per
//              8x32-pixel tile brings the halo tile into shared memory; the next tile's box is in
//              flight while the current tile is computed.  (The box starts 4 columns left of the
//              tile: the innermost start coordinate must be a multiple of 16 bytes -- measured with
//              tools/probes/tma_probe.cu, a box at x0-1 raises an illegal-instruction fault.)
//   warps 0-3  LayerNorm over channels per halo position, written straight into the UMMA K-major
//              operand layout [ci/4][position][ci%4] as tf32 hi and lo = a - hi (3xTF32 split)
//   warp 20    one thread issues tcgen05.mma kind::tf32 (M=128 positions, N=64 = [w_hi | w_lo] of a
//              32-channel output group, K=32 in 4 steps) into TMEM, double-buffered per group
//   warps 4-19 TMEM -> registers -> +bias, zero outside the image (the depthwise conv pads the 1x1
//              OUTPUT) -> shared [channel][row][36]; then the depthwise 3x3 + SiLU from shared
//              memory, four adjacent pixels per thread on the packed FP32 pipe (FFMA2), 16-byte stores
//
// Falls back to the cp.async / mma.sync kernel of pointwise.cu when the TMA preconditions do not
// hold (w % 4 != 0 or unaligned pointers).
#include <cuda.h>

#include <atomic>

#include "tc5_common.cuh"

namespace wm {
namespace pwdw {

using namespace wm::tc5;

constexpr int kTH = 8, kTW = 32;
constexpr int kBoxW = 40, kBoxH = kTH + 2;     // TMA box: columns tx0-4 .. tx0+35, rows ty0-1 .. ty0+8
constexpr int kBoxLeft = 4;                    // box column of image column tx0
constexpr int kRaw = kBoxW * kBoxH;            // 400 floats per channel in the TMA buffer
constexpr int kHW = kTW + 2;                   // halo row length 34 (columns tx0-1 .. tx0+32)
constexpr int kPos = kHW * kBoxH;              // 340 halo positions
constexpr int kPSW = 36;                       // row stride of the 1x1 output tile (16-byte rows)
constexpr int kPS = kPSW * kBoxH;              // 360 floats per channel
constexpr int kMPos = 384;                     // three M=128 MMAs
constexpr int kCin = 32;
constexpr int kWarpsA = 4, kWarpsB = 16;
constexpr int kThreadsA = 32 * kWarpsA, kThreadsB = 32 * kWarpsB;
constexpr int kWarpMma = kWarpsA + kWarpsB;
constexpr int kThreads = 32 * (kWarpMma + 1);  // 672
constexpr int kAccCols = 3 * 64;               // one accumulator set: 3 M tiles x [32 hi-sum | 32 lo]
constexpr uint32_t kBoxBytes = kRaw * kCin * 4;

template <int COUT>
struct Smem {
    static constexpr int G = COUT / 32;
    static constexpr size_t xraw = 0;                                  // [32][10][40] floats (TMA box)
    static constexpr size_t xhi = xraw + (size_t)kCin * kRaw * 4;      // [8][384] float4
    static constexpr size_t xlo = xhi + (size_t)8 * kMPos * 16;
    static constexpr size_t ps = xlo + (size_t)8 * kMPos * 16;         // [32][10][36] floats
    static constexpr size_t wsm = ps + (size_t)32 * kPS * 4;           // [G][8][64] float4
    static constexpr size_t cst = wsm + (size_t)G * 8 * 64 * 16;       // pwb[COUT] dww[COUT*9] dwb[COUT] lnw[32] lnb[32]
    static constexpr size_t bars = cst + (size_t)(COUT * 11 + 64) * 4; // 8 mbarriers + tmem slot
    static constexpr size_t total = bars + 8 * 8 + 16;
    static_assert(xhi % 128 == 0 && wsm % 16 == 0 && bars % 8 == 0, "alignment");
    static_assert(total <= 232448, "shared memory budget");
};

struct Args {
    const float *ln_w, *ln_b;
    float eps;
    const float *pw_w, *pw_b, *dw_w, *dw_b;
    float *y;
    int h, w, tiles_x, tiles_y, total_tiles;
    unsigned int *err;     // pipeline error word (mbar_wait_flag)
    long long *dbg;        // optional per-CTA cycle counters (wm_pw_dw_debug_timing), 12 per CTA:
                           // LN warps: [0] wait TMA [1] wait operand free [2] LN pass
                           // MMA: [3] wait operand [4] wait accumulators [5] issue
                           // epilogue warps: [6] wait MMA [7] TMEM -> ps [8] depthwise [9] tiles [10] total
};

static std::atomic<long long *> g_dbg{nullptr};

__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2,
                                             int c3, uint32_t mbar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(mbar)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

typedef unsigned long long f32x2;      // packed fp32 pair (Blackwell FFMA2)
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

__device__ __forceinline__ void named_bar(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int COUT, bool LN, bool SILU>
__global__ void __launch_bounds__(kThreads, 1)
pw_dw_tc5_kernel(const __grid_constant__ CUtensorMap tmap, const Args a)
{
    using S = Smem<COUT>;
    constexpr int G = S::G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *xraw = reinterpret_cast<float *>(smem_raw + S::xraw);
    float4 *xhi = reinterpret_cast<float4 *>(smem_raw + S::xhi);
    float4 *xlo = reinterpret_cast<float4 *>(smem_raw + S::xlo);
    float *ps = reinterpret_cast<float *>(smem_raw + S::ps);
    float4 *wsm = reinterpret_cast<float4 *>(smem_raw + S::wsm);
    float *pwb = reinterpret_cast<float *>(smem_raw + S::cst);
    float *dww = pwb + COUT;
    float *dwb = dww + COUT * 9;
    float *lnw = dwb + COUT;
    float *lnb = lnw + 32;
    const uint32_t bar0 = smem_u32(smem_raw + S::bars);
    const uint32_t xraw_full = bar0, xk_full = bar0 + 8, xk_empty = bar0 + 16;
    auto acc_full = [&](int i) { return bar0 + 24u + 8u * (uint32_t)i; };
    auto acc_empty = [&](int i) { return bar0 + 40u + 8u * (uint32_t)i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + S::bars + 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = a.h, w = a.w;
    const int64_t hw = (int64_t)h * w;

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(xraw_full, 1);
        mbar_init(xk_full, kThreadsA);
        mbar_init(xk_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kWarpsB); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // 1x1 weights -> UMMA K-major B operand [group][kc][hi co 0..31 | lo co 0..31], split at rna
    for (int i = tid; i < COUT * 8; i += kThreads) {
        const int co = i >> 3, kc = i & 7;
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v = __ldg(a.pw_w + co * kCin + kc * 4 + j);
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float rest = v - __uint_as_float(hb);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rest));
            hi[j] = __uint_as_float(hb);
            lo[j] = __uint_as_float(lb);
        }
        float4 *dst = wsm + ((co >> 5) * 8 + kc) * 64 + (co & 31);
        dst[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        dst[32] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int i = tid; i < COUT; i += kThreads) {
        pwb[i] = a.pw_b ? __ldg(a.pw_b + i) : 0.0f;
        dwb[i] = __ldg(a.dw_b + i);
    }
    for (int i = tid; i < COUT * 9; i += kThreads) dww[i] = __ldg(a.dw_w + i);
    if (LN && tid < 32) { lnw[tid] = __ldg(a.ln_w + tid); lnb[tid] = __ldg(a.ln_b + tid); }
    // rows 340..383 of the operand (read by the third M tile, results never used): defined values
    for (int i = tid; i < 8 * (kMPos - kPos); i += kThreads) {
        const int kc = i / (kMPos - kPos), r = i - kc * (kMPos - kPos);
        xhi[kc * kMPos + kPos + r] = make_float4(0.f, 0.f, 0.f, 0.f);
        xlo[kc * kMPos + kPos + r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int &tx0, int &ty0, int &b) {
        const int txi = tile % a.tiles_x, tyi = (tile / a.tiles_x) % a.tiles_y;
        b = tile / (a.tiles_x * a.tiles_y);
        tx0 = txi * kTW;
        ty0 = tyi * kTH;
    };

    if (warp < kWarpsA) {
        // =========================== LayerNorm + operand layout ==============================
        const CUtensorMap *tmap_ptr = &tmap;      // generic address of the __grid_constant__ parameter
        auto issue_tma = [&](int tile) {
            int tx0, ty0, b;
            tile_coords(tile, tx0, ty0, b);
            mbar_expect_tx(xraw_full, kBoxBytes);
            tma_load_box(smem_u32(xraw), tmap_ptr, tx0 - kBoxLeft, ty0 - 1, 0, b, xraw_full);
        };
        if (tid == 0 && (int)blockIdx.x < a.total_tiles) issue_tma(blockIdx.x);
        uint32_t it = 0;
        const bool timed = a.dbg != nullptr && tid == 0;
        long long ta[3] = {0, 0, 0}, tp = timed ? clock64() : 0;
#define WM_TICKA(k) do { if (timed) { const long long _t = clock64(); ta[k] += _t - tp; tp = _t; } } while (0)
#pragma unroll 1
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
            mbar_wait_flag(xraw_full, it & 1u, a.err, (1u << 24) | (0u << 16) | (it & 0xffffu));
            WM_TICKA(0);
            mbar_wait_flag(xk_empty, (it & 1u) ^ 1u, a.err, (1u << 24) | (2u << 16) | (it & 0xffffu));
            WM_TICKA(1);       // the previous tile's MMAs have read xhi/xlo
#pragma unroll 1
            for (int pos = tid; pos < kPos; pos += kThreadsA) {
                float v[kCin];
                const int prow = pos / kHW;
                const int raw = prow * kBoxW + (pos - prow * kHW) + (kBoxLeft - 1);   // box column 3 = tx0-1
#pragma unroll
                for (int c = 0; c < kCin; ++c) v[c] = xraw[c * kRaw + raw];
                if (LN) {
                    float mu = 0.0f;
#pragma unroll
                    for (int c = 0; c < kCin; ++c) mu += v[c];
                    mu *= (1.0f / kCin);
                    float var = 0.0f;
#pragma unroll
                    for (int c = 0; c < kCin; ++c) { const float d = v[c] - mu; var = fmaf(d, d, var); }
                    var *= (1.0f / kCin);
                    const float rstd = 1.0f / sqrtf(var + a.eps);
#pragma unroll
                    for (int c = 0; c < kCin; ++c) v[c] = fmaf((v[c] - mu) * rstd, lnw[c], lnb[c]);
                }
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {
                    const float4 t = make_float4(v[4 * kc], v[4 * kc + 1], v[4 * kc + 2], v[4 * kc + 3]);
                    xhi[kc * kMPos + pos] = t;
                    xlo[kc * kMPos + pos] = make_float4(tf32_lo(t.x), tf32_lo(t.y), tf32_lo(t.z), tf32_lo(t.w));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(xk_full);
            named_bar(1, kThreadsA);                   // every thread is done reading xraw
            if (tid == 0 && tile + (int)gridDim.x < a.total_tiles) issue_tma(tile + gridDim.x);
            WM_TICKA(2);
        }
        if (timed) for (int i = 0; i < 3; ++i) a.dbg[blockIdx.x * 12 + i] = ta[i];
#undef WM_TICKA
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer ==============================================
        if (lane == 0) {
            constexpr uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) |
                                         ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) |
                                         ((uint32_t)(128 >> 4) << 24);
            const uint64_t a_hi0 = make_desc(smem_u32(xhi), kMPos * 16u, 128u);
            const uint64_t a_lo0 = make_desc(smem_u32(xlo), kMPos * 16u, 128u);
            const uint64_t b_0 = make_desc(smem_u32(wsm), 64 * 16u, 128u);
            uint32_t it = 0, gcount = 0;
            const bool timed = a.dbg != nullptr;
            long long tm[3] = {0, 0, 0}, tp = timed ? clock64() : 0;
#define WM_TICKM(k) do { if (timed) { const long long _t = clock64(); tm[k] += _t - tp; tp = _t; } } while (0)
#pragma unroll 1
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
                mbar_wait_flag(xk_full, it & 1u, a.err, (2u << 24) | (1u << 16) | (it & 0xffffu));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                WM_TICKM(0);
#pragma unroll 1
                for (int g = 0; g < G; ++g, ++gcount) {
                    const int buf = gcount & 1;
                    mbar_wait_flag(acc_empty(buf), ((gcount >> 1) & 1u) ^ 1u, a.err, (2u << 24) | (4u << 16) | (gcount & 0xffffu));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    WM_TICKM(1);
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) {
                        const uint32_t d = tmem_base + (uint32_t)(buf * kAccCols + mt * 64);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t aoff = (uint32_t)(2 * ks * kMPos + mt * 128);
                            const uint32_t boff = (uint32_t)((g * 8 + 2 * ks) * 64);
                            // cols [0,32) += a_hi w_hi, [32,64) += a_hi w_lo ; cols [0,32) += a_lo w_hi
                            mma_tf32_ss(d, a_hi0 + aoff, b_0 + boff, idesc64, ks > 0 ? 1u : 0u);
                            mma_tf32_ss(d, a_lo0 + aoff, b_0 + boff, idesc32, 1u);
                        }
                    }
                    mma_commit(acc_full(buf));
                    WM_TICKM(2);
                }
                mma_commit(xk_empty);
            }
            if (timed) for (int i = 0; i < 3; ++i) a.dbg[blockIdx.x * 12 + 3 + i] = tm[i];
#undef WM_TICKM
        }
    } else {
        // =========================== epilogue + depthwise 3x3 ================================
        // 16 warps (the two phases are latency-bound: more warps hide the TMEM / MUFU / shared latencies)
        const int e = warp - kWarpsA;                  // 0..15
        // TMEM lane quarter (= warp % 4, the hardware rule), 16-channel half of the group, M-tile set:
        // set 0 drains M tiles 0 and 1, set 1 M tile 2 (positions 256..339)
        const int quarter = warp & 3, chalf = (e >> 2) & 1, mset = e >> 3;
        const int tb = tid - kThreadsA;                // 0..511
        uint32_t gcount = 0;
        const bool timed = a.dbg != nullptr && tb == 0;
        long long tbb[3] = {0, 0, 0}, t0 = timed ? clock64() : 0, tp = t0;
        int ntile = 0;
#define WM_TICKB(k) do { if (timed) { const long long _t = clock64(); tbb[k] += _t - tp; tp = _t; } } while (0)
#pragma unroll 1
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ntile) {
            int tx0, ty0, b;
            tile_coords(tile, tx0, ty0, b);
#pragma unroll 1
            for (int g = 0; g < G; ++g, ++gcount) {
                const int buf = gcount & 1;
                mbar_wait_flag(acc_full(buf), (gcount >> 1) & 1u, a.err, (3u << 24) | (3u << 16) | (gcount & 0xffffu));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                WM_TICKB(0);
#pragma unroll 1
                for (int mt = mset ? 2 : 0; mt < (mset ? 3 : 2); ++mt) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                           (uint32_t)(buf * kAccCols + mt * 64 + chalf * 16);
                    uint32_t acc[16], part[16];
                    tmem_ld16(taddr, acc);
                    tmem_ld16(taddr + 32u, part);
                    tmem_ld_wait();
                    const int pos = mt * 128 + quarter * 32 + lane;
                    if (pos < kPos) {
                        const int row = pos / kHW, col = pos - row * kHW;
                        const int gy = ty0 - 1 + row, gx = tx0 - 1 + col;
                        const bool valid = gy >= 0 && gy < h && gx >= 0 && gx < w;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float v = (__uint_as_float(acc[j]) + __uint_as_float(part[j])) +
                                            pwb[g * 32 + chalf * 16 + j];
                            ps[(chalf * 16 + j) * kPS + row * kPSW + col] = valid ? v : 0.0f;
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(buf));
                named_bar(2, kThreadsB);               // ps of this group is complete
                WM_TICKB(1);

                // depthwise 3x3 (+ SiLU): thread = (channel, 4 adjacent columns, 4 of the 8 rows), sliding
                // window of three input rows held as packed pairs (FFMA2: two outputs per instruction).
                // Per input row one LDS.128 + one LDS.64 bring the 6 values (v0..v5) four outputs need.
                {
                    const int rhalf = tb >> 8;                       // rows 0-3 or 4-7 of the tile
                    const int cl = (tb & 255) >> 3, j4 = (tb & 7) * 4;
                    const int co = g * 32 + cl;
                    f32x2 k2[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) { const float kv = dww[co * 9 + t]; k2[t] = pack2(kv, kv); }
                    const float bias = dwb[co];
                    const float *pc = ps + cl * kPS + rhalf * 4 * kPSW + j4;
                    // packed pairs of one input row: P0=(v0,v1) P1=(v2,v3) P2=(v4,v5) Q0=(v1,v2) Q1=(v3,v4)
                    f32x2 P[3][3], Q[3][2];
                    auto load_row = [&](int r, int slot) {
                        const float4 a4 = *reinterpret_cast<const float4 *>(pc + r * kPSW);
                        const float2 b2 = *reinterpret_cast<const float2 *>(pc + r * kPSW + 4);
                        P[slot][0] = pack2(a4.x, a4.y); P[slot][1] = pack2(a4.z, a4.w); P[slot][2] = pack2(b2.x, b2.y);
                        Q[slot][0] = pack2(a4.y, a4.z); Q[slot][1] = pack2(a4.w, b2.x);
                    };
                    load_row(0, 0);
                    load_row(1, 1);
                    const int gx = tx0 + j4, gy0 = ty0 + rhalf * 4;
                    float *yo = a.y + ((int64_t)b * COUT + co) * hw + (int64_t)gy0 * w + gx;
#pragma unroll
                    for (int row = 0; row < 4; ++row) {
                        load_row(row + 2, (row + 2) % 3);
                        f32x2 oa = pack2(bias, bias), ob = oa;
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const int sl = (row + dy) % 3;
                            oa = ffma2(k2[3 * dy + 0], P[sl][0], oa); ob = ffma2(k2[3 * dy + 0], P[sl][1], ob);
                            oa = ffma2(k2[3 * dy + 1], Q[sl][0], oa); ob = ffma2(k2[3 * dy + 1], Q[sl][1], ob);
                            oa = ffma2(k2[3 * dy + 2], P[sl][1], oa); ob = ffma2(k2[3 * dy + 2], P[sl][2], ob);
                        }
                        float o0, o1, o2, o3;
                        unpack2(oa, o0, o1);
                        unpack2(ob, o2, o3);
                        if (SILU) {   // SS2D.act (reference :487)
                            o0 = __fdividef(o0, 1.0f + __expf(-o0)); o1 = __fdividef(o1, 1.0f + __expf(-o1));
                            o2 = __fdividef(o2, 1.0f + __expf(-o2)); o3 = __fdividef(o3, 1.0f + __expf(-o3));
                        }
                        // w % 4 == 0 and gx % 4 == 0: the four columns are inside the image or outside together
                        if (gx < w && gy0 + row < h)
                            *reinterpret_cast<float4 *>(yo + (int64_t)row * w) = make_float4(o0, o1, o2, o3);
                    }
                }
                named_bar(2, kThreadsB);               // ps may be overwritten by the next group
                WM_TICKB(2);
            }
        }
        if (timed) {
            for (int i = 0; i < 3; ++i) a.dbg[blockIdx.x * 12 + 6 + i] = tbb[i];
            a.dbg[blockIdx.x * 12 + 9] = ntile;
            a.dbg[blockIdx.x * 12 + 10] = clock64() - t0;
        }
#undef WM_TICKB
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kWarpMma) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                     : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
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

// NCHW fp32 tensor (B, 32, h, w) as a 4-D tensor map with a 40 x 10 x 32 x 1 box
static bool make_tmap(CUtensorMap *tm, const float *x, int64_t B, int64_t h, int64_t w)
{
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)kCin, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)kCin * h * w * 4};
    const cuuint32_t box[4] = {kBoxW, kBoxH, kCin, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS;
}

template <int COUT, bool LN, bool SILU>
static int launch(const CUtensorMap &tm, const Args &a, cudaStream_t s)
{
    using S = Smem<COUT>;
    WM_CUDA_OK(cudaFuncSetAttribute(pw_dw_tc5_kernel<COUT, LN, SILU>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
    const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
    pw_dw_tc5_kernel<COUT, LN, SILU><<<grid, kThreads, S::total, s>>>(tm, a);
    WM_LAUNCH_OK("pw_dw (tcgen05)");
    return WM_OK;
}

// Returns WM_OK when the tcgen05 path ran, 1 when its preconditions do not hold (the caller then uses
// the cp.async / mma.sync kernel), or an error code.
int forward(const float *x, const float *ln_w, const float *ln_b, float eps, const float *pw_w,
            const float *pw_b, const float *dw_w, const float *dw_b, int act, float *y, int64_t B,
            int64_t Cout, int64_t h, int64_t w, cudaStream_t s)
{
    if (w % 4 != 0 || !aligned16(x) || !aligned16(y) || w < kTW || h < kTH) return 1;
    CUtensorMap tm;
    if (!make_tmap(&tm, x, B, h, w)) return 1;
    Args a;
    a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps; a.pw_w = pw_w; a.pw_b = pw_b; a.dw_w = dw_w; a.dw_b = dw_b;
    a.y = y; a.h = (int)h; a.w = (int)w;
    a.tiles_x = (int)((w + kTW - 1) / kTW);
    a.tiles_y = (int)((h + kTH - 1) / kTH);
    const int64_t total = (int64_t)a.tiles_x * a.tiles_y * B;
    if (total >= ((int64_t)1 << 31)) return 1;
    a.total_tiles = (int)total;
    a.err = pipeline_err_word();
    if (a.err == nullptr) return 1;
    a.dbg = g_dbg.load();
    const bool ln = ln_w != nullptr, silu = act == 1;
#define WM_PWDW_CASE(C)                                                                  \
    if (Cout == C) {                                                                     \
        if (ln && silu) return launch<C, true, true>(tm, a, s);                          \
        if (ln) return launch<C, true, false>(tm, a, s);                                 \
        if (silu) return launch<C, false, true>(tm, a, s);                               \
        return launch<C, false, false>(tm, a, s);                                        \
    }
    WM_PWDW_CASE(32)
    WM_PWDW_CASE(64)
    WM_PWDW_CASE(96)
#undef WM_PWDW_CASE
    return 1;
}

}  // namespace pwdw
}  // namespace wm

/* Developer aid: non-NULL device buffer of 12*SMs int64 -> per-CTA cycle counters of the pw_dw pipeline
 * (see Args::dbg); NULL disables. */
extern "C" int wm_pw_dw_debug_timing(void *device_buffer)
{
    wm::pwdw::g_dbg.store(static_cast<long long *>(device_buffer));
    return WM_OK;
}
