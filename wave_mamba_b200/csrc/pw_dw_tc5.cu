// y = act( dw3x3( pw1x1( LayerNorm_c?(x) ) ) ),  Cin = 32, Cout in {32, 64, 96}: the fused
// pointwise + depthwise groups of HFEBlock / LFSSBlock (reference wavemamba_arch.py:483-487
// in_proj+conv2d+SiLU, :226 ffn conv1+conv2, :729-732 project_in, :762-764 qkv+qkv_dwconv), as a
// persistent, warp-specialised sm_100a pipeline (one CTA per SM):
//
//   warp 21    (TMA) ten cp.async.bulk.tensor boxes (40 x 1 row x 32ch, zero fill outside the image) per
//              8x32-pixel tile bring the halo tile into shared memory, in three row groups with their
//              own mbarriers: a group of the NEXT tile is requested as soon as the LayerNorm pass has
//              consumed those rows of the current one, so the buffer works as a ring and the load
//              latency hides behind the pass.  (The box starts 4 columns left of the tile: the
//              innermost start coordinate must be a multiple of 16 bytes -- measured with
//              tools/probes/tma_probe.cu, a box at x0-1 raises an illegal-instruction fault.)
//   warps 0-3  LayerNorm over channels per halo position, written straight into the UMMA K-major
//              operand layout [ci/4][position][ci%4] as tf32 hi and lo = a - hi (3xTF32 split); pass i of
//              three writes exactly M tile i, with full / empty mbarriers per M tile, so the MMAs of a
//              tile run under the LayerNorm pass of the same tile and of the next one
//   warp 20    one thread issues tcgen05.mma kind::tf32 (M=128 positions, N=Cout (64+32 for 96), K=32 in
//              4 steps; the three 3xTF32 products a_lo w_hi + a_hi w_lo + a_hi w_hi accumulate into the
//              same TMEM columns), accumulators double-buffered per tile
//
//   warps 4-19 TMEM -> registers -> +bias, zero outside the image (the depthwise conv pads the 1x1
//              OUTPUT) -> shared [channel][row][36]; then the depthwise 3x3 + SiLU from shared
//              memory, four adjacent pixels per thread on the packed FP32 pipe (FFMA2), 16-byte stores
//
// The kernel is bound by SHARED-MEMORY bandwidth (ncu: LSU wavefronts + the MMAs' operand reads keep the
// 128 B/clk pipe ~80 % busy in the first version), so the layout choices below are about wavefronts:
// one wide MMA per product (the A operand is read once for all output channels), the M order of the
// halo positions has 36 per row = the row pitch of the 1x1 output tile (the TMEM drain stores to
// consecutive addresses: no bank conflicts at the row breaks), per-thread constants come in as 16-byte
// loads.
//
// Falls back to the cp.async / mma.sync kernel of pointwise.cu when the TMA preconditions do not
// hold (w % 4 != 0 or unaligned pointers).
#include <atomic>

#include "tma.cuh"

namespace wm {
namespace pwdw {

using namespace wm::tc5;

#ifndef WM_PWDW_HALVES
#define WM_PWDW_HALVES 1
#endif

constexpr int kTH = 8, kTW = 32;
constexpr int kBoxW = 40, kBoxH = kTH + 2;     // TMA box: columns tx0-4 .. tx0+35, rows ty0-1 .. ty0+8
constexpr int kBoxLeft = 4;                    // box column of image column tx0
constexpr int kRawRow = kBoxW * 32;            // floats per halo row in the TMA buffer: [row][32 ch][40]
                                               // (a TMA destination must be 128-byte aligned: no padding)
                                               // halo row: 34 columns tx0-1 .. tx0+32
constexpr int kPSW = 36;                       // row pitch of the 1x1 output tile (16-byte rows) and of the
                                               // M order: position m = row * 36 + col, col 34, 35 unused
constexpr int kPos = kPSW * kBoxH;             // 360 M positions
constexpr int kPS = kPSW * kBoxH;              // 360 floats per channel
constexpr int kMPos = 384;                     // three M=128 MMAs
constexpr int kCin = 32;
constexpr int kWarpsA = 4, kWarpsB = 16;
constexpr int kThreadsA = 32 * kWarpsA, kThreadsB = 32 * kWarpsB;
constexpr int kWarpMma = kWarpsA + kWarpsB;
constexpr int kWarpTma = kWarpMma + 1;
constexpr int kThreads = 32 * (kWarpTma + 1);  // 704
constexpr int kAccCols = 3 * 64;               // one accumulator buffer: 3 M tiles x up to 64 output channels
constexpr uint32_t kRowBytes = kBoxW * kCin * 4;      // one TMA box
// LayerNorm iteration i covers positions 128i .. 128i+127 = rows 0-3 | 3-7 | 7-9.  Row groups by LAST use
// (= when they can be refilled): rows 0-2 | 3-6 | 7-9.
__host__ __device__ constexpr int row_group_begin(int g) { return g == 0 ? 0 : (g == 1 ? 3 : 7); }
__host__ __device__ constexpr int row_group_end(int g) { return g == 0 ? 3 : (g == 1 ? 7 : kBoxH); }

template <int COUT>
struct Smem {
    static constexpr int G = COUT / 32;
    static constexpr size_t xraw = 0;                                  // [10][32][40] floats (TMA boxes)
    static constexpr size_t xhi = (xraw + (size_t)kBoxH * kRawRow * 4 + 127) / 128 * 128;   // [8][384] float4
    static constexpr size_t xlo = xhi + (size_t)8 * kMPos * 16;
    static constexpr size_t ps = xlo + (size_t)8 * kMPos * 16;         // [32][10][36] floats
    static constexpr size_t wsm = ps + (size_t)32 * kPS * 4;           // [8 kc][hi | lo][COUT] float4
    static constexpr size_t cst = wsm + (size_t)8 * 2 * COUT * 16;     // pwb[COUT] dwk[COUT][12] lnw[32] lnb[32]
    static constexpr size_t bars = cst + (size_t)(COUT * 13 + 64) * 4; // 8 mbarriers + tmem slot
    static constexpr size_t total = bars + 17 * 8 + 16;
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
    long long *dbg;        // optional per-CTA cycle counters (wm_pw_dw_debug_timing), 16 per CTA:
                           // LN warps: [0] wait TMA [1] wait operand free [2] LN pass
                           // MMA: [3] wait operand [4] wait accumulators [5] issue
                           // epilogue warps: [6] wait MMA [7] TMEM -> ps [8] depthwise [9] tiles [10] total
};

static std::atomic<long long *> g_dbg{nullptr};

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
    float *dwk = pwb + COUT;                       // [COUT][12]: 9 taps, bias, 2 pad (three 16-byte loads)
    float *lnw = dwk + COUT * 12;
    float *lnb = lnw + 32;
    // MMA groups: one N-wide MMA per product and group; 96 = 64 + 32 (TMEM holds 2 x 3 x 64 columns)
    constexpr int NMG = COUT == 96 ? 2 : 1;
    const uint32_t bar0 = smem_u32(smem_raw + S::bars);
    // operand barriers per M tile: LayerNorm iteration i writes exactly M tile i (positions 128i .. 128i+127)
    auto xk_full = [&](int mt) { return mt == 0 ? bar0 + 8u : bar0 + 96u + 8u * (uint32_t)mt; };    // 8, 104, 112
    auto xk_empty = [&](int mt) { return mt == 0 ? bar0 + 16u : bar0 + 112u + 8u * (uint32_t)mt; }; // 16, 120, 128
    auto xraw_full = [&](int g) { return g == 0 ? bar0 : bar0 + 56u + 8u * (uint32_t)g; };   // 0, 64, 72
    auto xraw_empty = [&](int g) { return bar0 + 80u + 8u * (uint32_t)g; };
    auto acc_full = [&](int i) { return bar0 + 24u + 8u * (uint32_t)i; };
    auto acc_empty = [&](int i) { return bar0 + 40u + 8u * (uint32_t)i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + S::bars + 136);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = a.h, w = a.w;
    const int64_t hw = (int64_t)h * w;

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        for (int g = 0; g < 3; ++g) { mbar_init(xraw_full(g), 1); mbar_init(xraw_empty(g), kWarpsA); }
        for (int mt = 0; mt < 3; ++mt) { mbar_init(xk_full(mt), kThreadsA); mbar_init(xk_empty(mt), 1); }
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
    // 1x1 weights -> UMMA K-major B operand [kc][hi co 0..COUT-1 | lo co 0..COUT-1], split at rna
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
        float4 *dst = wsm + kc * 2 * COUT + co;
        dst[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        dst[COUT] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int i = tid; i < COUT; i += kThreads) pwb[i] = a.pw_b ? __ldg(a.pw_b + i) : 0.0f;
    for (int i = tid; i < COUT * 12; i += kThreads) {
        const int co = i / 12, t = i - co * 12;
        dwk[i] = t < 9 ? __ldg(a.dw_w + co * 9 + t) : (t == 9 ? __ldg(a.dw_b + co) : 0.0f);
    }
    if (LN && tid < 32) { lnw[tid] = __ldg(a.ln_w + tid); lnb[tid] = __ldg(a.ln_b + tid); }
    // rows 360..383 of the operand (read by the third M tile, results never used): defined values
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
        uint32_t it = 0;
        const bool timed = a.dbg != nullptr && tid == 0;
        long long ta[3] = {0, 0, 0}, tp = timed ? clock64() : 0;
#define WM_TICKA(k) do { if (timed) { const long long _t = clock64(); ta[k] += _t - tp; tp = _t; } } while (0)
#pragma unroll 1
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {
                // rows this iteration touches: 0-3 | 3-7 | 7-9  ->  groups {0, 1} | {2} | {}
                if (i == 0) {
                    mbar_wait_flag(xraw_full(0), it & 1u, a.err, (1u << 24) | (0u << 16) | (it & 0xffffu));
                    mbar_wait_flag(xraw_full(1), it & 1u, a.err, (1u << 24) | (5u << 16) | (it & 0xffffu));
                    WM_TICKA(0);
                } else if (i == 1) {
                    mbar_wait_flag(xraw_full(2), it & 1u, a.err, (1u << 24) | (6u << 16) | (it & 0xffffu));
                    WM_TICKA(0);
                }
                // the previous tile's MMAs have read M tile i of xhi/xlo
                mbar_wait_flag(xk_empty(i), (it & 1u) ^ 1u, a.err, (1u << 24) | (2u << 16) | (it & 0xffffu));
                WM_TICKA(1);
                const int pos = i * kThreadsA + tid;
                if (pos < kPos) {
                    float v[kCin];
                    const int prow = pos / kPSW;
                    // box column 3 = image column tx0-1
                    const float *src = xraw + prow * kRawRow + (pos - prow * kPSW) + (kBoxLeft - 1);
#pragma unroll
                    for (int c = 0; c < kCin; ++c) v[c] = src[c * kBoxW];
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
                mbar_arrive(xk_full(i));
                WM_TICKA(2);
                __syncwarp();                          // this warp is done reading the rows of group i
                if (lane == 0) mbar_arrive(xraw_empty(i));
            }
        }
        if (timed) for (int i = 0; i < 3; ++i) a.dbg[blockIdx.x * 16 + i] = ta[i];
#undef WM_TICKA
    } else if (warp == kWarpTma) {
        // =========================== TMA producer ============================================
        // (its own warp: issuing the ten boxes of a tile takes the issuing thread a few thousand cycles)
        if (lane == 0) {
            const CUtensorMap *tmap_ptr = &tmap;      // generic address of the __grid_constant__ parameter
            uint32_t it = 0;
#pragma unroll 1
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
                int tx0, ty0, b;
                tile_coords(tile, tx0, ty0, b);
#pragma unroll 1
                for (int g = 0; g < 3; ++g) {
                    if (it > 0)       // the LayerNorm warps have consumed these rows of the previous tile
                        mbar_wait_flag(xraw_empty(g), (it - 1u) & 1u, a.err, (4u << 24) | ((uint32_t)g << 16) | (it & 0xffffu));
                    mbar_expect_tx(xraw_full(g), (uint32_t)(row_group_end(g) - row_group_begin(g)) * kRowBytes);
                    for (int r = row_group_begin(g); r < row_group_end(g); ++r)
                        tma::load_box(smem_u32(xraw + r * kRawRow), tmap_ptr, tx0 - kBoxLeft, ty0 - 1 + r, 0, b, xraw_full(g));
                }
            }
        }
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer ==============================================
        if (lane == 0) {
            auto idesc = [](int n) {
                return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            };
            const uint64_t a_hi0 = make_desc(smem_u32(xhi), kMPos * 16u, 128u);
            const uint64_t a_lo0 = make_desc(smem_u32(xlo), kMPos * 16u, 128u);
            const uint64_t b_0 = make_desc(smem_u32(wsm), 2 * COUT * 16u, 128u);
            uint32_t it = 0, mcount = 0;
            const bool timed = a.dbg != nullptr;
            long long tm[3] = {0, 0, 0}, tp = timed ? clock64() : 0;
#define WM_TICKM(k) do { if (timed) { const long long _t = clock64(); tm[k] += _t - tp; tp = _t; } } while (0)
#pragma unroll 1
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
                auto mmas = [&](int mt, int mg, int buf) {
                    const int ng = COUT == 96 ? (mg == 0 ? 64 : 32) : COUT;      // columns of this MMA group
                    const uint32_t id = idesc(ng);
                    const uint32_t d = tmem_base + (uint32_t)(buf * kAccCols + mt * ng);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t aoff = (uint32_t)(2 * ks * kMPos + mt * 128);
                        const uint32_t boff = (uint32_t)(2 * ks * 2 * COUT + mg * 64);
                        // small terms first: a_lo w_hi + a_hi w_lo + a_hi w_hi (w_lo sits COUT rows after w_hi)
                        mma_tf32_ss(d, a_lo0 + aoff, b_0 + boff, id, ks > 0 ? 1u : 0u);
                        mma_tf32_ss(d, a_hi0 + aoff, b_0 + boff + (uint32_t)COUT, id, 1u);
                        mma_tf32_ss(d, a_hi0 + aoff, b_0 + boff, id, 1u);
                    }
                };
                if (NMG == 1) {
                    // M tile by M tile, as the LayerNorm warps deliver them; each M tile of the operand is
                    // handed back as soon as its MMAs have completed
                    const int buf = mcount & 1;
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) {
                        mbar_wait_flag(xk_full(mt), it & 1u, a.err, (2u << 24) | (1u << 16) | (it & 0xffffu));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        WM_TICKM(0);
                        if (mt == 0) {
                            mbar_wait_flag(acc_empty(buf), ((mcount >> 1) & 1u) ^ 1u, a.err, (2u << 24) | (4u << 16) | (mcount & 0xffffu));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            WM_TICKM(1);
                        }
                        mmas(mt, 0, buf);
                        mma_commit(xk_empty(mt));
                        WM_TICKM(2);
                    }
                    mma_commit(acc_full(buf));
                    ++mcount;
                } else {
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt)
                        mbar_wait_flag(xk_full(mt), it & 1u, a.err, (2u << 24) | (1u << 16) | (it & 0xffffu));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    WM_TICKM(0);
#pragma unroll
                    for (int mg = 0; mg < NMG; ++mg, ++mcount) {
                        const int buf = mcount & 1;
                        mbar_wait_flag(acc_empty(buf), ((mcount >> 1) & 1u) ^ 1u, a.err, (2u << 24) | (4u << 16) | (mcount & 0xffffu));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        WM_TICKM(1);
#pragma unroll
                        for (int mt = 0; mt < 3; ++mt) mmas(mt, mg, buf);
                        mma_commit(acc_full(buf));
                        WM_TICKM(2);
                    }
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) mma_commit(xk_empty(mt));
                }
            }
            if (timed) for (int i = 0; i < 3; ++i) a.dbg[blockIdx.x * 16 + 3 + i] = tm[i];
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
#if WM_PWDW_HALVES
        // the two 16-channel halves of a group are independent in both phases (the depthwise conv is per
        // channel): each half is its own set of 8 warps with its own named barrier, so the halves drift
        // apart and one half's TMEM drain overlaps the other's depthwise phase
        const int ts = (mset * 4 + quarter) * 32 + lane;           // 0..255 inside my half
#define WM_EPI_BAR() named_bar(2 + chalf, kThreadsB / 2)
#else
#define WM_EPI_BAR() named_bar(2, kThreadsB)
#endif
        uint32_t mcount = 0;
        const bool timed = a.dbg != nullptr && tb == 0;
        long long tbb[3] = {0, 0, 0}, tf[5] = {0, 0, 0, 0, 0}, t0 = timed ? clock64() : 0, tp = t0, tq = t0;
        int ntile = 0;
#if WM_PWDW_STAGGER
        if (chalf) __nanosleep(WM_PWDW_STAGGER);       // experiment: start the two halves out of phase
#endif
#define WM_TICKB(k) do { if (timed) { const long long _t = clock64(); tbb[k] += _t - tp; tp = _t; } } while (0)
// finer split of the same warp's time: [11] TMEM loads [12] ps stores [13] barrier after the drain
// [14] depthwise loads + math + stores [15] barrier after the depthwise phase
#define WM_TICKF(k) do { if (timed) { const long long _t = clock64(); tf[k] += _t - tq; tq = _t; } } while (0)
#pragma unroll 1
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ntile) {
            int tx0, ty0, b;
            tile_coords(tile, tx0, ty0, b);
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                // 32-channel group g lives in MMA group mg at column offset gc; the accumulator buffer is
                // acquired at the first group of an MMA group and released after its last one
                const bool mg_first = g == 0 || (COUT == 96 && g == 2);
                const bool mg_last = g == G - 1 || (COUT == 96 && g == 1);
                const int ng = COUT == 96 ? (g < 2 ? 64 : 32) : COUT;
                const int gc = COUT == 96 ? (g < 2 ? g * 32 : 0) : g * 32;
                const int buf = mcount & 1;
                if (mg_first) {
                    mbar_wait_flag(acc_full(buf), (mcount >> 1) & 1u, a.err, (3u << 24) | (3u << 16) | (mcount & 0xffffu));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                WM_TICKB(0);
                float bias16[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(pwb + g * 32 + chalf * 16 + 4 * q);
                    bias16[4 * q] = b4.x; bias16[4 * q + 1] = b4.y; bias16[4 * q + 2] = b4.z; bias16[4 * q + 3] = b4.w;
                }
                if (timed) tq = clock64();
#pragma unroll 1
                for (int mt = mset ? 2 : 0; mt < (mset ? 3 : 2); ++mt) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                           (uint32_t)(buf * kAccCols + mt * ng + gc + chalf * 16);
                    uint32_t acc[16];
                    tmem_ld16(taddr, acc);
                    tmem_ld_wait();
                    WM_TICKF(0);
                    const int pos = mt * 128 + quarter * 32 + lane;
                    if (pos < kPos) {
                        const int row = pos / kPSW, col = pos - row * kPSW;
                        const int gy = ty0 - 1 + row, gx = tx0 - 1 + col;
                        const bool valid = gy >= 0 && gy < h && gx >= 0 && gx < w;     // (col 34, 35: never read)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float v = __uint_as_float(acc[j]) + bias16[j];
                            ps[(chalf * 16 + j) * kPS + pos] = valid ? v : 0.0f;       // consecutive lanes, no conflicts
                        }
                    }
                    WM_TICKF(1);
                }
                if (mg_last) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(buf));
                    ++mcount;
                }
                WM_EPI_BAR();                          // ps of this group (of my channel half) is complete
                WM_TICKB(1);
                WM_TICKF(2);

                // depthwise 3x3 (+ SiLU): thread = (channel, 4 adjacent columns, 4 of the 8 rows), sliding
                // window of three input rows held as packed pairs (FFMA2: two outputs per instruction).
                // Per input row one LDS.128 + one LDS.64 bring the 6 values (v0..v5) four outputs need.
                {
#if WM_PWDW_HALVES
                    const int rhalf = ts >> 7;                       // rows 0-3 or 4-7 of the tile
                    const int cl = chalf * 16 + ((ts & 127) >> 3), j4 = (ts & 7) * 4;
#else
                    const int rhalf = tb >> 8;                       // rows 0-3 or 4-7 of the tile
                    const int cl = (tb & 255) >> 3, j4 = (tb & 7) * 4;
#endif
                    const int co = g * 32 + cl;
                    f32x2 k2[9];
                    const float4 *kp = reinterpret_cast<const float4 *>(dwk + co * 12);
                    const float4 ka = kp[0], kb = kp[1], kc4 = kp[2];
                    k2[0] = pack2(ka.x, ka.x); k2[1] = pack2(ka.y, ka.y); k2[2] = pack2(ka.z, ka.z);
                    k2[3] = pack2(ka.w, ka.w); k2[4] = pack2(kb.x, kb.x); k2[5] = pack2(kb.y, kb.y);
                    k2[6] = pack2(kb.z, kb.z); k2[7] = pack2(kb.w, kb.w); k2[8] = pack2(kc4.x, kc4.x);
                    const float bias = kc4.y;
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
                WM_TICKF(3);
                WM_EPI_BAR();                          // ps may be overwritten by the next group
                WM_TICKB(2);
                WM_TICKF(4);
            }
        }
        if (timed) {
            for (int i = 0; i < 3; ++i) a.dbg[blockIdx.x * 16 + 6 + i] = tbb[i];
            a.dbg[blockIdx.x * 16 + 9] = ntile;
            a.dbg[blockIdx.x * 16 + 10] = clock64() - t0;
            for (int i = 0; i < 5; ++i) a.dbg[blockIdx.x * 16 + 11 + i] = tf[i];
        }
#undef WM_TICKB
#undef WM_TICKF
#undef WM_EPI_BAR
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kWarpMma) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                     : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------
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
    if (!tma::make_tmap_nchw(&tm, x, B, kCin, h, w, kBoxW, 1, kCin)) return 1;   // one halo row per box
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

/* Developer aid: non-NULL device buffer of 16*SMs int64 -> per-CTA cycle counters of the pw_dw pipeline
 * (see Args::dbg); NULL disables. */
extern "C" int wm_pw_dw_debug_timing(void *device_buffer)
{
    wm::pwdw::g_dbg.store(static_cast<long long *>(device_buffer));
    return WM_OK;
}
