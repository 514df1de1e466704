// Dense 3x3 convolution on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// accumulators in TMEM, fp32-accurate through the 3xTF32 split.  NCHW in/out, two-source input with
// a per-image channel gather (the reference's torch.cat + Matching gather, wavemamba_arch.py:716,
// :975), PAConv's 1x1 sigmoid gate fused as a 10th tap (:694-697).
//
// Implicit GEMM without im2col.  Shared memory holds the halo tile in the K-major *no-swizzle*
// UMMA canonical layout  X[kc = ci/4][position][ci%4]  (16 bytes per (kc, position)); with
// SBO = 128 B the row index of the MMA's M dimension is LINEAR in the position, so a 3x3 tap is
// just a different 16-byte-aligned start address in the shared-memory descriptor:
//     D[m][co] += sum_ci X[ci][q0 + m + (dy-1)*34 + (dx-1)] * W_tap[co][ci]
// M enumerates 256 consecutive halo positions (two M=128 MMAs) covering a 7x32-pixel tile; the
// two halo columns per row are computed and dropped (87.5 % useful rows).
//
// 3xTF32: the tensor core reads the top 19 bits of an fp32 word, so the "hi" operand is the raw
// activation and only lo = a - trunc(a) needs a second copy; weights are split (rna) at prepack.
//   [acc | acc2] += a_hi*[b_hi | b_lo]  (one MMA, N = 2*COUT) ; acc += a_lo*b  (N = COUT)
//   fp32 accumulate in TMEM, acc + acc2 in the epilogue.
// The kernel is bound by the shared-memory reads of the MMA operands, and the correction
// term a_lo*b is 2^-11 of the product, so it runs as a **bf16** MMA (kind::f16, K = 16 per instruction):
// a_lo and b rounded to bf16 cost 2^-9 of that term each = 2^-20 of the product, the size of the a_lo*b_lo
// term 3xTF32 drops anyway -- and the a_lo operand and its weights are read at half the bytes
// (11 -> 8.5 KB per 8 input channels and M tile for 32 outputs, 14 -> 11 KB for 64).
// The gated 64 -> 64 variant goes one step further ("KC"): BOTH correction terms share one bf16 MMA with
// K-concatenated operands ([a_lo | a_hi] x [b | b_lo], K = 16 per 8 input channels) and the tf32 MMA is
// the plain a_hi*b_hi with N = COUT.  That costs ~9 % more operand bytes than the merged form but halves
// the accumulator columns: 2 x (main + gate) x 2 M tiles = 512 columns fit TMEM twice, so the gated
// variant's epilogue -- until then exposed, a quarter of its tile time -- overlaps the next tile's MMAs.
//
// Warp-specialised, persistent (one CTA per SM, 15 warps), everything synchronised with mbarriers --
// no CTA-wide barrier inside the tile loop:
//   warps 0-3   X producers: a work unit is (tile, 32-channel K half); its halo slab is loaded
//               global -> registers (all 76 loads of a thread in flight) -> shared as hi AND lo
//               (the 3xTF32 split is fused into the staging), two slabs ring-buffered, so the next
//               unit is staged while the tensor core works on the current one
//   warp 12     weight producer: one TMA bulk copy (cp.async.bulk + mbarrier complete_tx) per
//               (tap, K half) chunk into a 3/4-deep ring of one or two chunks per stage; the prepacked
//               layout is the shared layout
//   warps 13,14 MMA issuers: one thread each (M tile 0 / M tile 1), tcgen05.mma; tcgen05.commit frees
//               ring stages / X slabs and publishes the accumulators
//   warps 4-11  epilogue: tcgen05.ld -> bias or k3*sigmoid(k2+b) -> NCHW / channel-quad stores; the
//               accumulators are double-buffered in TMEM when 2 x columns <= 512 (all variants but
//               the gated 64->64 and the 32->96 one), so the epilogue overlaps the next tile's MMAs
#include <stdlib.h>

#include <atomic>

#include "tc5_common.cuh"

namespace wm {
namespace tc5 {

constexpr int kR = 7, kTW = 32;             // tile: 7 rows x 32 columns
constexpr int kHW = kTW + 2;                // halo row length 34
constexpr int kHaloRows = kR + 2;           // 9
constexpr int kNPos = kHaloRows * kHW;      // 306 halo positions; rows m > 235 of the M window are
                                            // dropped: their reads run past a K-chunk block into the
                                            // next one (still inside this CTA's shared memory)
constexpr int kQ0 = kHW + 1;                // halo position of output (0,0) of the tile
constexpr int kKcSlab = 8;                  // 16-byte K chunks per slab (32 channels)
constexpr int kSlabF4 = kKcSlab * kNPos;    // float4 per slab and per part (hi or lo)
constexpr int kProdWarps = 4, kEpiWarps = 8;
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kWarpW = kProdWarps + kEpiWarps;       // weight producer warp
constexpr int kWarpMma = kWarpW + 1;                  // MMA issuer warps: kWarpMma (M tile 0), +1 (M tile 1)
constexpr int kMmaWarps = 2;
constexpr int kThreads = 32 * (kWarpMma + kMmaWarps); // 480

// optional per-CTA timing of the MMA thread (cycles), enabled with wm_conv3x3_debug_timing(ptr):
// [0] total  [1] wait weights  [2] wait X  [3] wait accumulator buffer  [4] issue  [5] tiles
static std::atomic<long long *> g_dbg{nullptr};

struct Args {
    long long *dbg;
    const float *in_a;
    int64_t a_bstride;
    int Ca;
    const float *in_b;
    int64_t b_bstride;
    const int *chan_map;
    const float4 *packed;    // [tap][part][kc 0..7][hi|lo][COUT] float4 (4 consecutive ci)
    const float *bias;
    const float *gate_bias;
    float *out;
    int h, w;
    // channel-quad layouts (B, C/4, h, w, 4): one 16-byte access moves the 4 channels of a K chunk.
    // in_c4 needs Ca == CIN (no second input); used between PAConv.k3 and k4.
    int in_c4, out_c4;
    unsigned int *err;       // pipeline error word (mbar_wait_flag)
};

template <int CIN, int COUT, bool GATE>
struct Cfg {
    static constexpr int NCH = CIN / 32;                       // K halves (slabs) per tile
    static constexpr int NTAPS = GATE ? 10 : 9;
    // K-concatenated bf16 corrections (see the header): the variants whose merged accumulators do not fit
    // TMEM twice -- the gated 64 -> 64 one and 32 -> 96
    static constexpr bool KC = GATE || COUT >= 96;
    // weight chunk: merged [kc 0..7][hi | lo][COUT] float4 + [kc8 0..3][COUT] 8 x bf16 of w;
    //               KC     [kc 0..7][hi][COUT] float4      + [kc8 0..3][w | w_lo][COUT] 8 x bf16
    static constexpr int kChunkTf32F4 = KC ? kKcSlab * COUT : kKcSlab * 2 * COUT;
    static constexpr int kChunkF4 = kChunkTf32F4 + (KC ? 8 : 4) * COUT;
    static constexpr int kLoSlabF4 = KC ? kSlabF4 : kSlabF4 / 2;      // 16-byte units of one slot of the bf16 slab
    static constexpr int kAccCols = KC ? COUT : 2 * COUT;             // TMEM columns per (M tile, region)
    static constexpr int kChunkBytes = kChunkF4 * 16;
    // weight ring depth (what fits next to the X slabs).  The issuing threads run ahead of the tensor
    // core until they meet a stage that is still being refilled, so their timers always show a wait on
    // the weights (~350 cycles per chunk at any depth): it is slack, not a stall of the tensor pipe.
    // A ring stage holds kUnits consecutive (K half, tap) chunks behind ONE pair of barriers: every stage
    // hand-over costs each issuing thread ~350 cycles whether or not the data is there (measured by running
    // the pipeline without refills and without waits: 19.3 k -> 16.2 k cycles per tile for 64 -> 32), about
    // as much as the 8 MMAs of a 32-output chunk take to issue.  Two chunks per stage where the ring has the
    // room (the 32-output variants).
    static constexpr int kUnits = COUT <= 32 ? 2 : 1;
    static constexpr int kStageF4 = kUnits * kChunkF4;         // float4 per ring stage
    static constexpr int kStages = COUT >= 96 ? 3 : 4;
    static constexpr int kTileUnits = NCH * NTAPS;             // chunks per tile, K half major
    static constexpr int kTileStages = (kTileUnits + kUnits - 1) / kUnits;
    static constexpr int kColsBuf = 2 * kAccCols * (GATE ? 2 : 1);    // TMEM columns of one accumulator set
    static constexpr int NACC = 2 * kColsBuf <= 512 ? 2 : 1;
    static constexpr int kColsNeed = NACC * kColsBuf;
    static constexpr int kCols = kColsNeed <= 32 ? 32 : kColsNeed <= 64 ? 64 : kColsNeed <= 128 ? 128
                                 : kColsNeed <= 256 ? 256 : 512;
    static constexpr int kNumBars = 2 * kStages + 4 + 2 * NACC;
    static constexpr size_t kSmem = sizeof(float4) * (size_t)(2 * kSlabF4 + 2 * kLoSlabF4 + kStages * kStageF4) +
                                    8 * kNumBars + 16;
    static_assert(CIN == 32 || CIN == 64 || CIN == 96, "CIN must be 32, 64 or 96");
    static_assert(2 * COUT <= 256 && (2 * COUT) % 16 == 0, "merged N must be a legal UMMA N");
    static_assert(kColsNeed <= 512, "accumulators do not fit TMEM");
    static_assert(kSmem <= 232448, "shared memory budget");
};

template <int CIN, int COUT, bool GATE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc5_kernel(const Args a, int tiles_x, int tiles_y, int total_tiles)
{
    using C = Cfg<CIN, COUT, GATE>;
    constexpr int NCH = C::NCH, NTAPS = C::NTAPS, kStages = C::kStages, NACC = C::NACC;
    extern __shared__ __align__(128) float smem[];
    float4 *xhi = reinterpret_cast<float4 *>(smem);                 // [2 slots][8 kc][kNPos]
    uint4 *xlo = reinterpret_cast<uint4 *>(xhi + 2 * kSlabF4);      // [2 slots][4 kc8]([lo | hi] if KC)[kNPos] 8 x bf16
    float4 *wbuf = xhi + 2 * kSlabF4 + 2 * C::kLoSlabF4;            // [kStages][kUnits][kChunkF4]
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbuf + kStages * C::kStageF4);
    const uint32_t bar0 = smem_u32(bars);
    // barrier map (8 bytes each)
    auto wfull = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto wempty = [&](int i) { return bar0 + 8u * (uint32_t)(kStages + i); };
    auto xfull = [&](int i) { return bar0 + 8u * (uint32_t)(2 * kStages + i); };
    auto xempty = [&](int i) { return bar0 + 8u * (uint32_t)(2 * kStages + 2 + i); };
    auto accfull = [&](int i) { return bar0 + 8u * (uint32_t)(2 * kStages + 4 + i); };
    auto accempty = [&](int i) { return bar0 + 8u * (uint32_t)(2 * kStages + 4 + NACC + i); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + C::kNumBars);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = a.h, w = a.w;
    const int64_t hw = (int64_t)h * w;

    // ---- one-time setup: barriers (one thread), TMEM allocation (the MMA warp) ---------------
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(wfull(i), 1); mbar_init(wempty(i), kMmaWarps); }
        for (int i = 0; i < 2; ++i) { mbar_init(xfull(i), kProdThreads); mbar_init(xempty(i), kMmaWarps); }
        for (int i = 0; i < NACC; ++i) { mbar_init(accfull(i), kMmaWarps); mbar_init(accempty(i), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"((uint32_t)C::kCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kProdWarps) {
        // =============================== X producers ========================================
        // warp pw stages K chunks 2pw, 2pw+1 of the slab: 18 halo rows, lanes along the row
        // (columns 0..31); the two tail columns of all 72 rows are spread over the 128 threads.
        const int pw = warp;
        uint32_t unit = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y;
            const int64_t b = tile / (tiles_x * tiles_y);
            const int tx0 = txi * kTW, ty0 = tyi * kR;
            const int gx = tx0 - 1 + lane;
            const bool okx = gx >= 0 && gx < w;
#pragma unroll 1
            for (int part = 0; part < NCH; ++part, ++unit) {
                const int slot = unit & 1;
                float4 v[18];
                float4 vt[2];
                // tail items: thread i < 72 -> (K chunks 2*(i/18), 2*(i/18)+1; py = (i % 18) / 2; column
                // 32 + (i & 1)): the 8 channels of one bf16 operand row
                const bool has_tail = tid < 72;
                int t_kc[2];
                const int t_py = (tid % 18) >> 1, t_px = 32 + (tid & 1);
                t_kc[0] = 2 * (tid / 18);
                t_kc[1] = t_kc[0] + 1;
                if (a.in_c4) {
                    const float4 *src4 = reinterpret_cast<const float4 *>(a.in_a + b * a.a_bstride);
#pragma unroll
                    for (int r = 0; r < 18; ++r) {
                        const int kcl = 2 * pw + r / 9, py = r % 9;
                        const int gy = ty0 - 1 + py;
                        const bool ok = okx && gy >= 0 && gy < h;
                        v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) v[r] = __ldg(src4 + (int64_t)(part * kKcSlab + kcl) * hw + (int64_t)gy * w + gx);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        vt[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_tail) {
                            const int gy = ty0 - 1 + t_py, gxt = tx0 - 1 + t_px;
                            if (gy >= 0 && gy < h && gxt < w)
                                vt[q] = __ldg(src4 + (int64_t)(part * kKcSlab + t_kc[q]) * hw + (int64_t)gy * w + gxt);
                        }
                    }
                } else {
                    // plane of input channel c: first Ca channels from in_a, the rest from in_b
                    // (optionally gathered through the per-image channel map)
                    auto plane_of = [&](int c) -> const float * {
                        if (c < a.Ca) return a.in_a + b * a.a_bstride + (int64_t)c * hw;
                        const int cb = a.chan_map ? __ldg(a.chan_map + b * (CIN - a.Ca) + (c - a.Ca)) : c - a.Ca;
                        return a.in_b + b * a.b_bstride + (int64_t)cb * hw;
                    };
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int c0 = part * 32 + (2 * pw + kk) * 4;
                        const float *p0 = plane_of(c0), *p1 = plane_of(c0 + 1), *p2 = plane_of(c0 + 2),
                                    *p3 = plane_of(c0 + 3);
#pragma unroll
                        for (int py = 0; py < 9; ++py) {
                            const int gy = ty0 - 1 + py;
                            const bool ok = okx && gy >= 0 && gy < h;
                            const int64_t off = (int64_t)gy * w + gx;
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ok) { t.x = __ldg(p0 + off); t.y = __ldg(p1 + off); t.z = __ldg(p2 + off); t.w = __ldg(p3 + off); }
                            v[kk * 9 + py] = t;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        vt[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_tail) {
                            const int gy = ty0 - 1 + t_py, gxt = tx0 - 1 + t_px;
                            if (gy >= 0 && gy < h && gxt < w) {
                                const int c0 = part * 32 + t_kc[q] * 4;
                                const int64_t off = (int64_t)gy * w + gxt;
                                vt[q].x = __ldg(plane_of(c0) + off); vt[q].y = __ldg(plane_of(c0 + 1) + off);
                                vt[q].z = __ldg(plane_of(c0 + 2) + off); vt[q].w = __ldg(plane_of(c0 + 3) + off);
                            }
                        }
                    }
                }
                // the loads above are in flight while we wait for the slab to be released
                mbar_wait_flag(xempty(slot), ((unit >> 1) & 1u) ^ 1u, a.err, (1u << 24) | (1u << 16) | (unit & 0xffffu));
                float4 *dhi = xhi + slot * kSlabF4;
                uint4 *dlo = xlo + slot * C::kLoSlabF4;
                constexpr int kLoRows = C::KC ? 2 : 1;       // 16-byte rows per K-chunk pair: lo (and bf16 of a)
                auto hi8 = [](const float4 &p, const float4 &q) {
                    return make_uint4(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w), pack_bf16x2(q.x, q.y),
                                      pack_bf16x2(q.z, q.w));
                };
                // lo = a - trunc_tf32(a) of 8 consecutive channels (K chunks 2pw, 2pw+1) -> one bf16 row
                auto lo8 = [](const float4 &p, const float4 &q) {
                    return make_uint4(pack_bf16x2(tf32_lo(p.x), tf32_lo(p.y)), pack_bf16x2(tf32_lo(p.z), tf32_lo(p.w)),
                                      pack_bf16x2(tf32_lo(q.x), tf32_lo(q.y)), pack_bf16x2(tf32_lo(q.z), tf32_lo(q.w)));
                };
#pragma unroll
                for (int r = 0; r < 18; ++r) dhi[(2 * pw + r / 9) * kNPos + (r % 9) * kHW + lane] = v[r];
#pragma unroll
                for (int py = 0; py < 9; ++py) {
                    dlo[pw * kLoRows * kNPos + py * kHW + lane] = lo8(v[py], v[9 + py]);
                    if (C::KC) dlo[(pw * 2 + 1) * kNPos + py * kHW + lane] = hi8(v[py], v[9 + py]);
                }
                if (has_tail) {
                    const int pos = t_py * kHW + t_px;
                    dhi[t_kc[0] * kNPos + pos] = vt[0];
                    dhi[t_kc[1] * kNPos + pos] = vt[1];
                    dlo[(t_kc[0] >> 1) * kLoRows * kNPos + pos] = lo8(vt[0], vt[1]);
                    if (C::KC) dlo[((t_kc[0] >> 1) * 2 + 1) * kNPos + pos] = hi8(vt[0], vt[1]);
                }
                // generic-proxy stores -> visible to the tensor core's async-proxy reads
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(xfull(slot));
            }
        }
    } else if (warp == kWarpW) {
        // =============================== weight producer ====================================
        if (lane == 0) {
            uint32_t cnt = 0;
#pragma unroll 1
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
#pragma unroll 1
                for (int sc = 0; sc < C::kTileStages; ++sc, ++cnt) {
                    const int st = cnt % kStages;
                    const int u0 = sc * C::kUnits;
                    const int nu = C::kTileUnits - u0 < C::kUnits ? C::kTileUnits - u0 : C::kUnits;
                    mbar_wait_flag(wempty(st), ((cnt / kStages) & 1u) ^ 1u, a.err, (4u << 24) | (2u << 16) | (cnt & 0xffffu));
                    mbar_expect_tx(wfull(st), (uint32_t)(nu * C::kChunkBytes));
                    for (int i = 0; i < nu; ++i) {
                        const int part = (u0 + i) / NTAPS, tap = (u0 + i) - part * NTAPS;
                        bulk_g2s(smem_u32(wbuf + st * C::kStageF4 + i * C::kChunkF4),
                                 a.packed + (int64_t)(tap * NCH + part) * C::kChunkF4,
                                 (uint32_t)C::kChunkBytes, wfull(st));
                    }
                }
            }
        }
    } else if (warp >= kWarpMma) {
        // =============================== MMA issuers ========================================
        // One thread needs ~54 cycles to issue a tcgen05.mma (measured: 16 MMAs per chunk in ~860 cycles)
        // -- about what the tensor core needs to EXECUTE one (48-64 cycles), so a single issuer that also
        // pays ~370 cycles of barrier latency per chunk starves the pipe.  Two issuer warps split the
        // two M tiles (disjoint TMEM columns, no ordering between them); every barrier they release
        // counts both.
        const int my_mt = warp - kWarpMma;
        if (lane == 0) {
            // instruction descriptors: D=f32, A=B=tf32, both K-major, M = 128; N = COUT (b_hi only)
            // or N = 2*COUT ([b_hi | b_lo]: the activation operand is read once for both)
            constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) |
                                        ((uint32_t)((2 * COUT) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // descriptors differ only in the 14-bit start-address field (bytes >> 4)
            const uint64_t a_hi0 = make_desc(smem_u32(xhi), kNPos * 16u, 128u);
            constexpr uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                         ((uint32_t)(128 >> 4) << 24);     // kind::f16: bf16 x bf16 -> f32
            const uint64_t a_lo0 = make_desc(smem_u32(xlo), kNPos * 16u, 128u);
            constexpr uint32_t idescN = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                        ((uint32_t)(128 >> 4) << 24);      // tf32, N = COUT (KC)
            const uint64_t b_00 = make_desc(smem_u32(wbuf), (C::KC ? 1 : 2) * COUT * 16u, 128u);
            const uint64_t b16_00 = make_desc(smem_u32(wbuf + C::kChunkTf32F4), COUT * 16u, 128u);
            uint32_t cnt = 0, unit = 0, tcount = 0;
            long long tacc[5] = {0, 0, 0, 0, 0}, t0 = 0, tp = 0;
            const bool timed = a.dbg != nullptr && my_mt == 0;
            if (timed) { t0 = clock64(); tp = t0; }
#define WM_TICK(k) do { if (timed) { const long long _t = clock64(); tacc[k] += _t - tp; tp = _t; } } while (0)
#pragma unroll 1
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                const int buf = tcount % NACC;
                mbar_wait_flag(accempty(buf), ((tcount / NACC) & 1u) ^ 1u, a.err, (2u << 24) | (4u << 16) | (tcount & 0xffffu));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                WM_TICK(3);
                const uint32_t dbase = tmem_base + (uint32_t)(buf * C::kColsBuf);
#pragma unroll 1
                for (int part = 0; part < NCH; ++part, ++unit) {
                    const int slot = unit & 1;
                    mbar_wait_flag(xfull(slot), (unit >> 1) & 1u, a.err, (2u << 24) | (1u << 16) | (unit & 0xffffu));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    WM_TICK(2);
#pragma unroll 1
                    for (int ti = 0; ti < NTAPS; ++ti) {
                        const int tap = ti;
                        const int u = part * NTAPS + ti;               // chunk of this tile
                        const int ui = u % C::kUnits;                  // position inside its ring stage
                        const int st = cnt % kStages;
                        if (ui == 0) {
                            mbar_wait_flag(wfull(st), (cnt / kStages) & 1u, a.err, (2u << 24) | (2u << 16) | (cnt & 0xffffu));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        WM_TICK(1);
                        const bool gate_tap = GATE && tap == 9;
                        const int dy = gate_tap ? 1 : tap / 3, dx = gate_tap ? 1 : tap - (tap / 3) * 3;
                        // offsets in 16-byte units (= one (kc, position) or (kc, co) element)
                        const uint32_t shift = (uint32_t)(slot * kSlabF4 + dy * kHW + dx);
                        const uint64_t b_hi0 = b_00 + (uint32_t)(st * C::kStageF4 + ui * C::kChunkF4);
                        const uint64_t b16_0 = b16_00 + (uint32_t)(st * C::kStageF4 + ui * C::kChunkF4);
                        const uint32_t shift16 = (uint32_t)(slot * C::kLoSlabF4 + dy * kHW + dx);
                        {
                            const int mt = my_mt;
                            const uint32_t dcol = dbase + (uint32_t)((gate_tap ? 2 * C::kAccCols : 0) + mt * C::kAccCols);
                            const uint32_t arow = shift + (uint32_t)mt * 128u;
#pragma unroll
                            for (int kl = 0; kl < 4; ++kl) {
                                const uint32_t aoff = (uint32_t)(2 * kl) * kNPos + arow;
                                const uint32_t boff = (uint32_t)(2 * kl) * (C::KC ? 1 : 2) * COUT;
                                // the first MMA into an accumulator region overwrites it: the main
                                // region at the first 3x3 tap of K half 0, the gate region at the gate tap
                                const uint32_t first = (part == 0 && kl == 0 && (gate_tap || ti == 0)) ? 0u : 1u;
                                // cols [0,COUT) += a_hi b_hi, [COUT,2COUT) += a_hi b_lo
                                // (KC: cols [0,COUT) += a_hi b_hi only)
                                mma_tf32_ss(dcol, a_hi0 + aoff, b_hi0 + boff, C::KC ? idescN : idesc2, first);
                            }
                            // cols [0,COUT) += a_lo b in bf16, 16 input channels per instruction
                            // (KC: += [a_lo | a] [b | b_lo], the two corrections of 8 input channels per instruction)
                            const uint32_t arow16 = shift16 + (uint32_t)mt * 128u;
#pragma unroll
                            for (int kl = 0; kl < (C::KC ? 4 : 2); ++kl)
                                mma_bf16_ss(dcol, a_lo0 + (uint32_t)(2 * kl) * kNPos + arow16,
                                            b16_0 + (uint32_t)(2 * kl) * COUT, idesc16, 1u);
                        }
                        if (ui == C::kUnits - 1 || u == C::kTileUnits - 1) {
                            mma_commit(wempty(st));
                            ++cnt;
                        }
                        WM_TICK(4);
                    }
                    mma_commit(xempty(slot));
                }
                mma_commit(accfull(buf));
            }
#undef WM_TICK
            if (timed) {
                a.dbg[blockIdx.x * 6 + 0] = clock64() - t0;
                for (int k = 1; k < 5; ++k) a.dbg[blockIdx.x * 6 + k] = tacc[k];
                a.dbg[blockIdx.x * 6 + 5] = tcount;
            }
        }
    } else {
        // =============================== epilogue ===========================================
        // warp e: TMEM lane quarter (warp % 4 -- the hardware restriction), M tile e / 4
        const int e = warp - kProdWarps;
        const int quarter = warp & 3, mt = e >> 2;
        constexpr int NG = COUT / 32;                   // 32-channel groups per accumulator
        uint32_t tcount = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y;
            const int64_t b = tile / (tiles_x * tiles_y);
            const int tx0 = txi * kTW, ty0 = tyi * kR;
            const int buf = tcount % NACC;
            const int m = mt * 128 + quarter * 32 + lane;
            const int q = kQ0 + m;
            const int py = q / kHW, px = q - py * kHW;
            const int gy = ty0 + py - 1, gx = tx0 + px - 1;
            const bool ok = px >= 1 && px <= kTW && py <= kR && gy < h && gx < w;
            mbar_wait_flag(accfull(buf), (tcount / NACC) & 1u, a.err, (3u << 24) | (3u << 16) | (tcount & 0xffffu));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                       (uint32_t)(buf * C::kColsBuf + mt * C::kAccCols);
#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                const int c0 = g * 32;
                uint32_t acc[32];
                if (!C::KC) {
                    uint32_t part[32];
                    tmem_ld32(lane_addr + (uint32_t)c0, acc);              // both loads in flight
                    tmem_ld32(lane_addr + (uint32_t)(COUT + c0), part);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(part[j]));
                } else if (GATE) {
                    // KC, gated: one accumulator per region; main and gate loads in flight together
                    uint32_t gt[32];
                    tmem_ld32(lane_addr + (uint32_t)c0, acc);
                    tmem_ld32(lane_addr + (uint32_t)(2 * C::kAccCols + c0), gt);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float z = __uint_as_float(gt[j]) + __ldg(a.gate_bias + c0 + j);
                        acc[j] = __float_as_uint(__fdividef(__uint_as_float(acc[j]), 1.0f + __expf(-z)));
                    }
                } else {
                    tmem_ld32(lane_addr + (uint32_t)c0, acc);          // KC, plain: one accumulator
                    tmem_ld_wait();
                }
                if (g == NG - 1) {
                    // every TMEM read of this warp is done: hand the accumulators back before the
                    // global stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(accempty(buf));
                }
                if (ok && a.out_c4) {
                    float4 *o4 = reinterpret_cast<float4 *>(a.out) +
                                 (b * (COUT / 4) + c0 / 4) * hw + (int64_t)gy * w + gx;
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[j] = __uint_as_float(acc[4 * qq + j]);
                            if (!GATE && a.bias) v[j] += __ldg(a.bias + c0 + 4 * qq + j);
                        }
                        o4[(int64_t)qq * hw] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                } else if (ok) {
                    float *o = a.out + (b * COUT + c0) * hw + (int64_t)gy * w + gx;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v = __uint_as_float(acc[j]);
                        if (!GATE && a.bias) v += __ldg(a.bias + c0 + j);
                        o[(int64_t)j * hw] = v;
                    }
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kWarpMma) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)C::kCols)
                     : "memory");
    }
}

// w3: (COUT, CIN, 3, 3); w1: (COUT, CIN) or null -> packed[tap][part]{ [kc 0..7][hi|lo][co] float4,
// [kc8 0..3][co] 8 x bf16 }: one contiguous chunk per (tap, K half) in exactly the shared-memory operand
// layout, so the weight producer moves a chunk with a single bulk copy.
__global__ void __launch_bounds__(256)
prepack_tc5_kernel(const float *__restrict__ w3, const float *__restrict__ w1,
                   float4 *__restrict__ out, int CIN, int COUT, int ntaps, bool kc_layout)
{
    const int KC = CIN / 4, NCH = CIN / 32;
    const int total = ntaps * KC * COUT;        // kc_layout: the chunk layout of the Cfg::KC variants
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int co = i % COUT;
        const int kc = (i / COUT) % KC;
        const int tap = i / (COUT * KC);
        float hi[4], lo[4], v_lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = kc * 4 + j;
            const float v = tap < 9 ? w3[((int64_t)co * CIN + ci) * 9 + tap] : w1[(int64_t)co * CIN + ci];
            uint32_t hbits, lbits;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hbits) : "f"(v));
            const float rest = v - __uint_as_float(hbits);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lbits) : "f"(rest));
            hi[j] = __uint_as_float(hbits);
            lo[j] = __uint_as_float(lbits);
            v_lo[j] = rest;
        }
        const int part = kc / kKcSlab, kcl = kc % kKcSlab;
        float wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = kc * 4 + j;
            wv[j] = tap < 9 ? w3[((int64_t)co * CIN + ci) * 9 + tap] : w1[(int64_t)co * CIN + ci];
        }
        const uint2 w16 = make_uint2(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]));
        if (!kc_layout) {
            float4 *chunk = out + (int64_t)(tap * NCH + part) * (kKcSlab * 2 + 4) * COUT;
            float4 *dst = chunk + kcl * 2 * COUT;
            dst[co] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            dst[COUT + co] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            // bf16 copy of the weights for the a_lo term: row co of K chunk pair kcl/2, this half of its 16 bytes
            reinterpret_cast<uint2 *>(chunk + kKcSlab * 2 * COUT + (kcl >> 1) * COUT + co)[kcl & 1] = w16;
        } else {
            // KC variants: [kc][hi][co] float4, then [kc8][w | w_lo][co] 8 x bf16
            float4 *chunk = out + (int64_t)(tap * NCH + part) * (kKcSlab + 8) * COUT;
            chunk[kcl * COUT + co] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            float4 *b16 = chunk + kKcSlab * COUT + (kcl >> 1) * 2 * COUT;
            reinterpret_cast<uint2 *>(b16 + co)[kcl & 1] = w16;
            reinterpret_cast<uint2 *>(b16 + COUT + co)[kcl & 1] =
                make_uint2(pack_bf16x2(v_lo[0], v_lo[1]), pack_bf16x2(v_lo[2], v_lo[3]));
        }
    }
}

template <int CIN, int COUT, bool GATE>
int launch(const Args &a, int64_t B, cudaStream_t s)
{
    using C = Cfg<CIN, COUT, GATE>;
    WM_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc5_kernel<CIN, COUT, GATE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    const int tiles_x = (a.w + kTW - 1) / kTW, tiles_y = (a.h + kR - 1) / kR;
    const int64_t total = (int64_t)tiles_x * tiles_y * B;
    WM_REQUIRE(total < (int64_t)1 << 31, "wm_conv3x3_fwd: too many tiles");
    const int grid = (int)(total < sm_count() ? total : sm_count());   // persistent: one CTA per SM
    conv3x3_tc5_kernel<CIN, COUT, GATE><<<grid, kThreads, C::kSmem, s>>>(a, tiles_x, tiles_y, (int)total);
    WM_LAUNCH_OK("conv3x3 tcgen05");
    return WM_OK;
}

}  // namespace tc5
}  // namespace wm

using namespace wm;

/* Developer aid (not part of the reference-facing surface): when `device_buffer` is non-NULL the
 * kernel's MMA thread adds cycle counts into it (6 int64 per CTA, 148 CTAs max). */
extern "C" int wm_conv3x3_debug_timing(void *device_buffer)
{
    tc5::g_dbg.store(static_cast<long long *>(device_buffer));
    return WM_OK;
}

extern "C" size_t wm_conv3x3_packed_bytes(int64_t Cin, int64_t Cout, int with_gate)
{
    if (Cin <= 0 || Cout <= 0 || Cin % 32 || Cout % 8) return 0;
    // per (tap, 32-channel K half): 16 * Cout float4 of tf32 hi | lo + 4 * Cout of bf16 (w);
    // gated and Cout >= 96 (Cfg::KC): 8 * Cout of tf32 hi + 8 * Cout of bf16 (w | w_lo)
    const bool kc = with_gate || Cout >= 96;
    return (size_t)(with_gate ? 10 : 9) * (kc ? 16 : 20) * (Cin / 32) * Cout * sizeof(float4);
}

extern "C" int wm_conv3x3_prepack(const float *w3x3, const float *w1x1, void *packed, int64_t Cin,
                                  int64_t Cout, wm_stream_t stream)
{
    WM_REQUIRE(w3x3 && packed, "wm_conv3x3_prepack: null pointer");
    WM_REQUIRE(Cin > 0 && Cout > 0 && Cin % 32 == 0 && Cout % 8 == 0 && Cin <= 512 && Cout <= 512,
               "wm_conv3x3_prepack: Cin=%lld must be a multiple of 32, Cout=%lld of 8", (long long)Cin,
               (long long)Cout);
    WM_REQUIRE(aligned16(packed), "wm_conv3x3_prepack: packed must be 16-byte aligned");
    const int ntaps = w1x1 ? 10 : 9;
    const int total = ntaps * (int)(Cin / 4) * (int)Cout;
    tc5::prepack_tc5_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        w3x3, w1x1, static_cast<float4 *>(packed), (int)Cin, (int)Cout, ntaps, w1x1 != nullptr || Cout >= 96);
    WM_LAUNCH_OK("conv3x3 prepack");
    return WM_OK;
}

extern "C" int wm_conv3x3_ex_fwd(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b,
                                 int64_t b_bstride, const int *chan_map, const void *packed,
                                 const float *bias, const float *gate_bias, float *out, int64_t B,
                                 int64_t Cin, int64_t Cout, int64_t h, int64_t w, int in_c4, int out_c4,
                                 wm_stream_t stream)
{
    using namespace wm::tc5;
    WM_REQUIRE(B >= 0 && B <= 65535 && h >= 0 && w >= 0 && h < (1 << 24) && w < (1 << 24),
               "wm_conv3x3_fwd: bad sizes");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(in_a && packed && out, "wm_conv3x3_fwd: null pointer");
    WM_REQUIRE(Ca > 0 && Ca <= Cin && (Ca == Cin || in_b != nullptr),
               "wm_conv3x3_fwd: Ca=%lld of Cin=%lld needs a second input", (long long)Ca, (long long)Cin);
    WM_REQUIRE((h + kR - 1) / kR <= 65535, "wm_conv3x3_fwd: image too tall");
    WM_REQUIRE(aligned16(packed), "wm_conv3x3_fwd: packed weights must be 16-byte aligned");
    WM_REQUIRE(!in_c4 || (Ca == Cin && aligned16(in_a) && a_bstride % 4 == 0),
               "wm_conv3x3_ex_fwd: the channel-quad input layout needs a single 16-byte aligned input");
    WM_REQUIRE(!out_c4 || aligned16(out), "wm_conv3x3_ex_fwd: channel-quad output must be 16-byte aligned");
    Args a;
    a.in_c4 = in_c4; a.out_c4 = out_c4;
    a.dbg = g_dbg.load();
    a.err = pipeline_err_word();
    WM_REQUIRE(a.err != nullptr, "wm_conv3x3_fwd: no CUDA device");

    a.in_a = in_a; a.a_bstride = a_bstride; a.Ca = (int)Ca; a.in_b = in_b; a.b_bstride = b_bstride;
    a.chan_map = chan_map; a.packed = static_cast<const float4 *>(packed); a.bias = bias;
    a.gate_bias = gate_bias; a.out = out; a.h = (int)h; a.w = (int)w;
    cudaStream_t s = (cudaStream_t)stream;
    if (gate_bias) {
        WM_REQUIRE(Cin == 64 && Cout == 64, "wm_conv3x3_fwd: gated mode supports 64->64 only");
        return launch<64, 64, true>(a, B, s);
    }
    if (Cin == 64 && Cout == 32) return launch<64, 32, false>(a, B, s);
    if (Cin == 64 && Cout == 64) return launch<64, 64, false>(a, B, s);
    if (Cin == 32 && Cout == 96) return launch<32, 96, false>(a, B, s);
    if (Cin == 32 && Cout == 32) return launch<32, 32, false>(a, B, s);
    // data gradients of the 64->32 and 32->96 convolutions (training path)
    if (Cin == 32 && Cout == 64) return launch<32, 64, false>(a, B, s);
    if (Cin == 96 && Cout == 32) return launch<96, 32, false>(a, B, s);
    WM_REQUIRE(false, "wm_conv3x3_fwd: Cin=%lld Cout=%lld unsupported", (long long)Cin, (long long)Cout);
    return WM_EINVAL;
}

extern "C" int wm_conv3x3_fwd(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b,
                              int64_t b_bstride, const int *chan_map, const void *packed,
                              const float *bias, const float *gate_bias, float *out, int64_t B,
                              int64_t Cin, int64_t Cout, int64_t h, int64_t w, wm_stream_t stream)
{
    return wm_conv3x3_ex_fwd(in_a, a_bstride, Ca, in_b, b_bstride, chan_map, packed, bias, gate_bias, out,
                             B, Cin, Cout, h, w, 0, 0, stream);
}
