// Dense 3x3 convolution on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// accumulators in TMEM, fp32-accurate through the 3xTF32 split.  Same contract as conv3x3.cu
// (NCHW in/out, two-source input with channel gather, PAConv gate fused); selected by
// wm_conv3x3_set_impl(1).
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
//   [acc | acc2] += a_hi*[b_hi | b_lo]  (one MMA, N = 2*COUT) ; acc += a_lo*b_hi  (N = COUT)
//   fp32 accumulate in TMEM, acc + acc2 in the epilogue.  The SS-mode MMA is bound by reading the
//   activation operand from shared memory, so the merged N halves the a_hi reads.
//
// Persistent: one CTA per SM (217 KB of shared memory, TMEM allocated once) loops over tiles.
// Per tile: finish the cp.async halo staging -> MMA pipeline over a 3-deep cp.async weight ring
// (one thread issues tcgen05.mma, tcgen05.commit -> mbarrier frees ring slots) -> start the NEXT
// tile's halo loads -> all 8 warps drain TMEM with tcgen05.ld.32x32b and run the epilogue while
// those loads are in flight.
#include "common.cuh"

namespace wm {
namespace tc5 {

constexpr int kR = 7, kTW = 32;             // tile: 7 rows x 32 columns
constexpr int kHW = kTW + 2;                // halo row length 34
constexpr int kHaloPos = (kR + 2) * kHW;    // 306 real halo positions
constexpr int kNPos = kHaloPos;             // rows m > 235 of the M window are dropped: their reads
                                            // may run past a K-chunk block into the next one (still
                                            // inside this CTA's shared memory), harmless garbage
constexpr int kQ0 = kHW + 1;                // halo position of output (0,0) of the tile
constexpr int kThreads = 256;
constexpr int kStages = 4;                  // weight-chunk ring depth (prefetch distance 3)

// optional per-CTA phase timing (cycles), enabled with wm_conv3x3_debug_timing(ptr != NULL):
// [0] wait for staged X  [1] xlo  [2] chunk loop  [3] drain  [4] issue next stage  [5] epilogue
static long long *g_dbg = nullptr;

struct Args {
    long long *dbg;
    const float *in_a;
    int64_t a_bstride;
    int Ca;
    const float *in_b;
    int64_t b_bstride;
    const int *chan_map;
    const float4 *packed;    // [ntaps][2 (hi,lo)][CIN/4][COUT] float4 (4 consecutive ci)
    const float *bias;
    const float *gate_bias;
    float *out;
    int h, w;
    // channel-quad layouts (B, C/4, h, w, 4): a 16-byte cp.async / store moves the 4 channels of one
    // K chunk at once.  in_c4 needs Ca == CIN (no second input); used between PAConv.k3 and k4.
    int in_c4, out_c4;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 8-row groups SBO bytes apart, the two 16-byte K chunks LBO bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version for sm_100
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u),
          "r"(0u)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok = 0;
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {   // non-blocking probe, bounded (~0.3 s)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();   // never hang the GPU on a protocol bug
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

template <int COUT, bool GATE>
constexpr int tmem_cols()
{
    // per M tile: COUT columns for a_hi*b_hi + a_lo*b_hi and COUT columns for a_hi*b_lo
    constexpr int need = 2 * 2 * COUT * (GATE ? 2 : 1);
    return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

template <int CIN, int COUT, bool GATE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc5_kernel(const Args a, int tiles_x, int tiles_y, int total_tiles)
{
    constexpr int KC = CIN / 4;                 // 16-byte K chunks
    constexpr int KS = CIN / 8;                 // MMA k-steps (K = 8 for tf32)
    constexpr int NTAPS = GATE ? 10 : 9;
    constexpr int kWF4 = KC * COUT;             // float4 per weight part (hi or lo) per tap
    constexpr int NCH = CIN >= 64 ? 2 : 1;      // weight chunks per tap (K split so the ring fits)
    constexpr int KSC = KS / NCH;               // k-steps per chunk
    constexpr int kCF4 = 2 * kWF4 / NCH;        // float4 per chunk: [hi | lo][2*KSC kc][COUT]
    constexpr int kCols = tmem_cols<COUT, GATE>();
    constexpr int NCHUNK = NTAPS * NCH;
    static_assert(NCHUNK >= 4, "pipeline prologue assumes at least four chunks");
    extern __shared__ __align__(128) float smem[];
    float4 *xhi = reinterpret_cast<float4 *>(smem);          // [KC][kNPos]
    float4 *xlo = xhi + KC * kNPos;                          // [KC][kNPos]
    float4 *wbuf = xlo + KC * kNPos;                         // [kStages][kCF4]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbuf + kStages * kCF4);   // [kStages]: chunk's MMAs done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + kStages);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = a.h, w = a.w;
    const int64_t hw = (int64_t)h * w;

    // chunk c of the packed weights: tap = c / NCH, K part = c % NCH.  Packed per tap as
    // [hi|lo][kc][co]; a chunk takes kc in [part*2*KSC, (part+1)*2*KSC) of both hi and lo.
    auto issue_chunk = [&](int c) {
        const int tap = c / NCH, part = c - tap * NCH;
        const float4 *src = a.packed + (int64_t)tap * 2 * kWF4 + part * (2 * KSC * COUT);
        float4 *dst = wbuf + (c % kStages) * kCF4;
        constexpr int kHalf = 2 * KSC * COUT;               // float4 of hi (or lo) per chunk
        for (int i = tid; i < kCF4; i += kThreads) {
            // shared layout [kc][hi co 0..COUT-1 | lo co 0..COUT-1]: one B operand of N = 2*COUT rows
            // (a_hi x [b_hi | b_lo] in a single MMA) whose first COUT rows are the b_hi operand
            const int hl = i / kHalf, r = i - hl * kHalf;
            const int kcl = r / COUT, co = r - kcl * COUT;
            const uint32_t d = smem_u32(dst + (kcl * 2 + hl) * COUT + co);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + hl * kWF4 + r)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Halo tile of `tile` -> xhi via 4-byte cp.async (zero fill outside the image): no register
    // staging, every load of the tile in flight at once.  Warp w takes channels w, w+8, ...; lanes
    // run along a halo row (34 floats: lanes 0,1 also take the tail).  One commit group.
    auto issue_stage = [&](int tile) {
        const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y;
        const int64_t b = tile / (tiles_x * tiles_y);
        const int tx0 = txi * kTW, ty0 = tyi * kR;
        const uint32_t xhi_base = smem_u32(xhi);
        const int gx0 = tx0 - 1 + lane, gx1 = tx0 + 31 + lane;
        const bool okx0 = gx0 >= 0 && gx0 < w;
        const bool okx1 = lane < 2 && gx1 < w;
        if (a.in_c4) {
            // (kc, position) elements are 16 contiguous bytes on both sides: one cp.async each
            const float4 *src4 = reinterpret_cast<const float4 *>(a.in_a + b * a.a_bstride);
#pragma unroll 1
            for (int kc = warp; kc < KC; kc += kThreads / 32) {
                const float4 *plane4 = src4 + (int64_t)kc * hw;
                const uint32_t dst_k = xhi_base + (uint32_t)(kc * kNPos) * 16u;
#pragma unroll
                for (int py = 0; py < kR + 2; ++py) {
                    const int gy = ty0 - 1 + py;
                    const bool oky = gy >= 0 && gy < h;
                    const float4 *row = plane4 + (int64_t)(oky ? gy : 0) * w;
                    const uint32_t dst = dst_k + (uint32_t)(py * kHW + lane) * 16u;
                    const bool ok0 = oky && okx0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                                 "l"(ok0 ? row + gx0 : plane4), "r"(ok0 ? 16u : 0u)
                                 : "memory");
                    if (lane < 2) {
                        const bool ok1 = oky && okx1;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 32u * 16u),
                                     "l"(ok1 ? row + gx1 : plane4), "r"(ok1 ? 16u : 0u)
                                     : "memory");
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
#pragma unroll 1
        for (int c = warp; c < CIN; c += kThreads / 32) {
            const float *plane;
            if (c < a.Ca) {
                plane = a.in_a + b * a.a_bstride + (int64_t)c * hw;
            } else {
                const int cb = a.chan_map ? __ldg(a.chan_map + b * (CIN - a.Ca) + (c - a.Ca)) : c - a.Ca;
                plane = a.in_b + b * a.b_bstride + (int64_t)cb * hw;
            }
            // element (c, pos) lives at float index ((c/4)*kNPos + pos)*4 + c%4
            const uint32_t dst_c = xhi_base + (uint32_t)(((c >> 2) * kNPos) * 4 + (c & 3)) * 4u;
#pragma unroll
            for (int py = 0; py < kR + 2; ++py) {
                const int gy = ty0 - 1 + py;
                const bool oky = gy >= 0 && gy < h;
                const float *row = plane + (int64_t)(oky ? gy : 0) * w;
                const uint32_t dst = dst_c + (uint32_t)(py * kHW + lane) * 16u;
                const bool ok0 = oky && okx0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst),
                             "l"(ok0 ? row + gx0 : plane), "r"(ok0 ? 4u : 0u)
                             : "memory");
                if (lane < 2) {
                    const bool ok1 = oky && okx1;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 32u * 16u),
                                 "l"(ok1 ? row + gx1 : plane), "r"(ok1 ? 4u : 0u)
                                 : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- one-time setup: TMEM allocation (warp 0), mbarrier init (one thread) ---------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"((uint32_t)kCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < kStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int tile = blockIdx.x;
    if (tile < total_tiles) issue_stage(tile);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptors: D=f32, A=B=tf32, both K-major, M = 128; N = COUT (b_hi only) or
    // N = 2*COUT ([b_hi | b_lo]: the activation operand is read from shared memory once for both)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * COUT) >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);
    static_assert(2 * COUT <= 256 && (2 * COUT) % 16 == 0, "merged N must be a legal UMMA N");
    // descriptors differ only in the 14-bit start-address field (bytes >> 4): build the bases once
    // and add offsets per MMA (all of this CTA's shared memory is below 256 KB, no carry out)
    const uint64_t a_hi0 = make_desc(smem_u32(xhi), kNPos * 16u, 128u);
    const uint64_t a_lo0 = make_desc(smem_u32(xlo), kNPos * 16u, 128u);
    const uint64_t b_00 = make_desc(smem_u32(wbuf), 2 * COUT * 16u, 128u);
    uint32_t phase_bits = 0u;                               // bit i: parity to wait for on mbar[i]
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tprev = clock64();
#define WM_TICK(k) do { if (a.dbg) { const long long _t = clock64(); tacc[k] += _t - tprev; tprev = _t; } } while (0)

    // Persistent loop over this CTA's tiles.  Per tile: [finish staging X] -> [MMA pipeline over
    // weight chunks] -> [start staging the NEXT tile's X, asynchronously] -> [epilogue from TMEM].
#pragma unroll 1
    for (; tile < total_tiles; tile += gridDim.x) {
        const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y;
        const int64_t b = tile / (tiles_x * tiles_y);
        const int tx0 = txi * kTW, ty0 = tyi * kR;

        issue_chunk(0);                          // first three weight chunks of this tile
        issue_chunk(1);
        issue_chunk(2);
        if (a.dbg) tprev = clock64();
        asm volatile("cp.async.wait_group 3;" ::: "memory");   // X(tile) landed (older than all three)
        __syncthreads();
        WM_TICK(0);
        for (int i = tid; i < KC * kNPos; i += kThreads) {      // xlo = a - trunc_tf32(a)
            const float4 v = xhi[i];
            float4 lo;
            lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            xlo[i] = lo;
        }
        WM_TICK(1);

        // 4-deep weight ring: the MMAs of chunk c run while chunks c+1..c+3 are copied; slot
        // (c+3)%4 is free once the MMAs of chunk c-1 (committed to mbar[(c-1)%4]) completed.
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
            // groups issued so far: chunks 0 .. min(c+2, NCHUNK-1); chunk c must have landed
            if (c + 2 < NCHUNK) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else if (c + 1 < NCHUNK) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();                                // chunk c weights (and X) visible
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int tap = c / NCH, part = c - tap * NCH;
                const bool gate_tap = GATE && tap == 9;
                const int dy = gate_tap ? 1 : tap / 3, dx = gate_tap ? 1 : tap - (tap / 3) * 3;
                // offsets in 16-byte units (= float4 = one (kc, position) or (kc, co) element)
                const uint32_t shift = (uint32_t)(dy * kHW + dx);
                const uint64_t b_hi0 = b_00 + (uint32_t)(c % kStages) * kCF4;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t dcol = tmem_base + (uint32_t)((gate_tap ? 4 * COUT : 0) + mt * 2 * COUT);
                    const uint32_t arow = shift + (uint32_t)mt * 128u;
#pragma unroll
                    for (int kl = 0; kl < KSC; ++kl) {
                        const int ks = part * KSC + kl;
                        const uint32_t aoff = (uint32_t)(2 * ks) * kNPos + arow;
                        const uint32_t boff = (uint32_t)(2 * kl) * 2 * COUT;
                        const uint64_t a_hi = a_hi0 + aoff, a_lo = a_lo0 + aoff;
                        const uint64_t b_hl = b_hi0 + boff;     // rows [0,COUT) = hi, [COUT,2COUT) = lo
                        const uint32_t first = (ks == 0 && (tap == 0 || gate_tap)) ? 0u : 1u;
                        mma_tf32_ss(dcol, a_hi, b_hl, idesc2, first);   // cols [0,COUT) += a_hi b_hi, [COUT,2COUT) += a_hi b_lo
                        mma_tf32_ss(dcol, a_lo, b_hl, idesc, 1u);       // cols [0,COUT) += a_lo b_hi
                    }
                }
                asm volatile(
                    "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                        smem_u32(mbar + (c % kStages)))
                    : "memory");
            }
            if (c + 3 < NCHUNK) {
                if (c >= 1) {                               // slot (c+3)%4 was read by chunk c-1
                    const int i = (c - 1) % kStages;
                    mbar_wait(smem_u32(mbar + i), (phase_bits >> i) & 1u);
                    phase_bits ^= 1u << i;
                }
                issue_chunk(c + 3);
            }
        }
        WM_TICK(2);
        // drain: chunks NCHUNK-4 .. NCHUNK-1 (in-loop waits covered 0 .. NCHUNK-5)
#pragma unroll 1
        for (int c = NCHUNK - 4; c < NCHUNK; ++c) {
            const int i = c % kStages;
            mbar_wait(smem_u32(mbar + i), (phase_bits >> i) & 1u);
            phase_bits ^= 1u << i;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WM_TICK(3);

        // every MMA of this tile has completed: X is free -> start fetching the next tile's halo
        // now; the loads fly while the accumulators are drained below
        if (tile + (int)gridDim.x < total_tiles) issue_stage(tile + gridDim.x);
        WM_TICK(4);

        // ---- epilogue: TMEM -> registers -> NCHW -------------------------------------------
        {
            // warp w drains TMEM lane quarter w % 4; warps w and w+4 split the 32-channel groups
            const int quarter = warp & 3, whalf = warp >> 2;
            constexpr int NG = COUT / 32;                   // 32-channel groups per accumulator
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt) {
                const int m = mt * 128 + quarter * 32 + lane;
                const int q = kQ0 + m;
                const int py = q / kHW, px = q - py * kHW;
                const int gy = ty0 + py - 1, gx = tx0 + px - 1;
                const bool ok = px >= 1 && px <= kTW && py <= kR && gy < h && gx < w;
                const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
                for (int g = whalf; g < NG; g += 2) {
                    const int c0 = g * 32;
                    uint32_t acc[32];
                    {
                        uint32_t part[32];
                        tmem_ld32(lane_addr + (uint32_t)(mt * 2 * COUT + c0), acc);
                        tmem_ld32(lane_addr + (uint32_t)(mt * 2 * COUT + COUT + c0), part);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(part[j]));
                    }
                    if (GATE) {
                        uint32_t gt[32], part[32];
                        tmem_ld32(lane_addr + (uint32_t)(4 * COUT + mt * 2 * COUT + c0), gt);
                        tmem_ld32(lane_addr + (uint32_t)(4 * COUT + mt * 2 * COUT + COUT + c0), part);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float z = (__uint_as_float(gt[j]) + __uint_as_float(part[j])) +
                                            __ldg(a.gate_bias + c0 + j);
                            acc[j] = __float_as_uint(__fdividef(__uint_as_float(acc[j]), 1.0f + __expf(-z)));
                        }
                    }
                    if (ok && a.out_c4) {
                        float4 *o4 = reinterpret_cast<float4 *>(a.out) +
                                     (b * (COUT / 4) + c0 / 4) * hw + (int64_t)gy * w + gx;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                v[j] = __uint_as_float(acc[4 * q + j]);
                                if (!GATE && a.bias) v[j] += __ldg(a.bias + c0 + 4 * q + j);
                            }
                            o4[(int64_t)q * hw] = make_float4(v[0], v[1], v[2], v[3]);
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
        // the next tile's first MMA (issued after the chunk-0 barrier) overwrites the accumulators:
        // order this warp's TMEM reads before that barrier
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        WM_TICK(5);
    }
    if (a.dbg && tid == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) a.dbg[blockIdx.x * 6 + k] = tacc[k];
    }
#undef WM_TICK

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)kCols)
                     : "memory");
    }
}

// w3: (COUT, CIN, 3, 3); w1: (COUT, CIN) or null -> packed[tap][hi|lo][ci/4][co] float4
__global__ void __launch_bounds__(256)
prepack_tc5_kernel(const float *__restrict__ w3, const float *__restrict__ w1,
                   float4 *__restrict__ out, int CIN, int COUT, int ntaps)
{
    const int KC = CIN / 4;
    const int total = ntaps * KC * COUT;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int co = i % COUT;
        const int kc = (i / COUT) % KC;
        const int tap = i / (COUT * KC);
        float hi[4], lo[4];
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
        }
        float4 *dst = out + (int64_t)tap * 2 * KC * COUT;
        dst[kc * COUT + co] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        dst[KC * COUT + kc * COUT + co] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

template <int CIN, int COUT, bool GATE>
int launch(const Args &a, int64_t B, cudaStream_t s)
{
    constexpr size_t smem = sizeof(float4) * (2 * (CIN / 4) * kNPos + kStages * (2 * (CIN / 4) * COUT / (CIN >= 64 ? 2 : 1))) + 64;
    WM_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc5_kernel<CIN, COUT, GATE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = (a.w + kTW - 1) / kTW, tiles_y = (a.h + kR - 1) / kR;
    const int64_t total = (int64_t)tiles_x * tiles_y * B;
    WM_REQUIRE(total < (int64_t)1 << 31, "wm_conv3x3_fwd: too many tiles");
    const int grid = (int)(total < sm_count() ? total : sm_count());   // persistent: one CTA per SM
    conv3x3_tc5_kernel<CIN, COUT, GATE><<<grid, kThreads, smem, s>>>(a, tiles_x, tiles_y, (int)total);
    WM_LAUNCH_OK("conv3x3 tcgen05");
    return WM_OK;
}

void set_debug(long long *p) { g_dbg = p; }

size_t packed_bytes(int64_t Cin, int64_t Cout, int with_gate)
{
    return (size_t)(with_gate ? 10 : 9) * 2 * (Cin / 4) * Cout * sizeof(float4);
}

int prepack(const float *w3x3, const float *w1x1, void *packed, int64_t Cin, int64_t Cout,
            cudaStream_t s)
{
    const int ntaps = w1x1 ? 10 : 9;
    const int total = ntaps * (int)(Cin / 4) * (int)Cout;
    prepack_tc5_kernel<<<(total + 255) / 256, 256, 0, s>>>(w3x3, w1x1, static_cast<float4 *>(packed),
                                                            (int)Cin, (int)Cout, ntaps);
    WM_LAUNCH_OK("conv3x3 tcgen05 prepack");
    return WM_OK;
}

int forward(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b, int64_t b_bstride,
            const int *chan_map, const void *packed, const float *bias, const float *gate_bias,
            float *out, int64_t B, int64_t Cin, int64_t Cout, int64_t h, int64_t w, int in_c4, int out_c4,
            cudaStream_t s)
{
    WM_REQUIRE((h + kR - 1) / kR <= 65535, "wm_conv3x3_fwd: image too tall");
    WM_REQUIRE(!in_c4 || (Ca == Cin && aligned16(in_a) && a_bstride % 4 == 0),
               "wm_conv3x3_ex_fwd: the channel-quad input layout needs a single 16-byte aligned input");
    WM_REQUIRE(!out_c4 || aligned16(out), "wm_conv3x3_ex_fwd: channel-quad output must be 16-byte aligned");
    Args a;
    a.in_c4 = in_c4; a.out_c4 = out_c4;
    a.dbg = g_dbg;
    a.in_a = in_a; a.a_bstride = a_bstride; a.Ca = (int)Ca; a.in_b = in_b; a.b_bstride = b_bstride;
    a.chan_map = chan_map; a.packed = static_cast<const float4 *>(packed); a.bias = bias;
    a.gate_bias = gate_bias; a.out = out; a.h = (int)h; a.w = (int)w;
    if (gate_bias) {
        WM_REQUIRE(Cin == 64 && Cout == 64, "wm_conv3x3_fwd: gated mode supports 64->64 only");
        return launch<64, 64, true>(a, B, s);
    }
    if (Cin == 64 && Cout == 32) return launch<64, 32, false>(a, B, s);
    if (Cin == 64 && Cout == 64) return launch<64, 64, false>(a, B, s);
    if (Cin == 32 && Cout == 96) return launch<32, 96, false>(a, B, s);
    if (Cin == 32 && Cout == 32) return launch<32, 32, false>(a, B, s);
    WM_REQUIRE(false, "wm_conv3x3_fwd: Cin=%lld Cout=%lld unsupported", (long long)Cin, (long long)Cout);
    return WM_EINVAL;
}

}  // namespace tc5
}  // namespace wm
