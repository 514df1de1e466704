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
//   acc += a_lo*b_hi ; acc += a_hi*b_lo ; acc += a_hi*b_hi     (fp32 accumulate in TMEM)
//
// One CTA per SM (201 KB of shared memory), 256 threads: all stage, one thread issues the MMAs
// (tcgen05.commit -> mbarrier), warps 0-3 drain TMEM with tcgen05.ld.32x32b and run the epilogue.
#include "common.cuh"

namespace wm {
namespace tc5 {

constexpr int kR = 7, kTW = 32;             // tile: 7 rows x 32 columns
constexpr int kHW = kTW + 2;                // halo row length 34
constexpr int kHaloPos = (kR + 2) * kHW;    // 306 real halo positions
constexpr int kNPos = 328;                  // + zero tail read by the dropped M rows (max 325)
constexpr int kQ0 = kHW + 1;                // halo position of output (0,0) of the tile
constexpr int kThreads = 256;
constexpr int kStages = 3;                  // weight-chunk ring depth

struct Args {
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
    constexpr int need = 2 * COUT * (GATE ? 2 : 1);
    return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

template <int CIN, int COUT, bool GATE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc5_kernel(const Args a)
{
    constexpr int KC = CIN / 4;                 // 16-byte K chunks
    constexpr int KS = CIN / 8;                 // MMA k-steps (K = 8 for tf32)
    constexpr int NTAPS = GATE ? 10 : 9;
    constexpr int kWF4 = KC * COUT;             // float4 per weight part (hi or lo) per tap
    constexpr int NCH = CIN >= 64 ? 2 : 1;      // weight chunks per tap (K split so 2 buffers fit)
    constexpr int KSC = KS / NCH;               // k-steps per chunk
    constexpr int kCF4 = 2 * kWF4 / NCH;        // float4 per chunk: [hi | lo][2*KSC kc][COUT]
    constexpr int kCols = tmem_cols<COUT, GATE>();
    extern __shared__ __align__(128) float smem[];
    float4 *xhi = reinterpret_cast<float4 *>(smem);          // [KC][kNPos]
    float4 *xlo = xhi + KC * kNPos;                          // [KC][kNPos]
    float4 *wbuf = xlo + KC * kNPos;                         // [kStages][kCF4]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbuf + kStages * kCF4);   // [kStages]: chunk's MMAs done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + kStages);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kR;
    const int64_t b = blockIdx.z;
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
            const int hl = i / kHalf, r = i - hl * kHalf;
            const uint32_t d = smem_u32(dst + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + hl * kWF4 + r)
                         : "memory");
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
    issue_chunk(0);                              // the first two weight chunks fly during staging
    issue_chunk(1);

    // ---- stage the halo tile: X[kc][pos] = 4 consecutive channels of one position ------------
    // warp w stages kc = w, w+8 (, ...); lanes run along positions; 4 positions x 4 channels of
    // loads are issued before the first use.
    for (int kc = warp; kc < KC; kc += kThreads / 32) {
        const float *plane[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = kc * 4 + j;
            if (c < a.Ca) {
                plane[j] = a.in_a + b * a.a_bstride + (int64_t)c * hw;
            } else {
                const int cb = a.chan_map ? __ldg(a.chan_map + b * (CIN - a.Ca) + (c - a.Ca)) : c - a.Ca;
                plane[j] = a.in_b + b * a.b_bstride + (int64_t)cb * hw;
            }
        }
#pragma unroll 1
        for (int p0 = 0; p0 < kNPos; p0 += 128) {
            float v[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pos = p0 + i * 32 + lane;
                const int py = pos / kHW, px = pos - py * kHW;
                const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
                const bool ok = pos < kHaloPos && gy >= 0 && gy < h && gx >= 0 && gx < w;
                const int64_t off = (int64_t)gy * w + gx;
#pragma unroll
                for (int j = 0; j < 4; ++j) v[i][j] = ok ? __ldg(plane[j] + off) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pos = p0 + i * 32 + lane;
                if (pos < kNPos) {
                    float lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        lo[j] = v[i][j] - __uint_as_float(__float_as_uint(v[i][j]) & 0xffffe000u);
                    xhi[kc * kNPos + pos] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
                    xlo[kc * kNPos + pos] = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptor: D=f32, A=B=tf32, both K-major, N = COUT, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    const uint32_t xhi_s = smem_u32(xhi), xlo_s = smem_u32(xlo);
    const uint32_t wbuf_s = smem_u32(wbuf);

    // Pipeline over weight chunks, 3-deep ring: the MMAs of chunk c run while chunks c+1 and c+2
    // are copied; buffer (c+2)%3 is free once the MMAs of chunk c-1 (committed to
    // mbar[(c-1)%3]) have completed.
    constexpr int NCHUNK = NTAPS * NCH;
    static_assert(NCHUNK >= 3, "pipeline prologue assumes at least three chunks");
    uint32_t phase_bits = 0u;                               // bit i: parity to wait for on mbar[i]
#pragma unroll 1
    for (int c = 0; c < NCHUNK; ++c) {
        if (c + 1 < NCHUNK) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                    // chunk c weights (and X) visible
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int tap = c / NCH, part = c - tap * NCH;
            const bool gate_tap = GATE && tap == 9;
            const int dy = gate_tap ? 1 : tap / 3, dx = gate_tap ? 1 : tap - (tap / 3) * 3;
            const uint32_t shift = (uint32_t)(dy * kHW + dx) * 16u;      // bytes
            const uint32_t whi_s = wbuf_s + (uint32_t)(c % kStages) * kCF4 * 16u;
            const uint32_t wlo_s = whi_s + (uint32_t)(2 * KSC * COUT) * 16u;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t dcol = tmem_base + (uint32_t)((gate_tap ? 2 * COUT : 0) + mt * COUT);
                const uint32_t arow = shift + (uint32_t)mt * 128u * 16u;
#pragma unroll
                for (int kl = 0; kl < KSC; ++kl) {
                    const int ks = part * KSC + kl;
                    const uint32_t aoff = (uint32_t)(2 * ks) * kNPos * 16u + arow;
                    const uint32_t boff = (uint32_t)(2 * kl) * COUT * 16u;
                    const uint64_t a_hi = make_desc(xhi_s + aoff, kNPos * 16u, 128u);
                    const uint64_t a_lo = make_desc(xlo_s + aoff, kNPos * 16u, 128u);
                    const uint64_t b_hi = make_desc(whi_s + boff, COUT * 16u, 128u);
                    const uint64_t b_lo = make_desc(wlo_s + boff, COUT * 16u, 128u);
                    const uint32_t first = (ks == 0 && (tap == 0 || gate_tap)) ? 0u : 1u;
                    mma_tf32_ss(dcol, a_lo, b_hi, idesc, first);
                    mma_tf32_ss(dcol, a_hi, b_lo, idesc, 1u);
                    mma_tf32_ss(dcol, a_hi, b_hi, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(mbar + (c % kStages)))
                         : "memory");
        }
        if (c + 2 < NCHUNK) {
            if (c >= 1) {                                   // buffer (c+2)%3 was read by chunk c-1
                const int i = (c - 1) % kStages;
                mbar_wait(smem_u32(mbar + i), (phase_bits >> i) & 1u);
                phase_bits ^= 1u << i;
            }
            issue_chunk(c + 2);
        }
    }
    // drain: chunks whose completion has not been observed yet are NCHUNK-3 .. NCHUNK-1
    // (in-loop waits covered chunks 0 .. NCHUNK-4)
#pragma unroll 1
    for (int c = NCHUNK - 3; c < NCHUNK; ++c) {
        const int i = c % kStages;
        mbar_wait(smem_u32(mbar + i), (phase_bits >> i) & 1u);
        phase_bits ^= 1u << i;
    }

    // ---- epilogue: TMEM -> registers -> NCHW -----------------------------------------------
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            const int m = mt * 128 + warp * 32 + lane;
            const int q = kQ0 + m;
            const int py = q / kHW, px = q - py * kHW;
            const int gy = ty0 + py - 1, gx = tx0 + px - 1;
            const bool ok = px >= 1 && px <= kTW && py <= kR && gy < h && gx < w;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(lane_addr + (uint32_t)(mt * COUT + c0), acc);
                if (GATE) {
                    uint32_t gt[32];
                    tmem_ld32(lane_addr + (uint32_t)(2 * COUT + mt * COUT + c0), gt);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float z = __uint_as_float(gt[j]) + __ldg(a.gate_bias + c0 + j);
                        acc[j] = __float_as_uint(__uint_as_float(acc[j]) * (1.0f / (1.0f + expf(-z))));
                    }
                }
                if (ok) {
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
    dim3 grid((a.w + kTW - 1) / kTW, (a.h + kR - 1) / kR, (unsigned)B);
    conv3x3_tc5_kernel<CIN, COUT, GATE><<<grid, kThreads, smem, s>>>(a);
    WM_LAUNCH_OK("conv3x3 tcgen05");
    return WM_OK;
}

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
            float *out, int64_t B, int64_t Cin, int64_t Cout, int64_t h, int64_t w, cudaStream_t s)
{
    WM_REQUIRE((h + kR - 1) / kR <= 65535, "wm_conv3x3_fwd: image too tall");
    Args a;
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
