// Dense 3x3 convolution (stride 1, zero padding 1) as an implicit GEMM on the tensor cores,
// fp32-accurate through the 3xTF32 split, NCHW in and out -- no layout transposes, no im2col.
//
// Serves (reference wavemamba_arch.py): PAConv.k3 / k4 (:689-698, 1.49 of ~2.05 TFLOP per 4K
// image), DownFRG.l_conv (:966,975) and upFRG.h_out_conv (:993,1005).  cuDNN runs these either in
// fp32 SIMT (66 ms per 4K image) or in TF32 (fast, but measured 2.8e-3 dB outside the 1e-3 dB PSNR
// budget); the split a = a_hi + a_lo, b = b_hi + b_lo with a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
// accumulated in fp32 keeps fp32 accuracy on the tensor pipe.
//
// GEMM view per CTA: M = 8 x 32 output pixels, N = COUT, K = 9 taps x CIN.
//   * the CIN-channel input tile with a 1-pixel halo is staged once in shared memory as
//     xs[ci][position] (row stride == 8 mod 32 so mma A-fragment loads are conflict-free);
//     a tap is just a constant offset into it;
//   * weights are pre-packed once per layer into mma B-fragment order, already split into
//     tf32 hi/lo (wm_conv3x3_prepack), and streamed tap by tap with cp.async double buffering;
//   * 16 warps: warp = (tile row, 16-pixel half) owns COUT/8 m16n8k8 accumulators in registers.
// Fusions: the input may come from two tensors with a per-batch channel gather for the second
//   (torch.cat([x, matched perception]) of Matching_transformation :716 is never materialised);
//   PAConv stage A adds the 1x1 k2 as a 10th "tap" and applies  k3(x) * sigmoid(k2(x) + b)  in
//   the epilogue (:694-697); plain mode adds the conv bias.
#include "common.cuh"

namespace wm {
namespace conv {

constexpr int kTH = 8, kTW = 32;
constexpr int kHW = kTW + 2;            // halo row length 34
constexpr int kHalo = (kTH + 2) * kHW;  // 340
constexpr int kPS = 360;                // xs row stride (== 8 mod 32)
constexpr int kThreads = 512;   // 16 warps: warp = (tile row, 16-pixel half)

struct Args {
    const float *in_a;       // first Ca channels: (B, >=Ca, h, w), batch stride a_bstride
    int64_t a_bstride;
    int Ca;
    const float *in_b;       // remaining Cin-Ca channels, gathered through chan_map (or identity)
    int64_t b_bstride;
    const int *chan_map;     // (B, Cin-Ca) channel indices into in_b, or null
    const float4 *packed;    // [ntaps][CIN/8][COUT/8][32] {b0_hi, b1_hi, b0_lo, b1_lo}
    const float *bias;       // (COUT) or null
    const float *gate_bias;  // (COUT): PAConv stage A
    float *out;              // (B, COUT, h, w)
    int h, w;
};

__device__ __forceinline__ uint32_t to_tf32(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// Activation-side split for 3xTF32: hi = a with the 13 low mantissa bits cleared (what the
// tensor core would read anyway), lo = a - hi (exact; the tensor core truncates it to tf32).
// One LOP3 + one FADD per element -- cvt.rna.tf32 has no native SASS on sm_100 (it expands to
// FSETP+IADD3+SEL+LOP3).  |a - hi - tf32(lo)| <= 2^-20 |a|, same order as the dropped lo*lo term.
__device__ __forceinline__ void split_tf32(float a, uint32_t &hi, uint32_t &lo)
{
    hi = __float_as_uint(a) & 0xffffe000u;
    lo = __float_as_uint(a - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CIN, int COUT>
__device__ __forceinline__ void issue_weights(float4 *dst, const float4 *src)
{
    constexpr int kF4 = (CIN / 8) * (COUT / 8) * 32;
    for (int i = threadIdx.x; i < kF4; i += kThreads) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// One tap: acc[nt] += A(tap) * W(tap) over all CIN, 3xTF32.  One 16-pixel m-tile per warp.
template <int CIN, int COUT>
__device__ __forceinline__ void tap_mma(const float *abase, const float4 *wb, int lane,
                                        float (&acc)[COUT / 8][4])
{
    constexpr int KS = CIN / 8, NT = COUT / 8;
#pragma unroll 2
    for (int ks = 0; ks < KS; ++ks) {
        uint32_t ahi[4], alo[4];
        const float *ap = abase + ks * 8 * kPS;
        const float av[4] = {ap[0], ap[8], ap[4 * kPS], ap[4 * kPS + 8]};
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(av[i], ahi[i], alo[i]);
        // the three partial products of one accumulator are issued NT MMAs apart so consecutive
        // tensor-core instructions never depend on each other
        float4 bw[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bw[nt] = wb[(ks * NT + nt) * 32 + lane];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            mma_tf32(acc[nt], alo, __float_as_uint(bw[nt].x), __float_as_uint(bw[nt].y));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            mma_tf32(acc[nt], ahi, __float_as_uint(bw[nt].z), __float_as_uint(bw[nt].w));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            mma_tf32(acc[nt], ahi, __float_as_uint(bw[nt].x), __float_as_uint(bw[nt].y));
    }
}

// GATE: PAConv stage A (10 taps, sigmoid gate).  Otherwise 9 taps (+ optional bias).
template <int CIN, int COUT, bool GATE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_kernel(const Args a)
{
    constexpr int KS = CIN / 8, NT = COUT / 8;
    constexpr int kF4 = KS * NT * 32;          // float4 per tap
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                           // [CIN][kPS]
    float4 *wbuf = reinterpret_cast<float4 *>(smem + CIN * kPS);   // [2][kF4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int trow = warp >> 1, mhalf = warp & 1;      // my output row in the tile, my 16-px half
    const int gq = lane >> 2, t4 = lane & 3;
    const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int64_t b = blockIdx.z;
    const int h = a.h, w = a.w;
    const int64_t hw = (int64_t)h * w;

    issue_weights<CIN, COUT>(wbuf, a.packed);   // tap 0 in flight during the halo load

    // ---- stage the input tile (zero padding outside the image) ----------------------------
    // Row-wise: a warp owns (channel, halo-row) pairs; lanes run along the row (34 floats: the
    // first two lanes take the tail), four rows of loads in flight before the first store.
    {
        constexpr int kRows = CIN * (kTH + 2);
        const int gx0 = tx0 - 1 + lane;           // lanes 0..31 -> halo columns 0..31
        const int gx1 = tx0 + 31 + lane;          // lanes 0,1   -> halo columns 32,33
        const bool ok0 = gx0 >= 0 && gx0 < w;
        const bool ok1 = lane < 2 && gx1 < w;
#pragma unroll 1
        for (int r0 = warp * 4; r0 < kRows; r0 += (kThreads / 32) * 4) {
            float v0[4], v1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + i;              // kRows (320 | 640) is a multiple of 16*4
                const int c = r / (kTH + 2), py = r - c * (kTH + 2);
                const int gy = ty0 - 1 + py;
                const float *plane;
                if (c < a.Ca) {
                    plane = a.in_a + b * a.a_bstride + (int64_t)c * hw;
                } else {
                    const int cb = a.chan_map ? __ldg(a.chan_map + b * (CIN - a.Ca) + (c - a.Ca))
                                              : c - a.Ca;
                    plane = a.in_b + b * a.b_bstride + (int64_t)cb * hw;
                }
                const bool oky = gy >= 0 && gy < h;
                v0[i] = (oky && ok0) ? __ldg(plane + (int64_t)gy * w + gx0) : 0.0f;
                v1[i] = (oky && ok1) ? __ldg(plane + (int64_t)gy * w + gx1) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + i;
                const int c = r / (kTH + 2), py = r - c * (kTH + 2);
                float *dst = xs + c * kPS + py * kHW;
                dst[lane] = v0[i];
                if (lane < 2) dst[32 + lane] = v1[i];
            }
        }
    }

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;

    constexpr int NTAPS = GATE ? 10 : 9;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();   // weights[tap] (and, at tap 0, the halo) visible; buffer (tap+1)&1 free
        if (tap + 1 < NTAPS)
            issue_weights<CIN, COUT>(wbuf + ((tap + 1) & 1) * kF4, a.packed + (int64_t)(tap + 1) * kF4);
        const int dy = tap / 3, dx = tap - dy * 3;
        const float *abase = xs + t4 * kPS + (trow + dy) * kHW + mhalf * 16 + gq + dx;
        tap_mma<CIN, COUT>(abase, wbuf + (tap & 1) * kF4, lane, acc);
    }

    if (GATE) {
        // 10th tap: the 1x1 k2 on the centre position, into its own accumulators
        float gate[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) gate[nt][i] = 0.0f;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float *abase = xs + t4 * kPS + (trow + 1) * kHW + mhalf * 16 + gq + 1;
        tap_mma<CIN, COUT>(abase, wbuf + (9 & 1) * kF4, lane, gate);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = nt * 8 + 2 * t4 + (i & 1);
                const float z = gate[nt][i] + __ldg(a.gate_bias + co);
                acc[nt][i] *= 1.0f / (1.0f + expf(-z));
            }
    }

    // ---- epilogue: fragments -> NCHW -----------------------------------------------------
    const int gy = ty0 + trow;
    if (gy < h) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = nt * 8 + 2 * t4 + (i & 1);
                const int gx = tx0 + mhalf * 16 + gq + ((i & 2) ? 8 : 0);
                if (gx < w) {
                    float v = acc[nt][i];
                    if (!GATE && a.bias) v += __ldg(a.bias + co);
                    a.out[(b * COUT + co) * hw + (int64_t)gy * w + gx] = v;
                }
            }
    }
}

// w3: (COUT, CIN, 3, 3); w1: (COUT, CIN) or null -> packed[tap][ks][nt][lane] float4
__global__ void __launch_bounds__(256)
prepack_kernel(const float *__restrict__ w3, const float *__restrict__ w1, float4 *__restrict__ out,
               int CIN, int COUT, int ntaps)
{
    const int KS = CIN / 8, NT = COUT / 8;
    const int total = ntaps * KS * NT * 32;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int lane = i & 31;
        int r = i >> 5;
        const int nt = r % NT; r /= NT;
        const int ks = r % KS;
        const int tap = r / KS;
        const int gq = lane >> 2, t4 = lane & 3;
        const int co = nt * 8 + gq, ci0 = ks * 8 + t4, ci1 = ci0 + 4;
        float v0, v1;
        if (tap < 9) {
            v0 = w3[((int64_t)co * CIN + ci0) * 9 + tap];
            v1 = w3[((int64_t)co * CIN + ci1) * 9 + tap];
        } else {
            v0 = w1[(int64_t)co * CIN + ci0];
            v1 = w1[(int64_t)co * CIN + ci1];
        }
        const uint32_t h0 = to_tf32(v0), h1 = to_tf32(v1);
        const uint32_t l0 = to_tf32(v0 - __uint_as_float(h0)), l1 = to_tf32(v1 - __uint_as_float(h1));
        out[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0),
                             __uint_as_float(l1));
    }
}

template <int CIN, int COUT, bool GATE>
int launch(const Args &a, int64_t B, cudaStream_t s)
{
    constexpr size_t smem = sizeof(float) * (CIN * kPS) + 2 * sizeof(float4) * (CIN / 8) * (COUT / 8) * 32;
    WM_CUDA_OK(cudaFuncSetAttribute(conv3x3_kernel<CIN, COUT, GATE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((a.w + kTW - 1) / kTW, (a.h + kTH - 1) / kTH, (unsigned)B);
    conv3x3_kernel<CIN, COUT, GATE><<<grid, kThreads, smem, s>>>(a);
    WM_LAUNCH_OK("conv3x3");
    return WM_OK;
}

}  // namespace conv
}  // namespace wm

namespace wm {
namespace tc5 {   // conv3x3_tc5.cu
size_t packed_bytes(int64_t Cin, int64_t Cout, int with_gate);
void set_debug(long long *p);
int prepack(const float *w3x3, const float *w1x1, void *packed, int64_t Cin, int64_t Cout,
            cudaStream_t s);
int forward(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b, int64_t b_bstride,
            const int *chan_map, const void *packed, const float *bias, const float *gate_bias,
            float *out, int64_t B, int64_t Cin, int64_t Cout, int64_t h, int64_t w, int in_c4, int out_c4,
            cudaStream_t s);
}  // namespace tc5
namespace conv {
static int g_impl = 1;   // 0: mma.sync m16n8k8 (legacy tensor path), 1: tcgen05 + TMEM (default)
inline size_t mma_part_bytes(int64_t Cin, int64_t Cout, int with_gate)
{
    const size_t n = (size_t)(with_gate ? 10 : 9) * (Cin / 8) * (Cout / 8) * 32 * sizeof(float4);
    return (n + 255) / 256 * 256;
}
}  // namespace conv
}  // namespace wm

extern "C" int wm_conv3x3_set_impl(int impl)
{
    if (impl != 0 && impl != 1) {
        wm::set_error("wm_conv3x3_set_impl: impl must be 0 (mma.sync) or 1 (tcgen05)");
        return WM_EINVAL;
    }
    wm::conv::g_impl = impl;
    return WM_OK;
}

extern "C" int wm_conv3x3_get_impl(void) { return wm::conv::g_impl; }

/* Developer aid (not part of the reference-facing surface): when `device_buffer` is non-NULL the
 * tcgen05 kernel adds per-CTA phase cycle counts into it (6 int64 per CTA, 148 CTAs max). */
extern "C" int wm_conv3x3_debug_timing(void *device_buffer)
{
    wm::tc5::set_debug(static_cast<long long *>(device_buffer));
    return WM_OK;
}

extern "C" size_t wm_conv3x3_packed_bytes(int64_t Cin, int64_t Cout, int with_gate)
{
    if (Cin <= 0 || Cout <= 0 || Cin % 8 || Cout % 8) return 0;
    return wm::conv::mma_part_bytes(Cin, Cout, with_gate) + wm::tc5::packed_bytes(Cin, Cout, with_gate);
}

extern "C" int wm_conv3x3_prepack(const float *w3x3, const float *w1x1, void *packed, int64_t Cin,
                                  int64_t Cout, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(w3x3 && packed, "wm_conv3x3_prepack: null pointer");
    WM_REQUIRE(Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0 && Cin <= 512 && Cout <= 512,
               "wm_conv3x3_prepack: Cin=%lld Cout=%lld must be multiples of 8", (long long)Cin,
               (long long)Cout);
    WM_REQUIRE(aligned16(packed), "wm_conv3x3_prepack: packed must be 16-byte aligned");
    const int ntaps = w1x1 ? 10 : 9;
    const int total = ntaps * (int)(Cin / 8) * (int)(Cout / 8) * 32;
    conv::prepack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        w3x3, w1x1, static_cast<float4 *>(packed), (int)Cin, (int)Cout, ntaps);
    WM_LAUNCH_OK("conv3x3 prepack");
    // second half of the buffer: the same weights in tcgen05 (UMMA K-major) order
    return tc5::prepack(w3x3, w1x1,
                        static_cast<char *>(packed) + conv::mma_part_bytes(Cin, Cout, w1x1 != nullptr),
                        Cin, Cout, (cudaStream_t)stream);
}

extern "C" int wm_conv3x3_ex_fwd(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b,
                                 int64_t b_bstride, const int *chan_map, const void *packed,
                                 const float *bias, const float *gate_bias, float *out, int64_t B,
                                 int64_t Cin, int64_t Cout, int64_t h, int64_t w, int in_c4, int out_c4,
                                 wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::conv;
    WM_REQUIRE(B >= 0 && B <= 65535 && h >= 0 && w >= 0 && h < (1 << 24) && w < (1 << 24),
               "wm_conv3x3_fwd: bad sizes");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(in_a && packed && out, "wm_conv3x3_fwd: null pointer");
    WM_REQUIRE(Ca > 0 && Ca <= Cin && (Ca == Cin || in_b != nullptr),
               "wm_conv3x3_fwd: Ca=%lld of Cin=%lld needs a second input", (long long)Ca, (long long)Cin);
    WM_REQUIRE((h + kTH - 1) / kTH <= 65535, "wm_conv3x3_fwd: image too tall");
    WM_REQUIRE(aligned16(packed), "wm_conv3x3_fwd: packed weights must be 16-byte aligned");
    if (g_impl == 1) {
        const void *tc = static_cast<const char *>(packed) + mma_part_bytes(Cin, Cout, gate_bias != nullptr);
        return tc5::forward(in_a, a_bstride, Ca, in_b, b_bstride, chan_map, tc, bias, gate_bias, out,
                            B, Cin, Cout, h, w, in_c4, out_c4, (cudaStream_t)stream);
    }
    WM_REQUIRE(!in_c4 && !out_c4, "wm_conv3x3_ex_fwd: channel-quad layouts need the tcgen05 implementation");
    Args a;
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
