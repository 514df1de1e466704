// Fused pointwise (1x1) / depthwise (3x3) convolution groups of HFEBlock and LFSSBlock.ffn
// for sm_100a.  NCHW float32 in and out, one HBM round trip per fused group.
//
// Reference: wavemamba_arch.py:214-231 (ffn), 535-543 (LayerNorm2d), 687-697 (PAConv k2 gate),
// 729-742 (FeedForward project_in / project_out), 762-764,775,797 (CMTAttention convs).
//
// Spatial kernels work on 8x32-pixel tiles with a one-pixel halo held in shared memory
// [channel][position]; lanes always run along the image row, so global accesses are
// coalesced 128-byte rows and shared-memory accesses are conflict-free.
#include <stdlib.h>

#include "common.cuh"

namespace wm {
namespace pwdw {   // pw_dw_tc5.cu: returns 1 when the TMA preconditions do not hold
int forward(const float *x, const float *ln_w, const float *ln_b, float eps, const float *pw_w,
            const float *pw_b, const float *dw_w, const float *dw_b, int act, float *y, int64_t B,
            int64_t Cout, int64_t h, int64_t w, cudaStream_t s);
}
namespace pw {

constexpr int kTH = 8, kTW = 32;               // interior tile
constexpr int kHH = kTH + 2, kHW = kTW + 2;    // halo tile 10 x 34
constexpr int kHalo = kHH * kHW;               // 340 positions
constexpr int kXP = 360;                       // padded position stride (== 8 mod 32: conflict-free
                                               // mma A-fragment loads from xs[channel][position])
constexpr int kThreads = 256;

// ---- tensor-core helpers for the 1x1 phase (3xTF32, fp32-accurate; see conv3x3.cu) ----------
__device__ __forceinline__ uint32_t to_tf32(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
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

// LayerNorm2d over `C` channels for one position held in smem column `pos` (stride kXP).
template <int C>
__device__ __forceinline__ void ln_column(float *col, const float *lnw, const float *lnb, float eps)
{
    float mu = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) mu += col[c * kXP];
    mu *= (1.0f / C);
    float var = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float dlt = col[c * kXP] - mu;
        var = fmaf(dlt, dlt, var);
    }
    var *= (1.0f / C);
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int c = 0; c < C; ++c) col[c * kXP] = fmaf((col[c * kXP] - mu) * rstd, lnw[c], lnb[c]);
}

// Load the 32-channel halo tile of image b at (ty0-1, tx0-1) into xs[c][pos]; zeros outside.
// 4-byte cp.async with zero fill: no register staging, every load of the tile in flight at once
// (the scalar-load form was bound by global latency).  Warp w takes halo rows w, w+8, ... of the
// C*10 (channel, row) pairs; lanes run along the row (34 floats: lanes 0,1 also take the tail).
// The caller waits with halo_wait() before its first __syncthreads.
__device__ __forceinline__ void load_halo32(const float *__restrict__ x, float *xs, int64_t b,
                                            int C, int h, int w, int ty0, int tx0)
{
    const int64_t hw = (int64_t)h * w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t xs_base = (uint32_t)__cvta_generic_to_shared(xs);
    const float *xb = x + (int64_t)b * C * hw;
    const int gx0 = tx0 - 1 + lane, gx1 = tx0 + 31 + lane;
    const bool okx0 = gx0 >= 0 && gx0 < w;
    const bool okx1 = lane < 2 && gx1 < w;
    // row offsets and validity are the same for every channel: computed once (the per-element
    // index arithmetic of the first version was a quarter of pw_dw's instructions)
    int64_t roff[kHH];
    uint32_t okrows = 0u;
#pragma unroll
    for (int py = 0; py < kHH; ++py) {
        const int gy = ty0 - 1 + py;
        const bool oky = gy >= 0 && gy < h;
        roff[py] = oky ? (int64_t)gy * w : 0;
        okrows |= oky ? (1u << py) : 0u;
    }
#pragma unroll 1
    for (int c = warp; c < C; c += kThreads / 32) {
        const float *plane = xb + (int64_t)c * hw;
        const uint32_t dst_c = xs_base + (uint32_t)(c * kXP + lane) * 4u;
#pragma unroll
        for (int py = 0; py < kHH; ++py) {
            const bool oky = (okrows >> py) & 1u;
            const float *row = plane + roff[py];
            const uint32_t dst = dst_c + (uint32_t)(py * kHW) * 4u;
            const bool ok0 = oky && okx0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(ok0 ? row + gx0 : xb),
                         "r"(ok0 ? 4u : 0u)
                         : "memory");
            if (lane < 2) {
                const bool ok1 = oky && okx1;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 32u * 4u),
                             "l"(ok1 ? row + gx1 : xb), "r"(ok1 ? 4u : 0u)
                             : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void halo_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// y = dw3x3( pw1x1( ln?(x) ) ),  Cin = 32, Cout = 32*G
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(kThreads, 2)
pw_dw_kernel(const float *__restrict__ x, const float *__restrict__ ln_w,
             const float *__restrict__ ln_b, float eps, const float *__restrict__ pw_w,
             const float *__restrict__ pw_b, const float *__restrict__ dw_w,
             const float *__restrict__ dw_b, int act, float *__restrict__ y, int h, int w)
{
    // 1x1 on the tensor cores: per 32-output group, D[pos][co] = X[pos][ci] W[co][ci] as
    // 22 m-tiles (340 halo positions) x 4 n-tiles x 4 k-steps of mma.sync m16n8k8 TF32 with the
    // 3xTF32 split (fp32-accurate).  The depthwise 3x3 then runs on the FMA pipe from shared memory.
    constexpr int CIN = 32;
    constexpr int kMT = (kHalo + 15) / 16;      // 22 m-tiles
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                       // [32][kXP]
    float *ps = xs + CIN * kXP;             // [32][kXP]  one output group after the 1x1
    float4 *wf = reinterpret_cast<float4 *>(ps + 32 * kXP);   // [4 ks][4 nt][32 lanes] B fragments hi|lo
    float *pb = reinterpret_cast<float *>(wf + 4 * 4 * 32);   // [COUT]
    float *dww = pb + COUT;                 // [COUT][9]
    float *dwb = dww + COUT * 9;            // [COUT]
    float *lnw = dwb + COUT;                // [32]
    float *lnb = lnw + CIN;                 // [32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int64_t b = blockIdx.z;
    const int64_t hw = (int64_t)h * w;

    for (int i = tid; i < COUT; i += kThreads) {
        pb[i] = pw_b ? __ldg(pw_b + i) : 0.0f;
        dwb[i] = __ldg(dw_b + i);
    }
    for (int i = tid; i < COUT * 9; i += kThreads) dww[i] = __ldg(dw_w + i);
    if (ln_w != nullptr && tid < CIN) { lnw[tid] = __ldg(ln_w + tid); lnb[tid] = __ldg(ln_b + tid); }
    load_halo32(x, xs, b, CIN, h, w, ty0, tx0);
    for (int i = tid; i < CIN * (kXP - kHalo); i += kThreads) {   // rows read by the last m-tile
        const int c = i / (kXP - kHalo), r = i - c * (kXP - kHalo);
        xs[c * kXP + kHalo + r] = 0.0f;
    }
    halo_wait();
    __syncthreads();

    if (ln_w != nullptr) {
        for (int pos = tid; pos < kHalo; pos += kThreads) ln_column<CIN>(xs + pos, lnw, lnb, eps);
        __syncthreads();
    }

    const int gq = lane >> 2, t4 = lane & 3;
    for (int g = 0; g < COUT / 32; ++g) {
        // ---- B fragments of this group's 32 x 32 weights, split into tf32 hi / lo ------------
        for (int i = tid; i < 4 * 4 * 32; i += kThreads) {
            const int ln_ = i & 31, nt = (i >> 5) & 3, ks = i >> 7;
            const int co = g * 32 + nt * 8 + (ln_ >> 2), ci = ks * 8 + (ln_ & 3);
            const float w0 = __ldg(pw_w + co * CIN + ci), w1 = __ldg(pw_w + co * CIN + ci + 4);
            const uint32_t h0 = to_tf32(w0), h1 = to_tf32(w1);
            wf[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1),
                                __uint_as_float(to_tf32(w0 - __uint_as_float(h0))),
                                __uint_as_float(to_tf32(w1 - __uint_as_float(h1))));
        }
        __syncthreads();   // also: every warp is done with the previous group's ps

        // ---- 1x1: warp w takes m-tiles w, w+8, w+16 ----------------------------------------
#pragma unroll 1
        for (int mt = warp; mt < kMT; mt += kThreads / 32) {
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
            const float *abase = xs + t4 * kXP + mt * 16 + gq;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const float *ap = abase + ks * 8 * kXP;
                const float av[4] = {ap[0], ap[8], ap[4 * kXP], ap[4 * kXP + 8]};
                uint32_t ahi[4], alo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(av[i], ahi[i], alo[i]);
                float4 bw[4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) bw[nt] = wf[(ks * 4 + nt) * 32 + lane];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_tf32(acc[nt], alo, __float_as_uint(bw[nt].x), __float_as_uint(bw[nt].y));
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_tf32(acc[nt], ahi, __float_as_uint(bw[nt].z), __float_as_uint(bw[nt].w));
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_tf32(acc[nt], ahi, __float_as_uint(bw[nt].x), __float_as_uint(bw[nt].y));
            }
            // fragments -> ps[co][pos]; the dw conv zero-pads the 1x1 OUTPUT (bias included)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int pos = mt * 16 + gq + half * 8;
                if (pos < kHalo) {
                    const int py = pos / kHW, px = pos - py * kHW;
                    const int gy = ty0 - 1 + py, gx = tx0 - 1 + px;
                    const bool valid = gy >= 0 && gy < h && gx >= 0 && gx < w;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int cl = nt * 8 + 2 * t4 + j;
                            ps[cl * kXP + pos] = valid ? acc[nt][half * 2 + j] + pb[g * 32 + cl] : 0.0f;
                        }
                }
            }
        }
        __syncthreads();

        // ---- depthwise 3x3 on the group: thread = (channel, tile column), 8 rows ----------
        const int col = tid & 31;
        const int gx = tx0 + col;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int cl = (tid >> 5) + 8 * i;       // channel inside the group
            const int co = g * 32 + cl;
            float k[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) k[t] = dww[co * 9 + t];
            const float bias = dwb[co];
            const float *pc = ps + cl * kXP + col;   // halo (py, px) = (row, col + dx)
            float r0[3], r1[3], r2[3];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) { r0[dx] = pc[dx]; r1[dx] = pc[kHW + dx]; }
            float *yo = y + ((int64_t)b * COUT + co) * hw + (int64_t)ty0 * w + gx;
#pragma unroll
            for (int row = 0; row < kTH; ++row) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) r2[dx] = pc[(row + 2) * kHW + dx];
                float o = bias;
                o = fmaf(k[0], r0[0], o); o = fmaf(k[1], r0[1], o); o = fmaf(k[2], r0[2], o);
                o = fmaf(k[3], r1[0], o); o = fmaf(k[4], r1[1], o); o = fmaf(k[5], r1[2], o);
                o = fmaf(k[6], r2[0], o); o = fmaf(k[7], r2[1], o); o = fmaf(k[8], r2[2], o);
                if (act == 1) o = __fdividef(o, 1.0f + __expf(-o));  // SiLU (SS2D.act, reference :487)
                if (gx < w && ty0 + row < h) yo[(int64_t)row * w] = o;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) { r0[dx] = r1[dx]; r1[dx] = r2[dx]; }
            }
        }
        // (the barrier at the top of the next group protects ps and wf)
    }
}

inline bool dims_ok(int64_t B, int64_t h, int64_t w)
{
    return B >= 0 && B <= 65535 && h >= 0 && w >= 0 && h < (1 << 24) && w < (1 << 24);
}

template <typename KernelT>
inline cudaError_t opt_in_smem(KernelT kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace pw
}  // namespace wm

using namespace wm;
using namespace wm::pw;

template <int COUT>
static int launch_pw_dw(const float *x, const float *ln_w, const float *ln_b, float eps,
                        const float *pw_w, const float *pw_b, const float *dw_w,
                        const float *dw_b, int act, float *y, int64_t B, int64_t h, int64_t w,
                        cudaStream_t s)
{
    const size_t smem = sizeof(float) * (32 * kXP * 2 + 4 * 4 * 32 * 4 + COUT + COUT * 9 + COUT + 64);
    WM_CUDA_OK(opt_in_smem(pw_dw_kernel<COUT>, smem));
    dim3 grid((unsigned)((w + kTW - 1) / kTW), (unsigned)((h + kTH - 1) / kTH), (unsigned)B);
    pw_dw_kernel<COUT><<<grid, kThreads, smem, s>>>(x, ln_w, ln_b, eps, pw_w, pw_b, dw_w, dw_b, act,
                                                    y, (int)h, (int)w);
    WM_LAUNCH_OK("pw_dw");
    return WM_OK;
}

extern "C" int wm_pw_dw_fwd(const float *x, const float *ln_w, const float *ln_b, float eps,
                            const float *pw_w, const float *pw_b, const float *dw_w,
                            const float *dw_b, int act, float *y, int64_t B, int64_t Cin,
                            int64_t Cout, int64_t h, int64_t w, wm_stream_t stream)
{
    WM_REQUIRE(x && pw_w && dw_w && dw_b && y, "wm_pw_dw_fwd: null pointer");
    WM_REQUIRE(act == 0 || act == 1, "wm_pw_dw_fwd: act must be 0 (none) or 1 (SiLU)");
    WM_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), "wm_pw_dw_fwd: ln_w/ln_b must come together");
    WM_REQUIRE(dims_ok(B, h, w), "wm_pw_dw_fwd: bad sizes");
    WM_REQUIRE(Cin == 32 && (Cout == 32 || Cout == 64 || Cout == 96),
               "wm_pw_dw_fwd: Cin=%lld Cout=%lld unsupported (32 -> 32|64|96)", (long long)Cin,
               (long long)Cout);
    WM_REQUIRE((h + kTH - 1) / kTH <= 65535, "wm_pw_dw_fwd: image too tall");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // TMA + tcgen05 pipeline (pw_dw_tc5.cu) whenever its preconditions hold; WM_PW_DW_LEGACY=1 is a
    // developer switch for A/B timing of the cp.async + mma.sync kernel below
    static const bool legacy = getenv("WM_PW_DW_LEGACY") != nullptr;
    if (!legacy) {
        const int rc = wm::pwdw::forward(x, ln_w, ln_b, eps, pw_w, pw_b, dw_w, dw_b, act, y, B, Cout, h, w, s);
        if (rc != 1) return rc;
    }
    if (Cout == 32) return launch_pw_dw<32>(x, ln_w, ln_b, eps, pw_w, pw_b, dw_w, dw_b, act, y, B, h, w, s);
    if (Cout == 64) return launch_pw_dw<64>(x, ln_w, ln_b, eps, pw_w, pw_b, dw_w, dw_b, act, y, B, h, w, s);
    return launch_pw_dw<96>(x, ln_w, ln_b, eps, pw_w, pw_b, dw_w, dw_b, act, y, B, h, w, s);
}
