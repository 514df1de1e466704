// Haar DWT / IWT for sm_100a -- one streaming pass each, HBM-bound.
//
// Replaces dwt_init / iwt_init (reference wavemamba_arch.py:97-130).  Algorithmic bytes per
// call = 2 * planes * H * W * 4 (every input element read once, every output written once).
// Arithmetic is bit-exact with the reference: the four taps are halved first (exact in
// binary floating point) and summed in the reference's left-to-right order.
//
// Vector path: each thread owns a 2 x 8 input patch (two 128-bit loads per row) and emits
// one 128-bit store per band (DWT), or the mirror image (IWT).  Loads use the read-only,
// no-L1-allocate path; stores are evict-first.  Persistent grid-stride over
// SMs x 8 CTAs x 256 threads keeps >= 128 KB of loads in flight per SM.
#include <stdlib.h>

#include "common.cuh"

namespace wm {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void haar_analysis(float a, float b, float c, float d, float &ll,
                                              float &hl, float &lh, float &hh)
{
    // a=(even row, even col) b=(odd row, even col) c=(even row, odd col) d=(odd row, odd col),
    // all already halved.  reference :105-108
    ll = ((a + b) + c) + d;
    hl = (((-a) - b) + c) + d;
    lh = (((-a) + b) - c) + d;
    hh = ((a - b) - c) + d;
}

__global__ void __launch_bounds__(kThreads)
dwt_vec_kernel(const float *__restrict__ x, float *__restrict__ ll, float *__restrict__ hl,
               float *__restrict__ lh, float *__restrict__ hh, int64_t out_rows, int w4, int64_t W)
{
    // out_rows = planes * h ; input row index of output row r is 2r (planes are contiguous).
    const int64_t items = out_rows * w4;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t it = (int64_t)blockIdx.x * kThreads + threadIdx.x; it < items; it += stride) {
        const int64_t r = it / w4;
        const int q = (int)(it - r * w4);
        const float *top = x + (2 * r) * W + 8 * q;
        const float4 t0 = ld_stream4(top), t1 = ld_stream4(top + 4);
        const float4 b0 = ld_stream4(top + W), b1 = ld_stream4(top + W + 4);
        float4 oll, ohl, olh, ohh;
        haar_analysis(t0.x * 0.5f, b0.x * 0.5f, t0.y * 0.5f, b0.y * 0.5f, oll.x, ohl.x, olh.x, ohh.x);
        haar_analysis(t0.z * 0.5f, b0.z * 0.5f, t0.w * 0.5f, b0.w * 0.5f, oll.y, ohl.y, olh.y, ohh.y);
        haar_analysis(t1.x * 0.5f, b1.x * 0.5f, t1.y * 0.5f, b1.y * 0.5f, oll.z, ohl.z, olh.z, ohh.z);
        haar_analysis(t1.z * 0.5f, b1.z * 0.5f, t1.w * 0.5f, b1.w * 0.5f, oll.w, ohl.w, olh.w, ohh.w);
        const int64_t o = r * (4 * (int64_t)w4) + 4 * q;
        st_stream4(ll + o, oll);
        st_stream4(hl + o, ohl);
        st_stream4(lh + o, olh);
        st_stream4(hh + o, ohh);
    }
}

// 256-bit form: a thread owns a 2 x 16 input patch (two 32-byte loads per row) and emits one 32-byte
// store per band (W % 16 == 0, 32-byte aligned tensors); half the load/store instructions.
__device__ __forceinline__ void haar_analysis8(const float8 &t, const float8 &b, float4 &oll, float4 &ohl,
                                               float4 &olh, float4 &ohh)
{
    haar_analysis(t.lo.x * 0.5f, b.lo.x * 0.5f, t.lo.y * 0.5f, b.lo.y * 0.5f, oll.x, ohl.x, olh.x, ohh.x);
    haar_analysis(t.lo.z * 0.5f, b.lo.z * 0.5f, t.lo.w * 0.5f, b.lo.w * 0.5f, oll.y, ohl.y, olh.y, ohh.y);
    haar_analysis(t.hi.x * 0.5f, b.hi.x * 0.5f, t.hi.y * 0.5f, b.hi.y * 0.5f, oll.z, ohl.z, olh.z, ohh.z);
    haar_analysis(t.hi.z * 0.5f, b.hi.z * 0.5f, t.hi.w * 0.5f, b.hi.w * 0.5f, oll.w, ohl.w, olh.w, ohh.w);
}

__global__ void __launch_bounds__(kThreads)
dwt_vec8_kernel(const float *__restrict__ x, float *__restrict__ ll, float *__restrict__ hl,
                float *__restrict__ lh, float *__restrict__ hh, int64_t out_rows, int w8, int64_t W)
{
    const int64_t items = out_rows * w8;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t it = (int64_t)blockIdx.x * kThreads + threadIdx.x; it < items; it += stride) {
        const int64_t r = it / w8;
        const int q = (int)(it - r * w8);
        const float *top = x + (2 * r) * W + 16 * q;
        const float8 t0 = ld_stream8(top), t1 = ld_stream8(top + 8);
        const float8 b0 = ld_stream8(top + W), b1 = ld_stream8(top + W + 8);
        float4 oll[2], ohl[2], olh[2], ohh[2];
        haar_analysis8(t0, b0, oll[0], ohl[0], olh[0], ohh[0]);
        haar_analysis8(t1, b1, oll[1], ohl[1], olh[1], ohh[1]);
        const int64_t o = r * (8 * (int64_t)w8) + 8 * q;
        st_stream8(ll + o, oll[0], oll[1]);
        st_stream8(hl + o, ohl[0], ohl[1]);
        st_stream8(lh + o, olh[0], olh[1]);
        st_stream8(hh + o, ohh[0], ohh[1]);
    }
}

// DWT with SKFF's global average pool in its epilogue (reference :939-948: U = (HL + LH) + HH is pooled
// right after the transform that produced the bands): grid (nblk, planes) as skff_pool_kernel, per-CTA
// sums in fp32 per thread and fp64 across threads, written in the layout wm_skff_apply_fwd reads.
__global__ void __launch_bounds__(kThreads)
dwt_vec8_pool_kernel(const float *__restrict__ x, float *__restrict__ ll, float *__restrict__ hl,
                     float *__restrict__ lh, float *__restrict__ hh, double *__restrict__ partial, int h, int w8,
                     int64_t W)
{
    const int nblk = gridDim.x;
    const int64_t plane = blockIdx.y;
    const int64_t hw = (int64_t)h * w8 * 8;
    const float *xp = x + plane * 4 * hw;
    const int items = h * w8;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    for (int it = blockIdx.x * kThreads + threadIdx.x; it < items; it += nblk * kThreads) {
        const int r = it / w8, q = it - r * w8;
        const float *top = xp + (int64_t)(2 * r) * W + 16 * q;
        const float8 t0 = ld_stream8(top), t1 = ld_stream8(top + 8);
        const float8 b0 = ld_stream8(top + W), b1 = ld_stream8(top + W + 8);
        float4 oll[2], ohl[2], olh[2], ohh[2];
        haar_analysis8(t0, b0, oll[0], ohl[0], olh[0], ohh[0]);
        haar_analysis8(t1, b1, oll[1], ohl[1], olh[1], ohh[1]);
        const int64_t o = plane * hw + (int64_t)r * (8 * w8) + 8 * q;
        st_stream8(ll + o, oll[0], oll[1]);
        st_stream8(hl + o, ohl[0], ohl[1]);
        st_stream8(lh + o, olh[0], olh[1]);
        st_stream8(hh + o, ohh[0], ohh[1]);
        s0.x += (ohl[0].x + olh[0].x) + ohh[0].x; s0.y += (ohl[0].y + olh[0].y) + ohh[0].y;
        s0.z += (ohl[0].z + olh[0].z) + ohh[0].z; s0.w += (ohl[0].w + olh[0].w) + ohh[0].w;
        s1.x += (ohl[1].x + olh[1].x) + ohh[1].x; s1.y += (ohl[1].y + olh[1].y) + ohh[1].y;
        s1.z += (ohl[1].z + olh[1].z) + ohh[1].z; s1.w += (ohl[1].w + olh[1].w) + ohh[1].w;
    }
    double d = (double)(((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w)));
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

__global__ void __launch_bounds__(kThreads)
dwt_scalar_kernel(const float *__restrict__ x, float *__restrict__ ll, float *__restrict__ hl,
                  float *__restrict__ lh, float *__restrict__ hh, int64_t out_rows, int64_t w,
                  int64_t W)
{
    const int64_t items = out_rows * w;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t it = (int64_t)blockIdx.x * kThreads + threadIdx.x; it < items; it += stride) {
        const int64_t r = it / w, j = it - r * w;
        const float *top = x + (2 * r) * W + 2 * j;
        float a = top[0] * 0.5f, c = top[1] * 0.5f, b = top[W] * 0.5f, d = top[W + 1] * 0.5f;
        haar_analysis(a, b, c, d, ll[it], hl[it], lh[it], hh[it]);
    }
}

__device__ __forceinline__ void haar_synthesis(float p, float q, float r, float s, float &ee,
                                               float &oe, float &eo, float &oo)
{
    // p,q,r,s = LL,HL,LH,HH halved.  ee=(even row, even col) oe=(odd row, even col) ...
    // reference :125-128
    ee = ((p - q) - r) + s;
    oe = ((p - q) + r) - s;
    eo = ((p + q) - r) - s;
    oo = ((p + q) + r) + s;
}

// blockIdx.y = plane (b, c); the x dimension strides over the (row, 4-pixel group) items of the plane
// with 32-bit index arithmetic (the flat 64-bit div/mod chain of the first version cost more
// instructions than the data movement: 118 M vs the DWT's 70 M at the 4K level-1 size).
__global__ void __launch_bounds__(kThreads)
iwt_vec_kernel(const float *__restrict__ low, int64_t low_bstride, const float *__restrict__ high,
               int64_t high_bstride, float *__restrict__ y, int C, int h, int w4)
{
    const int w = 4 * w4;
    const int64_t hw = (int64_t)h * w;
    const int plane = blockIdx.y;
    const int b = plane / C, c = plane - b * C;
    const float *lp = low + (int64_t)b * low_bstride + (int64_t)c * hw;
    const float *hq = high + (int64_t)b * high_bstride + (int64_t)c * hw;
    const float *hr = hq + (int64_t)C * hw, *hs = hq + 2 * (int64_t)C * hw;
    float *yp = y + (int64_t)plane * 4 * hw;
    const int items = h * w4;
    const int stride = gridDim.x * kThreads;
    for (int it = blockIdx.x * kThreads + threadIdx.x; it < items; it += stride) {
        const int i = it / w4, q = it - i * w4;
        const int in_off = i * w + 4 * q;
        const float4 vp = ld_stream4(lp + in_off);
        const float4 vq = ld_stream4(hq + in_off);
        const float4 vr = ld_stream4(hr + in_off);
        const float4 vs = ld_stream4(hs + in_off);
        float4 e0, e1, o0, o1;  // even output row (8 floats) / odd output row
        haar_synthesis(vp.x * 0.5f, vq.x * 0.5f, vr.x * 0.5f, vs.x * 0.5f, e0.x, o0.x, e0.y, o0.y);
        haar_synthesis(vp.y * 0.5f, vq.y * 0.5f, vr.y * 0.5f, vs.y * 0.5f, e0.z, o0.z, e0.w, o0.w);
        haar_synthesis(vp.z * 0.5f, vq.z * 0.5f, vr.z * 0.5f, vs.z * 0.5f, e1.x, o1.x, e1.y, o1.y);
        haar_synthesis(vp.w * 0.5f, vq.w * 0.5f, vr.w * 0.5f, vs.w * 0.5f, e1.z, o1.z, e1.w, o1.w);
        float *out = yp + (int64_t)(2 * i) * (2 * w) + 8 * q;
        st_stream4(out, e0);
        st_stream4(out + 4, e1);
        st_stream4(out + 2 * w, o0);
        st_stream4(out + 2 * w + 4, o1);
    }
}

// The same with 256-bit accesses: a thread owns 8 coefficient columns = one 32-byte load per band and two
// 32-byte stores per output row (w % 8 == 0, 32-byte aligned tensors).  Half the load/store
// instructions: the 128-bit form spends 27 % of its warp time stalled on the LSU queue (ncu lg_throttle).
__global__ void __launch_bounds__(kThreads)
iwt_vec8_kernel(const float *__restrict__ low, int64_t low_bstride, const float *__restrict__ high,
                int64_t high_bstride, float *__restrict__ y, int C, int h, int w8)
{
    const int w = 8 * w8;
    const int64_t hw = (int64_t)h * w;
    const int plane = blockIdx.y;
    const int b = plane / C, c = plane - b * C;
    const float *lp = low + (int64_t)b * low_bstride + (int64_t)c * hw;
    const float *hq = high + (int64_t)b * high_bstride + (int64_t)c * hw;
    const float *hr = hq + (int64_t)C * hw, *hs = hq + 2 * (int64_t)C * hw;
    float *yp = y + (int64_t)plane * 4 * hw;
    const int items = h * w8;
    const int stride = gridDim.x * kThreads;
    for (int it = blockIdx.x * kThreads + threadIdx.x; it < items; it += stride) {
        const int i = it / w8, q = it - i * w8;
        const int in_off = i * w + 8 * q;
        const float8 vp = ld_stream8(lp + in_off), vq = ld_stream8(hq + in_off);
        const float8 vr = ld_stream8(hr + in_off), vs = ld_stream8(hs + in_off);
        float4 e[4], o[4];   // even / odd output row, 16 floats each
        haar_synthesis(vp.lo.x * 0.5f, vq.lo.x * 0.5f, vr.lo.x * 0.5f, vs.lo.x * 0.5f, e[0].x, o[0].x, e[0].y, o[0].y);
        haar_synthesis(vp.lo.y * 0.5f, vq.lo.y * 0.5f, vr.lo.y * 0.5f, vs.lo.y * 0.5f, e[0].z, o[0].z, e[0].w, o[0].w);
        haar_synthesis(vp.lo.z * 0.5f, vq.lo.z * 0.5f, vr.lo.z * 0.5f, vs.lo.z * 0.5f, e[1].x, o[1].x, e[1].y, o[1].y);
        haar_synthesis(vp.lo.w * 0.5f, vq.lo.w * 0.5f, vr.lo.w * 0.5f, vs.lo.w * 0.5f, e[1].z, o[1].z, e[1].w, o[1].w);
        haar_synthesis(vp.hi.x * 0.5f, vq.hi.x * 0.5f, vr.hi.x * 0.5f, vs.hi.x * 0.5f, e[2].x, o[2].x, e[2].y, o[2].y);
        haar_synthesis(vp.hi.y * 0.5f, vq.hi.y * 0.5f, vr.hi.y * 0.5f, vs.hi.y * 0.5f, e[2].z, o[2].z, e[2].w, o[2].w);
        haar_synthesis(vp.hi.z * 0.5f, vq.hi.z * 0.5f, vr.hi.z * 0.5f, vs.hi.z * 0.5f, e[3].x, o[3].x, e[3].y, o[3].y);
        haar_synthesis(vp.hi.w * 0.5f, vq.hi.w * 0.5f, vr.hi.w * 0.5f, vs.hi.w * 0.5f, e[3].z, o[3].z, e[3].w, o[3].w);
        float *out = yp + (int64_t)(2 * i) * (2 * w) + 16 * q;
        st_stream8(out, e[0], e[1]);
        st_stream8(out + 8, e[2], e[3]);
        st_stream8(out + 2 * w, o[0], o[1]);
        st_stream8(out + 2 * w + 8, o[2], o[3]);
    }
}

__global__ void __launch_bounds__(kThreads)
iwt_scalar_kernel(const float *__restrict__ low, int64_t low_bstride,
                  const float *__restrict__ high, int64_t high_bstride, float *__restrict__ y,
                  int64_t B, int C, int h, int w)
{
    const int64_t hw = (int64_t)h * w;
    const int64_t items = B * C * hw;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t it = (int64_t)blockIdx.x * kThreads + threadIdx.x; it < items; it += stride) {
        int64_t t = it;
        const int j = (int)(t % w); t /= w;
        const int i = (int)(t % h); t /= h;
        const int c = (int)(t % C);
        const int64_t b = t / C;
        const int64_t in_off = (int64_t)i * w + j;
        const float p = low[b * low_bstride + c * hw + in_off] * 0.5f;
        const float *hb = high + b * high_bstride + in_off;
        const float q = hb[(int64_t)c * hw] * 0.5f;
        const float r = hb[(int64_t)(C + c) * hw] * 0.5f;
        const float s = hb[(int64_t)(2 * C + c) * hw] * 0.5f;
        float ee, oe, eo, oo;
        haar_synthesis(p, q, r, s, ee, oe, eo, oo);
        float *out = y + ((b * C + c) * (2 * (int64_t)h) + 2 * i) * (2 * (int64_t)w) + 2 * j;
        out[0] = ee;
        out[1] = eo;
        out[2 * w] = oe;
        out[2 * w + 1] = oo;
    }
}

inline int stream_grid(int64_t items)
{
    const int64_t want = (items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
namespace skff {   // skff.cu
int pool_blocks(int64_t hw);
int launch_pool(const float *f0, const float *f1, const float *f2, double *partial, int64_t planes, int64_t hw,
                cudaStream_t s);
}
}  // namespace wm

extern "C" int wm_dwt_haar_pool_fwd(const float *x, float *ll, float *hl, float *lh, float *hh,
                                    void *pool_partials, size_t workspace_bytes, int64_t planes, int64_t H,
                                    int64_t W, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(planes >= 0 && H >= 0 && W >= 0 && planes <= 65535, "wm_dwt_haar_pool_fwd: bad sizes");
    WM_REQUIRE(H % 2 == 0 && W % 2 == 0, "wm_dwt_haar_pool_fwd: H=%lld W=%lld must be even", (long long)H,
               (long long)W);
    if (planes == 0 || H == 0 || W == 0) return WM_OK;
    WM_REQUIRE(x && ll && hl && lh && hh && pool_partials, "wm_dwt_haar_pool_fwd: null pointer");
    const int64_t h = H / 2, w = W / 2;
    const int nblk = skff::pool_blocks(h * w);
    WM_REQUIRE(workspace_bytes >= (size_t)planes * nblk * sizeof(double) &&
                   (reinterpret_cast<uintptr_t>(pool_partials) & 7u) == 0,
               "wm_dwt_haar_pool_fwd: pool workspace too small or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    double *partial = static_cast<double *>(pool_partials);
    static const bool no256 = getenv("WM_IWT_128") != nullptr;
    if (!no256 && W % 16 == 0 && h * (w / 8) < ((int64_t)1 << 30) && aligned32(x) && aligned32(ll) && aligned32(hl) &&
        aligned32(lh) && aligned32(hh)) {
        dwt_vec8_pool_kernel<<<dim3(nblk, (unsigned)planes), kThreads, 0, s>>>(x, ll, hl, lh, hh, partial, (int)h,
                                                                               (int)(w / 8), W);
        WM_LAUNCH_OK("dwt + pool kernel");
        return WM_OK;
    }
    const int rc = wm_dwt_haar_fwd(x, ll, hl, lh, hh, planes, H, W, stream);
    if (rc != WM_OK) return rc;
    return skff::launch_pool(hl, lh, hh, partial, planes, h * w, s);
}

extern "C" int wm_dwt_haar_fwd(const float *x, float *ll, float *hl, float *lh, float *hh,
                               int64_t planes, int64_t H, int64_t W, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(planes >= 0 && H >= 0 && W >= 0, "wm_dwt_haar_fwd: negative size");
    WM_REQUIRE(H % 2 == 0 && W % 2 == 0, "wm_dwt_haar_fwd: H=%lld W=%lld must be even",
               (long long)H, (long long)W);
    if (planes == 0 || H == 0 || W == 0) return WM_OK;  // empty: nothing to do, pointers unused
    WM_REQUIRE(x && ll && hl && lh && hh, "wm_dwt_haar_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t h = H / 2, w = W / 2, out_rows = planes * h;
    const bool vec = (W % 8 == 0) && aligned16(x) && aligned16(ll) && aligned16(hl) &&
                     aligned16(lh) && aligned16(hh);
    static const bool no256 = getenv("WM_IWT_128") != nullptr;     // developer A/B switch (both transforms)
    if (vec && !no256 && W % 16 == 0 && aligned32(x) && aligned32(ll) && aligned32(hl) && aligned32(lh) && aligned32(hh)) {
        const int w8 = (int)(w / 8);
        dwt_vec8_kernel<<<stream_grid(out_rows * w8), kThreads, 0, s>>>(x, ll, hl, lh, hh, out_rows, w8, W);
    } else if (vec) {
        const int w4 = (int)(w / 4);
        dwt_vec_kernel<<<stream_grid(out_rows * w4), kThreads, 0, s>>>(x, ll, hl, lh, hh, out_rows,
                                                                        w4, W);
    } else {
        dwt_scalar_kernel<<<stream_grid(out_rows * w), kThreads, 0, s>>>(x, ll, hl, lh, hh,
                                                                          out_rows, w, W);
    }
    WM_LAUNCH_OK("dwt kernel");
    return WM_OK;
}

extern "C" int wm_iwt_haar_fwd(const float *low, int64_t low_bstride, const float *high,
                               int64_t high_bstride, float *y, int64_t B, int64_t C, int64_t h,
                               int64_t w, wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(B >= 0 && C >= 0 && h >= 0 && w >= 0, "wm_iwt_haar_fwd: negative size");
    WM_REQUIRE(C < (1 << 20) && h < (1 << 30) && w < (1 << 30), "wm_iwt_haar_fwd: size too large");
    if (B == 0 || C == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(low && high && y, "wm_iwt_haar_fwd: null pointer");
    WM_REQUIRE(low_bstride >= C * h * w && high_bstride >= 3 * C * h * w,
               "wm_iwt_haar_fwd: batch strides smaller than one sample");
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = (w % 4 == 0) && aligned16(low) && aligned16(high) && aligned16(y) &&
                     (low_bstride % 4 == 0) && (high_bstride % 4 == 0);
    // WM_IWT_128=1: developer A/B switch for the 128-bit form
    static const bool no256 = getenv("WM_IWT_128") != nullptr;
    const bool vec8 = vec && !no256 && (w % 8 == 0) && aligned32(low) && aligned32(high) && aligned32(y) &&
                      (low_bstride % 8 == 0) && (high_bstride % 8 == 0) && ((C * h * w) % 8 == 0);
    if (vec8 && B * C <= 65535 && h * (w / 8) < ((int64_t)1 << 30)) {
        const int w8 = (int)(w / 8);
        const int64_t planes = B * C;
        const int64_t per_plane = (h * w8 + kThreads - 1) / kThreads;
        int64_t gx = ((int64_t)sm_count() * 8 + planes - 1) / planes;      // ~8 CTAs per SM in total
        gx = gx < per_plane ? gx : per_plane;
        dim3 grid((unsigned)(gx > 0 ? gx : 1), (unsigned)planes);
        iwt_vec8_kernel<<<grid, kThreads, 0, s>>>(low, low_bstride, high, high_bstride, y, (int)C, (int)h, w8);
    } else if (vec && B * C <= 65535 && h * (w / 4) < ((int64_t)1 << 30)) {
        const int w4 = (int)(w / 4);
        const int64_t planes = B * C;
        const int64_t per_plane = (h * w4 + kThreads - 1) / kThreads;
        int64_t gx = ((int64_t)sm_count() * 8 + planes - 1) / planes;      // ~8 CTAs per SM in total
        gx = gx < per_plane ? gx : per_plane;
        dim3 grid((unsigned)(gx > 0 ? gx : 1), (unsigned)planes);
        iwt_vec_kernel<<<grid, kThreads, 0, s>>>(low, low_bstride, high, high_bstride, y, (int)C, (int)h, w4);
    } else if (vec) {
        iwt_scalar_kernel<<<stream_grid(B * C * h * w), kThreads, 0, s>>>(
            low, low_bstride, high, high_bstride, y, B, (int)C, (int)h, (int)w);
    } else {
        iwt_scalar_kernel<<<stream_grid(B * C * h * w), kThreads, 0, s>>>(
            low, low_bstride, high, high_bstride, y, B, (int)C, (int)h, (int)w);
    }
    WM_LAUNCH_OK("iwt kernel");
    return WM_OK;
}
