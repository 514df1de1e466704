// Warp-level 1x1 GEMM on mma.sync fragments, fp32-accurate (3xTF32 with the two correction terms as one bf16
// MMA), shared by the memory-shaped pipelines whose per-pixel GEMMs would otherwise be bound by the return
// path of shared-memory weight loads (lfss_out_tma.cu, spatial32.cu, pw_tma.cu).
#pragma once
#include "tc5_common.cuh"

namespace wm {
namespace frag {

using wm::tc5::pack_bf16x2;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[nt] += A(16 pixels x 8K channels) . W, fp32-accurate: per k-step one TF32 MMA on the raw fp32 words (the
// tensor core reads their top 19 bits) and one bf16 MMA whose K slots 0-7 carry a_lo x w and slots 8-15
// a x w_lo (slot 2t, 2t+1 <-> channels t, t+4 of the k-step).  a: [channel][pixel] with row pitch PA, already
// offset to this warp's 16 pixels; wf: float2 per (k-step, n-tile, lane) = (W[k = t][n = g], W[k = t+4][n = g]).
template <int KSTEPS, int NT, int PA>
__device__ __forceinline__ void gemm_frag(float (&acc)[NT][4], const float *a, const float2 *wf, int lane)
{
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        const float *ap = a + (8 * ks + t) * PA + g;
        const float av[4] = {ap[0], ap[8], ap[4 * PA], ap[4 * PA + 8]};
        uint32_t ahi[4];
        float alo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ahi[i] = __float_as_uint(av[i]);
            alo[i] = av[i] - __uint_as_float(ahi[i] & 0xffffe000u);
        }
        const uint32_t a16[4] = {pack_bf16x2(alo[0], alo[2]), pack_bf16x2(alo[1], alo[3]),
                                 pack_bf16x2(av[0], av[2]), pack_bf16x2(av[1], av[3])};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float2 w = wf[(ks * NT + nt) * 32 + lane];
            const float l0 = w.x - __uint_as_float(__float_as_uint(w.x) & 0xffffe000u);
            const float l1 = w.y - __uint_as_float(__float_as_uint(w.y) & 0xffffe000u);
            mma_bf16(acc[nt], a16, pack_bf16x2(w.x, w.y), pack_bf16x2(l0, l1));
            mma_tf32(acc[nt], ahi, __float_as_uint(w.x), __float_as_uint(w.y));
        }
    }
}

// Weight fragments of m16n8k8 for W (N outputs x K inputs, row-major, element W[n][k]): entry
// (ks, nt, lane) = (W[8 nt + lane / 4][8 ks + lane % 4], W[.][. + 4]); filled cooperatively by `nthreads` threads.
template <int KSTEPS, int NT>
__device__ __forceinline__ void fill_wfrag(float2 *wf, const float *__restrict__ w, int K, int tid, int nthreads)
{
    for (int i = tid; i < KSTEPS * NT * 32; i += nthreads) {
        const int ln = i & 31, nt = (i >> 5) % NT, ks = i / (32 * NT);
        const int n = 8 * nt + (ln >> 2), k = 8 * ks + (ln & 3);
        wf[i] = make_float2(__ldg(w + n * K + k), __ldg(w + n * K + k + 4));
    }
}

}  // namespace frag
}  // namespace wm
