// 32x32 Gram matrix + squared row norms of two (32, hw) channel stacks, one pass over HBM.
//
// Serves the two global reductions of HFEBlock (reference wavemamba_arch.py):
//   * Matching: torch.cdist(x, perception) (:664) is sqrt(|x_i|^2 + |p_j|^2 - 2 x_i.p_j) in its
//     default mm mode; the argmin over j (:624) only needs G = X P^T and the two norm vectors.
//   * CMTAttention: normalize(q) @ normalize(k)^T (:787-790) = (q k^T) / (|q| |k|^T).
// cuBLAS handles this shape (M=N=32, K = hw up to 2 M) with split-K SGEMMs at ~10 % of the HBM
// roofline; here every CTA streams a pixel range through shared memory, multiplies it on the tensor
// cores (mma.sync TF32 with the 3xTF32 split on both operands => fp32-accurate products), folds the
// fp32 accumulators into fp64 every 128 pixels per warp; per-CTA partials are reduced in a fixed
// order by a second tiny kernel => deterministic, and closer to the fp64 truth than SGEMM.
// Algorithmic bytes: 2 * 32 * hw * 4 per batch item.
#include "common.cuh"

namespace wm {
namespace gram {

constexpr int kC = 32;
constexpr int kTile = 128;           // pixels per smem tile
constexpr int kRS = 132;             // smem row stride (33 quads: conflict-free LDS.128 rows)
constexpr int kThreads = 256;
constexpr int kOut = kC * kC + 2 * kC;   // 1088 values per batch item: G | |X|^2 | |Y|^2

__device__ __forceinline__ float sq4(const float4 v, float acc)
{
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    return acc;
}

// 3xTF32 helpers (fp32-accurate products on the tensor cores; see conv3x3.cu)
__device__ __forceinline__ void split_tf32(float a, uint32_t &hi, uint32_t &lo)
{
    hi = __float_as_uint(a) & 0xffffe000u;
    lo = __float_as_uint(a - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kFlushTiles = 8;       // fp32 accumulators are folded into fp64 every 8 tiles
// The CTAs of a thread-block cluster add their 1088 partial outputs through distributed shared memory (fixed
// rank order) and only rank 0 writes them: the one-CTA reduce / tail kernel that follows is bound by how fast
// a single SM can pull the partials (148 x 8.7 KB took ~24 us per call, 24 calls per image), so every halving
// of the partial count helps.  Pairs pack onto any SM count; clusters of 4 do not (37 of them need a second wave
// on 148 SMs: match_index 1.03 -> 1.60 ms per image).
constexpr int kCluster = 2;

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double *p, uint32_t rank)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    uint32_t ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
constexpr size_t kSmemBytes = sizeof(double) * 8 * kC * kC;   // 64 KB: cross-warp reduction buffer
static_assert(kSmemBytes >= sizeof(float) * 2 * kC * kRS, "staging tiles must fit the reduction buffer");

// G = X Y^T on the tensor cores.  A CTA streams its pixel range in 128-pixel tiles: registers ->
// shared [row][pixel] (the next tile's loads are already in flight while this one is consumed);
// warp w multiplies the two 8-pixel k-steps {w, w+8} of the tile with mma.sync m16n8k8 TF32 and the
// 3xTF32 split on both operands (2 x 4 fragment tiles = the whole 32x32 output, fp32 accumulate),
// folds its accumulators into fp64 every 8 tiles (= 128 pixels per warp, as the tile sums of the
// fp32 version), and the eight warps are summed in a fixed order at the end.
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
gram_partial_kernel(const float *__restrict__ x, int64_t x_bstride, const float *__restrict__ y,
                    int64_t y_bstride, double *__restrict__ partial, int64_t hw, int chunk,
                    int nchunks, int vec)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);          // [32][kRS]
    float *ysm = xs + kC * kRS;                               // [32][kRS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int64_t b = blockIdx.y;
    const float *xb = x + b * x_bstride;
    const float *yb = y + b * y_bstride;
    const int64_t p_begin = (int64_t)blockIdx.x * chunk;
    int64_t p_end = p_begin + chunk;
    if (p_end > hw) p_end = hw;

    float acc[2][4][4];
    double g64[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[mt][nt][i] = 0.0f; g64[mt][nt][i] = 0.0; }
    // staging role: warp w loads rows w, w+8, w+16, w+24 (lane = quad); it also owns their norms
    double nx[4] = {0, 0, 0, 0}, ny[4] = {0, 0, 0, 0};
    float4 ra[4], rc[4];

    auto fetch = [&](int64_t p0) {
        const bool full = vec && p0 + kTile <= p_end;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = warp + 8 * r;
            if (full) {
                ra[r] = ld_stream4(xb + (int64_t)row * hw + p0 + 4 * lane);
                rc[r] = ld_stream4(yb + (int64_t)row * hw + p0 + 4 * lane);
            } else {
                float av[4], cv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t p = p0 + 4 * lane + e;
                    const bool ok = p < p_end;
                    av[e] = ok ? __ldg(xb + (int64_t)row * hw + p) : 0.0f;
                    cv[e] = ok ? __ldg(yb + (int64_t)row * hw + p) : 0.0f;
                }
                ra[r] = make_float4(av[0], av[1], av[2], av[3]);
                rc[r] = make_float4(cv[0], cv[1], cv[2], cv[3]);
            }
        }
    };

    if (p_begin < p_end) fetch(p_begin);
    int since_flush = 0;
    for (int64_t p0 = p_begin; p0 < p_end; p0 += kTile) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = warp + 8 * r;
            *reinterpret_cast<float4 *>(xs + row * kRS + 4 * lane) = ra[r];
            *reinterpret_cast<float4 *>(ysm + row * kRS + 4 * lane) = rc[r];
            nx[r] += (double)sq4(ra[r], 0.0f);
            ny[r] += (double)sq4(rc[r], 0.0f);
        }
        __syncthreads();
        if (p0 + kTile < p_end) fetch(p0 + kTile);          // in flight while this tile is multiplied
#pragma unroll
        for (int ksi = 0; ksi < 2; ++ksi) {
            const int pk = (warp + 8 * ksi) * 8;               // first pixel of the k-step
            uint32_t ahi[2][4], alo[2][4], bhi[4][2], blo[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float *ap = xs + (g + 16 * mt) * kRS + pk + t4;
                split_tf32(ap[0], ahi[mt][0], alo[mt][0]);
                split_tf32(ap[8 * kRS], ahi[mt][1], alo[mt][1]);
                split_tf32(ap[4], ahi[mt][2], alo[mt][2]);
                split_tf32(ap[8 * kRS + 4], ahi[mt][3], alo[mt][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float *bp = ysm + (g + 8 * nt) * kRS + pk + t4;
                split_tf32(bp[0], bhi[nt][0], blo[nt][0]);
                split_tf32(bp[4], bhi[nt][1], blo[nt][1]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mma_tf32(acc[mt][nt], alo[mt], bhi[nt]);
                    mma_tf32(acc[mt][nt], ahi[mt], blo[nt]);
                    mma_tf32(acc[mt][nt], ahi[mt], bhi[nt]);
                }
        }
        if (++since_flush == kFlushTiles || p0 + kTile >= p_end) {
            since_flush = 0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        g64[mt][nt][i] += (double)acc[mt][nt][i];
                        acc[mt][nt][i] = 0.0f;
                    }
        }
        __syncthreads();
    }

    // ---- fixed-order sum over the eight warps ------------------------------------------------
    double *red = reinterpret_cast<double *>(smem_raw);       // [8 warps][32*32]; staging tiles are dead
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = g + 16 * mt + ((i & 2) ? 8 : 0), col = 8 * nt + 2 * t4 + (i & 1);
                red[warp * kC * kC + row * kC + col] = g64[mt][nt][i];
            }
    __syncthreads();
    double tsum[kC * kC / kThreads];
#pragma unroll
    for (int j = 0; j < kC * kC / kThreads; ++j) {
        const int o = tid + j * kThreads;
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += red[w * kC * kC + o];
        tsum[j] = t;
    }
    __syncthreads();                                           // red is dead: its head becomes this CTA's 1088 outputs
    double *mine = red;
#pragma unroll
    for (int j = 0; j < kC * kC / kThreads; ++j) mine[tid + j * kThreads] = tsum[j];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double sx = nx[r], sy = ny[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
            sy += __shfl_xor_sync(0xffffffffu, sy, off);
        }
        if (lane == 0) {
            mine[kC * kC + warp + 8 * r] = sx;
            mine[kC * kC + kC + warp + 8 * r] = sy;
        }
    }
    // ---- fixed-order sum over the CTAs of the cluster, written by rank 0 ------------------------
    cluster_sync();
    if (cluster_rank() == 0) {
        double *out = partial + ((int64_t)b * (nchunks / kCluster) + blockIdx.x / kCluster) * kOut;
        for (int o = tid; o < kOut; o += kThreads) {
            double t = mine[o];
#pragma unroll
            for (int r = 1; r < kCluster; ++r) t += ld_dsmem_f64(mine + o, (uint32_t)r);
            out[o] = t;
        }
    }
    cluster_sync();                                            // the peers' shared memory outlives rank 0's reads
}

__global__ void __launch_bounds__(kThreads)
gram_reduce_kernel(const double *__restrict__ partial, float *__restrict__ out, int nchunks)
{
    const int64_t b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= kOut) return;
    const double *p = partial + b * nchunks * kOut + i;
    double acc = 0.0;
    for (int c = 0; c < nchunks; ++c) acc += p[(int64_t)c * kOut];
    out[b * kOut + i] = (float)acc;
}

// The reduce step with the consumer's 32x32 tail folded in (one CTA per image): the torch ops that
// followed wm_gram32_fwd in an HFEBlock -- ~10 launches of a few microseconds each on a 32x32 matrix per
// Matching / attention call, 128 of the ~400 launches of a forward -- become part of this kernel.
//   MODE 1  Matching (reference :664, :624): idx[i] = argmin_j (|x_i|^2 + |p_j|^2) - 2 x_i.p_j, evaluated
//           in fp32 exactly as the torch expression `nx[:, :, None] + ny[:, None, :] - 2.0 * gram`
//           (sqrt / clamp of cdist are monotone), first index on ties
//   MODE 2  CMTAttention (:787-797): attn = softmax_j( G_ij / (max(|q_i|, 1e-12) max(|k_j|, 1e-12)) * T ),
//           mixed = W_po . attn  (project_out folded with the attention: the per-image 1x1 weights)
constexpr int kTailThreads = 1024;
template <int MODE>
__global__ void __launch_bounds__(kTailThreads)
gram_reduce_tail_kernel(const double *__restrict__ partial, float *__restrict__ out, int nchunks,
                        int *__restrict__ idx_out, const float *__restrict__ temperature,
                        const float *__restrict__ w_po, float *__restrict__ mixed_out)
{
    __shared__ float G[kC * kC + 2 * kC];
    __shared__ float attn[kC * kC];
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    // one CTA per image sums the 74 x 1088 partials (one row per CTA pair of the partial kernel): thread = output, consecutive lanes = consecutive
    // outputs (coalesced 256-byte loads; a warp per output with lanes over the chunks touches one 32-byte
    // sector per lane and took ~90 us), chunks in order => deterministic and equal to gram_reduce_kernel
    for (int i = tid; i < kOut; i += kTailThreads) {
        const double *p = partial + b * nchunks * kOut + i;
        double acc = 0.0;
        int c = 0;
        for (; c + 8 <= nchunks; c += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(c + u) * kOut];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; c < nchunks; ++c) acc += p[(int64_t)c * kOut];
        G[i] = (float)acc;
        if (out != nullptr) out[b * kOut + i] = (float)acc;
    }
    __syncthreads();
    const float *nx = G + kC * kC, *ny = nx + kC;
    if (MODE == 1) {
        if (tid < kC) {
            float best = 0.0f;
            int bj = 0;
            for (int j = 0; j < kC; ++j) {
                const float d = __fsub_rn(__fadd_rn(nx[tid], ny[j]), __fmul_rn(2.0f, G[tid * kC + j]));
                if (j == 0 || d < best) { best = d; bj = j; }
            }
            idx_out[b * kC + tid] = bj;
        }
    } else {
        if (tid < kC) {
            const float T = __ldg(temperature);
            const float nq = fmaxf(sqrtf(nx[tid]), 1e-12f);
            float v[kC], m = -INFINITY;
#pragma unroll
            for (int j = 0; j < kC; ++j) {
                const float nk = fmaxf(sqrtf(ny[j]), 1e-12f);
                v[j] = __fmul_rn(__fdiv_rn(G[tid * kC + j], __fmul_rn(nq, nk)), T);
                m = fmaxf(m, v[j]);
            }
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < kC; ++j) { v[j] = expf(v[j] - m); sum += v[j]; }
#pragma unroll
            for (int j = 0; j < kC; ++j) attn[tid * kC + j] = __fdiv_rn(v[j], sum);
        }
        __syncthreads();
        for (int i = tid; i < kC * kC; i += kTailThreads) {
            const int o = i / kC, c = i - o * kC;
            float acc = 0.0f;
#pragma unroll
            for (int m2 = 0; m2 < kC; ++m2) acc = fmaf(__ldg(w_po + o * kC + m2), attn[m2 * kC + c], acc);
            mixed_out[b * kC * kC + i] = acc;
        }
    }
}

inline void plan(int64_t hw, int &chunk, int &nchunks)
{
    // one CTA per SM (64 KB of shared memory, ~180 registers), chunk a multiple of the tile
    const int64_t target = (int64_t)sm_count();
    int64_t c = (hw + target - 1) / target;
    c = (c + kTile - 1) / kTile * kTile;
    if (c < kTile) c = kTile;
    chunk = (int)c;
    nchunks = (int)((hw + c - 1) / c);
    nchunks = (nchunks + kCluster - 1) / kCluster * kCluster;     // whole clusters (a CTA past the end adds zeros)
}

}  // namespace gram
}  // namespace wm

extern "C" size_t wm_gram32_workspace_bytes(int64_t B, int64_t hw)
{
    if (B <= 0 || hw <= 0) return 0;
    int chunk, nchunks;
    wm::gram::plan(hw, chunk, nchunks);
    return (size_t)B * nchunks * wm::gram::kOut * sizeof(double);
}

static int gram_launch(const char *who, int mode, const float *x, int64_t x_bstride, const float *y,
                       int64_t y_bstride, float *out, int *idx_out, const float *temperature,
                       const float *w_po, float *mixed_out, void *workspace, size_t workspace_bytes, int64_t B,
                       int64_t hw, wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::gram;
    WM_REQUIRE(B >= 0 && hw >= 0 && B <= 65535, "%s: bad sizes", who);
    if (B == 0) return WM_OK;
    WM_REQUIRE(x && y, "%s: null pointer", who);
    cudaStream_t s = (cudaStream_t)stream;
    WM_REQUIRE(hw > 0 || mode == 0, "%s: empty maps", who);
    if (hw == 0) {
        WM_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)B * kOut * sizeof(float), s));
        return WM_OK;
    }
    WM_REQUIRE(x_bstride >= kC * hw && y_bstride >= kC * hw, "%s: batch stride too small", who);
    int chunk, nchunks;
    plan(hw, chunk, nchunks);
    const size_t need = (size_t)B * nchunks * kOut * sizeof(double);
    WM_REQUIRE(workspace && workspace_bytes >= need, "%s: workspace too small (%zu < %zu)", who,
               workspace_bytes, need);
    WM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "%s: workspace alignment", who);
    const int vec = (hw % 4 == 0 && aligned16(x) && aligned16(y) && x_bstride % 4 == 0 &&
                     y_bstride % 4 == 0) ? 1 : 0;
    dim3 grid(nchunks, (unsigned)B);
    WM_CUDA_OK(cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSmemBytes));
    gram_partial_kernel<<<grid, kThreads, kSmemBytes, s>>>(x, x_bstride, y, y_bstride,
                                                  static_cast<double *>(workspace), hw, chunk,
                                                  nchunks, vec);
    WM_LAUNCH_OK("gram partial");
    const double *part = static_cast<const double *>(workspace);
    const int nparts = nchunks / kCluster;                      // one row of partials per cluster
    if (mode == 0) {
        dim3 rgrid((kOut + kThreads - 1) / kThreads, (unsigned)B);
        gram_reduce_kernel<<<rgrid, kThreads, 0, s>>>(part, out, nparts);
    } else if (mode == 1) {
        gram_reduce_tail_kernel<1><<<(unsigned)B, kTailThreads, 0, s>>>(part, out, nparts, idx_out, nullptr, nullptr, nullptr);
    } else {
        gram_reduce_tail_kernel<2><<<(unsigned)B, kTailThreads, 0, s>>>(part, out, nparts, nullptr, temperature, w_po, mixed_out);
    }
    WM_LAUNCH_OK("gram reduce");
    return WM_OK;
}

extern "C" int wm_gram32_fwd(const float *x, int64_t x_bstride, const float *y, int64_t y_bstride,
                             float *out, void *workspace, size_t workspace_bytes, int64_t B,
                             int64_t hw, wm_stream_t stream)
{
    WM_REQUIRE(out != nullptr || B == 0, "wm_gram32_fwd: null pointer");
    return gram_launch("wm_gram32_fwd", 0, x, x_bstride, y, y_bstride, out, nullptr, nullptr, nullptr, nullptr,
                       workspace, workspace_bytes, B, hw, stream);
}

extern "C" int wm_gram32_match_fwd(const float *x, int64_t x_bstride, const float *p, int64_t p_bstride,
                                   int *idx, float *gram_out, void *workspace, size_t workspace_bytes,
                                   int64_t B, int64_t hw, wm_stream_t stream)
{
    WM_REQUIRE(idx != nullptr || B == 0, "wm_gram32_match_fwd: null pointer");
    return gram_launch("wm_gram32_match_fwd", 1, x, x_bstride, p, p_bstride, gram_out, idx, nullptr, nullptr, nullptr,
                       workspace, workspace_bytes, B, hw, stream);
}

extern "C" int wm_gram32_attn_fwd(const float *q, int64_t q_bstride, const float *k, int64_t k_bstride,
                                  const float *temperature, const float *w_po, float *mixed,
                                  float *gram_out, void *workspace, size_t workspace_bytes, int64_t B,
                                  int64_t hw, wm_stream_t stream)
{
    WM_REQUIRE((temperature && w_po && mixed) || B == 0, "wm_gram32_attn_fwd: null pointer");
    return gram_launch("wm_gram32_attn_fwd", 2, q, q_bstride, k, k_bstride, gram_out, nullptr, temperature, w_po, mixed,
                       workspace, workspace_bytes, B, hw, stream);
}
