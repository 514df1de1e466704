// Shared definitions of the SS2D kernels (ss2d.cu forward, ss2d_bwd.cu backward): constants, the
// chunk geometry, packed-fp32 / tensor-core helpers, tile addressing, tile loads and stores.
#pragma once

#include <initializer_list>

#include "common.cuh"

namespace wm {
namespace ss2d {

constexpr int kD = 64;       // d_inner
constexpr int kN = 16;       // d_state
constexpr int kK = 4;        // directions
constexpr int kProj = 34;    // dt_rank(2) + 2*d_state
// Tile layout of the including translation unit: a CTA owns kSeq neighbouring chunks ("strands") and
// walks them kTP steps per tile (always 64 positions per tile).  The forward (ss2d.cu) uses 8 x 8 --
// a scan thread then owns 4 channels x 8 states, which cuts the shared-memory bytes per state update
// from 5 to 3 --, the backward (ss2d_bwd.cu) 4 x 16.  The chunk plan (Geom) does not depend on it.
#ifndef WM_SS2D_SEQ
#error "define WM_SS2D_SEQ / WM_SS2D_TP before including ss2d_common.cuh"
#endif
constexpr int kSeq = WM_SS2D_SEQ;   // strands per CTA
constexpr int kTP = WM_SS2D_TP;     // steps per tile
constexpr int kPos = kSeq * kTP;    // 64 positions per tile
static_assert(kPos == 64 && kSeq % 4 == 0 && kTP % 4 == 0, "tile layout");
constexpr int kAlign = 16;          // chunk lengths are multiples of 16 steps in every layout
constexpr int kFwdSeq = 8;          // the forward's strands per CTA (Geom::row_ctas / col_ctas)
constexpr int kXS = 72;      // xs row stride  [channel][position]  (== 8 mod 32: mma A loads)
constexpr int kPJ = 36;      // pj row stride  [position][B16|C16|dt2|pad2]
constexpr int kDD = 132;     // dd row stride  [position][channel] float2 (dt, u)  (== 4 mod 32)
constexpr int kYS = 65;      // ys row stride  [channel][position]
constexpr int kThreads = 256;
constexpr int kChains = kD * kN;     // 1024 (d,n) chains per direction
constexpr int kNTiles = 5;           // mma n-tiles: B0-7, B8-15, C0-7, C8-15, dt(2)+pad

// shared memory carve-up (floats)
constexpr int kOffXs = 0;
constexpr int kOffPj = kOffXs + kD * kXS;               // 4608
constexpr int kOffDd = kOffPj + kPos * kPJ;             // +2304
constexpr int kOffYs = kOffDd + kPos * kDD;             // +8448
constexpr int kOffWf = kOffYs + kD * kYS;               // +4160
constexpr int kOffCst = kOffWf + 8 * kNTiles * 32 * 4;  // +5120
constexpr int kSmemFloats = kOffCst + 3 * kD;           // +192
constexpr int kTileFloats = kPos * kPJ + kPos * kDD;    // one projected tile [pj | dd]: 10752 floats
constexpr size_t kSmemBytes = sizeof(float) * kSmemFloats;   // 99,328 B -> 2 CTAs per SM

struct Geom {
    int B, h, w;
    int64_t L;
    int row_T;       // steps per row chunk (multiple of kTP)
    int row_chunks;  // ceil(L / row_T)
    int row_ctas;    // ceil(row_chunks / kSeq)
    int col_seg;     // steps per column chunk (multiple of kTP): a column is cut into ncolseg chunks
    int ncolseg;     // ceil(h / col_seg)
    int col_ctas;    // ceil(w / kSeq) * ncolseg
    int max_chunks;  // max(row_chunks, w * ncolseg): chunk stride of the aggregate arrays
    int cols_first;  // launch order: column-direction CTAs before row-direction CTAs
    int vec_rows;    // 1 when row tiles may use 16-byte global accesses (L % 4 == 0)
    int vec_cols;    // 1 when column tiles may (w % 4 == 0)
};

struct Launch {
    int ndirs;
    int dir[4];
    int cta_begin[5];  // blockIdx.x range of each direction
};

struct Params {
    const float *x;            // (B,64,L)
    const float *x_proj_w;     // (4,34,64)
    const float *dt_w;         // (4,64,2)
    const float *dt_b;         // (4,64)
    const float *A_logs;       // (256,16)
    const float *Ds;           // (256)
    float *planes;             // (4,B,64,L) per-direction outputs, pixel-major
    float *aggP;               // (B,4,max_chunks,1024)
    float *aggH;               // (B,4,max_chunks,1024)  pass 1: local end state; after carry: h_in
    long long *dbg;            // developer aid (wm_ss2d_debug_timing): per-CTA phase cycle sums
    float *tiles;              // replay scratch (null: pass 2 recomputes): pass 1 dumps the projected
                               // tile [pj | dd] of every (CTA, tile), pass 2 reads it back
    int tile_stride;           // tile slots per CTA
    float *hbuf;               // checkpoint pass (backward): state after every step of ONE direction,
                               // [b][chunk][step][1024 chains]
};

// chunks and steps per chunk of direction k (chunk index = position in the carry chain)
__host__ __device__ __forceinline__ int dir_chunks(const Geom &g, int k)
{
    return (k & 1) ? g.w * g.ncolseg : g.row_chunks;
}
__host__ __device__ __forceinline__ int dir_chunk_len(const Geom &g, int k)
{
    return (k & 1) ? g.col_seg : g.row_T;
}

// ---- packed fp32x2 helpers (Blackwell FFMA2 / FMUL2) ---------------------------------------
typedef unsigned long long f32x2;
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
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 ex2_2(f32x2 v)
{
    float lo, hi;
    unpack2(v, lo, hi);
    return pack2(ex2_approx(lo), ex2_approx(hi));
}

// ---- tensor-core helpers ------------------------------------------------------------------
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
// bf16 x bf16 -> fp32, m16n8k16 (the 3xTF32 correction terms: see the weight fragments in ss2d.cu)
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 -> one register of two bf16 (round to nearest even), `lo` in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// softplus (torch: beta 1, threshold 20) = max(v,0) + log1p(exp(-|v|)), with
// log1p(e) = 2 atanh(e / (2 + e)), e in (0,1]: odd series in s = e/(2+e) <= 1/3 up to s^13.
// Relative error <= 2e-7 for |v| < 15 (8e-7 worst case from the ex2 argument rounding beyond).
__device__ __forceinline__ float softplus_fast(float v)
{
    const float e = ex2_approx(-fabsf(v) * 1.4426950408889634f);
    const float s = __fdividef(e, 2.0f + e);
    const float t = s * s;
    float p = fmaf(t, 0.07692307692f, 0.09090909091f);
    p = fmaf(p, t, 0.11111111111f);
    p = fmaf(p, t, 0.14285714286f);
    p = fmaf(p, t, 0.2f);
    p = fmaf(p, t, 0.33333333333f);
    p = fmaf(p, t, 1.0f);
    const float sp = fmaf(2.0f * s, p, fmaxf(v, 0.0f));
    return v > 20.0f ? v : sp;
}

// Two softplus values at once on the packed FP32 pipe (same arithmetic per element as
// softplus_fast: identical results, ~2/3 of the instructions).
__device__ __forceinline__ void softplus_fast2(float v0, float v1, float &o0, float &o1)
{
    const float e0 = ex2_approx(-fabsf(v0) * 1.4426950408889634f);
    const float e1 = ex2_approx(-fabsf(v1) * 1.4426950408889634f);
    const f32x2 e = pack2(e0, e1);
    float d0, d1;
    unpack2(fadd2(e, pack2(2.0f, 2.0f)), d0, d1);
    // s = e / (2 + e) as e * rcp(2 + e), exactly what __fdividef does for these magnitudes
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
    const f32x2 sv = fmul2(e, pack2(r0, r1));
    const f32x2 t = fmul2(sv, sv);
    f32x2 p = ffma2(t, pack2(0.07692307692f, 0.07692307692f), pack2(0.09090909091f, 0.09090909091f));
    p = ffma2(p, t, pack2(0.11111111111f, 0.11111111111f));
    p = ffma2(p, t, pack2(0.14285714286f, 0.14285714286f));
    p = ffma2(p, t, pack2(0.2f, 0.2f));
    p = ffma2(p, t, pack2(0.33333333333f, 0.33333333333f));
    p = ffma2(p, t, pack2(1.0f, 1.0f));
    const f32x2 sp = ffma2(fadd2(sv, sv), p, pack2(fmaxf(v0, 0.0f), fmaxf(v1, 0.0f)));
    float s0, s1;
    unpack2(sp, s0, s1);
    o0 = v0 > 20.0f ? v0 : s0;
    o1 = v1 > 20.0f ? v1 : s1;
}

struct TileGeom {
    // this CTA's 4 strands live in one memory band; tile ti covers steps [16 ti, 16 ti + 16)
    int k;          // direction
    bool col;       // column-major direction
    bool fwd;       // forward direction (0 or 1)
    int chunk0;     // rows: sequence-order chunk index of strand 0; columns: image column of strand 0
    int seg, t0;    // columns: segment index inside the column and its first step (rows: 0)
    int maxlen;     // steps of the longest strand of this CTA
};

// number of valid steps of strand s
__device__ __forceinline__ int strand_len(const Geom &g, const TileGeom &tg, int s)
{
    const int c = tg.chunk0 + s;
    if (tg.col) return c < g.w ? min(g.col_seg, g.h - tg.t0) : 0;
    const int64_t rem = g.L - (int64_t)c * g.row_T;
    return rem <= 0 ? 0 : (rem < g.row_T ? (int)rem : g.row_T);
}

// element offset (inside a channel plane) of step t of strand s
__device__ __forceinline__ int64_t strand_elem(const Geom &g, const TileGeom &tg, int s, int t)
{
    const int c = tg.chunk0 + s;
    if (!tg.col) {
        const int64_t l = (int64_t)c * g.row_T + t;       // sequence index
        return tg.fwd ? l : g.L - 1 - l;
    }
    const int j = tg.fwd ? c : g.w - 1 - c;
    const int i = tg.fwd ? tg.t0 + t : g.h - 1 - tg.t0 - t;
    return (int64_t)i * g.w + j;
}

// smem position index of (strand s, step e within the tile): mirrors memory order so that
// 16-byte global chunks land contiguously
__device__ __forceinline__ int tile_pos(const TileGeom &tg, int s, int e)
{
    if (!tg.col) return s * kTP + (tg.fwd ? e : kTP - 1 - e);
    return e * kSeq + (tg.fwd ? s : kSeq - 1 - s);
}

// Is tile ti a full, 16-byte-addressable tile?  (all 4 strands present, 16 valid steps)
__device__ __forceinline__ bool tile_is_vec(const Geom &g, const TileGeom &tg, int ti)
{
    const int t_end = ti * kTP + kTP;
    if (tg.col)
        return g.vec_cols && tg.chunk0 + kSeq <= g.w && t_end <= min(g.col_seg, g.h - tg.t0);
    return g.vec_rows && (int64_t)(tg.chunk0 + kSeq - 1) * g.row_T + t_end <= g.L;
}

// Per-thread addressing of the four 16-byte chunks it moves per full tile (x in, y out).
struct ChunkMap {
    int64_t goff;      // element offset inside a channel plane for tile 0 (channel d0)
    int64_t gstep;     // added per tile
    uint32_t xs_dst;   // shared address of xs[d0][p0]
    int ys_idx;        // index of ys[d0][p0]
    int d0;            // first channel (chunk j: d0 + 16 j)
};

__device__ __forceinline__ ChunkMap make_chunk_map(const Geom &g, const TileGeom &tg, float *xs)
{
    ChunkMap cm;
    // thread -> (16-byte chunk cidx of the 16 per channel row, channel d0 of 16; +16j in the copy loops).
    // A quarter-warp covers 8 consecutive chunks of ONE channel (128 contiguous bytes: the cp.async / float4
    // side), a warp 4 adjacent channels: with the odd ys pitch the scalar reads of store_tile then touch 32
    // different banks (lanes on chunks c and c + 8 of one channel shared a bank).
    const int tid = threadIdx.x;
    const int lane = tid & 31, wp = tid >> 5;
    const int cidx = (lane & 7) + 8 * (wp & 1);
    cm.d0 = 4 * (wp >> 1) + (lane >> 3);
    int p0;
    if (!tg.col) {
        // rows: a strand's kTP steps are kTP/4 chunks of 16 bytes
        constexpr int kCps = kTP / 4;
        const int s = cidx / kCps, v = cidx % kCps;
        const int64_t l0 = (int64_t)(tg.chunk0 + s) * g.row_T;
        cm.goff = tg.fwd ? l0 + 4 * v : g.L - 1 - l0 - (kTP - 1) + 4 * v;
        cm.gstep = tg.fwd ? kTP : -kTP;
        p0 = s * kTP + 4 * v;
    } else {
        // columns: one image row of the kSeq adjacent columns is kSeq/4 chunks of 16 bytes
        constexpr int kCpr = kSeq / 4;
        const int e = cidx / kCpr, hv = cidx % kCpr;
        const int i = tg.fwd ? tg.t0 + e : g.h - 1 - tg.t0 - e;
        const int jlow = tg.fwd ? tg.chunk0 : g.w - kSeq - tg.chunk0;
        cm.goff = (int64_t)i * g.w + jlow + 4 * hv;
        cm.gstep = tg.fwd ? (int64_t)kTP * g.w : -(int64_t)kTP * g.w;
        p0 = e * kSeq + 4 * hv;
    }
    cm.xs_dst = (uint32_t)__cvta_generic_to_shared(xs + cm.d0 * kXS + p0);
    cm.ys_idx = cm.d0 * kYS + p0;
    return cm;
}

__device__ __forceinline__ void load_tile(const Geom &g, const TileGeom &tg, const ChunkMap &cm,
                                          int ti, const float *__restrict__ xb, float *xs)
{
    if (tile_is_vec(g, tg, ti)) {
        const float *src = xb + (int64_t)cm.d0 * g.L + cm.goff + (int64_t)ti * cm.gstep;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cm.xs_dst + j * 16 * kXS * 4),
                         "l"(src + (int64_t)j * 16 * g.L)
                         : "memory");
    } else {
        // ragged tile: scalar loads, zero fill
        const int tid = threadIdx.x;
#pragma unroll 1
        for (int r = 0; r < kD * kPos / kThreads; ++r) {
            const int idx = tid + r * kThreads;      // 0..4095 = (d, s, e)
            const int d = idx >> 6, s = (idx & 63) / kTP, e = (idx & 63) % kTP;
            const int t = ti * kTP + e;
            float v = 0.0f;
            if (t < strand_len(g, tg, s)) v = __ldg(xb + (int64_t)d * g.L + strand_elem(g, tg, s, t));
            xs[d * kXS + tile_pos(tg, s, e)] = v;
        }
    }
    cp_async_commit();
}

__device__ __forceinline__ void store_tile(const Geom &g, const TileGeom &tg, const ChunkMap &cm,
                                           int ti, float *__restrict__ ob, const float *ys)
{
    if (tile_is_vec(g, tg, ti)) {
        float *dst = ob + (int64_t)cm.d0 * g.L + cm.goff + (int64_t)ti * cm.gstep;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float *a = ys + cm.ys_idx + j * 16 * kYS;
            *reinterpret_cast<float4 *>(dst + (int64_t)j * 16 * g.L) = make_float4(a[0], a[1], a[2], a[3]);
        }
    } else {
        const int tid = threadIdx.x;
#pragma unroll 1
        for (int r = 0; r < kD * kPos / kThreads; ++r) {
            const int idx = tid + r * kThreads;
            const int d = idx >> 6, s = (idx & 63) / kTP, e = (idx & 63) % kTP;
            const int t = ti * kTP + e;
            if (t < strand_len(g, tg, s))
                ob[(int64_t)d * g.L + strand_elem(g, tg, s, t)] = ys[d * kYS + tile_pos(tg, s, e)];
        }
    }
}

// ---- host-side entry points of ss2d.cu used by the backward (ss2d_bwd.cu) ---------------------
Geom make_geom(int64_t B, int64_t h, int64_t w);
Launch make_launch(const Geom &g, std::initializer_list<int> dirs);
// mode 0: pass 1 (chunk aggregates from h = 0); 1: pass 2 (outputs); 2: checkpoint pass (every
// chunk from its true initial state, the state after EVERY step written to prm.hbuf)
int launch_pass(int mode, const Params &prm, const Geom &g, const Launch &ln, cudaStream_t s);
int launch_carry(const float *aggP, float *aggH, const Geom &g, cudaStream_t s);

}  // namespace ss2d
}  // namespace wm
