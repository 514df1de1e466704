// SS2D core for sm_100a: cross-scan + x_proj + dt_proj + softplus + selective scan + D skip +
// cross-merge, without ever materialising xs / x_dbl / dts / Bs / Cs / out_y.
//
// Replaces SS2D.forward_core and the y1+y2+y3+y4 of SS2D.forward
// (reference wavemamba_arch.py:446-478,490; direction index maps: SURVEY.md appendix A).
//
// Work decomposition ("strands")
//   A direction k is a sequence of L = h*w positions.  It is cut into chunks that are
//   contiguous in the sequence AND cheap to address in the NCHW map:
//     k=0/2 (row-major, forward/backward): chunk = row_T consecutive positions of the
//            flattened map (element stride +-1);
//     k=1/3 (column-major, forward/backward): chunk = one image column (element stride +-w).
//   One CTA (512 threads, 2 CTAs/SM) owns 4 neighbouring chunks ("strands") of one direction
//   and all 64 channels x 16 states of them.  thread = (strand s, channel d, state half):
//   8 states in registers, walked sequentially, 16 steps per tile.  For column directions the
//   four strands are four adjacent columns, so a tile row is one 16-byte segment per channel.
//
// Per tile (64 positions x 64 channels), five phases separated by __syncthreads:
//   load     cp.async 16-byte chunks -> xs[d][p]      (issued one tile ahead, overlaps the scan)
//   dt-low   pj[p][32..33] = W_k[0:2] (2x64) . x[:,p]   plain FP32 FMAs, 8 lanes per position
//   project+delta  (one phase, interleaved per warp so the tensor-pipe latency hides behind the
//            delta arithmetic)
//            pj[p][B16|C16] = W_k[2:34] (32x64) . x[:,p]  on the tensor cores: mma.sync m16n8k8
//            TF32 with the 3xTF32 split (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi), fp32 accumulate;
//            dd[d][p] = (dt, dt*u), dt = softplus(dt_proj . dt_low + bias); ys[d][p] = D*u
//   scan     h = exp2(dt*A2)*h + dt*u*B ; y += C.h    packed FFMA2/FMUL2, MUFU ex2.approx
//   store    ys[d][p] -> 16-byte coalesced stores into this direction's output plane (pass 2)
//
// Chunks are made independent with a three-phase carry scheme (the recurrence is linear):
//   pass 1  every chunk from h=0: aggregate (P = prod a = exp2(A2*sum dt), H = local end state)
//   carry   per (b,k,d,n): h_in[c] = P[c-1]*h_in[c-1] + H[c-1]          (tiny, sequential in c)
//   pass 2  every chunk again from its true h_in, emitting y into one plane per direction.
// The four planes are summed in the reference's order ((y0+y2)+y1)+y3 by the consumer
// (wm_lfss_out_fwd) or by the combine kernel of wm_ss2d_core_fwd.  No atomics: deterministic.
//
// Roofline: not HBM-bound.  Each state update costs one MUFU ex2 (16/clk/SM) and every exp is
// evaluated twice (pass 1 and pass 2); see DESIGN.md section 4.
#include <initializer_list>

#include "common.cuh"

namespace wm {
namespace ss2d {

constexpr int kD = 64;       // d_inner
constexpr int kN = 16;       // d_state
constexpr int kK = 4;        // directions
constexpr int kProj = 34;    // dt_rank(2) + 2*d_state
constexpr int kSeq = 4;      // strands per CTA
constexpr int kTP = 16;      // steps per tile
constexpr int kPos = kSeq * kTP;  // 64 positions per tile
constexpr int kXS = 72;      // xs row stride  [channel][position]  (== 8 mod 32: mma A loads)
constexpr int kPJ = 36;      // pj row stride  [position][B16|C16|dt2|pad2]
constexpr int kDS = 65;      // dd row stride  [channel][position] float2 (odd: scan reads)
constexpr int kYS = 65;      // ys row stride  [channel][position]
constexpr int kThreads = 512;
constexpr int kChains = kD * kN;     // 1024 (d,n) chains per direction
constexpr int kNTiles = 4;           // mma n-tiles: B0-7, B8-15, C0-7, C8-15 (dt rows: FP32 FMAs)

// shared memory carve-up (floats)
constexpr int kOffXs = 0;
constexpr int kOffPj = kOffXs + kD * kXS;               // 4608
constexpr int kOffDd = kOffPj + kPos * kPJ;             // +2304
constexpr int kOffYs = kOffDd + kD * kDS * 2;           // +8320
constexpr int kOffWf = kOffYs + kD * kYS;               // +4160
constexpr int kOffCst = kOffWf + 8 * kNTiles * 32 * 4;  // +5120
constexpr int kOffDlw = kOffCst + 4 * kD;               // +256
constexpr int kSmemFloats = kOffDlw + 2 * kD;           // +128
constexpr size_t kSmemBytes = sizeof(float) * kSmemFloats;   // 95,488 B -> 2 CTAs per SM

struct Geom {
    int B, h, w;
    int64_t L;
    int row_T;       // steps per row chunk (multiple of kTP)
    int row_chunks;  // ceil(L / row_T)
    int row_ctas;    // ceil(row_chunks / kSeq)
    int col_ctas;    // ceil(w / kSeq)
    int max_chunks;  // max(row_chunks, w): chunk stride of the aggregate arrays
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
};

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

struct TileGeom {
    // this CTA's 4 strands live in one memory band; tile ti covers steps [16 ti, 16 ti + 16)
    int k;          // direction
    bool col;       // column-major direction
    bool fwd;       // forward direction (0 or 1)
    int chunk0;     // sequence-order chunk index of strand 0
    int maxlen;     // steps per full chunk (row_T or h)
};

// number of valid steps of strand s
__device__ __forceinline__ int strand_len(const Geom &g, const TileGeom &tg, int s)
{
    const int c = tg.chunk0 + s;
    if (tg.col) return c < g.w ? g.h : 0;
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
    const int i = tg.fwd ? t : g.h - 1 - t;
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
        return g.vec_cols && tg.chunk0 + kSeq <= g.w && t_end <= g.h;
    return g.vec_rows && (int64_t)(tg.chunk0 + kSeq - 1) * g.row_T + t_end <= g.L;
}

// Global offset (inside a channel plane) of the 16-byte chunk `cidx` (0..15) of a vec tile and
// the smem position of its first element.  Rows: chunk = (s, v) -> 4 consecutive steps.
// Columns: chunk = e -> the 4 strands of one image row.
__device__ __forceinline__ void vec_chunk(const Geom &g, const TileGeom &tg, int ti, int cidx,
                                          int64_t &goff, int &p0)
{
    if (!tg.col) {
        const int s = cidx >> 2, v = cidx & 3;
        const int64_t l0 = (int64_t)(tg.chunk0 + s) * g.row_T + ti * kTP;  // first step of the tile
        // memory-ascending chunk v of the 16 elements of this strand's tile
        goff = tg.fwd ? l0 + 4 * v : g.L - 1 - l0 - (kTP - 1) + 4 * v;
        p0 = s * kTP + 4 * v;
    } else {
        const int e = cidx;
        const int i = tg.fwd ? ti * kTP + e : g.h - 1 - (ti * kTP + e);
        const int jlow = tg.fwd ? tg.chunk0 : g.w - kSeq - tg.chunk0;
        goff = (int64_t)i * g.w + jlow;
        p0 = e * kSeq;
    }
}

// Per-thread addressing of the two 16-byte chunks it moves per full tile (x in, y out).
struct ChunkMap {
    int64_t goff;      // element offset inside a channel plane for tile 0 (channel d0)
    int64_t gstep;     // added per tile
    uint32_t xs_dst;   // shared address of xs[d0][p0]
    int ys_idx;        // index of ys[d0][p0]
    int d0;            // first channel (second chunk: d0 + 32)
};

__device__ __forceinline__ ChunkMap make_chunk_map(const Geom &g, const TileGeom &tg, float *xs)
{
    ChunkMap cm;
    const int tid = threadIdx.x;
    const int cidx = tid & 15;
    cm.d0 = tid >> 4;
    int p0;
    if (!tg.col) {
        const int s = cidx >> 2, v = cidx & 3;
        const int64_t l0 = (int64_t)(tg.chunk0 + s) * g.row_T;
        cm.goff = tg.fwd ? l0 + 4 * v : g.L - 1 - l0 - (kTP - 1) + 4 * v;
        cm.gstep = tg.fwd ? kTP : -kTP;
        p0 = s * kTP + 4 * v;
    } else {
        const int e = cidx;
        const int i = tg.fwd ? e : g.h - 1 - e;
        const int jlow = tg.fwd ? tg.chunk0 : g.w - kSeq - tg.chunk0;
        cm.goff = (int64_t)i * g.w + jlow;
        cm.gstep = tg.fwd ? (int64_t)kTP * g.w : -(int64_t)kTP * g.w;
        p0 = e * kSeq;
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
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cm.xs_dst), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cm.xs_dst + 32 * kXS * 4),
                     "l"(src + 32 * g.L)
                     : "memory");
    } else {
        // ragged tile: scalar loads, zero fill
        const int tid = threadIdx.x;
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const int idx = tid + r * kThreads;      // 0..4095 = (d, s, e)
            const int d = idx >> 6, s = (idx >> 4) & 3, e = idx & 15;
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
        const float *a = ys + cm.ys_idx, *b = a + 32 * kYS;
        *reinterpret_cast<float4 *>(dst) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4 *>(dst + 32 * g.L) = make_float4(b[0], b[1], b[2], b[3]);
    } else {
        const int tid = threadIdx.x;
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const int idx = tid + r * kThreads;
            const int d = idx >> 6, s = (idx >> 4) & 3, e = idx & 15;
            const int t = ti * kTP + e;
            if (t < strand_len(g, tg, s))
                ob[(int64_t)d * g.L + strand_elem(g, tg, s, t)] = ys[d * kYS + tile_pos(tg, s, e)];
        }
    }
}

// One recurrence step.  DP = smem position stride per step (+1 row fwd, -1 row bwd, +4 column).
template <bool FINAL>
__device__ __forceinline__ void scan_step(const float2 *ddp, const float *pjp, float *ysp,
                                          f32x2 (&hst)[4], const f32x2 (&A2)[4], float &sdt,
                                          bool writer)
{
    const float2 dv = *ddp;                       // (dt, dt*u)
    const float4 b0 = *reinterpret_cast<const float4 *>(pjp);
    const float4 b1 = *reinterpret_cast<const float4 *>(pjp + 4);
    const f32x2 dt2 = pack2(dv.x, dv.x), du2 = pack2(dv.y, dv.y);
    const f32x2 bb[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
    if (!FINAL) sdt += dv.x;
    f32x2 cc[4];
    if (FINAL) {
        const float4 c0 = *reinterpret_cast<const float4 *>(pjp + 16);
        const float4 c1 = *reinterpret_cast<const float4 *>(pjp + 20);
        cc[0] = pack2(c0.x, c0.y); cc[1] = pack2(c0.z, c0.w);
        cc[2] = pack2(c1.x, c1.y); cc[3] = pack2(c1.z, c1.w);
    }
    f32x2 acc = pack2(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const f32x2 a = ex2_2(fmul2(dt2, A2[j]));
        hst[j] = ffma2(a, hst[j], fmul2(du2, bb[j]));
        if (FINAL) acc = ffma2(hst[j], cc[j], acc);
    }
    if (FINAL) {
        float lo, hi;
        unpack2(acc, lo, hi);
        float y = lo + hi;
        y += __shfl_xor_sync(0xffffffffu, y, 1);
        if (writer) *ysp += y;                    // ys held D*u
    }
}

// The whole tile loop of one CTA, specialised on the pass and on the scan's smem stride.
template <bool FINAL, int DP>
__device__ __forceinline__ void run_cta(const Params &prm, const Geom &g, const TileGeom &tg,
                                        float *smem, int b)
{
    float *xs = smem + kOffXs;
    float *pj = smem + kOffPj;
    float2 *dd = reinterpret_cast<float2 *>(smem + kOffDd);
    float *ys = smem + kOffYs;
    float4 *wf = reinterpret_cast<float4 *>(smem + kOffWf);
    float *cst = smem + kOffCst;   // [dtw0 | dtw1 | dtb | Dskip] x 64
    float *dlw = smem + kOffDlw;   // dt rows of W_k, [r 2][kq 4][i 16]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int k = tg.k;
    const int ntiles = (tg.maxlen + kTP - 1) / kTP;

    const float *xb = prm.x + (int64_t)b * kD * g.L;
    const ChunkMap cm = make_chunk_map(g, tg, xs);
    load_tile(g, tg, cm, 0, xb, xs);   // in flight while the weights are prepared

    // ---- mma B fragments of W_k, pre-split into tf32 hi/lo -----------------------------------
    // n-tile nt, column n (0..7) -> row 2 + 8 nt + n of x_proj_weight[k] (34,64) = [dt(2)|B(16)|C(16)]
    for (int i = tid; i < 8 * kNTiles * 32; i += kThreads) {
        const int ln_ = i & 31, nt = (i >> 5) % kNTiles, ks = i / (32 * kNTiles);
        const int gq = ln_ >> 2, t4 = ln_ & 3;
        const float *wr = prm.x_proj_w + ((int64_t)k * kProj + 2 + nt * 8 + gq) * kD + ks * 8;
        const float w0 = __ldg(wr + t4), w1 = __ldg(wr + t4 + 4);
        const uint32_t h0 = to_tf32(w0), h1 = to_tf32(w1);
        const uint32_t l0 = to_tf32(w0 - __uint_as_float(h0)), l1 = to_tf32(w1 - __uint_as_float(h1));
        wf[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0),
                            __uint_as_float(l1));
    }
    // dt rows for the FP32 dt-low phase: dlw[r][kq][i] = W_k[r][4 i + kq]
    if (tid < 2 * kD) {
        const int r = tid >> 6, kq = (tid >> 4) & 3, i = tid & 15;
        dlw[tid] = __ldg(prm.x_proj_w + ((int64_t)k * kProj + r) * kD + 4 * i + kq);
    }
    if (tid < kD) {
        const int chn = k * kD + tid;
        cst[tid] = __ldg(prm.dt_w + chn * 2 + 0);
        cst[kD + tid] = __ldg(prm.dt_w + chn * 2 + 1);
        cst[2 * kD + tid] = __ldg(prm.dt_b + chn);
        cst[3 * kD + tid] = __ldg(prm.Ds + chn);
    }

    // ---- scan-thread identity: (strand, channel, state half) -------------------------------
    const int s = tid >> 7, d = (tid & 127) >> 1, half = tid & 1;
    f32x2 A2[4];  // A * log2(e), A = -exp(A_log)   (reference :462)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float *ap = prm.A_logs + (int64_t)(k * kD + d) * kN + half * 8 + 2 * j;
        A2[j] = pack2(-expf(__ldg(ap)) * 1.4426950408889634f,
                      -expf(__ldg(ap + 1)) * 1.4426950408889634f);
    }
    const int my_len = strand_len(g, tg, s);
    const int my_chunk = tg.chunk0 + s;
    const int64_t agg_off =
        (((int64_t)b * kK + k) * g.max_chunks + my_chunk) * kChains + (int64_t)d * kN + half * 8;

    f32x2 hst[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hst[j] = pack2(0.0f, 0.0f);
    if (FINAL && my_len > 0 && my_chunk > 0) {
        const float4 *hp = reinterpret_cast<const float4 *>(prm.aggH + agg_off);
        const float4 f0 = hp[0], f1 = hp[1];
        hst[0] = pack2(f0.x, f0.y); hst[1] = pack2(f0.z, f0.w);
        hst[2] = pack2(f1.x, f1.y); hst[3] = pack2(f1.z, f1.w);
    }
    double sum_dt = 0.0;

    // smem position of step 0 of my strand; step e sits at p0 + e*DP
    const int p0 = tile_pos(tg, s, 0);
    const float2 *dd0 = dd + d * kDS + p0;
    const float *pj0 = pj + p0 * kPJ + half * 8;
    float *ys0 = ys + d * kYS + p0;
    const bool writer = half == 0;

    float *oplane = FINAL ? prm.planes + (((int64_t)k * g.B + b) * kD) * g.L : nullptr;

    // projection role of this warp: m-tile (16 positions) x one n-tile (pass 1 needs B only)
    const int mt = warp & 3, nt = warp >> 2;      // nt: B0-7 | B8-15 | C0-7 | C8-15
    const bool has_mma = FINAL || nt < 2;

    long long tacc[5] = {0, 0, 0, 0, 0}, tprev = 0;
    const bool timed = prm.dbg != nullptr && tid == 0;
#define WM_TICK(k) \
    if (timed) { const long long tn = clock64(); tacc[k] += tn - tprev; tprev = tn; }
    if (timed) tprev = clock64();

#pragma unroll 1
    for (int ti = 0; ti < ntiles; ++ti) {
        cp_async_wait_all();
        __syncthreads();                       // xs(ti) landed; previous tile fully consumed
        WM_TICK(0);

        // ---- dt-low: 2 x 64 dot products per position in plain FP32 (reference :453) -------
        // thread = (dt row r = tid/256, k-slice kq: channels kq, 4+kq, .., position p); a quad of
        // lanes 8 apart holds the four k-slices of one position (xs reads are conflict-free)
        {
            const int r = tid >> 8, kq = (tid >> 3) & 3, p = ((tid & 255) >> 5) * 8 + (tid & 7);
            const float *xp = xs + kq * kXS + p;
            const float4 *wq = reinterpret_cast<const float4 *>(dlw + (r * 4 + kq) * 16);
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w4 = wq[q];
                a0 = fmaf(w4.x, xp[(4 * q + 0) * 4 * kXS], a0);
                a1 = fmaf(w4.y, xp[(4 * q + 1) * 4 * kXS], a1);
                a0 = fmaf(w4.z, xp[(4 * q + 2) * 4 * kXS], a0);
                a1 = fmaf(w4.w, xp[(4 * q + 3) * 4 * kXS], a1);
            }
            a0 += a1;
            a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
            if (kq == 0) pj[p * kPJ + 32 + r] = a0;
        }
        __syncthreads();
        WM_TICK(1);

        // ---- B/C projection on tensor cores (reference :453-454), interleaved with the delta
        //      phase: dt = softplus(dt_proj . dt_low + bias)  (reference :455, scan_fn) ---------
        // mma: warp = (m-tile, n-tile).  delta: warp w owns channels 4w..4w+3, lane = position
        // within a 32-position round; item ks of the k-loop is (round ks/4, channel ks%4).
        {
            float c0[4] = {0.f, 0.f, 0.f, 0.f};
            const int gq = lane >> 2, t4 = lane & 3;
            const float *abase = xs + t4 * kXS + mt * 16 + gq;
            const float4 *wfp = wf + nt * 32 + lane;
            const float *xw = xs + warp * 4 * kXS + lane;
            float2 *dw = dd + warp * 4 * kDS + lane;
            float *yw = ys + warp * 4 * kYS + lane;
            const float *cw = cst + warp * 4;
            float2 dlow = make_float2(0.f, 0.f);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                if (has_mma) {
                    const float *ap = abase + ks * 8 * kXS;
                    const float av[4] = {ap[0], ap[8], ap[4 * kXS], ap[4 * kXS + 8]};
                    uint32_t ahi[4], alo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split_tf32(av[i], ahi[i], alo[i]);
                    const float4 bw = wfp[ks * kNTiles * 32];
                    mma_tf32(c0, alo, __float_as_uint(bw.x), __float_as_uint(bw.y));
                    mma_tf32(c0, ahi, __float_as_uint(bw.z), __float_as_uint(bw.w));
                    mma_tf32(c0, ahi, __float_as_uint(bw.x), __float_as_uint(bw.y));
                }
                const int rnd = ks >> 2, i = ks & 3;
                if (i == 0)
                    dlow = *reinterpret_cast<const float2 *>(pj + (rnd * 32 + lane) * kPJ + 32);
                const float u = xw[i * kXS + rnd * 32];
                const float raw = fmaf(cw[kD + i], dlow.y, cw[i] * dlow.x) + cw[2 * kD + i];
                const float dt = softplus_fast(raw);
                dw[i * kDS + rnd * 32] = make_float2(dt, dt * u);
                if (FINAL) yw[i * kYS + rnd * 32] = cw[3 * kD + i] * u;
            }
            if (has_mma) {
                float *pr = pj + (mt * 16 + gq) * kPJ + nt * 8 + 2 * t4;
                *reinterpret_cast<float2 *>(pr) = make_float2(c0[0], c0[1]);
                *reinterpret_cast<float2 *>(pr + 8 * kPJ) = make_float2(c0[2], c0[3]);
            }
        }
        __syncthreads();                       // xs is dead from here on
        WM_TICK(2);

        if (ti + 1 < ntiles) load_tile(g, tg, cm, ti + 1, xb, xs);

        // ---- recurrence over the 16 steps of this tile        (reference :465-471) --------
        {
            const int nvalid = my_len - ti * kTP;          // warp-uniform
            float sdt = 0.0f;
            if (nvalid >= kTP) {
#pragma unroll
                for (int e = 0; e < kTP; ++e)
                    scan_step<FINAL>(dd0 + e * DP, pj0 + e * DP * kPJ, ys0 + e * DP, hst, A2, sdt, writer);
            } else {
#pragma unroll 1
                for (int e = 0; e < nvalid; ++e)
                    scan_step<FINAL>(dd0 + e * DP, pj0 + e * DP * kPJ, ys0 + e * DP, hst, A2, sdt, writer);
            }
            if (!FINAL) sum_dt += (double)sdt;
        }
        if (FINAL) {
            __syncthreads();
            WM_TICK(3);
            store_tile(g, tg, cm, ti, oplane, ys);
            WM_TICK(4);
        }
    }
    if (!FINAL) { WM_TICK(3); }
#undef WM_TICK
    if (timed) {
        long long *o = prm.dbg + (blockIdx.x & 4095) * 6;
        for (int i = 0; i < 5; ++i) o[i] = tacc[i];
        o[5] = ntiles;
    }

    if (!FINAL && my_len > 0) {
        float4 *pp4 = reinterpret_cast<float4 *>(prm.aggP + agg_off);
        float4 *hp4 = reinterpret_cast<float4 *>(prm.aggH + agg_off);
        float pv[8], hv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a_lo, a_hi;
            unpack2(A2[j], a_lo, a_hi);
            pv[2 * j] = ex2_approx((float)((double)a_lo * sum_dt));
            pv[2 * j + 1] = ex2_approx((float)((double)a_hi * sum_dt));
            unpack2(hst[j], hv[2 * j], hv[2 * j + 1]);
        }
        pp4[0] = make_float4(pv[0], pv[1], pv[2], pv[3]);
        pp4[1] = make_float4(pv[4], pv[5], pv[6], pv[7]);
        hp4[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
        hp4[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
    }
}

// FINAL=false: pass 1 (aggregates).  FINAL=true: pass 2 (outputs).
template <bool FINAL>
__global__ void __launch_bounds__(kThreads, 2)
ss2d_pass_kernel(const Params prm, const Geom g, const Launch ln)
{
    extern __shared__ __align__(16) float smem[];
    int k = ln.dir[0], begin = 0;
    if (ln.ndirs > 1 && (int)blockIdx.x >= ln.cta_begin[1]) { k = ln.dir[1]; begin = ln.cta_begin[1]; }
    if (ln.ndirs > 2 && (int)blockIdx.x >= ln.cta_begin[2]) { k = ln.dir[2]; begin = ln.cta_begin[2]; }
    if (ln.ndirs > 3 && (int)blockIdx.x >= ln.cta_begin[3]) { k = ln.dir[3]; begin = ln.cta_begin[3]; }
    TileGeom tg;
    tg.k = k;
    tg.col = (k & 1) != 0;
    tg.fwd = k < 2;
    tg.chunk0 = (blockIdx.x - begin) * kSeq;
    tg.maxlen = tg.col ? g.h : g.row_T;
    const int b = blockIdx.y;
    if (tg.col) run_cta<FINAL, kSeq>(prm, g, tg, smem, b);
    else if (tg.fwd) run_cta<FINAL, 1>(prm, g, tg, smem, b);
    else run_cta<FINAL, -1>(prm, g, tg, smem, b);
}

// h_in[c] = P[c-1]*h_in[c-1] + H[c-1], h_in[0] = 0; written over aggH in place.
__global__ void __launch_bounds__(256)
ss2d_carry_kernel(const float *__restrict__ aggP, float *__restrict__ aggH, Geom g)
{
    const int chain = blockIdx.x * 256 + threadIdx.x;  // 0..1023
    const int k = blockIdx.y, b = blockIdx.z;
    const int nchunks = (k & 1) ? g.w : g.row_chunks;
    const int64_t off = (((int64_t)b * kK + k) * g.max_chunks) * kChains + chain;
    const float *P = aggP + off;
    float *H = aggH + off;
    float carry = 0.0f;
    int c = 0;
    for (; c + 8 <= nchunks; c += 8) {
        float p[8], hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            p[i] = P[(int64_t)(c + i) * kChains];
            hv[i] = H[(int64_t)(c + i) * kChains];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            H[(int64_t)(c + i) * kChains] = carry;
            carry = fmaf(p[i], carry, hv[i]);
        }
    }
    for (; c < nchunks; ++c) {
        const float p = P[(int64_t)c * kChains], hv = H[(int64_t)c * kChains];
        H[(int64_t)c * kChains] = carry;
        carry = fmaf(p, carry, hv);
    }
}

// y = ((p0 + p2) + p1) + p3   -- the reference's y1 + y2 + y3 + y4 order (:474-478,490)
__global__ void __launch_bounds__(256)
ss2d_combine_kernel(float *__restrict__ y, const float *__restrict__ planes, int64_t n, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * 256;
    const float *p0 = planes, *p1 = planes + n, *p2 = planes + 2 * n, *p3 = planes + 3 * n;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
            const float4 a = ld_stream4(p0 + 4 * i), b = ld_stream4(p1 + 4 * i);
            const float4 c = ld_stream4(p2 + 4 * i), dq = ld_stream4(p3 + 4 * i);
            float4 r;
            r.x = ((a.x + c.x) + b.x) + dq.x; r.y = ((a.y + c.y) + b.y) + dq.y;
            r.z = ((a.z + c.z) + b.z) + dq.z; r.w = ((a.w + c.w) + b.w) + dq.w;
            *reinterpret_cast<float4 *>(y + 4 * i) = r;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
            y[i] = ((p0[i] + p2[i]) + p1[i]) + p3[i];
    }
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

Geom make_geom(int64_t B, int64_t h, int64_t w)
{
    Geom g;
    g.B = (int)B; g.h = (int)h; g.w = (int)w;
    g.L = h * w;
    g.col_ctas = (int)((w + kSeq - 1) / kSeq);
    // Row chunks: about as many as there are columns so both orientations give similar CTA
    // counts and similar chunk lengths, and never shorter than 4 tiles.
    int64_t T = (g.L + w - 1) / w;       // = h
    if (T < 4 * kTP) T = 4 * kTP;
    T = align_up(T, kTP);
    g.row_T = (int)T;
    g.row_chunks = (int)((g.L + T - 1) / T);
    g.row_ctas = (g.row_chunks + kSeq - 1) / kSeq;
    g.max_chunks = g.row_chunks > g.w ? g.row_chunks : g.w;
    g.vec_rows = (g.L % 4 == 0) ? 1 : 0;
    g.vec_cols = (w % 4 == 0) ? 1 : 0;
    return g;
}

struct Workspace {
    int64_t planes_off, aggP_off, aggH_off, total;
};

Workspace plan_workspace(const Geom &g)
{
    Workspace ws;
    const int64_t planes_bytes = align_up((int64_t)kK * g.B * kD * g.L * 4, 256);
    const int64_t agg_bytes = align_up((int64_t)g.B * kK * g.max_chunks * kChains * 4, 256);
    ws.planes_off = 0;
    ws.aggP_off = planes_bytes;
    ws.aggH_off = planes_bytes + agg_bytes;
    ws.total = planes_bytes + 2 * agg_bytes;
    return ws;
}

Launch make_launch(const Geom &g, std::initializer_list<int> dirs)
{
    Launch ln;
    ln.ndirs = 0;
    int acc = 0;
    for (int k : dirs) {
        ln.dir[ln.ndirs] = k;
        ln.cta_begin[ln.ndirs] = acc;
        acc += (k & 1) ? g.col_ctas : g.row_ctas;
        ++ln.ndirs;
    }
    for (int i = ln.ndirs; i < 4; ++i) { ln.dir[i] = 0; ln.cta_begin[i] = acc; }
    ln.cta_begin[4] = acc;
    return ln;
}

long long *g_dbg = nullptr;   // wm_ss2d_debug_timing
int g_dbg_pad = 0;            // extra dynamic smem (forces one CTA per SM when > 0)

// Runs pass 1, carry, pass 2; leaves the four direction planes at workspace[0 : 4*B*64*L].
int run_dirs(const float *x, const float *x_proj_weight, const float *dt_projs_weight,
             const float *dt_projs_bias, const float *A_logs, const float *Ds, void *workspace,
             size_t workspace_bytes, int64_t B, int64_t h, int64_t w, cudaStream_t s,
             const char *who)
{
    WM_REQUIRE(x && x_proj_weight && dt_projs_weight && dt_projs_bias && A_logs && Ds,
               "%s: null pointer", who);
    WM_REQUIRE(B <= 65535, "%s: batch %lld exceeds 65535", who, (long long)B);
    WM_REQUIRE(h * w < (int64_t)1 << 31, "%s: h*w too large", who);
    WM_REQUIRE(aligned16(x), "%s: x must be 16-byte aligned", who);
    const Geom g = make_geom(B, h, w);
    const Workspace ws = plan_workspace(g);
    WM_REQUIRE(workspace && workspace_bytes >= (size_t)ws.total,
               "%s: workspace too small (%zu < %lld bytes)", who, workspace_bytes,
               (long long)ws.total);
    WM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "%s: workspace must be 256-byte aligned", who);
    char *wsb = static_cast<char *>(workspace);

    Params prm;
    prm.x = x; prm.x_proj_w = x_proj_weight; prm.dt_w = dt_projs_weight; prm.dt_b = dt_projs_bias;
    prm.A_logs = A_logs; prm.Ds = Ds;
    prm.planes = reinterpret_cast<float *>(wsb + ws.planes_off);
    prm.aggP = reinterpret_cast<float *>(wsb + ws.aggP_off);
    prm.aggH = reinterpret_cast<float *>(wsb + ws.aggH_off);
    prm.dbg = g_dbg;
    const size_t smem_bytes = kSmemBytes + (size_t)g_dbg_pad;

    WM_CUDA_OK(cudaFuncSetAttribute(ss2d_pass_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    WM_CUDA_OK(cudaFuncSetAttribute(ss2d_pass_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    // interleave row and column CTAs of the same cost so waves stay balanced
    const Launch ln = make_launch(g, {0, 1, 2, 3});
    dim3 grid(ln.cta_begin[4], (unsigned)B);
    ss2d_pass_kernel<false><<<grid, kThreads, smem_bytes, s>>>(prm, g, ln);
    WM_LAUNCH_OK("ss2d pass 1");
    dim3 cgrid(kChains / 256, kK, (unsigned)B);
    ss2d_carry_kernel<<<cgrid, 256, 0, s>>>(prm.aggP, prm.aggH, g);
    WM_LAUNCH_OK("ss2d carry");
    ss2d_pass_kernel<true><<<grid, kThreads, smem_bytes, s>>>(prm, g, ln);
    WM_LAUNCH_OK("ss2d pass 2");
    return WM_OK;
}

}  // namespace ss2d
}  // namespace wm

extern "C" int wm_ss2d_debug_timing(void *device_buffer)
{
    // low bit of a NULL-buffer call is not used; a buffer address with bit 0 set requests the
    // one-CTA-per-SM variant (smem padded) for occupancy experiments
    const uintptr_t v = reinterpret_cast<uintptr_t>(device_buffer);
    wm::ss2d::g_dbg_pad = (v & 1) ? 60 * 1024 : 0;
    wm::ss2d::g_dbg = reinterpret_cast<long long *>(v & ~(uintptr_t)1);
    return WM_OK;
}

extern "C" size_t wm_ss2d_core_workspace_bytes(int64_t B, int64_t h, int64_t w)
{
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)wm::ss2d::plan_workspace(wm::ss2d::make_geom(B, h, w)).total;
}

extern "C" int wm_ss2d_dirs_fwd(const float *x, const float *x_proj_weight,
                                const float *dt_projs_weight, const float *dt_projs_bias,
                                const float *A_logs, const float *Ds, void *workspace,
                                size_t workspace_bytes, int64_t B, int64_t h, int64_t w,
                                wm_stream_t stream)
{
    using namespace wm;
    WM_REQUIRE(B >= 0 && h >= 0 && w >= 0, "wm_ss2d_dirs_fwd: negative size");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    return ss2d::run_dirs(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, workspace,
                          workspace_bytes, B, h, w, (cudaStream_t)stream, "wm_ss2d_dirs_fwd");
}

extern "C" int wm_ss2d_core_fwd(const float *x, const float *x_proj_weight,
                                const float *dt_projs_weight, const float *dt_projs_bias,
                                const float *A_logs, const float *Ds, float *y, void *workspace,
                                size_t workspace_bytes, int64_t B, int64_t h, int64_t w,
                                wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::ss2d;
    WM_REQUIRE(B >= 0 && h >= 0 && w >= 0, "wm_ss2d_core_fwd: negative size");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(y != nullptr, "wm_ss2d_core_fwd: null pointer");
    const int rc = run_dirs(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, workspace,
                            workspace_bytes, B, h, w, (cudaStream_t)stream, "wm_ss2d_core_fwd");
    if (rc != WM_OK) return rc;
    const int64_t n = B * kD * h * w;
    const int vec = (n % 4 == 0 && aligned16(y)) ? 1 : 0;
    const int64_t want = ((vec ? n / 4 : n) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    ss2d_combine_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        y, static_cast<const float *>(workspace), n, vec);
    WM_LAUNCH_OK("ss2d combine");
    return WM_OK;
}
