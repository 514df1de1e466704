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
//   One CTA (256 threads) owns NSEQ=4 neighbouring chunks ("strands") of one direction and all
//   64 channels x 16 states of them: thread = (strand s, channel d) keeps h[16] in registers
//   and walks its strand sequentially, TP=16 steps per tile.  For column directions the four
//   strands are four adjacent columns, so a tile row is one 16-byte segment per channel.
//   Per tile: stage x (64 ch x 64 positions) in shared memory -> per-position projection
//   (34x64 mat-vec -> dt_low(2), B(16), C(16)) into shared memory -> 16 recurrence steps.
//
// Chunks are made independent with a three-phase carry scheme (the recurrence is linear):
//   pass 1  every chunk from h=0: aggregate (P = prod a = exp2(A2*sum dt), H = local end state)
//   carry   per (b,k,d,n): h_in[c] = P[c-1]*h_in[c-1] + H[c-1]          (tiny, sequential in c)
//   pass 2  every chunk again from its true h_in, emitting y.
// Merge order: launch A writes y <- dir0 and tmp <- dir1, launch B adds dir2 into y and dir3
// into tmp (same thread, same element: deterministic), then y += tmp.
//
// Roofline: not HBM-bound.  Each state update needs one MUFU ex2 and ~4 FMA-pipe ops and every
// exp is evaluated twice (pass 1 and pass 2); see DESIGN.md section 4.
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
constexpr int kXS = 68;      // smem row stride of the x tile  [position][channel]
constexpr int kPJ = 36;      // smem row stride of projections [position][B16|C16|dt2|pad2]
constexpr int kWT = 40;      // smem row stride of weights     [channel][B16|C16|dt2|pad6]
constexpr int kThreads = kSeq * kD;  // 256
constexpr int kChains = kD * kN;     // 1024 (d,n) chains per direction

struct Geom {
    int B, h, w;
    int64_t L;
    int row_T;       // steps per row chunk (multiple of kTP)
    int row_chunks;  // ceil(L / row_T)
    int row_ctas;    // ceil(row_chunks / kSeq)
    int col_ctas;    // ceil(w / kSeq)
    int max_chunks;  // max(row_chunks, w): chunk stride of the aggregate arrays
    int vec_rows;    // 1 when row tiles may use 128-bit global accesses
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
    float *y;                  // (B,64,L)  dirs 0,2
    float *tmp;                // (B,64,L)  dirs 1,3
    float *aggP;               // (B,4,max_chunks,1024)
    float *aggH;               // (B,4,max_chunks,1024)  pass 1: local end state; after carry: h_in
};

__device__ __forceinline__ float softplus_ref(float v)
{
    // torch softplus (beta 1, threshold 20) == mamba's `x <= 20 ? log1pf(expf(x)) : x`
    return v > 20.0f ? v : log1pf(expf(v));
}

struct Strand {
    int64_t base;  // element offset (inside one channel plane) of step 0
    int64_t step;  // element stride per step: +-1 or +-w
    int len;       // number of steps (0: strand does not exist)
    int chunk;     // chunk index in sequence order
};

__device__ __forceinline__ Strand make_strand(const Geom &g, int k, int q, int s)
{
    Strand st;
    st.chunk = q * kSeq + s;
    if ((k & 1) == 0) {  // row-major directions
        const int64_t start = (int64_t)st.chunk * g.row_T;
        int64_t rem = g.L - start;
        st.len = rem <= 0 ? 0 : (rem < g.row_T ? (int)rem : g.row_T);
        st.base = (k == 0) ? start : g.L - 1 - start;
        st.step = (k == 0) ? 1 : -1;
    } else {  // column-major directions: chunk = one column
        st.len = st.chunk < g.w ? g.h : 0;
        const int j = (k == 1) ? st.chunk : g.w - 1 - st.chunk;
        st.base = (k == 1) ? (int64_t)j : (int64_t)(g.h - 1) * g.w + j;
        st.step = (k == 1) ? (int64_t)g.w : -(int64_t)g.w;
    }
    return st;
}

__device__ __forceinline__ void load_tile(const float *__restrict__ plane, const Strand &st,
                                          int ti, bool vec_rows, float (&u)[kTP])
{
    const int t0 = ti * kTP;
    int nvalid = st.len - t0;
    nvalid = nvalid < 0 ? 0 : (nvalid > kTP ? kTP : nvalid);
    if (nvalid == kTP && vec_rows && st.step == 1) {
        const float4 *p = reinterpret_cast<const float4 *>(plane + st.base + t0);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float4 f = __ldg(p + v);
            u[4 * v + 0] = f.x; u[4 * v + 1] = f.y; u[4 * v + 2] = f.z; u[4 * v + 3] = f.w;
        }
    } else if (nvalid == kTP && vec_rows && st.step == -1) {
        const float4 *p = reinterpret_cast<const float4 *>(plane + st.base - t0 - (kTP - 1));
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float4 f = __ldg(p + v);  // ascending memory = descending step
            u[15 - 4 * v] = f.x; u[14 - 4 * v] = f.y; u[13 - 4 * v] = f.z; u[12 - 4 * v] = f.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < kTP; ++e)
            u[e] = e < nvalid ? __ldg(plane + st.base + (int64_t)(t0 + e) * st.step) : 0.0f;
    }
}

template <bool ACCUM>
__device__ __forceinline__ void store_tile(float *__restrict__ plane, const Strand &st, int ti,
                                           bool vec_rows, const float (&yv)[kTP])
{
    const int t0 = ti * kTP;
    int nvalid = st.len - t0;
    nvalid = nvalid < 0 ? 0 : (nvalid > kTP ? kTP : nvalid);
    if (nvalid == kTP && vec_rows && (st.step == 1 || st.step == -1)) {
        const bool fwd = st.step == 1;
        float4 *p = reinterpret_cast<float4 *>(fwd ? plane + st.base + t0
                                                   : plane + st.base - t0 - (kTP - 1));
        float4 old[4];
        if (ACCUM) {
#pragma unroll
            for (int v = 0; v < 4; ++v) old[v] = p[v];
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            float4 f;
            if (fwd) {
                f = make_float4(yv[4 * v], yv[4 * v + 1], yv[4 * v + 2], yv[4 * v + 3]);
            } else {
                f = make_float4(yv[15 - 4 * v], yv[14 - 4 * v], yv[13 - 4 * v], yv[12 - 4 * v]);
            }
            if (ACCUM) { f.x += old[v].x; f.y += old[v].y; f.z += old[v].z; f.w += old[v].w; }
            p[v] = f;
        }
    } else {
        float old[kTP];
        if (ACCUM) {
#pragma unroll
            for (int e = 0; e < kTP; ++e)
                old[e] = e < nvalid ? plane[st.base + (int64_t)(t0 + e) * st.step] : 0.0f;
        }
#pragma unroll
        for (int e = 0; e < kTP; ++e)
            if (e < nvalid)
                plane[st.base + (int64_t)(t0 + e) * st.step] = ACCUM ? old[e] + yv[e] : yv[e];
    }
}

// FINAL=false: pass 1 (aggregates).  FINAL=true: pass 2 (outputs); ACCUM adds into the
// destination instead of overwriting it.
template <bool FINAL, bool ACCUM>
__global__ void __launch_bounds__(kThreads, 2)
ss2d_pass_kernel(const Params prm, const Geom g, const Launch ln)
{
    __shared__ __align__(16) float xs[kPos * kXS];
    __shared__ __align__(16) float pj[kPos * kPJ];
    __shared__ __align__(16) float wt[kD * kWT];

    const int tid = threadIdx.x;
    const int s = tid >> 6, d = tid & 63;
    const int b = blockIdx.y;

    // which direction / which CTA inside that direction
    int slot = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < ln.ndirs && (int)blockIdx.x >= ln.cta_begin[i]) slot = i;
    const int k = ln.dir[slot];
    const int q = blockIdx.x - ln.cta_begin[slot];

    // ---- per-CTA weights: wt[d][0..15]=B rows, [16..31]=C rows, [32..33]=dt rows ------------
    // x_proj_weight[k] is (34,64) with rows [dt(2) | B(16) | C(16)]  (reference :454 split)
    for (int i = tid; i < kProj * kD; i += kThreads) {
        const int row = i / kD, col = i - row * kD;
        const int dst = row < 2 ? 32 + row : row - 2;
        wt[col * kWT + dst] = __ldg(prm.x_proj_w + (int64_t)k * kProj * kD + i);
    }

    // ---- per-thread constants ---------------------------------------------------------------
    const int ch = k * kD + d;
    float A2[kN];  // A * log2(e), A = -exp(A_log)   (reference :462)
#pragma unroll
    for (int n = 0; n < kN; ++n)
        A2[n] = -expf(__ldg(prm.A_logs + (int64_t)ch * kN + n)) * 1.4426950408889634f;
    const float dtw0 = __ldg(prm.dt_w + ch * 2 + 0), dtw1 = __ldg(prm.dt_w + ch * 2 + 1);
    const float dtb = __ldg(prm.dt_b + ch);
    const float skipD = __ldg(prm.Ds + ch);

    const Strand st = make_strand(g, k, q, s);
    const int maxlen = (k & 1) ? g.h : g.row_T;
    const int ntiles = (maxlen + kTP - 1) / kTP;
    const bool vec_rows = g.vec_rows != 0;

    const int64_t plane_off = ((int64_t)b * kD + d) * g.L;
    const float *xplane = prm.x + plane_off;
    const int64_t agg_off =
        (((int64_t)b * kK + k) * g.max_chunks + st.chunk) * kChains + (int64_t)d * kN;

    float hst[kN];
#pragma unroll
    for (int n = 0; n < kN; ++n) hst[n] = 0.0f;
    if (FINAL && st.len > 0 && st.chunk > 0) {
        const float4 *hp = reinterpret_cast<const float4 *>(prm.aggH + agg_off);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float4 f = hp[v];
            hst[4 * v] = f.x; hst[4 * v + 1] = f.y; hst[4 * v + 2] = f.z; hst[4 * v + 3] = f.w;
        }
    }
    double sum_dt = 0.0;

    float unext[kTP];
    load_tile(xplane, st, 0, vec_rows, unext);

    const int pp = tid & 63, og = tid >> 6;  // projection mapping: position, output octet

    for (int ti = 0; ti < ntiles; ++ti) {
        // stage this tile's x: xs[position][channel]
#pragma unroll
        for (int e = 0; e < kTP; ++e) xs[(s * kTP + e) * kXS + d] = unext[e];
        __syncthreads();
        if (ti + 1 < ntiles) load_tile(xplane, st, ti + 1, vec_rows, unext);

        // ---- projection: pj[p][c] = sum_d W[c][d] * x[d][p]            (reference :453) ----
        {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
            float accdt = 0.0f;
            const float *xrow = xs + pp * kXS;
#pragma unroll 4
            for (int d4 = 0; d4 < kD; d4 += 4) {
                const float4 xv = *reinterpret_cast<const float4 *>(xrow + d4);
                const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float *wr = wt + (d4 + j) * kWT;
                    const float4 w0 = *reinterpret_cast<const float4 *>(wr + og * 8);
                    const float4 w1 = *reinterpret_cast<const float4 *>(wr + og * 8 + 4);
                    acc[0] = fmaf(xe[j], w0.x, acc[0]); acc[1] = fmaf(xe[j], w0.y, acc[1]);
                    acc[2] = fmaf(xe[j], w0.z, acc[2]); acc[3] = fmaf(xe[j], w0.w, acc[3]);
                    acc[4] = fmaf(xe[j], w1.x, acc[4]); acc[5] = fmaf(xe[j], w1.y, acc[5]);
                    acc[6] = fmaf(xe[j], w1.z, acc[6]); acc[7] = fmaf(xe[j], w1.w, acc[7]);
                    if (og < 2) accdt = fmaf(xe[j], wr[32 + og], accdt);
                }
            }
            float *pr = pj + pp * kPJ;
            *reinterpret_cast<float4 *>(pr + og * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4 *>(pr + og * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            if (og < 2) pr[32 + og] = accdt;
        }
        __syncthreads();

        // ---- recurrence over the 16 steps of this tile                (reference :465-471) ----
        float yv[kTP];
        const int nvalid = st.len - ti * kTP;  // warp-uniform (a warp lies inside one strand)
#pragma unroll
        for (int e = 0; e < kTP; ++e) {
            yv[e] = 0.0f;
            if (e < nvalid) {
                const int p = s * kTP + e;
                const float *pr = pj + p * kPJ;
                const float u = xs[p * kXS + d];
                const float2 dlow = *reinterpret_cast<const float2 *>(pr + 32);
                // dt_proj (reference :455) then + bias, softplus (selective_scan_fn)
                const float dt = softplus_ref(fmaf(dtw1, dlow.y, dtw0 * dlow.x) + dtb);
                const float dtu = dt * u;
                if (!FINAL) sum_dt += (double)dt;
                float acc = 0.0f;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 bq = *reinterpret_cast<const float4 *>(pr + 4 * v);
                    const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
                    float cc[4] = {0.f, 0.f, 0.f, 0.f};
                    if (FINAL) {
                        const float4 cq = *reinterpret_cast<const float4 *>(pr + 16 + 4 * v);
                        cc[0] = cq.x; cc[1] = cq.y; cc[2] = cq.z; cc[3] = cq.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = 4 * v + j;
                        const float a = ex2_approx(dt * A2[n]);
                        hst[n] = fmaf(a, hst[n], dtu * bb[j]);
                        if (FINAL) acc = fmaf(hst[n], cc[j], acc);
                    }
                }
                if (FINAL) yv[e] = fmaf(skipD, u, acc);
            }
        }
        if (FINAL) {
            float *oplane = ((k & 1) ? prm.tmp : prm.y) + plane_off;
            store_tile<ACCUM>(oplane, st, ti, vec_rows, yv);
        }
        __syncthreads();  // xs / pj are rewritten by the next tile
    }

    if (!FINAL && st.len > 0) {
        float4 *pp4 = reinterpret_cast<float4 *>(prm.aggP + agg_off);
        float4 *hp4 = reinterpret_cast<float4 *>(prm.aggH + agg_off);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            float pv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                pv[j] = ex2_approx((float)((double)A2[4 * v + j] * sum_dt));
            pp4[v] = make_float4(pv[0], pv[1], pv[2], pv[3]);
            hp4[v] = make_float4(hst[4 * v], hst[4 * v + 1], hst[4 * v + 2], hst[4 * v + 3]);
        }
    }
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
    for (; c + 4 <= nchunks; c += 4) {
        float p[4], hv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            p[i] = P[(int64_t)(c + i) * kChains];
            hv[i] = H[(int64_t)(c + i) * kChains];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
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

__global__ void __launch_bounds__(256)
ss2d_combine_kernel(float *__restrict__ y, const float *__restrict__ tmp, int64_t n, int vec)
{
    const int64_t stride = (int64_t)gridDim.x * 256;
    if (vec) {
        const int64_t n4 = n >> 2;
        float4 *y4 = reinterpret_cast<float4 *>(y);
        const float4 *t4 = reinterpret_cast<const float4 *>(tmp);
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
            float4 a = y4[i];
            const float4 t = ld_stream4(reinterpret_cast<const float *>(t4 + i));
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            y4[i] = a;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) y[i] += tmp[i];
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
    return g;
}

struct Workspace {
    int64_t tmp_off, aggP_off, aggH_off, total;
};

Workspace plan_workspace(const Geom &g)
{
    Workspace ws;
    const int64_t plane_bytes = align_up((int64_t)g.B * kD * g.L * 4, 256);
    const int64_t agg_bytes = align_up((int64_t)g.B * kK * g.max_chunks * kChains * 4, 256);
    ws.tmp_off = 0;
    ws.aggP_off = plane_bytes;
    ws.aggH_off = plane_bytes + agg_bytes;
    ws.total = plane_bytes + 2 * agg_bytes;
    return ws;
}

}  // namespace ss2d
}  // namespace wm

extern "C" size_t wm_ss2d_core_workspace_bytes(int64_t B, int64_t h, int64_t w)
{
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)wm::ss2d::plan_workspace(wm::ss2d::make_geom(B, h, w)).total;
}

extern "C" int wm_ss2d_core_fwd(const float *x, const float *x_proj_weight,
                                const float *dt_projs_weight, const float *dt_projs_bias,
                                const float *A_logs, const float *Ds, float *y, void *workspace,
                                size_t workspace_bytes, int64_t B, int64_t h, int64_t w,
                                wm_stream_t stream)
{
    using namespace wm;
    using namespace wm::ss2d;
    WM_REQUIRE(x && x_proj_weight && dt_projs_weight && dt_projs_bias && A_logs && Ds && y,
               "wm_ss2d_core_fwd: null pointer");
    WM_REQUIRE(B >= 0 && h >= 0 && w >= 0, "wm_ss2d_core_fwd: negative size");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(B <= 65535, "wm_ss2d_core_fwd: batch %lld exceeds 65535", (long long)B);
    WM_REQUIRE(h * w < (int64_t)1 << 31, "wm_ss2d_core_fwd: h*w too large");
    WM_REQUIRE(aligned16(x) && aligned16(y) && aligned16(A_logs),
               "wm_ss2d_core_fwd: x, y and A_logs must be 16-byte aligned");
    const Geom g = make_geom(B, h, w);
    const Workspace ws = plan_workspace(g);
    WM_REQUIRE(workspace && workspace_bytes >= (size_t)ws.total,
               "wm_ss2d_core_fwd: workspace too small (%zu < %lld bytes)", workspace_bytes,
               (long long)ws.total);
    WM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "wm_ss2d_core_fwd: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    char *wsb = static_cast<char *>(workspace);

    Params prm;
    prm.x = x; prm.x_proj_w = x_proj_weight; prm.dt_w = dt_projs_weight; prm.dt_b = dt_projs_bias;
    prm.A_logs = A_logs; prm.Ds = Ds; prm.y = y;
    prm.tmp = reinterpret_cast<float *>(wsb + ws.tmp_off);
    prm.aggP = reinterpret_cast<float *>(wsb + ws.aggP_off);
    prm.aggH = reinterpret_cast<float *>(wsb + ws.aggH_off);

    auto make_launch = [&](std::initializer_list<int> dirs) {
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
    };

    {   // pass 1: all four directions
        const Launch ln = make_launch({0, 2, 1, 3});
        dim3 grid(ln.cta_begin[4], (unsigned)B);
        ss2d_pass_kernel<false, false><<<grid, kThreads, 0, s>>>(prm, g, ln);
        WM_LAUNCH_OK("ss2d pass 1");
    }
    {
        dim3 grid(kChains / 256, kK, (unsigned)B);
        ss2d_carry_kernel<<<grid, 256, 0, s>>>(prm.aggP, prm.aggH, g);
        WM_LAUNCH_OK("ss2d carry");
    }
    {   // pass 2A: dir 0 -> y, dir 1 -> tmp
        const Launch ln = make_launch({0, 1});
        dim3 grid(ln.cta_begin[4], (unsigned)B);
        ss2d_pass_kernel<true, false><<<grid, kThreads, 0, s>>>(prm, g, ln);
        WM_LAUNCH_OK("ss2d pass 2A");
    }
    {   // pass 2B: dir 2 += y, dir 3 += tmp
        const Launch ln = make_launch({2, 3});
        dim3 grid(ln.cta_begin[4], (unsigned)B);
        ss2d_pass_kernel<true, true><<<grid, kThreads, 0, s>>>(prm, g, ln);
        WM_LAUNCH_OK("ss2d pass 2B");
    }
    {
        const int64_t n = B * kD * g.L;
        const int vec = (n % 4 == 0) ? 1 : 0;
        const int64_t want = ((vec ? n / 4 : n) + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        ss2d_combine_kernel<<<(int)(want < cap ? want : cap), 256, 0, s>>>(y, prm.tmp, n, vec);
        WM_LAUNCH_OK("ss2d combine");
    }
    return WM_OK;
}
