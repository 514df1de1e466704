// Backward of the SS2D core (SS2D.forward_core + the 4-way sum) for sm_100a -- the gradient the
// reference obtains from mamba_ssm's selective_scan_fn autograd plus autograd through the einsum
// projections and the cross-scan / cross-merge (wavemamba_arch.py:446-478,490; training step
// basicsr/models/femasr_model.py:157-185).
//
// Per direction k, channel d, state n, with dt = softplus(pre), a = exp(dt A):
//   forward   h_l = a_l h_{l-1} + dt_l u_l B_l ,  y_l = sum_n C_l h_l + D u_l
//   backward  g_l = C_l gy_l + a_{l+1} g_{l+1}                      (reverse scan, g = dL/dh_l)
//             dC_l[n] = sum_d gy h ;  dB_l[n] = sum_d g dt u
//             d dt_l[d] = sum_n g (A (h_l - dt u B) + u B) ;  du_l[d] = sum_n g dt B + D gy
//             dA[d,n] = sum_l g dt (h_l - dt u B) ;  dD[d] = sum_l gy u
//   then through softplus (d pre = d dt (1 - exp(-dt))), dt_proj, x_proj and the index maps.
//
// Same chunk geometry as the forward (ss2d.cu): chunks are decoupled with the carry scheme in BOTH
// directions of time.  Call sequence:
//   forward pass 1 + carry            -> true initial state of every chunk           (ss2d.cu)
//   reverse pass 1 (MODE 0) + carry   -> true "incoming" q = a g of every chunk (aggregates stored at
//                                        the mirrored chunk index, so the forward carry kernel is reused)
//   per direction: checkpoint pass    -> the state after every step in a scratch buffer (ss2d.cu)
//                  main pass (MODE 1) -> reverse scan, all position-wise products, the transposed
//                                        projections, dx accumulated over the directions, per-CTA partial
//                                        sums of the weight gradients
//                  reduce             -> fixed-order sum of the partials (fp64): deterministic
// Correctness first: one CTA per SM, scalar fp32 arithmetic, about 4x the forward's time.
#define WM_SS2D_SEQ 4
#define WM_SS2D_TP 16
#include "ss2d_common.cuh"

namespace wm {
namespace ss2d {

constexpr int kGS = kXS;     // gs row stride [channel][position] (gy tile, later the dx tile)
constexpr int kDT = 65;      // ddt / dus row stride [position][channel]
constexpr int kB_Xs = 0;
constexpr int kB_Gs = kB_Xs + kD * kXS;
constexpr int kB_Pj = kB_Gs + kD * kGS;
constexpr int kB_Dd = kB_Pj + kPos * kPJ;
constexpr int kB_Ddt = kB_Dd + kPos * kDD;
constexpr int kB_Dus = kB_Ddt + kPos * kDT;
constexpr int kB_Dpj = kB_Dus + kPos * kDT;       // [position][dB16 | dC16 | dr2 | pad2]
constexpr int kB_Wx = kB_Dpj + kPos * kPJ;        // x_proj_weight[k] (34, 64)
constexpr int kB_Cst = kB_Wx + kProj * kD;        // dt_w col 0 | dt_w col 1 | dt bias | D   (4 x 64)
constexpr int kB_Drp = kB_Cst + 4 * kD;           // partial dr: [4 groups][64 positions][2]
constexpr int kBwdFloats = kB_Drp + 4 * kPos * 2;
constexpr size_t kBwdSmem = sizeof(float) * kBwdFloats;   // ~134 KB: one CTA per SM

// per-CTA partial sums of the parameter gradients of one direction
constexpr int kP_Wx = 0;                          // (34, 64)
constexpr int kP_Wdt = kP_Wx + kProj * kD;        // (64, 2)
constexpr int kP_Bias = kP_Wdt + 2 * kD;          // (64)
constexpr int kP_A = kP_Bias + kD;                // (64, 16)  d A_logs
constexpr int kP_D = kP_A + kChains;              // (64)
constexpr int kPart = kP_D + kD;                  // 3456 floats

struct BwdParams {
    const float *x, *gy;                 // (B,64,L) input map and upstream gradient (pixel order)
    const float *x_proj_w, *dt_w, *dt_b, *A_logs, *Ds;
    const float *aggH;                   // forward: true initial state per chunk (after the carry)
    float *aggP2, *aggQ;                 // reverse aggregates at the mirrored chunk index
    const float *hbuf;                   // checkpoint rows of the current direction
    float *dx;                           // (B,64,L)
    float *part;                         // [B * ctas][kPart]
    int accumulate;                      // dx += (directions after the first)
};

// pj column of output row o of x_proj_weight[k] = [dt(2) | B(16) | C(16)]
__device__ __forceinline__ int pj_col(int o) { return o < 2 ? 32 + o : o - 2; }

__device__ __forceinline__ void store_tile_acc(const Geom &g, const TileGeom &tg, const ChunkMap &cm, int ti,
                                               float *__restrict__ ob, const float *gs, bool accumulate)
{
    if (tile_is_vec(g, tg, ti)) {
        float *dst = ob + (int64_t)cm.d0 * g.L + cm.goff + (int64_t)ti * cm.gstep;
        const float *src = gs + cm.d0 * kGS + (cm.ys_idx - cm.d0 * kYS);   // same p0 as the xs chunk
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 v = *reinterpret_cast<const float4 *>(src + j * 16 * kGS);
            float4 *o4 = reinterpret_cast<float4 *>(dst + (int64_t)j * 16 * g.L);
            if (accumulate) { const float4 old = *o4; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
            *o4 = v;
        }
    } else {
        const int tid = threadIdx.x;
#pragma unroll 1
        for (int r = 0; r < kD * kPos / kThreads; ++r) {
            const int idx = tid + r * kThreads;
            const int d = idx >> 6, s = (idx & 63) / kTP, e = (idx & 63) % kTP;
            const int t = ti * kTP + e;
            if (t < strand_len(g, tg, s)) {
                float *o = ob + (int64_t)d * g.L + strand_elem(g, tg, s, t);
                const float v = gs[d * kGS + tile_pos(tg, s, e)];
                *o = accumulate ? *o + v : v;
            }
        }
    }
}

// MODE 0: reverse pass 1 (chunk aggregates of q from q_in = 0).  MODE 1: main backward pass.
template <int MODE, int DP>
__device__ __forceinline__ void run_bwd_cta(const BwdParams &prm, const Geom &g, const TileGeom &tg,
                                            float *smem, int b)
{
    float *xs = smem + kB_Xs, *gs = smem + kB_Gs, *pj = smem + kB_Pj, *dd = smem + kB_Dd;
    float *ddt = smem + kB_Ddt, *dus = smem + kB_Dus, *dpj = smem + kB_Dpj, *wx = smem + kB_Wx;
    float *cst = smem + kB_Cst, *drp = smem + kB_Drp;

    const int tid = threadIdx.x, lane = tid & 31;
    const int k = tg.k;
    const int ntiles = (tg.maxlen + kTP - 1) / kTP;
    const float *xb = prm.x + (int64_t)b * kD * g.L;
    const float *gyb = prm.gy + (int64_t)b * kD * g.L;
    const ChunkMap cmx = make_chunk_map(g, tg, xs);
    const ChunkMap cmg = make_chunk_map(g, tg, gs);

    for (int i = tid; i < kProj * kD; i += kThreads) wx[i] = __ldg(prm.x_proj_w + (int64_t)k * kProj * kD + i);
    if (tid < kD) {
        const int chn = k * kD + tid;
        cst[tid] = __ldg(prm.dt_w + chn * 2 + 0);
        cst[kD + tid] = __ldg(prm.dt_w + chn * 2 + 1);
        cst[2 * kD + tid] = __ldg(prm.dt_b + chn);
        cst[3 * kD + tid] = __ldg(prm.Ds + chn);
    }

    // ---- scan-thread identity (as the forward): (strand, channel pair, state half) -------------
    const int s = tid >> 6, cp = (tid & 63) >> 1;
    const int half = tid & 1, hoff = half * 8;
    float A2[2][8];      // A log2(e)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            A2[c][j] = -expf(__ldg(prm.A_logs + (int64_t)(k * kD + 2 * cp + c) * kN + hoff + j)) *
                       1.4426950408889634f;
    const int my_len = strand_len(g, tg, s);
    const int my_chunk = tg.col ? (tg.chunk0 + s) * g.ncolseg + tg.seg : tg.chunk0 + s;
    const int nch = dir_chunks(g, k);
    const int64_t chain0 = (int64_t)(2 * cp) * kN + hoff;
    const int64_t agg_mirror =
        (((int64_t)b * kK + k) * g.max_chunks + (nch - 1 - my_chunk)) * kChains + chain0;
    const float *hck = MODE == 1 ? prm.hbuf + (((int64_t)b * nch + my_chunk) * dir_chunk_len(g, k)) * kChains + chain0
                                 : nullptr;

    float q[2][8];
    double dA[2][8];     // sums over every step of the chunk: fp64 (long sums with cancellation)
    double dDacc = 0.0;
    double sum_dt[2] = {0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) { q[c][j] = 0.0f; dA[c][j] = 0.0; }
    // state after step t of my chunk (checkpoint row t); t = -1: the chunk's true initial state
    const float *h_init = prm.aggH + (((int64_t)b * kK + k) * g.max_chunks + my_chunk) * kChains + chain0;
    auto load_state = [&](int t, float (&dst)[2][8]) {
        if (t < 0 && my_chunk == 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[c][j] = 0.0f;
            return;
        }
        const float *hp = t < 0 ? h_init : hck + (int64_t)t * kChains;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float4 f0 = *reinterpret_cast<const float4 *>(hp + c * kN);
            const float4 f1 = *reinterpret_cast<const float4 *>(hp + c * kN + 4);
            dst[c][0] = f0.x; dst[c][1] = f0.y; dst[c][2] = f0.z; dst[c][3] = f0.w;
            dst[c][4] = f1.x; dst[c][5] = f1.y; dst[c][6] = f1.z; dst[c][7] = f1.w;
        }
    };
    if (MODE == 1 && my_len > 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float4 *qp = reinterpret_cast<const float4 *>(prm.aggQ + agg_mirror + c * kN);
            const float4 f0 = qp[0], f1 = qp[1];
            q[c][0] = f0.x; q[c][1] = f0.y; q[c][2] = f0.z; q[c][3] = f0.w;
            q[c][4] = f1.x; q[c][5] = f1.y; q[c][6] = f1.z; q[c][7] = f1.w;
        }
    }

    const int p0 = tile_pos(tg, s, 0);
    const float *dd0 = dd + p0 * kDD + 4 * cp;
    const float *pj0 = pj + p0 * kPJ + hoff;

    // dense-phase identity: (position, group of 16 channels / outputs)
    const int dp = tid & 63, grp = tid >> 6;
    double accw[9];      // d x_proj_weight rows grp, grp+4, ... for input channel dp (fp32 per tile, fp64 across)
#pragma unroll
    for (int i = 0; i < 9; ++i) accw[i] = 0.0;
    double acc_misc = 0.0;   // grp 0: d dt_w[:,0], grp 1: d dt_w[:,1], grp 2: d dt_bias   (channel dp)

#pragma unroll 1
    for (int ti = ntiles - 1; ti >= 0; --ti) {
        __syncthreads();                       // previous tile fully consumed
        load_tile(g, tg, cmx, ti, xb, xs);
        load_tile(g, tg, cmg, ti, gyb, gs);
        if (MODE == 1) {
            for (int i = tid; i < kPos * kDT; i += kThreads) { ddt[i] = 0.0f; dus[i] = 0.0f; }
            for (int i = tid; i < kPos * kPJ; i += kThreads) dpj[i] = 0.0f;
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- projections (FMA pipe, plain fp32): pj[p] = W_k x[:,p] -----------------------------
#pragma unroll 1
        for (int o = grp; o < kProj; o += 4) {
            if (MODE == 0 && o >= 2 && o < 18) continue;       // B is not needed for the q aggregates
            const float *wr = wx + o * kD;
            float acc = 0.0f;
#pragma unroll 16
            for (int d = 0; d < kD; ++d) acc = fmaf(wr[d], xs[d * kXS + dp], acc);
            pj[dp * kPJ + pj_col(o)] = acc;
        }
        __syncthreads();
        // ---- delta phase: dd[p][d] = (dt, u) ----------------------------------------------------
        {
            const float r0 = pj[dp * kPJ + 32], r1 = pj[dp * kPJ + 33];
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int d = grp * 16 + i;
                const float pre = fmaf(cst[kD + d], r1, cst[d] * r0) + cst[2 * kD + d];
                *reinterpret_cast<float2 *>(dd + dp * kDD + 2 * d) = make_float2(softplus_fast(pre), xs[d * kXS + dp]);
            }
        }
        __syncthreads();

        // ---- reverse recurrence over the valid steps of this tile --------------------------------
        {
            const int nvalid = min(my_len - ti * kTP, kTP);       // warp-uniform
            float hcur[2][8], hprev[2][8];     // states after steps t and t-1 of the step being processed
            if (MODE == 1 && nvalid > 0) {
                load_state(ti * kTP + nvalid - 1, hcur);
                load_state(ti * kTP + nvalid - 2, hprev);
            }
#pragma unroll 1
            for (int e = nvalid - 1; e >= 0; --e) {
                const int p = p0 + e * DP;
                const float4 dv = *reinterpret_cast<const float4 *>(dd0 + e * DP * kDD);
                const float *pjp = pj0 + e * DP * kPJ;
                float Bv[8], Cv[8];
                {
                    const float4 c0 = *reinterpret_cast<const float4 *>(pjp + 16);
                    const float4 c1 = *reinterpret_cast<const float4 *>(pjp + 20);
                    Cv[0] = c0.x; Cv[1] = c0.y; Cv[2] = c0.z; Cv[3] = c0.w;
                    Cv[4] = c1.x; Cv[5] = c1.y; Cv[6] = c1.z; Cv[7] = c1.w;
                }
                const float dtv[2] = {dv.x, dv.z}, uv[2] = {dv.y, dv.w};
                const float gyv[2] = {gs[(2 * cp) * kGS + p], gs[(2 * cp + 1) * kGS + p]};
                if (MODE == 0) {
                    sum_dt[0] += (double)dtv[0];
                    sum_dt[1] += (double)dtv[1];
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float a = ex2_approx(dtv[c] * A2[c][j]);
                            q[c][j] = a * fmaf(Cv[j], gyv[c], q[c][j]);
                        }
                    continue;
                }
                {
                    const float4 b0 = *reinterpret_cast<const float4 *>(pjp);
                    const float4 b1 = *reinterpret_cast<const float4 *>(pjp + 4);
                    Bv[0] = b0.x; Bv[1] = b0.y; Bv[2] = b0.z; Bv[3] = b0.w;
                    Bv[4] = b1.x; Bv[5] = b1.y; Bv[6] = b1.z; Bv[7] = b1.w;
                }
                float hpp[2][8];                // prefetch: the state two steps back (next step's hprev)
                if (e > 0) load_state(ti * kTP + e - 2, hpp);
                float v[16];                    // dB[0..7] | dC[0..7] partial sums over my 2 channels
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.0f;
                float sdt[2], sdu[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float dtu = dtv[c] * uv[c];
                    float s_dt = 0.0f, s_du = 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a = ex2_approx(dtv[c] * A2[c][j]);
                        const float gq = fmaf(Cv[j], gyv[c], q[c][j]);                 // g_l
                        const float hm = a * hprev[c][j];                              // a_l h_{l-1}
                        const float An = A2[c][j] * 0.6931471805599453f;               // A
                        s_dt = fmaf(gq, fmaf(An, hm, uv[c] * Bv[j]), s_dt);
                        s_du = fmaf(gq, Bv[j], s_du);
                        dA[c][j] += (double)(gq * dtv[c]) * (double)hm;
                        v[j] = fmaf(gq, dtu, v[j]);
                        v[8 + j] = fmaf(gyv[c], hcur[c][j], v[8 + j]);
                        q[c][j] = a * gq;
                    }
                    sdt[c] = s_dt;
                    sdu[c] = s_du * dtv[c];
                }
                // sum over the two state halves (partner lane), then lane `half` owns channel 2cp+half
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    sdt[c] += __shfl_xor_sync(0xffffffffu, sdt[c], 1);
                    sdu[c] += __shfl_xor_sync(0xffffffffu, sdu[c], 1);
                }
                {
                    const int c = half, d = 2 * cp + c;
                    ddt[p * kDT + d] = c ? sdt[1] : sdt[0];
                    const float gyc = c ? gyv[1] : gyv[0], uc = c ? uv[1] : uv[0];
                    dus[p * kDT + d] = fmaf(cst[3 * kD + d], gyc, c ? sdu[1] : sdu[0]);
                    dDacc += (double)gyc * (double)uc;
                }
                // dB / dC: reduce-scatter over the 16 channel-pair lanes of this warp with my state half
                // (lane bits 1..4): 8 + 4 + 2 + 1 shuffles, every lane ends with one finished value
                float w8[8], w4[4], w2[2], w1;
                {
                    const bool up = (lane & 16) != 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float send = up ? v[i] : v[8 + i], keep = up ? v[8 + i] : v[i];
                        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                }
                {
                    const bool up = (lane & 8) != 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float send = up ? w8[i] : w8[4 + i], keep = up ? w8[4 + i] : w8[i];
                        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                }
                {
                    const bool up = (lane & 4) != 0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = up ? w4[i] : w4[2 + i], keep = up ? w4[2 + i] : w4[i];
                        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                }
                {
                    const bool up = (lane & 2) != 0;
                    const float send = up ? w2[0] : w2[1], keep = up ? w2[1] : w2[0];
                    w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
                const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                // idx < 8: dB[hoff + idx], else dC[hoff + idx - 8]; the strand's other warp adds the
                // other 16 channel pairs (two addends on a zeroed cell: order-independent)
                atomicAdd(dpj + p * kPJ + (idx < 8 ? hoff + idx : 16 + hoff + (idx - 8)), w1);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 8; ++j) { hcur[c][j] = hprev[c][j]; hprev[c][j] = hpp[c][j]; }
            }
        }
        if (MODE == 0) continue;
        __syncthreads();

        // ---- through softplus and dt_proj --------------------------------------------------------
        {
            float r0p = 0.0f, r1p = 0.0f;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int d = grp * 16 + i;
                const float dt = dd[dp * kDD + 2 * d];
                // softplus'(pre) = sigmoid(pre) = 1 - exp(-dt); expm1 keeps the tiny-dt end accurate
                const float dpre = ddt[dp * kDT + d] * -expm1f(-dt);
                ddt[dp * kDT + d] = dpre;
                r0p = fmaf(cst[d], dpre, r0p);
                r1p = fmaf(cst[kD + d], dpre, r1p);
            }
            drp[(grp * kPos + dp) * 2 + 0] = r0p;
            drp[(grp * kPos + dp) * 2 + 1] = r1p;
        }
        __syncthreads();
        if (tid < kPos) {
            float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) { r0 += drp[(gq * kPos + tid) * 2]; r1 += drp[(gq * kPos + tid) * 2 + 1]; }
            dpj[tid * kPJ + 32] = r0;
            dpj[tid * kPJ + 33] = r1;
        }
        __syncthreads();
        // ---- dx tile = du + W_k^T dproj  (into gs; gy is dead) and the weight-gradient sums ------
        {
            float dpv[kProj];
#pragma unroll
            for (int o = 0; o < kProj; ++o) dpv[o] = dpj[dp * kPJ + pj_col(o)];
#pragma unroll 1
            for (int i = 0; i < 16; ++i) {
                const int d = grp * 16 + i;
                float acc = dus[dp * kDT + d];
#pragma unroll
                for (int o = 0; o < kProj; ++o) acc = fmaf(wx[o * kD + d], dpv[o], acc);
                gs[d * kGS + dp] = acc;
            }
        }
        {
            // thread = (input channel dp, output rows grp, grp+4, ...): sums over the 64 positions
            float tw[9], tm = 0.0f;
#pragma unroll
            for (int i = 0; i < 9; ++i) tw[i] = 0.0f;
#pragma unroll 1
            for (int p = 0; p < kPos; ++p) {
                const float xv = xs[dp * kXS + p];
                const float *dr = dpj + p * kPJ;
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int o = grp + 4 * i;
                    if (o < kProj) tw[i] = fmaf(dr[pj_col(o)], xv, tw[i]);
                }
                const float dpre = ddt[p * kDT + dp];
                if (grp == 0) tm = fmaf(dpre, pj[p * kPJ + 32], tm);
                else if (grp == 1) tm = fmaf(dpre, pj[p * kPJ + 33], tm);
                else if (grp == 2) tm += dpre;
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) accw[i] += (double)tw[i];
            acc_misc += (double)tm;
        }
        __syncthreads();
        store_tile_acc(g, tg, cmg, ti, prm.dx + (int64_t)b * kD * g.L, gs, prm.accumulate != 0);
    }

    if (MODE == 0) {
        if (my_len > 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float pv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pv[j] = ex2_approx((float)((double)A2[c][j] * sum_dt[c]));
                float4 *pp4 = reinterpret_cast<float4 *>(prm.aggP2 + agg_mirror + c * kN);
                float4 *qp4 = reinterpret_cast<float4 *>(prm.aggQ + agg_mirror + c * kN);
                pp4[0] = make_float4(pv[0], pv[1], pv[2], pv[3]);
                pp4[1] = make_float4(pv[4], pv[5], pv[6], pv[7]);
                qp4[0] = make_float4(q[c][0], q[c][1], q[c][2], q[c][3]);
                qp4[1] = make_float4(q[c][4], q[c][5], q[c][6], q[c][7]);
            }
        }
        return;
    }

    // ---- per-CTA partial sums -> global (fixed layout, summed by the reduce kernel) ---------------
    __syncthreads();
    float *stage = smem;                               // 4 strands x (1024 dA + 64 dD), over xs/gs
    {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                stage[s * (kChains + kD) + (2 * cp + c) * kN + hoff + j] =
                    (float)(dA[c][j] * (double)(A2[c][j] * 0.6931471805599453f));   // d A_log = dA * A
        // lane `half` accumulated dD of channel 2cp+half
        stage[s * (kChains + kD) + kChains + 2 * cp + half] = (float)dDacc;
    }
    __syncthreads();
    float *part = prm.part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * kPart;
    for (int i = tid; i < kChains + kD; i += kThreads) {
        const float v = (stage[i] + stage[(kChains + kD) + i]) + (stage[2 * (kChains + kD) + i] + stage[3 * (kChains + kD) + i]);
        part[kP_A + i] = v;                            // kP_D follows kP_A directly
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int o = grp + 4 * i;
        if (o < kProj) part[kP_Wx + o * kD + dp] = (float)accw[i];
    }
    if (grp == 0) part[kP_Wdt + dp * 2 + 0] = (float)acc_misc;
    else if (grp == 1) part[kP_Wdt + dp * 2 + 1] = (float)acc_misc;
    else if (grp == 2) part[kP_Bias + dp] = (float)acc_misc;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
ss2d_bwd_kernel(const BwdParams prm, const Geom g, const Launch ln)
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
    const int idx = blockIdx.x - begin;
    if (tg.col) {
        tg.seg = idx % g.ncolseg;
        tg.chunk0 = (idx / g.ncolseg) * kSeq;
        tg.t0 = tg.seg * g.col_seg;
        tg.maxlen = min(g.col_seg, g.h - tg.t0);
    } else {
        tg.seg = 0;
        tg.t0 = 0;
        tg.chunk0 = idx * kSeq;
        tg.maxlen = g.row_T;
    }
    const int b = blockIdx.y;
    if (tg.col) run_bwd_cta<MODE, kSeq>(prm, g, tg, smem, b);
    else if (tg.fwd) run_bwd_cta<MODE, 1>(prm, g, tg, smem, b);
    else run_bwd_cta<MODE, -1>(prm, g, tg, smem, b);
}

// Fixed-order (fp64) sum of the per-CTA partials of direction k into the parameter gradients.
__global__ void __launch_bounds__(256)
ss2d_bwd_reduce_kernel(const float *__restrict__ part, int nparts, int k, float *__restrict__ d_xproj,
                       float *__restrict__ d_dtw, float *__restrict__ d_dtb, float *__restrict__ d_alogs,
                       float *__restrict__ d_ds)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= kPart) return;
    double acc = 0.0;
    for (int c = 0; c < nparts; ++c) acc += (double)part[(int64_t)c * kPart + i];
    const float v = (float)acc;
    if (i < kP_Wdt) d_xproj[(int64_t)k * kProj * kD + i] = v;
    else if (i < kP_Bias) d_dtw[(int64_t)k * 2 * kD + (i - kP_Wdt)] = v;
    else if (i < kP_A) d_dtb[(int64_t)k * kD + (i - kP_Bias)] = v;
    else if (i < kP_D) d_alogs[(int64_t)k * kChains + (i - kP_A)] = v;
    else d_ds[(int64_t)k * kD + (i - kP_D)] = v;
}

// CTA ranges of the backward's own tile layout (kSeq = 4 strands per CTA) over the shared chunk plan
static Launch make_launch_bwd(const Geom &g, std::initializer_list<int> dirs)
{
    Launch ln;
    ln.ndirs = 0;
    int acc = 0;
    const int row_ctas = (g.row_chunks + kSeq - 1) / kSeq;
    const int col_ctas = ((g.w + kSeq - 1) / kSeq) * g.ncolseg;
    for (int k : dirs) {
        ln.dir[ln.ndirs] = k;
        ln.cta_begin[ln.ndirs] = acc;
        acc += (k & 1) ? col_ctas : row_ctas;
        ++ln.ndirs;
    }
    for (int i = ln.ndirs; i < 4; ++i) { ln.dir[i] = 0; ln.cta_begin[i] = acc; }
    ln.cta_begin[4] = acc;
    return ln;
}

struct BwdWorkspace {
    int64_t agg, hbuf_off, part_off, total;
};

static BwdWorkspace plan_bwd(const Geom &g)
{
    BwdWorkspace ws;
    auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
    ws.agg = up((int64_t)g.B * kK * g.max_chunks * kChains * 4);          // x4: P, H, P2, Q
    int64_t rows = 0, ctas = 0;
    for (int k = 0; k < 2; ++k) {
        const int64_t r = (int64_t)dir_chunks(g, k) * dir_chunk_len(g, k);
        rows = r > rows ? r : rows;
        const int64_t c = (k & 1) ? (int64_t)((g.w + kSeq - 1) / kSeq) * g.ncolseg : (g.row_chunks + kSeq - 1) / kSeq;
        ctas = c > ctas ? c : ctas;
    }
    ws.hbuf_off = 4 * ws.agg;
    ws.part_off = ws.hbuf_off + up((int64_t)g.B * rows * kChains * 4);
    ws.total = ws.part_off + up((int64_t)g.B * ctas * kPart * 4);
    return ws;
}

}  // namespace ss2d
}  // namespace wm

using namespace wm;
using namespace wm::ss2d;

extern "C" size_t wm_ss2d_core_bwd_workspace_bytes(int64_t B, int64_t h, int64_t w)
{
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)plan_bwd(make_geom(B, h, w)).total;
}

extern "C" int wm_ss2d_core_bwd(const float *x, const float *x_proj_weight, const float *dt_projs_weight,
                                const float *dt_projs_bias, const float *A_logs, const float *Ds,
                                const float *grad_y, float *grad_x, float *grad_x_proj_weight,
                                float *grad_dt_projs_weight, float *grad_dt_projs_bias, float *grad_A_logs,
                                float *grad_Ds, void *workspace, size_t workspace_bytes, int64_t B,
                                int64_t h, int64_t w, wm_stream_t stream)
{
    WM_REQUIRE(B >= 0 && h >= 0 && w >= 0, "wm_ss2d_core_bwd: negative size");
    if (B == 0 || h == 0 || w == 0) return WM_OK;
    WM_REQUIRE(x && x_proj_weight && dt_projs_weight && dt_projs_bias && A_logs && Ds && grad_y && grad_x &&
                   grad_x_proj_weight && grad_dt_projs_weight && grad_dt_projs_bias && grad_A_logs && grad_Ds,
               "wm_ss2d_core_bwd: null pointer");
    WM_REQUIRE(B <= 65535 && h * w < ((int64_t)1 << 31), "wm_ss2d_core_bwd: size out of range");
    WM_REQUIRE(aligned16(x) && aligned16(grad_y) && aligned16(grad_x), "wm_ss2d_core_bwd: tensors must be 16-byte aligned");
    const Geom g = make_geom(B, h, w);
    const BwdWorkspace ws = plan_bwd(g);
    WM_REQUIRE(workspace && workspace_bytes >= (size_t)ws.total, "wm_ss2d_core_bwd: workspace too small (%zu < %lld bytes)",
               workspace_bytes, (long long)ws.total);
    WM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "wm_ss2d_core_bwd: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    char *wsb = static_cast<char *>(workspace);
    float *aggP = reinterpret_cast<float *>(wsb), *aggH = reinterpret_cast<float *>(wsb + ws.agg);
    float *aggP2 = reinterpret_cast<float *>(wsb + 2 * ws.agg), *aggQ = reinterpret_cast<float *>(wsb + 3 * ws.agg);
    float *hbuf = reinterpret_cast<float *>(wsb + ws.hbuf_off);
    float *part = reinterpret_cast<float *>(wsb + ws.part_off);

    Params fp;
    fp.x = x; fp.x_proj_w = x_proj_weight; fp.dt_w = dt_projs_weight; fp.dt_b = dt_projs_bias;
    fp.A_logs = A_logs; fp.Ds = Ds; fp.planes = nullptr; fp.aggP = aggP; fp.aggH = aggH; fp.dbg = nullptr;
    fp.hbuf = hbuf;
    fp.tiles = nullptr; fp.tile_stride = 0;
    BwdParams bp;
    bp.x = x; bp.gy = grad_y; bp.x_proj_w = x_proj_weight; bp.dt_w = dt_projs_weight; bp.dt_b = dt_projs_bias;
    bp.A_logs = A_logs; bp.Ds = Ds; bp.aggH = aggH; bp.aggP2 = aggP2; bp.aggQ = aggQ; bp.hbuf = hbuf;
    bp.dx = grad_x; bp.part = part; bp.accumulate = 0;

    WM_CUDA_OK(cudaFuncSetAttribute(ss2d_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    WM_CUDA_OK(cudaFuncSetAttribute(ss2d_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));

    const Launch all = g.cols_first ? make_launch(g, {1, 3, 0, 2}) : make_launch(g, {0, 1, 2, 3});
    int rc = launch_pass(0, fp, g, all, s);                 // chunk aggregates of h
    if (rc != WM_OK) return rc;
    rc = launch_carry(aggP, aggH, g, s);                    // aggH <- true initial states
    if (rc != WM_OK) return rc;
    {
        const Launch ball = make_launch_bwd(g, {0, 1, 2, 3});
        dim3 grid(ball.cta_begin[4], (unsigned)B);
        ss2d_bwd_kernel<0><<<grid, kThreads, kBwdSmem, s>>>(bp, g, ball);   // chunk aggregates of q (mirrored)
        WM_LAUNCH_OK("ss2d backward pass 1");
    }
    rc = launch_carry(aggP2, aggQ, g, s);                   // aggQ <- true incoming q
    if (rc != WM_OK) return rc;
    for (int k = 0; k < kK; ++k) {
        const Launch one = make_launch(g, {k});
        rc = launch_pass(2, fp, g, one, s);                 // states after every step of direction k
        if (rc != WM_OK) return rc;
        bp.accumulate = k > 0 ? 1 : 0;
        const Launch bone = make_launch_bwd(g, {k});
        dim3 grid(bone.cta_begin[4], (unsigned)B);
        ss2d_bwd_kernel<1><<<grid, kThreads, kBwdSmem, s>>>(bp, g, bone);
        WM_LAUNCH_OK("ss2d backward main pass");
        ss2d_bwd_reduce_kernel<<<(kPart + 255) / 256, 256, 0, s>>>(part, (int)(bone.cta_begin[4] * B), k,
                                                                   grad_x_proj_weight, grad_dt_projs_weight,
                                                                   grad_dt_projs_bias, grad_A_logs, grad_Ds);
        WM_LAUNCH_OK("ss2d backward reduce");
    }
    return WM_OK;
}
