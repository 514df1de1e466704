/*
 * wavemamba_b200 -- C ABI of the B200-native Wave-Mamba forward hot path.
 *
 * This is the drop-in boundary.  The reference has no native code;
 * its hot path is Python in basicsr/archs/wavemamba_arch.py plus one third-party CUDA
 * extension (mamba_ssm.selective_scan_fn).  Each entry point below replaces the Python
 * function (file:line in /root/reference) named in its comment.  The reference-side
 * binding is a ctypes stub; wave_mamba_b200/_cabi.py is that stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to float32 data on the current CUDA device,
 *     contiguous NCHW unless stated; the caller owns all buffers (inputs, outputs,
 *     workspaces); nothing is allocated or freed by the library;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); kernels
 *     are enqueued asynchronously on it and the call returns immediately;
 *   - return value: 0 on success, a negative WM_E* code on failure; the message for the
 *     calling thread's last failure is wm_last_error().  There is no CPU fallback.
 */
#ifndef WAVEMAMBA_B200_H
#define WAVEMAMBA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WM_OK 0
#define WM_EINVAL (-1)    /* bad shape / null pointer / misaligned or too-small workspace */
#define WM_ECUDA (-2)     /* a CUDA runtime call or kernel launch failed               */
#define WM_ENODEVICE (-3) /* no sm_100-class CUDA device is current                     */

#define WM_ABI_VERSION 15

typedef void *wm_stream_t;

int wm_abi_version(void);
const char *wm_last_error(void);
/* 0 when the current device can run this library (compute capability 10.x), else WM_ENODEVICE. */
int wm_device_check(void);

/* ---- Haar DWT -- dwt_init / DWT.forward, wavemamba_arch.py:97-110,133-139 -------------
 * x: (planes, H, W) with planes = B*C, H and W even.  ll/hl/lh/hh: (planes, H/2, W/2).
 * Bit-exact with the reference (same /2 and the same left-to-right association). */
int wm_dwt_haar_fwd(const float *x, float *ll, float *hl, float *lh, float *hh,
                    int64_t planes, int64_t H, int64_t W, wm_stream_t stream);
/* The same transform with SKFF's global average pool (reference :939-948) in its epilogue: per-CTA partial
 * sums of (HL + LH) + HH per plane go to `pool_partials` (>= wm_skff_workspace_bytes(B, H/2, W/2) bytes for
 * planes = B*32, 8-byte aligned), in the layout wm_skff_apply_fwd reads. */
int wm_dwt_haar_pool_fwd(const float *x, float *ll, float *hl, float *lh, float *hh, void *pool_partials,
                         size_t workspace_bytes, int64_t planes, int64_t H, int64_t W, wm_stream_t stream);

/* ---- Haar IWT -- iwt_init / IWT.forward, wavemamba_arch.py:113-130,142-148 -------------
 * The reference takes torch.cat([x_l, x_h], 1) (B,4C,h,w) (wavemamba_arch.py:1006); here
 * the two halves are passed separately so the cat copy is never made:
 *   low : B planes-groups of C planes (LL),  batch stride low_bstride  (floats)
 *   high: B groups of 3C planes (HL|LH|HH),  batch stride high_bstride (floats)
 * Passing low = cat, high = cat + C*h*w, both strides 4*C*h*w reproduces the cat form.
 * y: (B, C, 2h, 2w) contiguous.  Bit-exact with the reference. */
int wm_iwt_haar_fwd(const float *low, int64_t low_bstride, const float *high,
                    int64_t high_bstride, float *y, int64_t B, int64_t C, int64_t h, int64_t w,
                    wm_stream_t stream);

/* ---- SS2D core -- SS2D.forward_core + the 4-way sum of SS2D.forward,
 *      wavemamba_arch.py:446-478,490 (cross-scan, x_proj, dt_proj, softplus,
 *      selective_scan_fn, flips/transposes, y1+y2+y3+y4) ------------------------------------
 * x: (B, 64, h, w) (output of conv2d+SiLU).  y: (B, 64, h, w) = merged scan output.
 * x_proj_weight (4,34,64), dt_projs_weight (4,64,2), dt_projs_bias (4,64), A_logs (256,16),
 * Ds (256): the module's parameters as stored in the checkpoint.
 * workspace: >= wm_ss2d_core_workspace_bytes(B,h,w) bytes, 256-byte aligned.
 * Fixed model constants: d_inner 64, d_state 16, dt_rank 2, 4 directions. */
size_t wm_ss2d_core_workspace_bytes(int64_t B, int64_t h, int64_t w);
/* Developer aid: non-NULL device buffer of 8192*6 int64 -> every pass kernel CTA writes its phase
 * cycle sums [wait-x, projection, delta, scan, store, tiles] at row blockIdx.x % 4096 (the output pass)
 * or 4096 + blockIdx.x % 4096 (pass 1). */
int wm_ss2d_debug_timing(void *device_buffer);
/* Developer aid: the chunk plan chosen for (B,h,w): out6 = {row chunk steps, row CTAs per
 * direction, column segment steps, segments per column, column CTAs per direction, columns first}. */
int wm_ss2d_debug_geometry(int64_t B, int64_t h, int64_t w, int *out6);
/* Same computation without the final 4-way sum: on return the first 4*B*64*h*w floats of
 * `workspace` hold the four direction outputs as planes[k][b][d][i][j] (pixel order, k = 0..3 in
 * the reference's direction order); the consumer sums them as ((p0 + p2) + p1) + p3, which is the
 * reference's y1+y2+y3+y4 (wavemamba_arch.py:474-478,490).  wm_lfss_out_fwd does that sum. */
int wm_ss2d_dirs_fwd(const float *x, const float *x_proj_weight, const float *dt_projs_weight,
                     const float *dt_projs_bias, const float *A_logs, const float *Ds,
                     void *workspace, size_t workspace_bytes, int64_t B, int64_t h, int64_t w,
                     wm_stream_t stream);
int wm_ss2d_core_fwd(const float *x, const float *x_proj_weight, const float *dt_projs_weight,
                     const float *dt_projs_bias, const float *A_logs, const float *Ds, float *y,
                     void *workspace, size_t workspace_bytes, int64_t B, int64_t h, int64_t w,
                     wm_stream_t stream);

/* Backward of wm_ss2d_core_fwd (training: the gradient the reference gets from autograd through
 * selective_scan_fn, the einsum projections and the cross-scan / cross-merge, :446-478,490).
 * grad_y (B,64,h,w) -> grad_x (B,64,h,w) and the parameter gradients (same shapes as the
 * parameters; overwritten, not accumulated).  Deterministic (fixed-order reductions).
 * workspace: wm_ss2d_core_bwd_workspace_bytes(B,h,w) bytes, 256-byte aligned. */
size_t wm_ss2d_core_bwd_workspace_bytes(int64_t B, int64_t h, int64_t w);
int wm_ss2d_core_bwd(const float *x, const float *x_proj_weight, const float *dt_projs_weight,
                     const float *dt_projs_bias, const float *A_logs, const float *Ds,
                     const float *grad_y, float *grad_x, float *grad_x_proj_weight,
                     float *grad_dt_projs_weight, float *grad_dt_projs_bias, float *grad_A_logs,
                     float *grad_Ds, void *workspace, size_t workspace_bytes, int64_t B, int64_t h,
                     int64_t w, wm_stream_t stream);

/* ---- fused pointwise (1x1) / depthwise (3x3) convolution groups ----------------------------
 * HFEBlock / CMTAttention / FeedForward / PAConv / LFSSBlock.ffn pieces,
 * wavemamba_arch.py:214-231,687,694-697,729-742,756-764,775,797,826-851.
 * All tensors NCHW float32; weights in PyTorch conv layout (Cout,Cin,1,1) / (C,1,3,3).
 * `ln_w`/`ln_b` non-NULL fuses LayerNorm2d (wavemamba_arch.py:535-543, biased variance over
 * the channel axis, eps) in front of the 1x1.  Depthwise convs use zero padding 1.          */

/* LayerNorm2d alone: y = ln(x).  C <= 64. */
int wm_layernorm2d_fwd(const float *x, const float *ln_w, const float *ln_b, float eps, float *y,
                       int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream);

/* y = act( dw3x3( pw1x1( ln?(x) ) ) ) : Cin=32 -> Cout in {32,64,96}; pw_b may be NULL;
 * act: 0 none, 1 SiLU.
 * Used for qkv+qkv_dwconv (:775), FeedForward.project_in (:729-732,745), ffn.conv1+conv2 (:226)
 * and, with ln_1 + in_proj[:64] + conv2d + SiLU, the x branch of SS2D.forward (:483-487, :524). */
int wm_pw_dw_fwd(const float *x, const float *ln_w, const float *ln_b, float eps,
                 const float *pw_w, const float *pw_b, const float *dw_w, const float *dw_b,
                 int act, float *y, int64_t B, int64_t Cin, int64_t Cout, int64_t h, int64_t w,
                 wm_stream_t stream);

/* y = residual? + pw1x1( act( dw3x3(x) ) ), act: 0 none, 1 exact (erf) GELU.   C=32 -> 32.
 * FeedForward.project_out (:739-742,750) fused with the HFEBlock residual add (:851). */
int wm_dw_act_pw_fwd(const float *x, const float *dw_w, const float *dw_b, const float *pw_w,
                     const float *pw_b, int act, const float *residual, float *y, int64_t B,
                     int64_t C, int64_t h, int64_t w, wm_stream_t stream);

/* y = residual?*res_scale? + pw1x1(x) + bias? : Cin -> Cout, both <= 64.
 * CMTAttention.project_out (:797) fused with the HFEBlock residual add (:849);
 * ffn.conv3 after the gate fused with `x*skip_scale2 +` (:526).  gate_mode 0: plain; 1: input is
 * (B,2*Cin,h,w) and the 1x1 sees gelu(x[:, :Cin]) * x[:, Cin:]  (ffn gate, :227-228).
 * res_scale: optional (Cout) per-channel multiplier of the residual.
 * x_bstride: batch stride of x in floats (0 = dense; lets x be a channel slice of a wider tensor);
 * w_bstride: batch stride of pw_w in floats (0 = one weight matrix; non-zero = per-image weights,
 * used for the CxC attention folded into project_out, :791-797). */
int wm_pw_fwd(const float *x, int64_t x_bstride, const float *pw_w, int64_t w_bstride,
              const float *pw_b, int gate_mode, const float *residual, const float *res_scale,
              float *y, int64_t B, int64_t Cin, int64_t Cout, int64_t h, int64_t w,
              wm_stream_t stream);

/* SS2D z branch: zs = silu( in_proj.weight[64:128] . LayerNorm_c(x) ), x (B,32,h,w) -> (B,64,h,w)
 * (ln_1 :524, in_proj + chunk :483-484, F.silu(z) :493).  w_z points at row 64 of in_proj.weight. */
int wm_lfss_z_fwd(const float *x, const float *ln_w, const float *ln_b, float eps, const float *w_z,
                  float *zs, int64_t B, int64_t h, int64_t w, wm_stream_t stream);

/* SS2D tail + LFSSBlock residual:
 *   out = x*skip_scale + out_proj( out_norm(((y + ya) + yb) + yc) * zs )
 * (4-way sum :490, out_norm :492, gate :493, out_proj :494, residual :525).  ya/yb/yc are optional
 * (NULL) extra addends: pass the planes of wm_ss2d_dirs_fwd as y=p0, ya=p2, yb=p1, yc=p3.
 * y*, zs: (B,64,h,w); x, out: (B,32,h,w). */
int wm_lfss_out_fwd(const float *y, const float *ya, const float *yb, const float *yc,
                    const float *zs, const float *on_w, const float *on_b, float eps,
                    const float *w_out, const float *x, const float *skip_scale, float *out,
                    int64_t B, int64_t h, int64_t w, wm_stream_t stream);

/* The two calls above in one: the gate zs = silu(w_z . LayerNorm_32(x)) is computed inside the tail kernel
 * (one kernel and 224 channel planes of traffic less per LFSSBlock).  All four planes are required.
 * Shapes whose pixel count is not a multiple of 4 (or unaligned tensors) run as wm_lfss_z_fwd +
 * wm_lfss_out_fwd through `zs_scratch` (B,64,h,w), which may be NULL otherwise (then such a shape is
 * WM_EINVAL).  Replaces wavemamba_arch.py:483-484 (z half), :490-494, :524-525. */
int wm_lfss_tail_fwd(const float *y, const float *ya, const float *yb, const float *yc, const float *x,
                     const float *ln1_w, const float *ln1_b, float ln1_eps, const float *w_z,
                     const float *on_w, const float *on_b, float on_eps, const float *w_out,
                     const float *skip_scale, float *zs_scratch, float *out, int64_t B, int64_t h,
                     int64_t w, wm_stream_t stream);

/* ---- 32x32 Gram matrix + squared norms of two 32-channel stacks, one pass -----------------
 * out[b] = [ G (32x32, row-major, G[i][j] = sum_p X[b][i][p]*Y[b][j][p]) | |X_i|^2 (32) | |Y_j|^2 (32) ]
 * (1088 floats per batch item).  X, Y: 32 channel planes of hw contiguous floats each, batch
 * strides x_bstride / y_bstride in floats (so a 32-channel slice of a wider NCHW tensor works).
 * Replaces torch.cdist in Matching (:664; dist^2 = |x|^2 + |p|^2 - 2G, argmin :624) and
 * normalize(q) @ normalize(k)^T in CMTAttention (:787-790).  Accumulates 128-pixel tiles in fp32
 * and tile sums in fp64; deterministic.  workspace >= wm_gram32_workspace_bytes(B, hw), 8-aligned. */
size_t wm_gram32_workspace_bytes(int64_t B, int64_t hw);
int wm_gram32_fwd(const float *x, int64_t x_bstride, const float *y, int64_t y_bstride, float *out,
                  void *workspace, size_t workspace_bytes, int64_t B, int64_t hw,
                  wm_stream_t stream);

/* wm_gram32_fwd with the consumer's 32x32 tail folded into the reduce kernel (one launch instead of the
 * ~5-8 tiny torch launches that followed it in an HFEBlock).  gram_out: optional (B,1088) copy of the raw
 * Gram output, may be NULL.  Workspace as wm_gram32_fwd.
 *   match: idx[b][i] = argmin_j (|x_i|^2 + |p_j|^2) - 2 x_i.p_j   -- Matching :618-680 with match_factor 1
 *          (torch.cdist's mm mode + topk(k=1, largest=False); same fp32 expression, first index on ties)
 *   attn:  mixed[b] = W_po . softmax_j( q_i.k_j / (max(|q_i|,1e-12) max(|k_j|,1e-12)) * temperature )
 *          -- CMTAttention :787-797: F.normalize, @, * temperature, softmax, and project_out folded into
 *          per-image 1x1 weights for wm_pw_fwd. */
int wm_gram32_match_fwd(const float *x, int64_t x_bstride, const float *p, int64_t p_bstride, int *idx,
                        float *gram_out, void *workspace, size_t workspace_bytes, int64_t B, int64_t hw,
                        wm_stream_t stream);
int wm_gram32_attn_fwd(const float *q, int64_t q_bstride, const float *k, int64_t k_bstride,
                       const float *temperature, const float *w_po, float *mixed, float *gram_out,
                       void *workspace, size_t workspace_bytes, int64_t B, int64_t hw, wm_stream_t stream);

/* ---- dense 3x3 convolution (stride 1, zero pad 1), implicit GEMM on tensor cores, 3xTF32 ----
 * PAConv.k3/k4 (:689-698), DownFRG.l_conv (:966,975), upFRG.h_out_conv (:993,1005).
 * Weights are pre-packed once per layer (mma fragment order, tf32 hi/lo split):
 *   wm_conv3x3_prepack(w3x3 (Cout,Cin,3,3), w1x1 (Cout,Cin) or NULL, packed, Cin, Cout)
 *   with `packed` >= wm_conv3x3_packed_bytes(Cin, Cout, w1x1 != NULL) bytes, 16-byte aligned.
 * Forward: the Cin input channels are the first Ca channels of in_a followed by Cin-Ca channels of
 * in_b selected per batch item through chan_map (B, Cin-Ca) int32 (NULL = identity) -- the
 * torch.cat([x, matched]) of :716 / cat([LL, x_d]) of :975 is never materialised.
 *   gate_bias == NULL: out = conv3x3(in) + bias?            (Cin,Cout) in {(64,32),(64,64),(32,96),(32,32)}
 *   gate_bias != NULL: out = conv3x3(in) * sigmoid(conv1x1(in) + gate_bias)   PAConv :694-697, 64->64,
 *                      `packed` must have been built with the 1x1 weights.
 * fp32 accuracy (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate). */
size_t wm_conv3x3_packed_bytes(int64_t Cin, int64_t Cout, int with_gate);
/* Developer aid: non-NULL device buffer of 6*SMs int64 -> the kernel's MMA thread writes per-CTA
 * cycle counts (total, wait-weights, wait-X, wait-accumulators, issue, tiles); NULL disables. */
int wm_conv3x3_debug_timing(void *device_buffer);
/* Developer aid: synchronises the device, returns and clears the word in which the mbarrier pipelines
 * (dense 3x3 conv, pw_dw) record a wait that timed out: 0 = none, else
 * 0x80000000 | role << 24 | barrier << 16 | iteration of the first one. */
int wm_debug_pipeline_error(unsigned int *out);
/* Developer aid: non-NULL device buffer of 16*SMs int64 -> per-CTA cycle counters of the pw_dw
 * pipeline (waits and phases of its three warp roles); NULL disables. */
int wm_pw_dw_debug_timing(void *device_buffer);
int wm_conv3x3_prepack(const float *w3x3, const float *w1x1, void *packed, int64_t Cin, int64_t Cout,
                       wm_stream_t stream);
int wm_conv3x3_fwd(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b,
                   int64_t b_bstride, const int *chan_map, const void *packed, const float *bias,
                   const float *gate_bias, float *out, int64_t B, int64_t Cin, int64_t Cout,
                   int64_t h, int64_t w, wm_stream_t stream);
/* Same, with channel-quad tensor layouts (B, C/4, h, w, 4) for the input (in_c4, needs Ca == Cin)
 * and/or the output (out_c4): the intermediate between PAConv.k3 and k4 is only ever read by the
 * next conv, and 16-byte (4-channel) elements are what its shared-memory operand layout wants.
 * tcgen05 implementation only. */
int wm_conv3x3_ex_fwd(const float *in_a, int64_t a_bstride, int64_t Ca, const float *in_b,
                   int64_t b_bstride, const int *chan_map, const void *packed, const float *bias,
                   const float *gate_bias, float *out, int64_t B, int64_t Cin, int64_t Cout,
                   int64_t h, int64_t w, int in_c4, int out_c4,
                     wm_stream_t stream);

/* Stem / head 3x3 convs with 3 channels on one side (direct FFMA, fp32):
 *   stem: UNet.conv_01 (:1026,1048)  x (B,3,h,w)  -> y (B,32,h,w) = conv3x3(x) + bias
 *   head: UNet.last + global residual (:1039,1061)  x (B,32,h,w) -> y (B,3,h,w) = conv3x3(x)+bias+residual? */
int wm_stem_conv3x3_fwd(const float *x, const float *w3x3, const float *bias, float *y, int64_t B,
                        int64_t h, int64_t w, wm_stream_t stream);
int wm_head_conv3x3_fwd(const float *x, const float *w3x3, const float *bias, const float *residual,
                        float *y, int64_t B, int64_t h, int64_t w, wm_stream_t stream);

/* PAConv gate: y = k3out * sigmoid( pw1x1(x) + b ), x and k3out and y all (B,64,h,w)
 * (PAConv.k2 + sigmoid + mul, :694-697).  In-place on k3out allowed (y == k3out). */
int wm_paconv_gate_fwd(const float *x, const float *k2_w, const float *k2_b, const float *k3out,
                       float *y, int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream);

/* ---- SKFF band fusion -- SKFF.forward, wavemamba_arch.py:939-959 (called :981) -------------
 * f0,f1,f2: the HL, LH, HH bands, each (B,32,h,w) contiguous.  out (B,32,h,w) =
 * (f0*a0 + f1*a1) + f2*a2 with a = softmax over the three bands of fcs_k(PReLU(conv_du(mean_hw(
 * (f0+f1)+f2)))).  w_du (4,32) = conv_du.0.weight, prelu_weight (1) = conv_du.1.weight,
 * w_fck (32,4) = fcs.k.weight (all bias-free, as in the reference).
 * workspace: >= wm_skff_workspace_bytes(B,h,w) bytes, 8-byte aligned (per-CTA pool partials). */
size_t wm_skff_workspace_bytes(int64_t B, int64_t h, int64_t w);
int wm_skff_fwd(const float *f0, const float *f1, const float *f2, const float *w_du,
                const float *prelu_weight, const float *w_fc0, const float *w_fc1,
                const float *w_fc2, float *out, void *workspace, size_t workspace_bytes, int64_t B,
                int64_t C, int64_t h, int64_t w, wm_stream_t stream);
/* SKFF with its pool pass done elsewhere: `pool_partials` = the buffer wm_dwt_haar_pool_fwd filled for the
 * DWT that produced f0, f1, f2 (= HL, LH, HH).  One streaming pass instead of two. */
int wm_skff_apply_fwd(const float *f0, const float *f1, const float *f2, const float *w_du,
                      const float *prelu_weight, const float *w_fc0, const float *w_fc1,
                      const float *w_fc2, float *out, const void *pool_partials, size_t workspace_bytes,
                      int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream);

/* ---- side inputs -- UNet.ps_down{1,2,3} = PixelUnshuffle(r) + Conv2d(3 r^2, 32, 1),
 *      wavemamba_arch.py:1014-1025, called :1043-1045 ----------------------------------------
 * x: (B,3,H,W), H and W multiples of r (2, 4 or 8), 16-byte aligned, W % 4 == 0.
 * weight (32, 3 r^2) with the unshuffled channel order ci*r*r + dy*r + dx, bias (32) or NULL.
 * y: (B,32,H/r,W/r).  The unshuffled tensor is never materialised. */
int wm_ps_down_fwd(const float *x, const float *weight, const float *bias, float *y, int64_t B,
                   int64_t H, int64_t W, int r, wm_stream_t stream);

/* ---- training-only kernels (reference training step femasr_model.py:157-185) ---------
 * The backward pass is built from forward kernels with transposed weights, Gram matrices for the
 * weight gradients (wm_gram32_fwd), the Haar adjoints, wm_ss2d_core_bwd, and these three. */
/* Depthwise 3x3, zero pad 1, any C; wgt (C,9); bias may be NULL.  flip != 0: taps rotated by 180
 * degrees (the data gradient of the same convolution). */
int wm_dw3x3_fwd(const float *x, const float *wgt, const float *bias, float *y, int64_t B, int64_t C,
                 int64_t h, int64_t w, int flip, wm_stream_t stream);
/* Scratch for wm_dw3x3_wgrad / wm_layernorm2d_bwd (16-byte aligned). */
size_t wm_train_workspace_bytes(int64_t B, int64_t C, int64_t h, int64_t w);
/* dwgt[c][t] = sum dy[b,c,p] x[b,c,p+t], dbias[c] = sum dy (dbias may be NULL). Deterministic. */
int wm_dw3x3_wgrad(const float *dy, const float *x, float *dwgt, float *dbias, void *workspace,
                   size_t workspace_bytes, int64_t B, int64_t C, int64_t h, int64_t w, wm_stream_t stream);
/* Backward of wm_layernorm2d_fwd (C = 32 or 64): dx, dln_w, dln_b.  Deterministic. */
int wm_layernorm2d_bwd(const float *x, const float *ln_w, const float *dy, float eps, float *dx,
                       float *dln_w, float *dln_b, void *workspace, size_t workspace_bytes, int64_t B,
                       int64_t C, int64_t h, int64_t w, wm_stream_t stream);

/* ---- image I/O edges of the inference loop -- img2tensor + "/255." + check_image_size
 *      (basicsr/utils/img_util.py:9-33, inference_wavemamba.py:28-36,102-106) and the crop +
 *      tensor2img on the way out (inference_wavemamba.py:112-113, img_util.py:36-98) -----------
 * img: (B,H,W,3) uint8, BGR (a cv2 image), 4-byte aligned.  out: (B,3,Hp,Wp) float32 RGB in [0,1],
 * reflect-padded at the bottom / right (Hp-H < H, Wp-W < W), 16-byte aligned.  Bit-exact:
 * reciprocal == 0: byte / 255.0f (IEEE division -- what the reference's line computes on a CPU tensor);
 * reciprocal != 0: byte * (1.0f / 255.0f) -- what the same line computes on a CUDA tensor (torch's
 * tensor / python-scalar kernel multiplies by the reciprocal), i.e. the reference script's GPU run. */
int wm_img_u8_to_f32_fwd(const uint8_t *img, float *out, int64_t B, int64_t H, int64_t W, int64_t Hp,
                         int64_t Wp, int reciprocal, wm_stream_t stream);
/* x: (B,3,Hs,Ws) float32 RGB.  img: (B,h,w,3) uint8 BGR = round_half_even(clamp(x[:, :, :h, :w], 0, 1)
 * * 255).  Bit-exact with tensor2img. */
int wm_img_f32_to_u8_fwd(const float *x, uint8_t *img, int64_t B, int64_t h, int64_t w, int64_t Hs,
                         int64_t Ws, wm_stream_t stream);

/* ---- the two metrics the inference loop prints per image (inference_wavemamba.py:117-118) --------
 * comput_psnr_ssim.py calculate_psnr :387-438 and calculate_ssim :596-668 with their defaults
 * (input_order 'HWC', test_y_channel True): img1, img2 (B,H,W,3) uint8 BGR on the device, cropped by
 * crop_border on every edge, Y channel per to_y_channel :374-385 / bgr2ycbcr :210-238, PSNR from the
 * fp64 mean of the fp32 squared differences (+inf if 0), SSIM per _ssim_cly :559-592 (11x11 Gaussian
 * sigma 1.5, replicate border, fp64).  out: (B,2) float64 = [psnr, ssim] per image.  Deterministic.
 * workspace: wm_psnr_ssim_y_workspace_bytes(B,H,W,crop_border) bytes, 8-byte aligned. */
size_t wm_psnr_ssim_y_workspace_bytes(int64_t B, int64_t H, int64_t W, int crop_border);
int wm_psnr_ssim_y_u8(const uint8_t *img1, const uint8_t *img2, double *out, void *workspace,
                      size_t workspace_bytes, int64_t B, int64_t H, int64_t W, int crop_border,
                      wm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVEMAMBA_B200_H */
