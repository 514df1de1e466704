"""Host-side mirror of the reference's Wave-Mamba arch interface, driving the sm_100a kernels.

Drop-in contract (reference basicsr/archs/wavemamba_arch.py:1066-1176):
  * ``WaveMamba(*, in_chn, wf, n_l_blocks, n_h_blocks, ffn_scale, **ignored)`` with the
    attribute ``restoration_network`` (an nn.Module called directly by inference_wavemamba.py),
    methods ``forward / test / test_tile / encode_and_decode / check_image_size / print_network``;
  * the parameter tree has exactly the reference's 591 state-dict keys and shapes, so the shipped
    checkpoints load with ``strict=True``.

Module classes below are parameter containers whose names reproduce that key schema; their
forwards call ``wave_mamba_b200.ops`` (hand-written CUDA through the C ABI) for everything that
touches a feature map: Haar DWT/IWT, the SS2D core, the pointwise + depthwise groups, the dense
3x3 convolutions (tcgen05), channel matching / CxC attention (Gram kernel), SKFF, the stem / head
convolutions.  Only 32x32-element bookkeeping (argmin, softmax, folding the attention into the
1x1 weights) is left to torch.

Two paths share the parameters:
  * inference (``torch.no_grad()`` / ``.eval()`` without grad): the fused kernels;
  * training (grad enabled): an unfused composition of ``wave_mamba_b200.autograd`` Functions --
    every convolution, LayerNorm, the wavelets and the SS2D core run this repo's CUDA kernels in
    both directions; pointwise glue is torch elementwise autograd.
There is no CPU path: tensors that are not on a CUDA device raise.
"""
from __future__ import annotations

import math
import os
from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as ag
from . import ops
from ._cabi import WaveMambaNativeError

__all__ = ["WaveMamba", "UNet", "DownFRG", "upFRG", "LFSSBlock", "SS2D", "HFEBlock", "SKFF",
           "DWT", "IWT"]


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise WaveMambaNativeError(
            f"input is on {t.device}: the B200 Wave-Mamba path has no CPU fallback; move the "
            "module and its input to a CUDA device")


def _wants_grad(module: nn.Module, t: torch.Tensor) -> bool:
    """The differentiable (training) path is taken whenever autograd could record something: grad mode
    on and either the input or any parameter of the network requires a gradient."""
    return torch.is_grad_enabled() and (t.requires_grad or any(p.requires_grad for p in module.parameters()))


# --------------------------------------------------------------------------------------
# wavelets                                                      reference :133-148
# --------------------------------------------------------------------------------------
class _DWTFn(torch.autograd.Function):
    """The reference's Haar pair is orthonormal (4 taps of +-1/2), so the adjoint of the analysis is
    the synthesis and vice versa: both backward passes reuse the forward kernels."""

    @staticmethod
    def forward(ctx, x):
        return ops.dwt_haar(x.contiguous())

    @staticmethod
    def backward(ctx, g_ll, g_hl, g_lh, g_hh):
        high = torch.cat([g_hl, g_lh, g_hh], dim=1)
        return ops.iwt_haar(g_ll.contiguous(), high)


class _IWTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, low, high):
        ctx.channels = low.shape[1]
        return ops.iwt_haar(low.contiguous(), high.contiguous())

    @staticmethod
    def backward(ctx, g):
        g_ll, g_hl, g_lh, g_hh = ops.dwt_haar(g.contiguous())
        return g_ll, torch.cat([g_hl, g_lh, g_hh], dim=1)


class DWT(nn.Module):
    """Differentiable on its own (the whole network's training step still needs the SS2D backward)."""

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            return _DWTFn.apply(x)
        return ops.dwt_haar(x.contiguous())


class IWT(nn.Module):
    """Accepts the reference's concatenated (B,4C,h,w) tensor, or (low, high) without a cat."""

    def forward(self, x, high=None):
        if high is None:
            if torch.is_grad_enabled() and x.requires_grad:
                c = x.shape[1] // 4
                return _IWTFn.apply(x[:, :c], x[:, c:])
            return ops.iwt_haar_cat(x.contiguous())
        if torch.is_grad_enabled() and (x.requires_grad or high.requires_grad):
            return _IWTFn.apply(x, high)
        return ops.iwt_haar(x.contiguous(), high.contiguous())


# --------------------------------------------------------------------------------------
# low-frequency branch                                          reference :214-231,316-528
# --------------------------------------------------------------------------------------
class _FFN(nn.Module):
    """reference ``ffn`` (:214-231): 1x1 C->2C, dw3x3, gelu(x1)*x2, 1x1 C->C."""

    def __init__(self, num_feat: int, ffn_expand: int = 2):
        super().__init__()
        mid = num_feat * ffn_expand
        self.conv1 = nn.Conv2d(num_feat, mid, 1)
        self.conv2 = nn.Conv2d(mid, mid, 3, padding=1, groups=mid)
        self.conv3 = nn.Conv2d(mid // 2, num_feat, 1)

    def forward_train(self, x):
        """reference ffn.forward (:225-230), differentiable."""
        t = ag.PW.apply(x, self.conv1.weight, self.conv1.bias)
        t = ag.DW.apply(t, self.conv2.weight, self.conv2.bias)
        x1, x2 = t.chunk(2, dim=1)
        return ag.PW.apply(F.gelu(x1) * x2, self.conv3.weight, self.conv3.bias)

    def forward(self, x, ln_w=None, ln_b=None, eps=1e-5, residual=None, res_scale=None):
        """ln?(x) -> conv1 -> conv2 -> gate -> conv3 (+ residual*res_scale), two kernels."""
        t = ops.pw_dw(x, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                      ln_w, ln_b, eps)
        return ops.pw(t, self.conv3.weight, self.conv3.bias, gate=True, residual=residual,
                      res_scale=res_scale)


class SS2D(nn.Module):
    """reference ``SS2D`` (:316-497).  Parameters and their initialisation follow :345-387."""

    def __init__(self, d_model: int, d_state: int = 16, d_conv: int = 3, expand: float = 2.0,
                 dt_min: float = 0.001, dt_max: float = 0.1, dt_scale: float = 1.0,
                 dt_init_floor: float = 1e-4, **_unused):
        super().__init__()
        self.d_model, self.d_state = d_model, d_state
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16)
        D, R, N, K = self.d_inner, self.dt_rank, d_state, 4

        self.in_proj = nn.Linear(d_model, 2 * D, bias=False)
        self.conv2d = nn.Conv2d(D, D, d_conv, padding=(d_conv - 1) // 2, groups=D, bias=True)

        bound = 1.0 / math.sqrt(D)  # nn.Linear default init of the four x_proj layers (:357-363)
        self.x_proj_weight = nn.Parameter(torch.empty(K, R + 2 * N, D).uniform_(-bound, bound))
        std = R ** -0.5 * dt_scale  # :395-399
        self.dt_projs_weight = nn.Parameter(torch.empty(K, D, R).uniform_(-std, std))
        dt = torch.exp(torch.rand(K, D) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)        # :404-407
        self.dt_projs_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))  # inverse softplus :409
        a = torch.arange(1, N + 1, dtype=torch.float32).log()               # S4D-real :420-425
        self.A_logs = nn.Parameter(a.repeat(K * D, 1))
        self.Ds = nn.Parameter(torch.ones(K * D))
        self.A_logs._no_weight_decay = True
        self.Ds._no_weight_decay = True

        self.out_norm = nn.LayerNorm(D)
        self.out_proj = nn.Linear(D, d_model, bias=False)

    def forward_core(self, x: torch.Tensor) -> torch.Tensor:
        """(B,D,h,w) -> merged (B,D,h,w): reference forward_core + y1+y2+y3+y4 (:446-478,490)."""
        return ops.ss2d_core(x.contiguous(), self.x_proj_weight, self.dt_projs_weight,
                             self.dt_projs_bias, self.A_logs, self.Ds)

    def forward_nchw(self, x, ln_w, ln_b, ln_eps, skip_scale):
        """Fused NCHW form of ``x*skip_scale + SS2D(ln_1(x))`` (reference :524-525, :480-497):
        two kernels around the scan, no permutes, no channels-last round trips.
            xc = silu(dwconv3x3(in_proj[:D] . ln(x)))          [pw_dw, LN prologue, SiLU epilogue]
            p  = the four direction planes of the scan           [ss2d_dirs]
            zs = silu(in_proj[D:] . ln(x))                      } [lfss_tail: one kernel,
            out = x*skip_scale + out_proj(out_norm(sum p) * zs)  }  zs never leaves the SM]"""
        D = self.d_inner
        xc = ops.pw_dw(x, self.in_proj.weight[:D], None, self.conv2d.weight, self.conv2d.bias,
                       ln_w, ln_b, ln_eps, act="silu")
        p = ops.ss2d_dirs(xc, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias,
                          self.A_logs, self.Ds)
        # reference sum order y1+y2+y3+y4 = ((dir0 + dir2) + dir1) + dir3, folded into the tail kernel,
        # which also computes the gate zs from x (no z tensor)
        return ops.lfss_tail((p[0], p[2], p[1], p[3]), x, ln_w, ln_b, ln_eps, self.in_proj.weight,
                             self.out_norm.weight, self.out_norm.bias, self.out_norm.eps,
                             self.out_proj.weight, skip_scale)

    def forward_train(self, xn: torch.Tensor) -> torch.Tensor:
        """SS2D.forward (:480-497) on an already normalised NCHW map, differentiable: in_proj -> dwconv ->
        SiLU -> core -> out_norm -> * silu(z) -> out_proj."""
        D = self.d_inner
        xz = ag.PW.apply(xn, self.in_proj.weight[:D], None), ag.PW.apply(xn, self.in_proj.weight[D:], None)
        xc = F.silu(ag.DW.apply(xz[0], self.conv2d.weight, self.conv2d.bias))
        y = ag.SS2DCore.apply(xc, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias,
                              self.A_logs, self.Ds)
        y = ag.LN2d.apply(y, self.out_norm.weight, self.out_norm.bias, self.out_norm.eps)
        return ag.PW.apply(y * F.silu(xz[1]), self.out_proj.weight, None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference calling convention: x (B,h,w,C) channels-last -> (B,h,w,C) (:480-497)."""
        # The fused kernels take the LayerNorm with them, so this un-normalised entry point runs
        # the glue through library ops; the network itself uses forward_nchw.
        xz = F.linear(x, self.in_proj.weight)
        xp, z = xz.chunk(2, dim=-1)
        xc = xp.permute(0, 3, 1, 2).contiguous()
        xc = F.silu(F.conv2d(xc, self.conv2d.weight, self.conv2d.bias, padding=1,
                             groups=self.d_inner))
        y = self.forward_core(xc).permute(0, 2, 3, 1)
        y = F.layer_norm(y, (self.d_inner,), self.out_norm.weight, self.out_norm.bias, 1e-5)
        return F.linear(y * F.silu(z), self.out_proj.weight)


class LFSSBlock(nn.Module):
    """reference ``LFSSBlock`` (:499-528)."""

    def __init__(self, hidden_dim: int, d_state: int = 16, expand: float = 2.0, **_unused):
        super().__init__()
        self.ln_1 = nn.LayerNorm(hidden_dim, eps=1e-6)
        self.self_attention = SS2D(d_model=hidden_dim, d_state=d_state, expand=expand)
        self.skip_scale = nn.Parameter(torch.ones(hidden_dim))
        self.conv_blk = _FFN(hidden_dim)
        self.ln_2 = nn.LayerNorm(hidden_dim)
        self.skip_scale2 = nn.Parameter(torch.ones(hidden_dim))

    def forward_train(self, x: torch.Tensor) -> torch.Tensor:
        """LFSSBlock.forward (:520-528) on NCHW, differentiable."""
        s1 = self.skip_scale.view(1, -1, 1, 1)
        s2 = self.skip_scale2.view(1, -1, 1, 1)
        x = x * s1 + self.self_attention.forward_train(
            ag.LN2d.apply(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps))
        return x * s2 + self.conv_blk.forward_train(
            ag.LN2d.apply(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps))

    def forward_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """(B,C,h,w) -> (B,C,h,w); five kernels + the scan, all NCHW."""
        if _wants_grad(self, x):
            return self.forward_train(x)
        x = self.self_attention.forward_nchw(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps,
                                             self.skip_scale)                       # :524-525
        return self.conv_blk(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps,
                             residual=x, res_scale=self.skip_scale2)                # :526

    def forward(self, inp: torch.Tensor, x_size: Sequence[int]) -> torch.Tensor:
        """Reference calling convention: (B, h*w, C) in and out."""
        B, L, C = inp.shape
        x = inp.view(B, x_size[0], x_size[1], C).permute(0, 3, 1, 2).contiguous()
        return self.forward_nchw(x).permute(0, 2, 3, 1).reshape(B, L, C)


# --------------------------------------------------------------------------------------
# high-frequency branch                                         reference :560-854
# --------------------------------------------------------------------------------------
class LayerNorm2d(nn.Module):
    def __init__(self, channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.eps = eps

    def forward(self, x):
        if _wants_grad(self, x):
            return ag.LN2d.apply(x, self.weight, self.bias, self.eps)
        return ops.layernorm2d(x.contiguous(), self.weight, self.bias, self.eps)


class PAConv(nn.Module):
    """reference :683-700.  Two tensor-core kernels: k3 (+ the 1x1 k2 sigmoid gate) and k4."""

    def __init__(self, nf: int):
        super().__init__()
        self.k2 = nn.Conv2d(nf, nf, 1)
        self.k3 = nn.Conv2d(nf, nf, 3, padding=1, bias=False)
        self.k4 = nn.Conv2d(nf, nf // 2, 3, padding=1, bias=False)

    def forward_train(self, x):
        """PAConv.forward (:692-699) on the concatenated input, differentiable."""
        y = torch.sigmoid(ag.PW.apply(x, self.k2.weight, self.k2.bias))
        out = ag.Conv3.apply(x, self.k3.weight, None) * y
        return ag.Conv3.apply(out, self.k4.weight, None)

    def forward(self, x, x_b=None, chan_map=None):
        """x (+ x_b gathered by chan_map) are the 2*dim input channels (the reference's cat)."""
        # t is only read by k4: it stays in the channel-quad layout (16-byte stores / loads)
        t = ops.conv3x3(x, self.k3.weight, x_b=x_b, chan_map=chan_map, gate_w=self.k2.weight,
                        gate_b=self.k2.bias, out_c4=True)                      # :694-697
        return ops.conv3x3(t, self.k4.weight, in_c4=True)                      # :698


def nearest_channel_index(x: torch.Tensor, perception: torch.Tensor) -> torch.Tensor:
    """reference Matching (:618-680) with match_factor=1: argmin over candidate channel maps of
    the L2 distance between whole maps, from one Gram pass (fp64 tile accumulation)."""
    gram, nx, ny = ops.gram32(x, perception)
    # torch.cdist's mm mode: dist^2 = |x_i|^2 + |p_j|^2 - 2 x_i.p_j (sqrt/clamp are monotone)
    dist2 = nx[:, :, None] + ny[:, None, :] - 2.0 * gram
    return dist2.topk(k=1, largest=False).indices.squeeze(-1)


class Matching_transformation(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.paconv = PAConv(dim * 2)
        self.last_index = None  # kept for parity tests (argmin indices)

    def forward(self, x, perception):
        train = _wants_grad(self, x) or perception.requires_grad
        with torch.no_grad():
            if train:
                idx = nearest_channel_index(x.detach().contiguous(), perception.detach().contiguous())
            else:   # Gram pass + argmin in two launches, int32 indices
                idx = ops.match_index(x.contiguous(), perception.contiguous())
        self.last_index = idx
        if train:
            # the gather and the cat are data movement: torch autograd scatters the gradient back
            B, C, h, w = perception.shape
            cand = torch.gather(perception, 1, idx[:, :, None, None].expand(-1, -1, h, w))
            return self.paconv.forward_train(torch.cat([x, cand], dim=1))
        # cat([x, perception[idx]]) (:716) is expressed as a channel gather inside the conv
        return self.paconv(x, perception, idx)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.project_in = nn.Sequential(nn.Conv2d(dim, dim, 1), nn.Conv2d(dim, dim, 3, padding=1, groups=dim))
        self.matching_transformation = Matching_transformation(dim)
        self.project_out = nn.Sequential(nn.Conv2d(dim, dim, 3, padding=1, groups=dim), nn.GELU(),
                                         nn.Conv2d(dim, dim, 1))

    def forward_train(self, x, perception, norm: LayerNorm2d, residual):
        """x + FeedForward(norm2(x), per) (:745-751, :851), differentiable."""
        pi0, pi1 = self.project_in[0], self.project_in[1]
        t = ag.PW.apply(norm(x), pi0.weight, pi0.bias)
        t = ag.DW.apply(t, pi1.weight, pi1.bias)
        t = self.matching_transformation(t, perception)
        po0, po2 = self.project_out[0], self.project_out[2]
        t = F.gelu(ag.DW.apply(t, po0.weight, po0.bias))
        return residual + ag.PW.apply(t, po2.weight, po2.bias)

    def forward(self, x, perception, norm: LayerNorm2d, residual):
        if _wants_grad(self, x) or perception.requires_grad:
            return self.forward_train(x, perception, norm, residual)
        pi0, pi1 = self.project_in[0], self.project_in[1]
        t = ops.pw_dw(x, pi0.weight, pi0.bias, pi1.weight, pi1.bias, norm.weight, norm.bias, norm.eps)
        t = self.matching_transformation(t, perception)
        po0, po2 = self.project_out[0], self.project_out[2]
        return ops.dw_act_pw(t, po0.weight, po0.bias, po2.weight, po2.bias, "gelu", residual)


class CMTAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int = 1):
        super().__init__()
        if num_heads != 1:
            raise NotImplementedError("Wave-Mamba uses a single head")
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, padding=1, groups=dim * 3)
        self.project_out = nn.Conv2d(dim, dim, 1)
        self.matching_transformation = Matching_transformation(dim)

    def forward_train(self, x, perception, norm: LayerNorm2d, residual):
        """x + CMTAttention(norm1(x), per) (:772-798, :849), differentiable.  Same algebra as the
        inference path: one Gram matrix + norms instead of the normalised copies, the attention folded
        into the project_out weights."""
        B, C, h, w = x.shape
        qkv = ag.DW.apply(ag.PW.apply(norm(x), self.qkv.weight, self.qkv.bias),
                          self.qkv_dwconv.weight, self.qkv_dwconv.bias)
        q, k, v = qkv.chunk(3, dim=1)
        q = self.matching_transformation(q, perception)
        gram, nq2, nk2 = ag.Gram32.apply(q, k)
        nq = nq2.sqrt().clamp_min(1e-12)
        nk = nk2.sqrt().clamp_min(1e-12)
        attn = (gram / (nq[:, :, None] * nk[:, None, :]) * self.temperature).softmax(dim=-1)
        mixed = (self.project_out.weight.view(1, C, C, 1) * attn.unsqueeze(1)).sum(2)
        return ag.PWPerImage.apply(v, mixed, self.project_out.bias, residual)

    def forward(self, x, perception, norm: LayerNorm2d, residual):
        if _wants_grad(self, x) or perception.requires_grad:
            return self.forward_train(x, perception, norm, residual)
        B, C, h, w = x.shape
        qkv = ops.pw_dw(x, self.qkv.weight, self.qkv.bias, self.qkv_dwconv.weight,
                        self.qkv_dwconv.bias, norm.weight, norm.bias, norm.eps)
        q, k, v = qkv.chunk(3, dim=1)
        q = self.matching_transformation(q, perception)
        # normalize(q) @ normalize(k)^T (:787-790) == (q @ k^T) / (|q| |k|^T): one Gram matrix and
        # two norm reductions instead of materialising the normalised copies (eps 1e-12 as F.normalize)
        # and the 32x32 tail is folded into the Gram reduce kernel: normalisation, * temperature, softmax,
        # and project_out(attn @ v) == (W_po @ attn) @ v, i.e. the CxC attention folded into per-image 1x1
        # weights, so `attn @ v` (:793), project_out (:797) and the residual (:849) are ONE pass over v
        mixed = ops.attn_mixed(q, k, self.temperature, self.project_out.weight)        # (B, C, C)
        return ops.pw(v, mixed, self.project_out.bias, residual=residual)


class HFEBlock(nn.Module):
    def __init__(self, dim: int, **_unused):
        super().__init__()
        self.norm1 = LayerNorm2d(dim)
        self.attn = CMTAttention(dim)
        self.norm2 = LayerNorm2d(dim)
        self.ffn = FeedForward(dim)
        self.LayerNorm = LayerNorm2d(dim)

    def forward(self, x, perception):
        x = x.contiguous()
        per = self.LayerNorm(perception)
        x = self.attn(x, per, self.norm1, x)      # x + attn(norm1(x), per)   (:849)
        x = self.ffn(x, per, self.norm2, x)       # x + ffn(norm2(x), per)    (:851)
        return x


class SKFF(nn.Module):
    """reference :923-959 as two streaming kernels (wm_skff_fwd)."""

    def __init__(self, in_channels: int, height: int = 3, reduction: int = 8):
        super().__init__()
        d = max(int(in_channels / reduction), 4)
        self.conv_du = nn.Sequential(nn.Conv2d(in_channels, d, 1, bias=False), nn.PReLU())
        self.fcs = nn.ModuleList([nn.Conv2d(d, in_channels, 1, bias=False) for _ in range(height)])

    def forward_train(self, feats: List[torch.Tensor]):
        """SKFF.forward (:939-959), differentiable: three-band sum, global average pool, 32->4->3x32
        squeeze-excite on (B,32) vectors, softmax over the bands, weighted sum -- all elementwise /
        reductions (no convolution touches a feature map), left to torch autograd."""
        B, C = feats[0].shape[:2]
        stacked = torch.stack(feats, dim=1)
        pooled = stacked.sum(1).mean(dim=(2, 3))                                    # (B, C)
        z = F.prelu(pooled @ self.conv_du[0].weight.view(-1, C).t(), self.conv_du[1].weight)
        att = torch.stack([z @ fc.weight.view(C, -1).t() for fc in self.fcs], dim=1)  # (B, 3, C)
        att = att.softmax(dim=1)
        return (stacked * att[:, :, :, None, None]).sum(1)

    def forward(self, feats: List[torch.Tensor], pool=None):
        if _wants_grad(self, feats[0]):
            return self.forward_train(feats)
        return ops.skff(feats[0], feats[1], feats[2], self.conv_du[0].weight, self.conv_du[1].weight,
                        self.fcs[0].weight, self.fcs[1].weight, self.fcs[2].weight, pool=pool)


# --------------------------------------------------------------------------------------
# groups and the network                                        reference :962-1176
# --------------------------------------------------------------------------------------
def _run_low(blocks, x):
    """reference :976-979 / :998-1001 without the NCHW <-> (B,L,C) round trips."""
    x = x.contiguous()
    for blk in blocks:
        x = blk.forward_nchw(x)
    return x


class DownFRG(nn.Module):
    def __init__(self, dim, n_l_blocks=1, n_h_blocks=1, expand=2):
        super().__init__()
        self.dwt = DWT()
        self.l_conv = nn.Conv2d(dim * 2, dim, 3, 1, 1)
        self.l_blk = nn.Sequential(*[LFSSBlock(dim, expand=expand) for _ in range(n_l_blocks)])
        self.h_fusion = SKFF(dim, height=3, reduction=8)
        self.h_blk = nn.Sequential(*[HFEBlock(dim) for _ in range(n_h_blocks)])

    def forward(self, x, x_d):
        if _wants_grad(self, x):
            ll, hl, lh, hh = _DWTFn.apply(x)
            low = ag.Conv3.apply(torch.cat([ll, x_d], dim=1), self.l_conv.weight, self.l_conv.bias)
            low = _run_low(self.l_blk, low)
            high = self.h_fusion([hl, lh, hh])
            for blk in self.h_blk:
                high = blk(high, low)
            return low, high
        # the DWT also produces SKFF's pooled sums of its three high bands (:939-948) in its epilogue
        ll, hl, lh, hh, pool = ops.dwt_haar_pool(x.contiguous())
        low = ops.conv3x3(ll, self.l_conv.weight, self.l_conv.bias, x_b=x_d.contiguous())  # cat-free :975
        low = _run_low(self.l_blk, low)
        high = self.h_fusion([hl, lh, hh], pool=pool)
        for blk in self.h_blk:
            high = blk(high, low)
        return low, high


class upFRG(nn.Module):
    def __init__(self, dim, n_l_blocks=1, n_h_blocks=1, expand=2):
        super().__init__()
        self.iwt = IWT()
        self.l_blk = nn.Sequential(*[LFSSBlock(dim, expand=expand) for _ in range(n_l_blocks)])
        self.h_out_conv = nn.Conv2d(dim, dim * 3, 3, 1, 1)
        self.h_blk = nn.Sequential(*[HFEBlock(dim) for _ in range(n_h_blocks)])

    def forward(self, x_l, x_h):
        x_l = _run_low(self.l_blk, x_l)
        for blk in self.h_blk:
            x_h = blk(x_h, x_l)
        if _wants_grad(self, x_h):
            x_h = ag.Conv3.apply(x_h, self.h_out_conv.weight, self.h_out_conv.bias)
            return _IWTFn.apply(x_l, x_h)
        x_h = ops.conv3x3(x_h, self.h_out_conv.weight, self.h_out_conv.bias)       # :1005
        return self.iwt(x_l, x_h)  # IWT of cat([x_l, x_h]) without the cat (:1006)


def _activation_storage(module):
    """``module.activation_storage`` (None | torch.bfloat16 | "bf16"), else the environment variable
    WM_ACT_STORAGE=bf16 -- the switch for the reference's unmodified basicsr/train.py."""
    v = getattr(module, "activation_storage", None)
    if v is None:
        v = os.environ.get("WM_ACT_STORAGE") or None
    if v is None or v in ("fp32", "float32", torch.float32):
        return None
    if v in ("bf16", "bfloat16", torch.bfloat16):
        return torch.bfloat16
    raise ValueError(f"activation_storage: expected None, 'fp32' or 'bf16', got {v!r}")


class UNet(nn.Module):
    def __init__(self, in_chn=3, wf=48, n_l_blocks=(1, 1, 2), n_h_blocks=(1, 1, 1), ffn_scale=2):
        super().__init__()
        if wf != 32 or float(ffn_scale) != 2.0:
            raise NotImplementedError(
                "the sm_100a kernels are specialised for wf=32, ffn_scale=2 (d_inner=64), the "
                "configuration of every shipped checkpoint and YAML")
        for lvl, r in ((1, 2), (2, 4), (3, 8)):
            setattr(self, f"ps_down{lvl}", nn.Sequential(nn.PixelUnshuffle(r),
                                                          nn.Conv2d(r * r * in_chn, wf, 1, 1, 0)))
        self.conv_01 = nn.Conv2d(in_chn, wf, 3, 1, 1)
        for i in range(3):
            setattr(self, f"down_group{i + 1}", DownFRG(wf, n_l_blocks[i], n_h_blocks[i], ffn_scale))
        for i in (2, 1, 0):
            setattr(self, f"up_group{i + 1}", upFRG(wf, n_l_blocks[i], n_h_blocks[i], ffn_scale))
        self.last = nn.Conv2d(wf, in_chn, 3, 1, 1, bias=True)

    def forward_train(self, x):
        """UNet.forward (:1041-1063) with autograd: the same graph through wave_mamba_b200.autograd."""
        x = x.contiguous()
        side = []
        for lvl, r in ((1, 2), (2, 4), (3, 8)):
            conv = getattr(self, f"ps_down{lvl}")[1]
            side.append(ag.PSDown.apply(x, conv.weight, conv.bias, r))
        t = ag.StemConv.apply(x, self.conv_01.weight, self.conv_01.bias)
        low, h1 = self.down_group1(t, side[0])
        low, h2 = self.down_group2(low, side[1])
        low, h3 = self.down_group3(low, side[2])
        low = self.up_group3(low, h3)
        low = self.up_group2(low, h2)
        low = self.up_group1(low, h1)
        return ag.HeadConv.apply(low, self.last.weight, self.last.bias, x)

    def forward(self, x):
        _require_cuda(x)
        if x.dtype != torch.float32:
            raise TypeError(f"expected a float32 image tensor, got {x.dtype}")
        if x.dim() != 4 or x.shape[2] % 8 or x.shape[3] % 8:
            raise ValueError(f"input must be (B,C,H,W) with H and W multiples of 8, got {tuple(x.shape)}")
        if _wants_grad(self, x):
            if _activation_storage(self) is torch.bfloat16:
                with ag.bf16_activation_storage():
                    return self.forward_train(x)
            return self.forward_train(x)
        with torch.no_grad():
            x = x.contiguous()
            side = [ops.ps_down(x, getattr(self, f"ps_down{l}")[1].weight,
                                getattr(self, f"ps_down{l}")[1].bias, r)
                    for l, r in ((1, 2), (2, 4), (3, 8))]                       # :1043-1045
            t = ops.stem_conv3x3(x, self.conv_01.weight, self.conv_01.bias)
            low, h1 = self.down_group1(t, side[0])
            low, h2 = self.down_group2(low, side[1])
            low, h3 = self.down_group3(low, side[2])
            low = self.up_group3(low, h3)
            low = self.up_group2(low, h2)
            low = self.up_group1(low, h1)
            return ops.head_conv3x3(low, self.last.weight, self.last.bias, residual=x)


class WaveMamba(nn.Module):
    """Registry class; basicsr/archs/wavemamba_arch.py in plugin/ registers it in ARCH_REGISTRY."""

    def __init__(self, *, in_chn, wf, n_l_blocks=[1, 1, 2], n_h_blocks=[1, 1, 1], ffn_scale=2.0,
                 **ignore_kwargs):
        super().__init__()
        self.restoration_network = UNet(in_chn=in_chn, wf=wf, n_l_blocks=n_l_blocks,
                                        n_h_blocks=n_h_blocks, ffn_scale=ffn_scale)

    def print_network(self, model):
        print(model)
        print("The number of parameters: {}".format(sum(p.numel() for p in model.parameters())))

    def encode_and_decode(self, input, current_iter=None):
        return self.restoration_network(input)

    def check_image_size(self, x, window_size=8):
        _, _, h, w = x.size()
        pad_h = (window_size - h % window_size) % window_size
        pad_w = (window_size - w % window_size) % window_size
        return F.pad(x, (0, pad_w, 0, pad_h), "reflect")

    @torch.no_grad()
    def test(self, input):
        return self.encode_and_decode(input)

    @torch.no_grad()
    def test_tile(self, input, tile_size=240, tile_pad=16):
        """Spatial tiling (reference :1091-1151 reads an undefined ``self.scale_factor`` and
        crashes; the enhancement network is 1:1, so the scale factor is 1 here)."""
        B, C, H, W = input.shape
        out = input.new_zeros(B, C, H, W)
        for y0 in range(0, H, tile_size):
            for x0 in range(0, W, tile_size):
                y1, x1 = min(y0 + tile_size, H), min(x0 + tile_size, W)
                py0, px0 = max(y0 - tile_pad, 0), max(x0 - tile_pad, 0)
                py1, px1 = min(y1 + tile_pad, H), min(x1 + tile_pad, W)
                tile = input[:, :, py0:py1, px0:px1]
                th, tw = tile.shape[2:]
                tile = F.pad(tile, (0, (8 - tw % 8) % 8, 0, (8 - th % 8) % 8), "reflect")
                res = self.test(tile)[:, :, :th, :tw]
                out[:, :, y0:y1, x0:x1] = res[:, :, y0 - py0:y0 - py0 + (y1 - y0),
                                              x0 - px0:x0 - px0 + (x1 - x0)]
        return out

    def forward(self, input):
        return self.encode_and_decode(input)
