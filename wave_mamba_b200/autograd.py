"""Differentiable building blocks of the training path (the reference trains through
plain PyTorch autograd + mamba_ssm's selective_scan_fn, basicsr/models/femasr_model.py:157-185).

Every Function here runs hand-written CUDA in both directions:
  * data gradients of the 1x1 / dense 3x3 / depthwise convolutions are the SAME forward kernels run with
    transposed (and 180-degree rotated) weights;
  * weight gradients of the 1x1 and dense 3x3 convolutions are Gram matrices over the pixels
    (``ops.gram32``: 3xTF32 tensor-core products, fp64 accumulation) of the upstream gradient against the
    (shifted) input, 32 channels at a time;
  * LayerNorm-over-channels, the depthwise taps and the SS2D core have their own backward kernels
    (csrc/train.cu, csrc/ss2d_bwd.cu); the Haar pair is orthonormal, so each transform is the other's
    adjoint.
Pointwise glue (SiLU, GELU, gates, residual adds, the 32x32 softmax) stays with torch's elementwise
autograd.  This is an unfused, functional path -- correctness first; the inference path is the tuned one.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def gram_blocks(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sum over batch and pixels of a[:, i] * b[:, j]: (B,Ca,h,w), (B,Cb,h,w) -> (Ca,Cb); channel counts
    are padded up to multiples of 32 (ops.gram32 works on 32-channel stacks)."""
    Ca, Cb = a.shape[1], b.shape[1]
    pa, pb = (-Ca) % 32, (-Cb) % 32
    if pa:
        a = F.pad(a, (0, 0, 0, 0, 0, pa))
    if pb:
        b = F.pad(b, (0, 0, 0, 0, 0, pb))
    a, b = _c(a), _c(b)
    out = a.new_empty(a.shape[1], b.shape[1])
    for i in range(0, a.shape[1], 32):
        for j in range(0, b.shape[1], 32):
            g, _, _ = ops.gram32(a[:, i:i + 32], b[:, j:j + 32])
            out[i:i + 32, j:j + 32] = g.sum(0)
    return out[:Ca, :Cb]


def shifted_stack(x: torch.Tensor, sign: int = 1) -> torch.Tensor:
    """(B,C,h,w) -> (B,9C,h,w): tap t = 3*dy+dx holds x shifted so that entry p is x[p + sign*(t - center)]
    (zero outside), tap-major.  sign=+1: the im2col of a 3x3 / pad 1 convolution."""
    B, C, h, w = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    taps = []
    for dy in range(3):
        for dx in range(3):
            oy, ox = (dy, dx) if sign > 0 else (2 - dy, 2 - dx)
            taps.append(xp[:, :, oy:oy + h, ox:ox + w])
    return torch.cat(taps, dim=1)


def pw_data(x: torch.Tensor, w2d: torch.Tensor, bias=None) -> torch.Tensor:
    """1x1 convolution through ops.pw; Cin = 96 (the qkv data gradient) is summed 32 channels at a time."""
    Cout, Cin = w2d.shape
    if Cin == 96:
        y = None
        for i in range(0, 96, 32):
            y = ops.pw(x[:, i:i + 32], _c(w2d[:, i:i + 32]), bias if i == 0 else None, residual=y)
        return y
    return ops.pw(_c(x), _c(w2d), bias)


# --------------------------------------------------------------------------------------------
# bf16 activation storage: what autograd keeps for the backward pass is stored in bf16,
# everything that is computed -- forward values, scan state, gradients, weights -- stays fp32.
# --------------------------------------------------------------------------------------------
_BF16_MIN_NUMEL = 1 << 16      # feature maps only: 32x32 matrices, norms, weights and indices stay as they are


class _Bf16Saved:
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def _pack_bf16(t: torch.Tensor):
    if (t.dtype == torch.float32 and t.is_cuda and t.numel() >= _BF16_MIN_NUMEL
            and not isinstance(t, torch.nn.Parameter)):
        return _Bf16Saved(t.to(torch.bfloat16))
    return t


def _unpack_bf16(v):
    return v.t.to(torch.float32) if isinstance(v, _Bf16Saved) else v


def bf16_activation_storage():
    """Context manager for a training forward: every float32 feature map saved for backward (by the
    Functions below and by torch's elementwise glue alike) is kept as bf16 and widened again when the
    backward pass asks for it.  The forward result is unchanged bit for bit; gradients see activations
    rounded to 8 bits of mantissa (relative 2^-9), i.e. they differ from the fp32-storage gradients at the
    1e-2 level per element while pointing the same way (tests/test_train_gpu.py measures both)."""
    return torch.autograd.graph.saved_tensors_hooks(_pack_bf16, _unpack_bf16)


class PW(torch.autograd.Function):
    """y = conv1x1(x; w) + b."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _c(x)
        w2 = w.reshape(w.shape[0], w.shape[1])
        ctx.save_for_backward(x, w2)
        ctx.has_bias, ctx.wshape = b is not None, w.shape
        return pw_data(x, w2, b)

    @staticmethod
    def backward(ctx, gy):
        x, w2 = ctx.saved_tensors
        gy = _c(gy)
        gx = pw_data(gy, _c(w2.t())) if ctx.needs_input_grad[0] else None
        gw = gram_blocks(gy, x).reshape(ctx.wshape) if ctx.needs_input_grad[1] else None
        gb = gy.sum((0, 2, 3)) if ctx.has_bias else None
        return gx, gw, gb


class PWPerImage(torch.autograd.Function):
    """y[b] = W[b] x[b] + bias + residual[b]  (the CxC attention folded into project_out, :791-797,849)."""

    @staticmethod
    def forward(ctx, x, w, bias, residual):
        x, w = _c(x), _c(w)
        ctx.save_for_backward(x, w)
        return ops.pw(x, w, bias, residual=_c(residual))

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _c(gy)
        gx = ops.pw(gy, _c(w.transpose(1, 2)))
        gw, _, _ = ops.gram32(gy, x)                     # (B,32,32): sum_p gy[i,p] x[j,p]
        return gx, gw, gy.sum((0, 2, 3)), gy


class DW(torch.autograd.Function):
    """Depthwise 3x3 with bias."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _c(x)
        ctx.save_for_backward(x, w)
        return ops.dw3x3(x, w, b)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _c(gy)
        gx = ops.dw3x3(gy, w, None, flip=True)
        gw, gb = ops.dw3x3_wgrad(gy, x)
        return gx, gw.reshape(w.shape), gb


class LN2d(torch.autograd.Function):
    """LayerNorm over the channels of an NCHW tensor (C = 32 or 64)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        x = _c(x)
        ctx.save_for_backward(x, w)
        ctx.eps = eps
        return ops.layernorm2d(x, w, b, eps)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx, gw, gb = ops.layernorm2d_bwd(x, w, _c(gy), ctx.eps)
        return gx, gw, gb, None


class Conv3(torch.autograd.Function):
    """Dense 3x3 convolution, zero pad 1, optional bias (tcgen05 kernel in both directions)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _c(x)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return ops.conv3x3(x, w, b)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _c(gy)
        gx = ops.conv3x3(gy, _c(w.transpose(0, 1).flip(2, 3))) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            Cout, Cin = w.shape[0], w.shape[1]
            g = gram_blocks(gy, shifted_stack(x))            # (Cout, 9*Cin), tap-major
            gw = _c(g.reshape(Cout, 9, Cin).permute(0, 2, 1)).reshape(Cout, Cin, 3, 3)
        gb = gy.sum((0, 2, 3)) if ctx.has_bias else None
        return gx, gw, gb


class StemConv(torch.autograd.Function):
    """UNet.conv_01: 3 -> 32, 3x3 (the input image needs no gradient)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _c(x)
        ctx.save_for_backward(x)
        return ops.stem_conv3x3(x, w, b)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = _c(gy)
        g = gram_blocks(gy, shifted_stack(x))                # (32, 27)
        gw = _c(g.reshape(32, 9, 3).permute(0, 2, 1)).reshape(32, 3, 3, 3)
        return None, gw, gy.sum((0, 2, 3))


class HeadConv(torch.autograd.Function):
    """UNet.last + the global residual: 32 -> 3, 3x3."""

    @staticmethod
    def forward(ctx, x, w, b, image):
        x = _c(x)
        ctx.save_for_backward(x, w)
        return ops.head_conv3x3(x, w, b, residual=image)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _c(gy)
        gx = ops.stem_conv3x3(gy, _c(w.transpose(0, 1).flip(2, 3)))       # 3 -> 32 with rotated taps
        # dW[co,ci,t] = sum_p gy[co,p] x[ci,p+t]: the same Gram with the roles swapped
        g = gram_blocks(x, shifted_stack(gy, sign=-1))       # (32, 27): [ci, t*3 + co]
        gw = _c(g.reshape(32, 9, 3).permute(2, 0, 1)).reshape(3, 32, 3, 3)
        return gx, gw, gy.sum((0, 2, 3)), None


class PSDown(torch.autograd.Function):
    """PixelUnshuffle(r) + 1x1 conv on the input image (weights only need gradients)."""

    @staticmethod
    def forward(ctx, x, w, b, r):
        x = _c(x)
        ctx.save_for_backward(x)
        ctx.r, ctx.wshape = r, w.shape
        return ops.ps_down(x, w, b, r)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = _c(gy)
        gw = gram_blocks(gy, F.pixel_unshuffle(x, ctx.r)).reshape(ctx.wshape)
        return None, gw, gy.sum((0, 2, 3)), None


class SS2DCore(torch.autograd.Function):
    """SS2D.forward_core + the 4-way sum (wm_ss2d_core_fwd / wm_ss2d_core_bwd)."""

    @staticmethod
    def forward(ctx, x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds):
        x = _c(x)
        ctx.save_for_backward(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds)
        return ops.ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds)

    @staticmethod
    def backward(ctx, gy):
        return ops.ss2d_core_bwd(*ctx.saved_tensors, _c(gy))


class Gram32(torch.autograd.Function):
    """G = X Y^T over the pixels plus the squared row norms of both (ops.gram32)."""

    @staticmethod
    def forward(ctx, x, y):
        x, y = _c(x), _c(y)
        ctx.save_for_backward(x, y)
        g, nx, ny = ops.gram32(x, y)
        return g.clone(), nx.clone(), ny.clone()

    @staticmethod
    def backward(ctx, gg, gnx, gny):
        x, y = ctx.saved_tensors
        gg = _c(gg)
        gx = ops.pw(y, gg) + 2.0 * gnx[:, :, None, None] * x
        gy = ops.pw(x, _c(gg.transpose(1, 2))) + 2.0 * gny[:, :, None, None] * y
        return gx, gy
