"""Tensor-level wrappers over the C ABI (include/wavemamba_b200.h).

PyTorch is plumbing here: it owns device memory (caching allocator) and the current stream.
Every function checks that its tensors are CUDA / float32 / contiguous and raises otherwise;
nothing silently falls back to a PyTorch or CPU implementation.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch

from . import _cabi

# number of CUDA kernels this module has enqueued (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n: int) -> None:
    global launch_count
    launch_count += n


def _chk(t: torch.Tensor, name: str, shape: Optional[Tuple[int, ...]] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise _cabi.WaveMambaNativeError(
            f"{name} is on {t.device}; wave_mamba_b200 runs on CUDA (sm_100a) only -- no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor, got strides {t.stride()}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def dwt_haar(x: torch.Tensor):
    """(B,C,H,W) -> LL, HL, LH, HH each (B,C,H/2,W/2).  reference dwt_init :97-110."""
    _chk(x, "x")
    if x.dim() != 4:
        raise ValueError(f"x: expected 4 dims, got {x.dim()}")
    B, C, H, W = x.shape
    if H % 2 or W % 2:
        raise ValueError(f"DWT needs even H and W, got {H}x{W}")
    outs = [torch.empty(B, C, H // 2, W // 2, device=x.device, dtype=x.dtype) for _ in range(4)]
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_dwt_haar_fwd(x.data_ptr(), *[o.data_ptr() for o in outs], B * C, H, W, _stream(x))
    _cabi.check(rc, "wm_dwt_haar_fwd")
    _count(1)
    return tuple(outs)


def dwt_haar_pool(x: torch.Tensor):
    """dwt_haar plus SKFF's pool pass in the transform's epilogue (reference :939-948 pools (HL + LH) + HH
    right after the DWT): returns (LL, HL, LH, HH, pool) where ``pool`` is the per-CTA partial-sum buffer
    that ``skff(..., pool=pool)`` consumes instead of re-reading the three bands."""
    _chk(x, "x")
    if x.dim() != 4:
        raise ValueError(f"x: expected 4 dims, got {x.dim()}")
    B, C, H, W = x.shape
    if H % 2 or W % 2:
        raise ValueError(f"DWT needs even H and W, got {H}x{W}")
    if C != 32:
        raise ValueError(f"dwt_haar_pool: C={C} unsupported (32: the SKFF band width)")
    outs = [torch.empty(B, C, H // 2, W // 2, device=x.device, dtype=x.dtype) for _ in range(4)]
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        nbytes = lib.wm_skff_workspace_bytes(B, H // 2, W // 2)
        pool = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=x.device)
        rc = lib.wm_dwt_haar_pool_fwd(x.data_ptr(), *[o.data_ptr() for o in outs], pool.data_ptr(), nbytes, B * C,
                                      H, W, _stream(x))
    _cabi.check(rc, "wm_dwt_haar_pool_fwd")
    _count(1)
    return (*outs, pool)


def iwt_haar(low: torch.Tensor, high: torch.Tensor) -> torch.Tensor:
    """low (B,C,h,w) = LL, high (B,3C,h,w) = [HL|LH|HH] -> (B,C,2h,2w).
    reference iwt_init :113-130 applied to cat([low, high], 1) (:1006), without the cat."""
    _chk(low, "low")
    _chk(high, "high")
    B, C, h, w = low.shape
    if tuple(high.shape) != (B, 3 * C, h, w):
        raise ValueError(f"high: expected {(B, 3 * C, h, w)}, got {tuple(high.shape)}")
    y = torch.empty(B, C, 2 * h, 2 * w, device=low.device, dtype=low.dtype)
    lib = _cabi.load()
    with torch.cuda.device(low.device):
        rc = lib.wm_iwt_haar_fwd(low.data_ptr(), C * h * w, high.data_ptr(), 3 * C * h * w,
                                 y.data_ptr(), B, C, h, w, _stream(low))
    _cabi.check(rc, "wm_iwt_haar_fwd")
    _count(1)
    return y


def iwt_haar_cat(x: torch.Tensor) -> torch.Tensor:
    """The reference's calling convention: x = (B,4C,h,w) = [LL|HL|LH|HH]."""
    _chk(x, "x")
    B, C4, h, w = x.shape
    if C4 % 4:
        raise ValueError("IWT input channels must be a multiple of 4")
    C = C4 // 4
    y = torch.empty(B, C, 2 * h, 2 * w, device=x.device, dtype=x.dtype)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_iwt_haar_fwd(x.data_ptr(), C4 * h * w, x.data_ptr() + 4 * C * h * w,
                                 C4 * h * w, y.data_ptr(), B, C, h, w, _stream(x))
    _cabi.check(rc, "wm_iwt_haar_fwd")
    _count(1)
    return y


def ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds) -> torch.Tensor:
    """SS2D.forward_core + 4-way sum (reference :446-478,490).  x (B,64,h,w) -> y (B,64,h,w)."""
    _chk(x, "x")
    B, D, h, w = x.shape
    if D != 64:
        raise ValueError(f"ss2d_core supports d_inner=64 (wf=32, expand=2); got {D}")
    _chk(x_proj_weight, "x_proj_weight", (4, 34, 64))
    _chk(dt_projs_weight, "dt_projs_weight", (4, 64, 2))
    _chk(dt_projs_bias, "dt_projs_bias", (4, 64))
    _chk(A_logs, "A_logs", (256, 16))
    _chk(Ds, "Ds", (256,))
    lib = _cabi.load()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        nbytes = lib.wm_ss2d_core_workspace_bytes(B, h, w)
        ws = torch.empty(max(nbytes, 256), device=x.device, dtype=torch.uint8)
        rc = lib.wm_ss2d_core_fwd(x.data_ptr(), x_proj_weight.data_ptr(), dt_projs_weight.data_ptr(),
                                  dt_projs_bias.data_ptr(), A_logs.data_ptr(), Ds.data_ptr(),
                                  y.data_ptr(), ws.data_ptr(), nbytes, B, h, w, _stream(x))
    _cabi.check(rc, "wm_ss2d_core_fwd")
    _count(4)
    return y


def ss2d_core_bwd(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, grad_y):
    """Backward of ``ss2d_core`` (training; the reference gets it from autograd through
    selective_scan_fn and the einsums, :446-478,490).  Returns (grad_x, grad_x_proj_weight,
    grad_dt_projs_weight, grad_dt_projs_bias, grad_A_logs, grad_Ds)."""
    _chk(x, "x")
    B, D, h, w = x.shape
    if D != 64:
        raise ValueError(f"ss2d_core_bwd supports d_inner=64; got {D}")
    _chk(grad_y, "grad_y", (B, D, h, w))
    _chk(x_proj_weight, "x_proj_weight", (4, 34, 64))
    _chk(dt_projs_weight, "dt_projs_weight", (4, 64, 2))
    _chk(dt_projs_bias, "dt_projs_bias", (4, 64))
    _chk(A_logs, "A_logs", (256, 16))
    _chk(Ds, "Ds", (256,))
    lib = _cabi.load()
    gx = torch.empty_like(x)
    grads = [torch.empty_like(t) for t in (x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds)]
    with torch.cuda.device(x.device):
        nbytes = lib.wm_ss2d_core_bwd_workspace_bytes(B, h, w)
        ws = torch.empty(max(nbytes, 256), device=x.device, dtype=torch.uint8)
        rc = lib.wm_ss2d_core_bwd(x.data_ptr(), x_proj_weight.data_ptr(), dt_projs_weight.data_ptr(),
                                  dt_projs_bias.data_ptr(), A_logs.data_ptr(), Ds.data_ptr(),
                                  grad_y.data_ptr(), gx.data_ptr(), *[t.data_ptr() for t in grads],
                                  ws.data_ptr(), nbytes, B, h, w, _stream(x))
    _cabi.check(rc, "wm_ss2d_core_bwd")
    _count(16)
    return (gx, *grads)


def pipeline_error() -> int:
    """Developer aid: 0 when every mbarrier wait of the tcgen05 pipelines completed since the last
    call, else the code of the first wait that timed out (synchronises the device)."""
    import ctypes
    out = ctypes.c_uint(0)
    _cabi.check(_cabi.load().wm_debug_pipeline_error(ctypes.byref(out)), "wm_debug_pipeline_error")
    return int(out.value)


def ss2d_dirs(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds) -> torch.Tensor:
    """Un-merged SS2D core: returns the four direction outputs as a (4,B,64,h,w) view of the
    workspace (pixel order).  The reference sum y1+y2+y3+y4 is ((p[0]+p[2])+p[1])+p[3]."""
    _chk(x, "x")
    B, D, h, w = x.shape
    if D != 64:
        raise ValueError(f"ss2d_dirs supports d_inner=64 (wf=32, expand=2); got {D}")
    _chk(x_proj_weight, "x_proj_weight", (4, 34, 64))
    _chk(dt_projs_weight, "dt_projs_weight", (4, 64, 2))
    _chk(dt_projs_bias, "dt_projs_bias", (4, 64))
    _chk(A_logs, "A_logs", (256, 16))
    _chk(Ds, "Ds", (256,))
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        nbytes = lib.wm_ss2d_core_workspace_bytes(B, h, w)
        ws = torch.empty(max(nbytes, 256) // 4, device=x.device, dtype=torch.float32)
        rc = lib.wm_ss2d_dirs_fwd(x.data_ptr(), x_proj_weight.data_ptr(), dt_projs_weight.data_ptr(),
                                  dt_projs_bias.data_ptr(), A_logs.data_ptr(), Ds.data_ptr(),
                                  ws.data_ptr(), nbytes, B, h, w, _stream(x))
    _cabi.check(rc, "wm_ss2d_dirs_fwd")
    _count(3)
    return ws[:4 * B * D * h * w].view(4, B, D, h, w)


def layernorm2d(x, weight, bias, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm2d (reference :535-543) on NCHW."""
    _chk(x, "x")
    B, C, h, w = x.shape
    _chk(weight, "weight", (C,))
    _chk(bias, "bias", (C,))
    y = torch.empty_like(x)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_layernorm2d_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), eps,
                                    y.data_ptr(), B, C, h, w, _stream(x))
    _cabi.check(rc, "wm_layernorm2d_fwd")
    _count(1)
    return y


def pw_dw(x, pw_w, pw_b, dw_w, dw_b, ln_w=None, ln_b=None, eps: float = 1e-6,
          act: str = "none") -> torch.Tensor:
    """y = act(dw3x3(pw1x1(ln?(x)))): (B,32,h,w) -> (B,Cout,h,w), Cout in {32,64,96}.
    pw_w may be (Cout,Cin,1,1) or (Cout,Cin); pw_b may be None; act in {"none","silu"}."""
    _chk(x, "x")
    B, Cin, h, w = x.shape
    Cout = pw_w.shape[0]
    if pw_w.dim() == 2:
        _chk(pw_w, "pw_w", (Cout, Cin))
    else:
        _chk(pw_w, "pw_w", (Cout, Cin, 1, 1))
    if pw_b is not None:
        _chk(pw_b, "pw_b", (Cout,))
    _chk(dw_w, "dw_w", (Cout, 1, 3, 3))
    _chk(dw_b, "dw_b", (Cout,))
    if ln_w is not None:
        _chk(ln_w, "ln_w", (Cin,))
        _chk(ln_b, "ln_b", (Cin,))
    y = torch.empty(B, Cout, h, w, device=x.device, dtype=x.dtype)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_pw_dw_fwd(x.data_ptr(), _ptr(ln_w), _ptr(ln_b), eps, pw_w.data_ptr(),
                              _ptr(pw_b), dw_w.data_ptr(), dw_b.data_ptr(),
                              {"none": 0, "silu": 1}[act], y.data_ptr(), B, Cin, Cout, h, w,
                              _stream(x))
    _cabi.check(rc, "wm_pw_dw_fwd")
    _count(1)
    return y


def dw_act_pw(x, dw_w, dw_b, pw_w, pw_b, act: str = "gelu", residual=None) -> torch.Tensor:
    """y = residual? + pw1x1(act(dw3x3(x))), C=32."""
    _chk(x, "x")
    B, C, h, w = x.shape
    _chk(dw_w, "dw_w", (C, 1, 3, 3))
    _chk(dw_b, "dw_b", (C,))
    _chk(pw_w, "pw_w", (C, C, 1, 1))
    _chk(pw_b, "pw_b", (C,))
    if residual is not None:
        _chk(residual, "residual", tuple(x.shape))
    y = torch.empty_like(x)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_dw_act_pw_fwd(x.data_ptr(), dw_w.data_ptr(), dw_b.data_ptr(), pw_w.data_ptr(),
                                  pw_b.data_ptr(), {"none": 0, "gelu": 1}[act], _ptr(residual),
                                  y.data_ptr(), B, C, h, w, _stream(x))
    _cabi.check(rc, "wm_dw_act_pw_fwd")
    _count(1)
    return y


def pw(x, pw_w, pw_b=None, gate: bool = False, residual=None, res_scale=None, out=None) -> torch.Tensor:
    """y = residual?*res_scale? + pw1x1(x) (+bias).  gate=True: x is (B,2*Cin,h,w) and the conv
    sees gelu(x[:, :Cin]) * x[:, Cin:]  (reference ffn :227-228).
    x may be a channel slice of a wider tensor (planes contiguous); pw_w may be (Cout,Cin[,1,1]) or
    per-image (B,Cout,Cin)."""
    _chk_planes(x, "x")
    B, Cx, h, w = x.shape
    per_image = pw_w.dim() == 3
    Cout, Cin = (pw_w.shape[1], pw_w.shape[2]) if per_image else (pw_w.shape[0], pw_w.shape[1])
    if Cx != (2 * Cin if gate else Cin):
        raise ValueError(f"x has {Cx} channels, weight expects {2 * Cin if gate else Cin}")
    if per_image:
        _chk(pw_w, "pw_w", (B, Cout, Cin))
    else:
        _chk(pw_w, "pw_w", (Cout, Cin) if pw_w.dim() == 2 else (Cout, Cin, 1, 1))
    if pw_b is not None:
        _chk(pw_b, "pw_b", (Cout,))
    if residual is not None:
        _chk(residual, "residual", (B, Cout, h, w))
    if res_scale is not None:
        _chk(res_scale, "res_scale", (Cout,))
    if out is None:
        y = torch.empty(B, Cout, h, w, device=x.device, dtype=x.dtype)
    else:
        y = _chk(out, "out", (B, Cout, h, w))
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_pw_fwd(x.data_ptr(), x.stride(0) if B > 1 else 0, pw_w.data_ptr(),
                           Cout * Cin if per_image else 0, _ptr(pw_b), 1 if gate else 0,
                           _ptr(residual), _ptr(res_scale), y.data_ptr(), B, Cin, Cout, h, w,
                           _stream(x))
    _cabi.check(rc, "wm_pw_fwd")
    _count(1)
    return y


def paconv_gate(x, k2_w, k2_b, k3out, inplace: bool = True) -> torch.Tensor:
    """y = k3out * sigmoid(pw1x1(x) + b)  (reference PAConv :694-697), all (B,64,h,w)."""
    _chk(x, "x")
    B, C, h, w = x.shape
    _chk(k2_w, "k2_w", (C, C, 1, 1))
    _chk(k2_b, "k2_b", (C,))
    _chk(k3out, "k3out", tuple(x.shape))
    y = k3out if inplace else torch.empty_like(k3out)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_paconv_gate_fwd(x.data_ptr(), k2_w.data_ptr(), k2_b.data_ptr(),
                                    k3out.data_ptr(), y.data_ptr(), B, C, h, w, _stream(x))
    _cabi.check(rc, "wm_paconv_gate_fwd")
    _count(1)
    return y


def lfss_z(x, ln_w, ln_b, eps, in_proj_weight) -> torch.Tensor:
    """zs = silu(in_proj.weight[64:] . LayerNorm_c(x)): (B,32,h,w) -> (B,64,h,w)
    (reference ln_1 :524, in_proj/chunk :483-484, F.silu(z) :493)."""
    _chk(x, "x")
    B, C, h, w = x.shape
    _chk(in_proj_weight, "in_proj_weight", (4 * C, C))
    _chk(ln_w, "ln_w", (C,))
    _chk(ln_b, "ln_b", (C,))
    zs = torch.empty(B, 2 * C, h, w, device=x.device, dtype=x.dtype)
    lib = _cabi.load()
    w_z = in_proj_weight.data_ptr() + 2 * C * C * 4  # rows 64..127
    with torch.cuda.device(x.device):
        rc = lib.wm_lfss_z_fwd(x.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), eps, w_z,
                               zs.data_ptr(), B, h, w, _stream(x))
    _cabi.check(rc, "wm_lfss_z_fwd")
    _count(1)
    return zs


def lfss_out(y, zs, on_w, on_b, eps, out_proj_weight, x, skip_scale, extra=()) -> torch.Tensor:
    """out = x*skip_scale + out_proj(out_norm(((y + e0) + e1) + e2) * zs)  (reference :490-494,
    :525).  ``extra``: up to three more (B,64,h,w) addends, summed in the given order."""
    _chk(y, "y")
    B, D, h, w = y.shape
    _chk(zs, "zs", (B, D, h, w))
    extra = list(extra) + [None] * (3 - len(extra))
    for i, t in enumerate(extra):
        if t is not None:
            _chk(t, f"extra[{i}]", (B, D, h, w))
    C = out_proj_weight.shape[0]
    _chk(out_proj_weight, "out_proj_weight", (C, D))
    _chk(on_w, "on_w", (D,))
    _chk(on_b, "on_b", (D,))
    _chk(x, "x", (B, C, h, w))
    _chk(skip_scale, "skip_scale", (C,))
    out = torch.empty_like(x)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_lfss_out_fwd(y.data_ptr(), _ptr(extra[0]), _ptr(extra[1]), _ptr(extra[2]),
                                 zs.data_ptr(), on_w.data_ptr(),
                                 on_b.data_ptr(), eps, out_proj_weight.data_ptr(), x.data_ptr(),
                                 skip_scale.data_ptr(), out.data_ptr(), B, h, w, _stream(x))
    _cabi.check(rc, "wm_lfss_out_fwd")
    _count(1)
    return out


def lfss_tail(planes, x, ln_w, ln_b, ln_eps, in_proj_weight, on_w, on_b, on_eps, out_proj_weight,
              skip_scale) -> torch.Tensor:
    """out = x*skip_scale + out_proj(out_norm(((p0 + p1) + p2) + p3) * silu(in_proj.weight[64:] . ln_1(x)))
    -- lfss_z and lfss_out in one kernel (reference :483-484, :490-494, :524-525).  ``planes``: the four
    (B,64,h,w) direction outputs in summation order."""
    _chk(x, "x")
    B, C, h, w = x.shape
    D = 2 * C
    if len(planes) != 4:
        raise ValueError("lfss_tail: four direction planes expected")
    for i, t in enumerate(planes):
        _chk(t, f"planes[{i}]", (B, D, h, w))
    _chk(in_proj_weight, "in_proj_weight", (4 * C, C))
    _chk(out_proj_weight, "out_proj_weight", (C, D))
    for name, t, n in (("ln_w", ln_w, C), ("ln_b", ln_b, C), ("on_w", on_w, D), ("on_b", on_b, D),
                       ("skip_scale", skip_scale, C)):
        _chk(t, name, (n,))
    out = torch.empty_like(x)
    # the one-kernel form needs h*w % 4 == 0 and 16-byte aligned tensors; other shapes go through a scratch z
    fused = (h * w) % 4 == 0 and h * w >= 64 and all(t.data_ptr() % 16 == 0 for t in (*planes, x, out))
    scratch = None if fused else torch.empty(B, D, h, w, device=x.device, dtype=x.dtype)
    lib = _cabi.load()
    w_z = in_proj_weight.data_ptr() + 2 * C * C * 4  # rows 64..127
    with torch.cuda.device(x.device):
        rc = lib.wm_lfss_tail_fwd(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(),
                                  planes[3].data_ptr(), x.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), ln_eps,
                                  w_z, on_w.data_ptr(), on_b.data_ptr(), on_eps, out_proj_weight.data_ptr(),
                                  skip_scale.data_ptr(), _ptr(scratch), out.data_ptr(), B, h, w, _stream(x))
    _cabi.check(rc, "wm_lfss_tail_fwd")
    _count(1 if fused else 2)
    return out


def _chk_planes32(t: torch.Tensor, name: str):
    """(B,32,h,w) float32 CUDA tensor whose 32 planes are contiguous (batch stride free)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _cabi.WaveMambaNativeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] != 32:
        raise TypeError(f"{name}: expected float32 (B,32,h,w), got {t.dtype} {tuple(t.shape)}")
    B, C, h, w = t.shape
    if t.numel() and (t.stride(3) != 1 or t.stride(2) != w or t.stride(1) != h * w):
        raise ValueError(f"{name}: channel planes must be contiguous, got strides {t.stride()}")
    return t


def gram32(x: torch.Tensor, y: torch.Tensor):
    """G = X Y^T over all pixels plus squared row norms.  x, y: (B,32,h,w) (channel slices of a
    wider tensor are fine).  Returns G (B,32,32), |x_i|^2 (B,32), |y_j|^2 (B,32).
    Reference: the mm-mode torch.cdist of Matching (:664) and the q.k^T / F.normalize of
    CMTAttention (:787-790)."""
    _chk_planes32(x, "x")
    _chk_planes32(y, "y")
    if x.shape != y.shape:
        raise ValueError(f"x {tuple(x.shape)} and y {tuple(y.shape)} must match")
    B, C, h, w = x.shape
    hw = h * w
    lib = _cabi.load()
    out = torch.empty(B, 32 * 32 + 64, device=x.device, dtype=torch.float32)
    xb = x.stride(0) if B > 1 else 32 * hw
    yb = y.stride(0) if B > 1 else 32 * hw
    with torch.cuda.device(x.device):
        nbytes = lib.wm_gram32_workspace_bytes(B, hw)
        ws = torch.empty(max(nbytes, 8) // 8, device=x.device, dtype=torch.float64)
        rc = lib.wm_gram32_fwd(x.data_ptr(), xb, y.data_ptr(), yb, out.data_ptr(), ws.data_ptr(),
                               nbytes, B, hw, _stream(x))
    _cabi.check(rc, "wm_gram32_fwd")
    _count(2)
    return out[:, :1024].view(B, 32, 32), out[:, 1024:1056], out[:, 1056:1088]


def _gram_call(fn_name, x, y, extra_args):
    _chk_planes32(x, "x")
    _chk_planes32(y, "y")
    if x.shape != y.shape:
        raise ValueError(f"x {tuple(x.shape)} and y {tuple(y.shape)} must match")
    B, C, h, w = x.shape
    hw = h * w
    lib = _cabi.load()
    xb = x.stride(0) if B > 1 else 32 * hw
    yb = y.stride(0) if B > 1 else 32 * hw
    with torch.cuda.device(x.device):
        nbytes = lib.wm_gram32_workspace_bytes(B, hw)
        ws = torch.empty(max(nbytes, 8) // 8, device=x.device, dtype=torch.float64)
        rc = getattr(lib, fn_name)(x.data_ptr(), xb, y.data_ptr(), yb, *extra_args, None, ws.data_ptr(), nbytes,
                                   B, hw, _stream(x))
    _cabi.check(rc, fn_name)
    _count(2)


def match_index(x: torch.Tensor, perception: torch.Tensor) -> torch.Tensor:
    """Matching (reference :618-680, match_factor 1): for every channel map of x the index of the nearest
    channel map of ``perception`` (L2 over the whole map), (B,32) int32 -- the Gram pass and the argmin
    of `|x_i|^2 + |p_j|^2 - 2 x_i.p_j` in two launches."""
    idx = torch.empty(x.shape[0], 32, device=x.device, dtype=torch.int32)
    _gram_call("wm_gram32_match_fwd", x, perception, (idx.data_ptr(),))
    return idx


def attn_mixed(q: torch.Tensor, k: torch.Tensor, temperature: torch.Tensor, w_po: torch.Tensor) -> torch.Tensor:
    """CMTAttention (reference :787-797): W_po . softmax(normalize(q) normalize(k)^T * temperature), (B,32,32):
    the per-image 1x1 weights that fold attention and project_out into one pass over v (``ops.pw``)."""
    _chk(temperature, "temperature")
    if temperature.numel() != 1:
        raise ValueError("attn_mixed: single-head temperature expected")
    w2 = w_po.reshape(32, 32)
    _chk(w2, "w_po", (32, 32))
    mixed = torch.empty(q.shape[0], 32, 32, device=q.device, dtype=torch.float32)
    _gram_call("wm_gram32_attn_fwd", q, k, (temperature.data_ptr(), w2.data_ptr(), mixed.data_ptr()))
    return mixed


_packed_cache = {}   # id(weight) -> (weakref, (data_ptr, device, version) of w3x3 and w1x1, packed)


def _w_state(t: Optional[torch.Tensor]):
    """What a cached pack is valid for.  ``_version`` alone misses writes through ``.data`` /
    ``module.to()``, which keep the Parameter object: the storage address and device are compared
    too, and code that rewrites weights behind autograd's back (``p.data.copy_``) must call
    ``clear_pack_cache()`` (parallel.broadcast_parameters does)."""
    return None if t is None else (t.data_ptr(), t.device, t._version)


def clear_pack_cache() -> None:
    """Drop every cached pre-packed conv weight (call after modifying weights through ``.data``)."""
    _packed_cache.clear()


def conv3x3_pack(w3x3: torch.Tensor, w1x1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pre-pack (Cout,Cin,3,3) weights (+ optional (Cout,Cin[,1,1]) 1x1 gate weights) into the
    tcgen05 operand order with the tf32 hi/lo split.  Cached per weight tensor object (weakly);
    a hit is validated against the storage address, device and ``_version`` of both tensors."""
    _chk(w3x3, "w3x3")
    Cout, Cin = w3x3.shape[0], w3x3.shape[1]
    key = id(w3x3)
    state = (_w_state(w3x3), _w_state(w1x1))
    hit = _packed_cache.get(key)
    if hit is not None:
        ref3, ref1, st, packed = hit
        same_gate = (ref1 is None and w1x1 is None) or (ref1 is not None and ref1() is w1x1)
        if ref3() is w3x3 and same_gate and st == state and packed.device == w3x3.device:
            return packed
    if w1x1 is not None:
        _chk(w1x1, "w1x1")
        if w1x1.numel() != Cout * Cin:
            raise ValueError("w1x1 must be (Cout, Cin[,1,1])")
    lib = _cabi.load()
    nbytes = lib.wm_conv3x3_packed_bytes(Cin, Cout, 1 if w1x1 is not None else 0)
    if nbytes == 0:
        raise ValueError(f"conv3x3: Cin={Cin} must be a multiple of 32 and Cout={Cout} of 8")
    packed = torch.empty(nbytes // 4, device=w3x3.device, dtype=torch.float32)
    with torch.cuda.device(w3x3.device):
        rc = lib.wm_conv3x3_prepack(w3x3.data_ptr(), _ptr(w1x1), packed.data_ptr(), Cin, Cout,
                                    _stream(w3x3))
    _cabi.check(rc, "wm_conv3x3_prepack")
    _count(1)
    _packed_cache[key] = (weakref.ref(w3x3, lambda _r, k=key: _packed_cache.pop(k, None)),
                          None if w1x1 is None else weakref.ref(w1x1), state, packed)
    return packed


def conv3x3(x_a, w3x3, bias=None, x_b=None, chan_map=None, gate_w=None, gate_b=None,
            in_c4: bool = False, out_c4: bool = False) -> torch.Tensor:
    """Dense 3x3 conv, stride 1, zero pad 1 (3xTF32 tensor-core implicit GEMM, fp32-accurate).

    ``in_c4`` / ``out_c4``: x_a / the result use the channel-quad layout (B, C/4, h, w, 4) (tcgen05
    implementation only; an intermediate that only the next conv reads, e.g. PAConv k3 -> k4).

    Input channels = x_a's channels followed by x_b's (optionally gathered per batch item through
    ``chan_map`` (B, Cb) int32) -- the reference's torch.cat is not materialised.
    ``gate_w``/``gate_b``: PAConv stage A, returns conv3x3(x) * sigmoid(conv1x1(x; gate_w) + gate_b)."""
    if in_c4:
        _chk(x_a, "x_a")
        if x_a.dim() != 5 or x_a.shape[4] != 4 or x_b is not None:
            raise ValueError("in_c4: x_a must be (B, C/4, h, w, 4) and the only input")
        B, Ca, h, w = x_a.shape[0], x_a.shape[1] * 4, x_a.shape[2], x_a.shape[3]
    else:
        _chk_planes(x_a, "x_a")
        B, Ca, h, w = x_a.shape
    Cout, Cin = w3x3.shape[0], w3x3.shape[1]
    Cb = Cin - Ca
    if Cb < 0 or (Cb > 0 and x_b is None):
        raise ValueError(f"conv3x3: weight expects {Cin} input channels, got {Ca} (+ x_b)")
    if x_b is not None:
        _chk_planes(x_b, "x_b")
        if x_b.shape[0] != B or tuple(x_b.shape[2:]) != (h, w):
            raise ValueError("x_b must match x_a in batch and spatial size")
        if chan_map is None and x_b.shape[1] != Cb:
            raise ValueError(f"x_b has {x_b.shape[1]} channels, expected {Cb}")
    if chan_map is not None:
        if chan_map.dtype != torch.int32 or tuple(chan_map.shape) != (B, Cb) or not chan_map.is_cuda \
                or not chan_map.is_contiguous():
            raise ValueError("chan_map must be a contiguous CUDA int32 tensor of shape (B, Cin - Ca)")
    if bias is not None:
        _chk(bias, "bias", (Cout,))
    if (gate_w is None) != (gate_b is None):
        raise ValueError("gate_w and gate_b come together")
    if gate_b is not None:
        _chk(gate_b, "gate_b", (Cout,))
    packed = conv3x3_pack(w3x3, gate_w)
    out = torch.empty((B, Cout // 4, h, w, 4) if out_c4 else (B, Cout, h, w), device=x_a.device,
                      dtype=torch.float32)
    lib = _cabi.load()
    with torch.cuda.device(x_a.device):
        rc = lib.wm_conv3x3_ex_fwd(x_a.data_ptr(), x_a.stride(0) if B > 1 else Ca * h * w, Ca,
                                   _ptr(x_b), 0 if x_b is None else (x_b.stride(0) if B > 1 else 0),
                                   _ptr(chan_map), packed.data_ptr(), _ptr(bias), _ptr(gate_b),
                                   out.data_ptr(), B, Cin, Cout, h, w, int(in_c4), int(out_c4),
                                   _stream(x_a))
    _cabi.check(rc, "wm_conv3x3_ex_fwd")
    _count(1)
    return out


def _chk_planes(t: torch.Tensor, name: str):
    """(B,C,h,w) float32 CUDA tensor whose channel planes are contiguous (batch stride free)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _cabi.WaveMambaNativeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 4:
        raise TypeError(f"{name}: expected float32 (B,C,h,w), got {t.dtype} {tuple(t.shape)}")
    B, C, h, w = t.shape
    if t.numel() and (t.stride(3) != 1 or t.stride(2) != w or t.stride(1) != h * w):
        raise ValueError(f"{name}: channel planes must be contiguous, got strides {t.stride()}")
    return t


def stem_conv3x3(x, weight, bias=None) -> torch.Tensor:
    """UNet.conv_01 (reference :1026,1048): (B,3,h,w) -> (B,32,h,w)."""
    _chk(x, "x")
    B, C, h, w = x.shape
    _chk(weight, "weight", (32, 3, 3, 3))
    if C != 3:
        raise ValueError(f"stem conv expects 3 input channels, got {C}")
    if bias is not None:
        _chk(bias, "bias", (32,))
    y = torch.empty(B, 32, h, w, device=x.device, dtype=x.dtype)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_stem_conv3x3_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), y.data_ptr(),
                                     B, h, w, _stream(x))
    _cabi.check(rc, "wm_stem_conv3x3_fwd")
    _count(1)
    return y


def head_conv3x3(x, weight, bias=None, residual=None) -> torch.Tensor:
    """UNet.last + global residual (reference :1039,1061): (B,32,h,w) -> (B,3,h,w)."""
    _chk(x, "x")
    B, C, h, w = x.shape
    _chk(weight, "weight", (3, 32, 3, 3))
    if bias is not None:
        _chk(bias, "bias", (3,))
    if residual is not None:
        _chk(residual, "residual", (B, 3, h, w))
    y = torch.empty(B, 3, h, w, device=x.device, dtype=x.dtype)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_head_conv3x3_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), _ptr(residual),
                                     y.data_ptr(), B, h, w, _stream(x))
    _cabi.check(rc, "wm_head_conv3x3_fwd")
    _count(1)
    return y


def skff(f0, f1, f2, w_du, prelu_weight, w_fc0, w_fc1, w_fc2, pool=None) -> torch.Tensor:
    """SKFF over the three high-frequency bands (reference :939-959): two streaming kernels, or one when
    ``pool`` (from ``dwt_haar_pool`` of the same bands) already holds the pool pass."""
    _chk(f0, "f0")
    B, C, h, w = f0.shape
    _chk(f1, "f1", (B, C, h, w))
    _chk(f2, "f2", (B, C, h, w))
    d = w_du.shape[0]
    if C != 32 or d != 4:
        raise ValueError(f"skff: C={C}, d={d} unsupported (32, 4)")
    w_du = _chk(w_du.reshape(d, C), "w_du")
    fcs = [_chk(t.reshape(C, d), "w_fc") for t in (w_fc0, w_fc1, w_fc2)]
    _chk(prelu_weight, "prelu_weight", (1,))
    out = torch.empty_like(f0)
    lib = _cabi.load()
    with torch.cuda.device(f0.device):
        nbytes = lib.wm_skff_workspace_bytes(B, h, w)
        if pool is not None:
            if pool.dtype != torch.float64 or pool.numel() * 8 < nbytes or pool.device != f0.device:
                raise ValueError("skff: pool is not the buffer dwt_haar_pool returned for these bands")
            rc = lib.wm_skff_apply_fwd(f0.data_ptr(), f1.data_ptr(), f2.data_ptr(), w_du.data_ptr(),
                                       prelu_weight.data_ptr(), fcs[0].data_ptr(), fcs[1].data_ptr(),
                                       fcs[2].data_ptr(), out.data_ptr(), pool.data_ptr(), nbytes, B, C, h, w,
                                       _stream(f0))
            _cabi.check(rc, "wm_skff_apply_fwd")
        else:
            ws = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=f0.device)
            rc = lib.wm_skff_fwd(f0.data_ptr(), f1.data_ptr(), f2.data_ptr(), w_du.data_ptr(),
                                 prelu_weight.data_ptr(), fcs[0].data_ptr(), fcs[1].data_ptr(),
                                 fcs[2].data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, B, C, h, w,
                                 _stream(f0))
            _cabi.check(rc, "wm_skff_fwd")
    if B and h and w:
        _count(1 if pool is not None else 2)
    return out


def ps_down(x, weight, bias, r: int) -> torch.Tensor:
    """PixelUnshuffle(r) + 1x1 conv 3 r^2 -> 32 (reference :1014-1025) without the unshuffled copy."""
    _chk(x, "x")
    B, C, H, W = x.shape
    if C != 3:
        raise ValueError(f"ps_down: expected 3 input channels, got {C}")
    weight = _chk(weight.reshape(32, 3 * r * r), "weight")
    if bias is not None:
        _chk(bias, "bias", (32,))
    if H % r or W % r:
        raise ValueError(f"ps_down: H, W must be multiples of r={r}, got {(H, W)}")
    y = torch.empty(B, 32, H // r, W // r, dtype=x.dtype, device=x.device)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_ps_down_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), y.data_ptr(), B, H, W, r,
                                _stream(x))
    _cabi.check(rc, "wm_ps_down_fwd")
    if B and H and W:
        _count(1)
    return y


def _chk_u8(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8:
        raise TypeError(f"{name}: expected a uint8 torch.Tensor")
    if not t.is_cuda:
        raise _cabi.WaveMambaNativeError(f"{name} is on {t.device}; wave_mamba_b200 runs on CUDA only")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


def img_u8_to_f32(img: torch.Tensor, window: int = 128, cuda_division: bool = False) -> torch.Tensor:
    """(B,H,W,3) uint8 BGR -> (B,3,Hp,Wp) float32 RGB in [0,1], reflect-padded up to multiples of
    `window`: img2tensor + /255. + check_image_size of inference_wavemamba.py, bit-exact.
    ``cuda_division=False``: IEEE ``byte / 255`` (the line evaluated on a CPU tensor, and the oracle);
    ``True``: ``byte * (1/255)``, what torch computes for the same line on a CUDA tensor -- the
    reference script's own GPU run (a few bytes per million differ by one ulp)."""
    _chk_u8(img, "img")
    if img.dim() != 4 or img.shape[3] != 3:
        raise ValueError(f"img: expected (B,H,W,3), got {tuple(img.shape)}")
    B, H, W, _ = img.shape
    Hp, Wp = -(-H // window) * window, -(-W // window) * window
    if B and H and W and (Hp - H >= H or Wp - W >= W):
        raise ValueError("reflect padding must be smaller than the image")
    out = torch.empty(B, 3, Hp, Wp, dtype=torch.float32, device=img.device)
    lib = _cabi.load()
    with torch.cuda.device(img.device):
        rc = lib.wm_img_u8_to_f32_fwd(img.data_ptr(), out.data_ptr(), B, H, W, Hp, Wp,
                                      1 if cuda_division else 0, _stream(img))
    _cabi.check(rc, "wm_img_u8_to_f32_fwd")
    if B and H and W:
        _count(1)
    return out


def img_f32_to_u8(x: torch.Tensor, h: Optional[int] = None, w: Optional[int] = None) -> torch.Tensor:
    """(B,3,Hs,Ws) float32 RGB -> (B,h,w,3) uint8 BGR: crop + tensor2img, bit-exact."""
    _chk(x, "x")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"x: expected (B,3,H,W), got {tuple(x.shape)}")
    B, _, Hs, Ws = x.shape
    h = Hs if h is None else h
    w = Ws if w is None else w
    if h > Hs or w > Ws:
        raise ValueError("crop larger than the source")
    img = torch.empty(B, h, w, 3, dtype=torch.uint8, device=x.device)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        rc = lib.wm_img_f32_to_u8_fwd(x.data_ptr(), img.data_ptr(), B, h, w, Hs, Ws, _stream(x))
    _cabi.check(rc, "wm_img_f32_to_u8_fwd")
    if B and h and w:
        _count(1)
    return img


def psnr_ssim_y(img1: torch.Tensor, img2: torch.Tensor, crop_border: int = 1) -> torch.Tensor:
    """PSNR and SSIM on the Y channel of two (B,H,W,3) uint8 BGR image batches on the device, as
    comput_psnr_ssim.py calculate_psnr :387-438 / calculate_ssim :596-668 compute them with their
    defaults.  Returns (B,2) float64 = [psnr, ssim] per image (psnr = +inf for identical Y planes)."""
    for t, name in ((img1, "img1"), (img2, "img2")):
        if not t.is_cuda:
            raise _cabi.WaveMambaNativeError(f"{name}: the B200 path needs a CUDA tensor")
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3 or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous (B,H,W,3) uint8 tensor, got {t.dtype} {tuple(t.shape)}")
    if img1.shape != img2.shape:
        raise ValueError(f"Image shapes are different: {tuple(img1.shape)}, {tuple(img2.shape)}")
    if img1.device != img2.device:
        raise ValueError("img1 and img2 live on different devices")
    B, H, W, _ = img1.shape
    crop_border = int(crop_border)
    if crop_border < 0 or H - 2 * crop_border < 1 or W - 2 * crop_border < 1:
        raise ValueError(f"crop_border={crop_border} leaves nothing of a {H}x{W} image")
    out = torch.empty(B, 2, dtype=torch.float64, device=img1.device)
    lib = _cabi.load()
    nbytes = lib.wm_psnr_ssim_y_workspace_bytes(B, H, W, crop_border)
    ws = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=img1.device)
    with torch.cuda.device(img1.device):
        rc = lib.wm_psnr_ssim_y_u8(img1.data_ptr(), img2.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes,
                                   B, H, W, crop_border, _stream(img1))
    _cabi.check(rc, "wm_psnr_ssim_y_u8")
    if B:
        _count(2)
    return out


# --------------------------------------------------------------------------------------------
# training-only kernels (csrc/train.cu)
# --------------------------------------------------------------------------------------------
def dw3x3(x, weight, bias=None, flip: bool = False) -> torch.Tensor:
    """Depthwise 3x3, zero pad 1: x (B,C,h,w), weight (C,1,3,3) or (C,9).  ``flip`` rotates the taps by
    180 degrees (the data gradient of the same convolution)."""
    _chk(x, "x")
    B, C, h, w = x.shape
    weight = _chk(weight.reshape(C, 9), "weight")
    if bias is not None:
        _chk(bias, "bias", (C,))
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _cabi.load().wm_dw3x3_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), y.data_ptr(), B, C, h, w,
                                       1 if flip else 0, _stream(x))
    _cabi.check(rc, "wm_dw3x3_fwd")
    _count(1)
    return y


def _train_ws(x: torch.Tensor, C: int):
    B, _, h, w = x.shape
    nbytes = _cabi.load().wm_train_workspace_bytes(B, C, h, w)
    return torch.empty(max(nbytes, 16) // 8 + 1, device=x.device, dtype=torch.float64), nbytes


def dw3x3_wgrad(grad_y, x):
    """Tap and bias gradients of ``dw3x3``: returns (dweight (C,1,3,3), dbias (C))."""
    _chk(grad_y, "grad_y")
    _chk(x, "x", tuple(grad_y.shape))
    B, C, h, w = x.shape
    dw = torch.empty(C, 1, 3, 3, device=x.device, dtype=torch.float32)
    db = torch.empty(C, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        ws, nbytes = _train_ws(x, C)
        rc = _cabi.load().wm_dw3x3_wgrad(grad_y.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                         ws.data_ptr(), nbytes, B, C, h, w, _stream(x))
    _cabi.check(rc, "wm_dw3x3_wgrad")
    _count(2)
    return dw, db


def layernorm2d_bwd(x, weight, grad_y, eps: float):
    """Backward of ``layernorm2d``: returns (dx, dweight, dbias)."""
    _chk(x, "x")
    _chk(grad_y, "grad_y", tuple(x.shape))
    B, C, h, w = x.shape
    _chk(weight, "weight", (C,))
    dx = torch.empty_like(x)
    dw = torch.empty(C, device=x.device, dtype=torch.float32)
    db = torch.empty(C, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        ws, nbytes = _train_ws(x, C)
        rc = _cabi.load().wm_layernorm2d_bwd(x.data_ptr(), weight.data_ptr(), grad_y.data_ptr(), eps,
                                             dx.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                             nbytes, B, C, h, w, _stream(x))
    _cabi.check(rc, "wm_layernorm2d_bwd")
    _count(3)
    return dx, dw, db
