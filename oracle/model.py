"""Functional CPU restatement of the Wave-Mamba forward -- TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py for who may import this.  Every function cites the lines of
/root/reference/basicsr/archs/wavemamba_arch.py (written ``ref:A-B``) it follows.  The
model is expressed as plain functions over a flat ``{name: tensor}`` parameter dict (the
checkpoint's ``params`` with the ``restoration_network.`` prefix stripped), not as an
nn.Module tree, so it shares no structure with the reference source.

dtype: all arithmetic runs in the dtype of the inputs/parameters (use ``cast_params`` to
get the float64 arbiter).  Pinning: tests/test_oracle.py checks these functions against
golden tensors produced by the unmodified reference file (tools/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from . import scan as _scan

Params = Dict[str, torch.Tensor]

D_STATE = 16  # ref:320 default d_state, never overridden by UNet (ref:967,512)


def strip_prefix(params: Params, prefix: str = "restoration_network.") -> Params:
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in params.items()}


def cast_params(params: Params, dtype: torch.dtype) -> Params:
    return {k: v.detach().to("cpu", dtype) for k, v in params.items()}


def sub(params: Params, prefix: str) -> Params:
    """Parameters below ``prefix.`` with that prefix removed."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in params.items() if k.startswith(p)}


def count_blocks(params: Params, prefix: str) -> int:
    idx = {int(k[len(prefix) + 1:].split(".")[0]) for k in params if k.startswith(prefix + ".")}
    return len(idx)


# --------------------------------------------------------------------------------------
# Haar DWT / IWT                                                          ref:97-130
# --------------------------------------------------------------------------------------
def haar_dwt(x: torch.Tensor):
    """ref:97-110.  a,b,c,d = the four taps of each 2x2 block, each halved first;
    sums associate left to right exactly as the reference's Python expression does."""
    a = x[:, :, 0::2, 0::2] / 2  # even row, even col   (x1, ref:99,101)
    b = x[:, :, 1::2, 0::2] / 2  # odd row,  even col   (x2, ref:100,102)
    c = x[:, :, 0::2, 1::2] / 2  # even row, odd col    (x3, ref:103)
    d = x[:, :, 1::2, 1::2] / 2  # odd row,  odd col    (x4, ref:104)
    ll = ((a + b) + c) + d
    hl = (((-a) - b) + c) + d
    lh = (((-a) + b) - c) + d
    hh = ((a - b) - c) + d
    return ll, hl, lh, hh


def haar_iwt(x: torch.Tensor) -> torch.Tensor:
    """ref:113-130.  x = [LL | HL | LH | HH] stacked on channels; output (B, C/4, 2h, 2w)."""
    bsz, c4, h, w = x.shape
    c = c4 // 4
    p = x[:, 0 * c:1 * c] / 2
    q = x[:, 1 * c:2 * c] / 2
    r = x[:, 2 * c:3 * c] / 2
    s = x[:, 3 * c:4 * c] / 2
    out = x.new_zeros(bsz, c, 2 * h, 2 * w)
    out[:, :, 0::2, 0::2] = ((p - q) - r) + s
    out[:, :, 1::2, 0::2] = ((p - q) + r) - s
    out[:, :, 0::2, 1::2] = ((p + q) - r) - s
    out[:, :, 1::2, 1::2] = ((p + q) + r) + s
    return out


# --------------------------------------------------------------------------------------
# SS2D                                                                    ref:446-497
# --------------------------------------------------------------------------------------
def ss2d_core(x: torch.Tensor, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds,
              scan_fn: Optional[Callable] = None, trace: Optional[dict] = None) -> torch.Tensor:
    """``SS2D.forward_core`` + the 4-way sum of ``SS2D.forward`` (ref:446-478, 490).

    x: (B, D, h, w) -> merged y: (B, D, h, w).  Direction index maps are SURVEY.md
    appendix A: dir0 l=i*w+j, dir1 l=j*h+i, dir2/dir3 = dir0/dir1 reversed.
    Sum order matches ref:490: ((y_dir0 + y_dir2) + y_dir1) + y_dir3.
    """
    scan_fn = scan_fn or _scan.selective_scan_c
    bsz, dch, h, w = x.shape
    L = h * w
    K = 4
    rank = dt_projs_weight.shape[-1]
    nstate = A_logs.shape[-1]

    row_major = x.reshape(bsz, dch, L)                             # ref:451 x.view(B,-1,L)
    col_major = x.transpose(2, 3).reshape(bsz, dch, L)             # ref:451 transpose(2,3)
    seqs = torch.stack([row_major, col_major, row_major.flip(-1), col_major.flip(-1)], 1)

    # ref:453-455 -- per-direction projections
    proj = torch.einsum("bkdl,kcd->bkcl", seqs, x_proj_weight)
    dt_low, Bmat, Cmat = torch.split(proj, [rank, nstate, nstate], dim=2)
    delta = torch.einsum("bkrl,kdr->bkdl", dt_low, dt_projs_weight)

    u = seqs.reshape(bsz, K * dch, L)                              # ref:457
    delta = delta.reshape(bsz, K * dch, L)                         # ref:458
    A = -torch.exp(A_logs)                                         # ref:462
    out = scan_fn(u, delta, A, Bmat.contiguous(), Cmat.contiguous(), Ds.reshape(-1),
                  dt_projs_bias.reshape(-1))                       # ref:465-471
    if trace is not None:
        trace["scan_in"] = dict(u=u, delta=delta, A=A, B=Bmat, C=Cmat)
        trace["scan_out"] = out
    out = out.to(x.dtype).reshape(bsz, K, dch, L)

    def col_to_row(t):                                             # ref:475-476
        return t.reshape(bsz, dch, w, h).transpose(2, 3).reshape(bsz, dch, L)

    y0 = out[:, 0]
    y2 = out[:, 2].flip(-1)                                        # ref:474
    y1 = col_to_row(out[:, 1])
    y3 = col_to_row(out[:, 3].flip(-1))
    y = ((y0 + y2) + y1) + y3                                      # ref:490
    return y.reshape(bsz, dch, h, w)


def ss2d(p: Params, x: torch.Tensor, scan_fn=None, trace=None) -> torch.Tensor:
    """``SS2D.forward`` (ref:480-497).  x: (B, h, w, C) channels-last -> same shape."""
    xz = F.linear(x, p["in_proj.weight"])                          # ref:483
    xpart, z = xz.chunk(2, dim=-1)                                 # ref:484
    xc = xpart.permute(0, 3, 1, 2)
    dch = xc.shape[1]
    xc = F.silu(F.conv2d(xc, p["conv2d.weight"], p["conv2d.bias"], padding=1, groups=dch))
    if trace is not None:
        trace["core_in"] = xc
    y = ss2d_core(xc, p["x_proj_weight"], p["dt_projs_weight"], p["dt_projs_bias"],
                  p["A_logs"], p["Ds"], scan_fn, trace)
    if trace is not None:
        trace["core_out"] = y
    y = y.permute(0, 2, 3, 1)                                      # ref:491
    y = F.layer_norm(y, (dch,), p["out_norm.weight"], p["out_norm.bias"], 1e-5)  # ref:385,492
    y = y * F.silu(z)                                              # ref:493
    return F.linear(y, p["out_proj.weight"])                       # ref:494


def lfss_ffn(p: Params, x: torch.Tensor) -> torch.Tensor:
    """``ffn`` (ref:214-231); x NCHW."""
    t = F.conv2d(x, p["conv1.weight"], p["conv1.bias"])
    t = F.conv2d(t, p["conv2.weight"], p["conv2.bias"], padding=1, groups=t.shape[1])
    g, v = t.chunk(2, dim=1)
    return F.conv2d(F.gelu(g) * v, p["conv3.weight"], p["conv3.bias"])


def lfss_block(p: Params, x: torch.Tensor, scan_fn=None, trace=None) -> torch.Tensor:
    """``LFSSBlock.forward`` (ref:520-528).  x: (B, h, w, C) channels-last."""
    C = x.shape[-1]
    t = F.layer_norm(x, (C,), p["ln_1.weight"], p["ln_1.bias"], 1e-6)           # ref:504,511
    x = x * p["skip_scale"] + ss2d(sub(p, "self_attention"), t, scan_fn, trace)  # ref:525
    t = F.layer_norm(x, (C,), p["ln_2.weight"], p["ln_2.bias"], 1e-5)           # ref:516
    f = lfss_ffn(sub(p, "conv_blk"), t.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    return x * p["skip_scale2"] + f                                              # ref:526


# --------------------------------------------------------------------------------------
# HFEBlock                                                                ref:532-854
# --------------------------------------------------------------------------------------
def layer_norm_2d(x: torch.Tensor, weight, bias, eps: float = 1e-6) -> torch.Tensor:
    """``LayerNormFunction.forward`` (ref:535-543): biased variance over channels."""
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    y = (x - mu) / (var + eps).sqrt()
    return weight.view(1, -1, 1, 1) * y + bias.view(1, -1, 1, 1)


def channel_match(x: torch.Tensor, perception: torch.Tensor, trace=None, tag=""):
    """``Matching`` with match_factor=1 (ref:618-680): for every channel map of x pick the
    nearest (L2 over all pixels) channel map of ``perception``.  With num_matches == C the
    sort/mask of ref:630-641 keeps every entry in order, so the result is a plain gather."""
    xf = x.flatten(2, 3)
    pf = perception.flatten(2, 3)
    dist = torch.cdist(xf, pf)                                      # ref:664
    idx = dist.topk(k=1, largest=False).indices.squeeze(-1)         # ref:624
    if trace is not None:
        trace.setdefault("match_idx", {})[tag] = idx
        trace.setdefault("match_dist", {})[tag] = dist
    gathered = torch.gather(pf, 1, idx[:, :, None].expand(-1, -1, pf.shape[-1]))
    return gathered.reshape(x.shape)


def paconv(p: Params, x: torch.Tensor) -> torch.Tensor:
    """``PAConv`` (ref:683-700)."""
    gate = torch.sigmoid(F.conv2d(x, p["k2.weight"], p["k2.bias"]))
    t = F.conv2d(x, p["k3.weight"], None, padding=1) * gate
    return F.conv2d(t, p["k4.weight"], None, padding=1)


def matching_transformation(p: Params, x, perception, trace=None, tag=""):
    """ref:702-719."""
    cand = channel_match(x, perception, trace, tag)
    return paconv(sub(p, "paconv"), torch.cat([x, cand], dim=1))


def cmt_attention(p: Params, x, perception, trace=None, tag=""):
    """``CMTAttention`` with num_heads=1 (ref:756-798)."""
    bsz, c, h, w = x.shape
    qkv = F.conv2d(x, p["qkv.weight"], p["qkv.bias"])
    qkv = F.conv2d(qkv, p["qkv_dwconv.weight"], p["qkv_dwconv.bias"], padding=1, groups=3 * c)
    q, k, v = qkv.chunk(3, dim=1)
    q = matching_transformation(sub(p, "matching_transformation"), q, perception, trace, tag + "attn")
    q = F.normalize(q.flatten(2, 3), dim=-1)                        # ref:787 (eps 1e-12)
    k = F.normalize(k.flatten(2, 3), dim=-1)
    v = v.flatten(2, 3)
    attn = (q @ k.transpose(-2, -1)) * p["temperature"]             # ref:790, one head
    attn = attn.softmax(dim=-1)
    out = (attn @ v).reshape(bsz, c, h, w)
    return F.conv2d(out, p["project_out.weight"], p["project_out.bias"])


def hfe_feed_forward(p: Params, x, perception, trace=None, tag=""):
    """``FeedForward`` (ref:721-751)."""
    c = x.shape[1]
    t = F.conv2d(x, p["project_in.0.weight"], p["project_in.0.bias"])
    t = F.conv2d(t, p["project_in.1.weight"], p["project_in.1.bias"], padding=1, groups=c)
    t = matching_transformation(sub(p, "matching_transformation"), t, perception, trace, tag + "ffn")
    t = F.conv2d(t, p["project_out.0.weight"], p["project_out.0.bias"], padding=1, groups=c)
    t = F.gelu(t)
    return F.conv2d(t, p["project_out.2.weight"], p["project_out.2.bias"])


def hfe_block(p: Params, x, perception, trace=None, tag=""):
    """``HFEBlock.forward`` (ref:847-854)."""
    per = layer_norm_2d(perception, p["LayerNorm.weight"], p["LayerNorm.bias"])
    x = x + cmt_attention(sub(p, "attn"), layer_norm_2d(x, p["norm1.weight"], p["norm1.bias"]),
                          per, trace, tag)
    x = x + hfe_feed_forward(sub(p, "ffn"), layer_norm_2d(x, p["norm2.weight"], p["norm2.bias"]),
                             per, trace, tag)
    return x


# --------------------------------------------------------------------------------------
# SKFF, groups, UNet                                                      ref:923-1063
# --------------------------------------------------------------------------------------
def skff(p: Params, feats: List[torch.Tensor]) -> torch.Tensor:
    """``SKFF`` with height=3 (ref:939-959)."""
    stacked = torch.stack(feats, dim=1)                             # (B, 3, C, h, w)
    pooled = stacked.sum(1).mean(dim=(2, 3), keepdim=True)          # ref:947-948
    z = F.prelu(F.conv2d(pooled, p["conv_du.0.weight"]), p["conv_du.1.weight"])
    att = torch.stack([F.conv2d(z, p[f"fcs.{i}.weight"]) for i in range(3)], dim=1)
    att = att.softmax(dim=1)                                        # ref:955
    return (stacked * att).sum(1)                                   # ref:957


def _run_lfss_chain(params: Params, prefix: str, x_nchw, scan_fn, trace, tag):
    """ref:976-979 / 998-1001: NCHW -> (B,h,w,C) -> blocks -> NCHW."""
    t = x_nchw.permute(0, 2, 3, 1)
    for i in range(count_blocks(params, prefix)):
        tr = None
        if trace is not None:
            tr = trace.setdefault("lfss", {}).setdefault(f"{tag}.{i}", {})
        t = lfss_block(sub(params, f"{prefix}.{i}"), t, scan_fn, tr)
    return t.permute(0, 3, 1, 2)


def down_group(params: Params, name: str, x, x_side, scan_fn=None, trace=None):
    """``DownFRG.forward`` (ref:972-985)."""
    p = sub(params, name)
    ll, hl, lh, hh = haar_dwt(x)
    low = F.conv2d(torch.cat([ll, x_side], dim=1), p["l_conv.weight"], p["l_conv.bias"], padding=1)
    low = _run_lfss_chain(p, "l_blk", low, scan_fn, trace, name)
    high = skff(sub(p, "h_fusion"), [hl, lh, hh])
    for i in range(count_blocks(p, "h_blk")):
        high = hfe_block(sub(p, f"h_blk.{i}"), high, low, trace, f"{name}.h{i}.")
    return low, high


def up_group(params: Params, name: str, low, high, scan_fn=None, trace=None):
    """``upFRG.forward`` (ref:996-1008)."""
    p = sub(params, name)
    low = _run_lfss_chain(p, "l_blk", low, scan_fn, trace, name)
    for i in range(count_blocks(p, "h_blk")):
        high = hfe_block(sub(p, f"h_blk.{i}"), high, low, trace, f"{name}.h{i}.")
    high = F.conv2d(high, p["h_out_conv.weight"], p["h_out_conv.bias"], padding=1)
    return haar_iwt(torch.cat([low, high], dim=1))


def unet_forward(params: Params, x: torch.Tensor, scan_fn=None, trace=None) -> torch.Tensor:
    """``UNet.forward`` (ref:1041-1063).  ``params`` = checkpoint['params'] (prefix optional)."""
    params = strip_prefix(params)
    side = []
    for lvl, r in ((1, 2), (2, 4), (3, 8)):                         # ref:1014-1025,1043-1045
        side.append(F.conv2d(F.pixel_unshuffle(x, r), params[f"ps_down{lvl}.1.weight"],
                             params[f"ps_down{lvl}.1.bias"]))
    t = F.conv2d(x, params["conv_01.weight"], params["conv_01.bias"], padding=1)  # ref:1048
    low, h1 = down_group(params, "down_group1", t, side[0], scan_fn, trace)
    low, h2 = down_group(params, "down_group2", low, side[1], scan_fn, trace)
    low, h3 = down_group(params, "down_group3", low, side[2], scan_fn, trace)
    low = up_group(params, "up_group3", low, h3, scan_fn, trace)
    low = up_group(params, "up_group2", low, h2, scan_fn, trace)
    low = up_group(params, "up_group1", low, h1, scan_fn, trace)
    return F.conv2d(low, params["last.weight"], params["last.bias"], padding=1) + x  # ref:1061


# --------------------------------------------------------------------------------------
# Synthetic inputs and the PSNR definition used by the parity criterion
# --------------------------------------------------------------------------------------
def synth_lowlight(bsz: int, H: int, W: int, seed: int):
    """SURVEY.md section 8d synthetic low-light generator.  Returns (x, pseudo_gt)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(bsz, 3, max(H // 16, 1), max(W // 16, 1), generator=g)
    base = F.interpolate(base, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    x = (base * 0.15 + 0.02 * torch.rand(bsz, 3, H, W, generator=g)).clamp(0, 1)
    return x.contiguous(), base.contiguous()


def to_uint8_bgr(t: torch.Tensor):
    """basicsr/utils/img_util.py:36-98 ``tensor2img`` for one (1|-,3,H,W) tensor:
    clamp [0,1], CHW RGB -> HWC BGR, *255, round, uint8."""
    import numpy as np
    t = t.detach().float().cpu().clamp(0, 1)
    if t.dim() == 4:
        t = t[0]
    img = t.numpy().transpose(1, 2, 0)[:, :, ::-1]
    return (img * 255.0).round().astype(np.uint8)


def _y_channel(im: "np.ndarray") -> "np.ndarray":
    """to_y_channel (comput_psnr_ssim.py:374-385) of a float64 BGR image in [0,255]: float32 in, fp64 dot
    with the BT.601 BGR weights (bgr2ycbcr y_only, :210-238), back to float32 in [0,255]."""
    import numpy as np
    f = im.astype(np.float32) / 255.0
    y = np.dot(f, [24.966, 128.553, 65.481]) + 16.0   # BGR order weights
    y = (y / 255.0).astype(np.float32)                 # _convert_output_type_range(float32)
    return y * 255.0                                   # to_y_channel: back to [0,255]


def _crop_f64(img: "np.ndarray", crop_border: int) -> "np.ndarray":
    import numpy as np
    a = img.astype(np.float64)
    if crop_border:
        a = a[crop_border:-crop_border, crop_border:-crop_border]
    return a


def psnr_y(img: "np.ndarray", ref: "np.ndarray", crop_border: int = 1) -> float:
    """comput_psnr_ssim.py:387-438 with the inference defaults (crop 1, Y channel).
    Y from BGR uint8 per comput_psnr_ssim.py:210-238 (bgr2ycbcr y_only) / 374-385."""
    import numpy as np
    a, b = _y_channel(_crop_f64(img, crop_border)), _y_channel(_crop_f64(ref, crop_border))
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    return float(20.0 * math.log10(255.0 / math.sqrt(mse)))


def gaussian_window_11() -> "np.ndarray":
    """cv2.getGaussianKernel(11, 1.5) (comput_psnr_ssim.py:573): exp(-(i-5)^2 / (2 sigma^2)), normalised."""
    import numpy as np
    x = np.arange(11, dtype=np.float64) - 5.0
    k = np.exp(-0.5 / (1.5 * 1.5) * x * x)
    return k * (1.0 / k.sum())


def ssim_y(img: "np.ndarray", ref: "np.ndarray", crop_border: int = 1) -> float:
    """comput_psnr_ssim.py:596-668 with the inference defaults -> _ssim_cly :559-592: the 11x11 Gaussian
    window correlated (cv2.filter2D, BORDER_REPLICATE, same size) with Y1, Y2, Y1^2, Y2^2, Y1*Y2 in fp64,
    the SSIM map with C1 = (0.01*255)^2, C2 = (0.03*255)^2, its mean.  Plain numpy: the 2-D window is
    applied tap by tap on an edge-padded copy."""
    import numpy as np
    a = _y_channel(_crop_f64(img, crop_border)).astype(np.float64)
    b = _y_channel(_crop_f64(ref, crop_border)).astype(np.float64)
    k = gaussian_window_11()
    window = np.outer(k, k)

    def filt(x):
        h, w = x.shape
        p = np.pad(x, 5, mode="edge")
        out = np.zeros_like(x)
        for i in range(11):
            for j in range(11):
                out += window[i, j] * p[i:i + h, j:j + w]
        return out

    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    mu1, mu2 = filt(a), filt(b)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = filt(a ** 2) - mu1_sq
    s2 = filt(b ** 2) - mu2_sq
    s12 = filt(a * b) - mu1_mu2
    ssim_map = ((2 * mu1_mu2 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    return float(ssim_map.mean())


# --------------------------------------------------------------------------------------
# image I/O edges of the inference loop (reference inference_wavemamba.py:101-113)
# --------------------------------------------------------------------------------------
def img_u8_to_f32(img_bgr_u8: torch.Tensor, window: int = 128) -> torch.Tensor:
    """(B,H,W,3) uint8 BGR -> (B,3,Hp,Wp) float32 RGB: ``img2tensor`` (basicsr/utils/img_util.py:9-33:
    BGR->RGB, HWC->CHW, float), ``/ 255.`` (inference_wavemamba.py:102) and ``check_image_size``
    (inference_wavemamba.py:28-36: reflect pad bottom/right to a multiple of ``window``)."""
    x = img_bgr_u8.flip(-1).permute(0, 3, 1, 2).float() / 255.
    h, w = x.shape[2:]
    pad_h = (window - h % window) % window
    pad_w = (window - w % window) % window
    return F.pad(x, (0, pad_w, 0, pad_h), "reflect")


def img_f32_to_u8(x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """(B,3,Hs,Ws) float32 RGB -> (B,h,w,3) uint8 BGR: the crop at inference_wavemamba.py:112 and
    ``tensor2img`` (img_util.py:36-98: clamp to [0,1], HWC, RGB->BGR, ``(x * 255.0).round()``, uint8)."""
    t = x[:, :, :h, :w].float().clamp(0, 1).permute(0, 2, 3, 1).flip(-1)
    return torch.from_numpy((t.numpy() * 255.0).round().astype("uint8"))
