"""The per-image body of the reference's inference loop with uint8 edges.

``inference_wavemamba.py:101-113`` does, per image: cv2 image (H,W,3 uint8 BGR) -> img2tensor ->
/255. -> reflect-pad to a multiple of 128 -> ``restoration_network`` -> crop -> tensor2img (clamp,
*255, round, uint8, RGB->BGR).  Here the uint8 image is what crosses PCIe (3 bytes per pixel each
way instead of 12) and the conversions run on the device (wm_img_u8_to_f32_fwd /
wm_img_f32_to_u8_fwd, bit-exact with img2tensor / tensor2img).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


@torch.no_grad()
def enhance_bgr_u8(net, img: torch.Tensor, window: int = 128, out: Optional[torch.Tensor] = None,
                   device: Optional[torch.device] = None, sync: bool = False,
                   cuda_division: bool = False) -> torch.Tensor:
    """img: (H,W,3) or (B,H,W,3) uint8 BGR, on the host (ideally pinned) or already on the GPU.
    Returns the enhanced uint8 BGR image(s) with the input's leading shape: on the GPU, or copied
    into ``out`` (a host uint8 tensor of the same shape, ideally pinned) when given.
    ``net`` is a ``WaveMamba`` (its ``restoration_network`` is used, as the reference does).

    ``cuda_division=True`` reproduces the reference script's GPU run bit for bit (its ``/ 255.`` on a
    CUDA tensor is a multiplication by the rounded reciprocal); the default is the IEEE division the
    CPU reference and the oracle compute.

    Stream contract: everything is enqueued on the current CUDA stream.  With a pinned ``img`` /
    ``out`` the two PCIe copies are asynchronous: do not overwrite ``img`` or read ``out`` before the
    stream has been synchronised -- pass ``sync=True`` for the reference loop's blocking behaviour
    (``tensor2img`` returns finished data)."""
    if img.dtype != torch.uint8 or img.dim() not in (3, 4) or img.shape[-1] != 3:
        raise ValueError(f"expected a (H,W,3) or (B,H,W,3) uint8 image, got {tuple(img.shape)} {img.dtype}")
    squeeze = img.dim() == 3
    if squeeze:
        img = img.unsqueeze(0)
    if not img.is_cuda:
        if device is None:
            device = next(net.parameters()).device
        img = img.to(device, non_blocking=True)
    img = img.contiguous()
    _, H, W, _ = img.shape
    fwd = getattr(net, "restoration_network", net)
    x = ops.img_u8_to_f32(img, window, cuda_division)                 # img2tensor, /255., check_image_size
    y = fwd(x)
    res = ops.img_f32_to_u8(y, H, W)                   # [:, :, :h, :w] + tensor2img
    if squeeze:
        res = res[0]
    if out is not None:
        out.copy_(res, non_blocking=True)
        res = out
    if sync:
        torch.cuda.current_stream(img.device).synchronize()
    return res
