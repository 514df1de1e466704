"""Synthetic low-light inputs for bench.py, torch only.

bench.py's product arm must not import ``oracle`` (test infrastructure), so the generator lives
here; ``oracle.model.synth_lowlight`` is the same arithmetic and tests/test_oracle.py checks that
the two produce identical tensors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def synth_lowlight(bsz: int, H: int, W: int, seed: int):
    """Smooth random structure at 1/16 resolution, bicubic x16, scaled to 0.15, plus 2 % uniform
    sensor noise: dark images in the regime the checkpoints were trained on.  Returns (x, pseudo_gt)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(bsz, 3, max(H // 16, 1), max(W // 16, 1), generator=g)
    base = F.interpolate(base, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    x = (base * 0.15 + 0.02 * torch.rand(bsz, 3, H, W, generator=g)).clamp(0, 1)
    return x.contiguous(), base.contiguous()


def f32_to_u8_bgr(x: torch.Tensor) -> torch.Tensor:
    """(B,3,H,W) float32 RGB in [0,1] -> (B,H,W,3) uint8 BGR (the cv2 array inference_wavemamba.py
    reads from disk): clamp, *255, round."""
    t = x.float().clamp(0, 1).permute(0, 2, 3, 1).flip(-1)
    return (t * 255.0).round().to(torch.uint8).contiguous()
