"""One launch of each memory-shaped kernel at its 4K size (the command ncu wraps; GPU box only)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
h, w, H, W = 1080, 1920, 2160, 3840
x, res = r(1, 32, h, w), r(1, 32, h, w)
ops.dw_act_pw(x, r(32, 1, 3, 3) * 0.3, r(32) * 0.1, r(32, 32, 1, 1) * 0.2, r(32) * 0.1, "gelu", res)
img = r(1, 3, H, W)
ops.stem_conv3x3(img, r(32, 3, 3, 3) * 0.2, r(32) * 0.1)
ops.head_conv3x3(r(1, 32, H, W), r(3, 32, 3, 3) * 0.1, r(3) * 0.1, img)
planes = [r(1, 64, h, w) for _ in range(4)]
ones, zeros = torch.ones(32, device=dev), torch.zeros(32, device=dev)
ops.lfss_tail(planes, x, ones, zeros, 1e-6, r(128, 32) * 0.2, torch.ones(64, device=dev), torch.zeros(64, device=dev),
              1e-5, r(32, 64) * 0.1, r(32))
ops.pw(r(1, 64, h, w), r(32, 32, 1, 1) * 0.2, r(32) * 0.1, gate=True, residual=res, res_scale=r(32))
qkv = r(1, 96, h, w)
ops.pw(qkv[:, 64:], r(1, 32, 32) * 0.2, r(32) * 0.1, residual=res)
ops.iwt_haar(x, r(1, 96, h, w))
ll, hl, lh, hh = ops.dwt_haar(r(1, 32, H, W))
ops.skff(hl, lh, hh, r(4, 32, 1, 1) * 0.2, torch.tensor([0.25], device=dev), *[r(32, 4, 1, 1) * 0.2 for _ in range(3)])
ops.layernorm2d(x, ones, zeros, 1e-6)
torch.cuda.synchronize()
