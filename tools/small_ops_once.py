"""One launch of each small kernel at its 4K size (for ncu)."""
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
ops.lfss_z(x, torch.ones(32, device=dev), torch.zeros(32, device=dev), 1e-6, r(128, 32) * 0.2)
ops.pw(x, r(32, 32, 1, 1) * 0.2, r(32) * 0.1, residual=res)
torch.cuda.synchronize()
