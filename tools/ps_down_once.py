import sys, torch
sys.path.insert(0, ".")
from wave_mamba_b200 import ops
img = torch.randn(1, 3, 2160, 3840, device="cuda")
for rr in (2, 4, 8):
    w = torch.randn(32, 3 * rr * rr, 1, 1, device="cuda") * 0.2
    b = torch.randn(32, device="cuda")
    for _ in range(2):
        ops.ps_down(img, w, b, rr)
torch.cuda.synchronize()
