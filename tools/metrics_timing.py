"""Developer aid: device time of wm_psnr_ssim_y_u8 on a 3840x2160 image pair.  Run on the GPU box."""
import sys

import torch

sys.path.insert(0, ".")
from wave_mamba_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
a = torch.randint(0, 256, (1, 2160, 3840, 3), generator=g, dtype=torch.uint8).cuda()
b = torch.randint(0, 256, (1, 2160, 3840, 3), generator=g, dtype=torch.uint8).cuda()
for _ in range(3):
    r = ops.psnr_ssim_y(a, b, 1)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    r = ops.psnr_ssim_y(a, b, 1)
e.record()
torch.cuda.synchronize()
print(f"psnr_ssim_y 3840x2160: {s.elapsed_time(e) / 10:.3f} ms per image pair; psnr {float(r[0, 0]):.4f} ssim {float(r[0, 1]):.6f}")
