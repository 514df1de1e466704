"""Developer aid: run the two mbarrier pipelines on small inputs and report a timed-out wait."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from wave_mamba_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
for (B, h, w, cout) in ((2, 8, 32, 32), (1, 40, 96, 64), (2, 203, 400, 96)):
    x = torch.randn(B, 32, h, w, device=dev)
    pw_w, pw_b = torch.randn(cout, 32, device=dev) * 0.2, torch.randn(cout, device=dev) * 0.1
    dw_w, dw_b = torch.randn(cout, 1, 3, 3, device=dev) * 0.3, torch.randn(cout, device=dev) * 0.1
    try:
        y = ops.pw_dw(x, pw_w, pw_b, dw_w, dw_b)
        code = ops.pipeline_error()
        want = F.conv2d(F.conv2d(x.double(), pw_w.double()[:, :, None, None], pw_b.double()), dw_w.double(),
                        dw_b.double(), padding=1, groups=cout)
        err = (y.double() - want).abs().max().item()
        print(f"pw_dw {B}x32x{h}x{w} -> {cout}: pipeline error word {code:#010x}, max err {err:.3e}", flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"pw_dw {B}x32x{h}x{w} -> {cout}: EXCEPTION {type(exc).__name__}: {str(exc)[:200]}", flush=True)
        break
