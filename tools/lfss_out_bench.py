"""Time wm_lfss_out_fwd at the three 4K level sizes (GPU box only)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops
dev = torch.device("cuda:0")
for h, w in ((1080, 1920), (540, 960), (270, 480)):
    planes = [torch.randn(1, 64, h, w, device=dev) for _ in range(4)]
    zs = torch.randn(1, 64, h, w, device=dev); x = torch.randn(1, 32, h, w, device=dev)
    on_w, on_b = torch.randn(64, device=dev), torch.randn(64, device=dev)
    W = torch.randn(32, 64, device=dev) * 0.1; sk = torch.randn(32, device=dev)
    f = lambda: ops.lfss_out(planes[0], zs, on_w, on_b, 1e-5, W, x, sk, extra=planes[1:])
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    gb = (4 * 64 + 64 + 32 + 32) * h * w * 4 / 1e9
    print(f"lfss_out {h}x{w}: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s algorithmic")
