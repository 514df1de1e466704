"""Time the two production forms of wm_pw_fwd (32 -> 32) at the three 4K level sizes (GPU box only)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops
dev = torch.device("cuda:0")
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for h, w in ((1080, 1920), (540, 960), (270, 480)):
    x = torch.randn(1, 64, h, w, device=dev); res = torch.randn(1, 32, h, w, device=dev)
    W = torch.randn(32, 32, 1, 1, device=dev) * 0.2; b = torch.randn(32, device=dev); sc = torch.randn(32, device=dev)
    ms = t(lambda: ops.pw(x, W, b, gate=True, residual=res, res_scale=sc))
    gb = 128 * h * w * 4 / 1e9
    qkv = torch.randn(1, 96, h, w, device=dev); Wb = torch.randn(1, 32, 32, device=dev) * 0.2
    ms2 = t(lambda: ops.pw(qkv[:, 64:], Wb, b, residual=res))
    gb2 = 96 * h * w * 4 / 1e9
    print(f"pw gate {h}x{w}: {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s)   pw per-image {ms2:.3f} ms ({gb2 / ms2 * 1e3:.0f} GB/s)")
