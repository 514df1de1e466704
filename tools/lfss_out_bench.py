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
    x32 = x
    ln_w, ln_b = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    w_in = torch.randn(128, 32, device=dev) * 0.2
    f2 = lambda: ops.lfss_tail(planes, x32, ln_w, ln_b, 1e-6, w_in, on_w, on_b, 1e-5, W, sk)
    for _ in range(3): f2()
    torch.cuda.synchronize()
    s.record()
    for _ in range(10): f2()
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / 10
    f3 = lambda: ops.lfss_z(x32, ln_w, ln_b, 1e-6, w_in)
    for _ in range(3): f3()
    torch.cuda.synchronize()
    s.record()
    for _ in range(10): f3()
    e.record(); torch.cuda.synchronize()
    ms3 = s.elapsed_time(e) / 10
    gb2 = (4 * 64 + 32 + 32) * h * w * 4 / 1e9
    print(f"lfss_tail {h}x{w}: {ms2:.3f} ms ({gb2 / ms2 * 1e3:.0f} GB/s algorithmic) vs lfss_z {ms3:.3f} + lfss_out {ms:.3f} = {ms3 + ms:.3f} ms")
