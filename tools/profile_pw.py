"""Run the pointwise/depthwise group kernels at the 4K level-1 size (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wave_mamba_b200 import ops
dev = torch.device("cuda:0")
h, w = 1080, 1920
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(1, 32, h, w, device=dev)
pw_w = torch.randn(64, 32, device=dev) * 0.2
dw_w, dw_b = torch.randn(64, 1, 3, 3, device=dev) * 0.3, torch.randn(64, device=dev) * 0.1
ln_w, ln_b = torch.ones(32, device=dev), torch.zeros(32, device=dev)
y4 = [torch.randn(1, 64, h, w, device=dev) for _ in range(4)]
zs = torch.randn(1, 64, h, w, device=dev)
on_w, on_b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
w_out = torch.randn(32, 64, device=dev) * 0.2
skip = torch.ones(32, device=dev)
w_in = torch.randn(128, 32, device=dev) * 0.2
def run():
    ops.pw_dw(x, pw_w, None, dw_w, dw_b, ln_w, ln_b, 1e-6, act="silu")
    ops.lfss_out(y4[0], zs, on_w, on_b, 1e-5, w_out, x, skip, extra=(y4[2], y4[1], y4[3]))
    ops.lfss_z(x, ln_w, ln_b, 1e-6, w_in)
    ops.gram32(x, x)
run(); torch.cuda.synchronize()
for name, fn, nbytes in (
    ("pw_dw<64>", lambda: ops.pw_dw(x, pw_w, None, dw_w, dw_b, ln_w, ln_b, 1e-6, act="silu"), (32 + 64) * h * w * 4),
    ("lfss_out", lambda: ops.lfss_out(y4[0], zs, on_w, on_b, 1e-5, w_out, x, skip, extra=(y4[2], y4[1], y4[3])), (4 * 64 + 64 + 32 + 32) * h * w * 4),
    ("lfss_z", lambda: ops.lfss_z(x, ln_w, ln_b, 1e-6, w_in), (32 + 64) * h * w * 4),
    ("gram32", lambda: ops.gram32(x, x), 64 * h * w * 4)):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"{name}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s algorithmic")
