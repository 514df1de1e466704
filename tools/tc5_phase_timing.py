"""Developer aid: per-tile cycle split of the tcgen05 conv's MMA thread (wm_conv3x3_debug_timing) and
the call time of every variant at the 4K level-1 size.  Run on the GPU box."""
import sys

import torch

sys.path.insert(0, ".")
from wave_mamba_b200 import _cabi, ops  # noqa: E402

lib = _cabi.load()
dev = "cuda"
x = torch.randn(1, 64, 1080, 1920, device=dev)
x32 = torch.randn(1, 32, 1080, 1920, device=dev)
w3 = torch.randn(64, 64, 3, 3, device=dev) * 0.1
w1 = torch.randn(64, 64, 1, 1, device=dev) * 0.1
b = torch.zeros(64, device=dev)
w4 = torch.randn(32, 64, 3, 3, device=dev) * 0.1
wh = torch.randn(96, 32, 3, 3, device=dev) * 0.1
t = ops.conv3x3(x, w3, gate_w=w1, gate_b=b, out_c4=True)
names = ["total", "wait_w", "wait_x", "wait_acc", "issue", "tiles"]
cases = (("gate 64->64 (nchw in, c4 out)", lambda: ops.conv3x3(x, w3, gate_w=w1, gate_b=b, out_c4=True), 64 * 64 * 10),
         ("k4 64->32 (c4 in)", lambda: ops.conv3x3(t, w4, in_c4=True), 64 * 32 * 9),
         ("l_conv 64->32 (nchw)", lambda: ops.conv3x3(x, w4), 64 * 32 * 9),
         ("h_out 32->96", lambda: ops.conv3x3(x32, wh), 32 * 96 * 9))
for label, fn, mac_px in cases:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    dbg = torch.zeros(148 * 6, dtype=torch.int64, device=dev)
    lib.wm_conv3x3_debug_timing(dbg.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.wm_conv3x3_debug_timing(None)
    d = dbg.view(148, 6).double()
    tiles = d[:, 5].clamp(min=1)
    per = {n: int((d[:, i] / tiles).mean().item()) for i, n in enumerate(names[:5])}
    tf = 2.0 * mac_px * 1080 * 1920 / (ms * 1e-3) / 1e12
    print(f"{label}: {ms:.3f} ms, {tf:.1f} TFLOP/s fp32-equivalent; MMA-thread cycles per tile: {per}")
