"""Developer aid: pw_dw call times at the 4K level-1 size, new pipeline vs WM_PW_DW_LEGACY (run twice)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from wave_mamba_b200 import ops  # noqa: E402

dev = "cuda"
x = torch.randn(1, 32, 1080, 1920, device=dev)
ln_w, ln_b = torch.ones(32, device=dev), torch.zeros(32, device=dev)
for cout, act in ((64, "silu"), (64, "none"), (96, "none"), (32, "none")):
    pw_w = torch.randn(cout, 32, device=dev) * 0.2
    dw_w, dw_b = torch.randn(cout, 1, 3, 3, device=dev) * 0.3, torch.zeros(cout, device=dev)
    for _ in range(3):
        ops.pw_dw(x, pw_w, None, dw_w, dw_b, ln_w, ln_b, 1e-6, act=act)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.pw_dw(x, pw_w, None, dw_w, dw_b, ln_w, ln_b, 1e-6, act=act)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    gb = (32 + cout) * 1080 * 1920 * 4 / 1e9
    if not os.environ.get("WM_PW_DW_LEGACY"):
        from wave_mamba_b200 import _cabi
        dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        _cabi.load().wm_pw_dw_debug_timing(dbg.data_ptr())
        ops.pw_dw(x, pw_w, None, dw_w, dw_b, ln_w, ln_b, 1e-6, act=act)
        torch.cuda.synchronize()
        _cabi.load().wm_pw_dw_debug_timing(None)
        d = dbg.view(148, 16).double()
        tiles = d[:, 9].clamp(min=1)
        names = ["A:wait_tma", "A:wait_xk_free", "A:ln_pass", "M:wait_xk", "M:wait_acc", "M:issue",
                 "B:wait_mma", "B:tmem_to_ps", "B:depthwise"]
        print("   cycles per tile:", {n: int((d[:, i] / tiles).mean().item()) for i, n in enumerate(names)},
              "total", int((d[:, 10] / tiles).mean().item()),
              "| epilogue warp 0:", {n: int((d[:, 11 + i] / tiles).mean().item()) for i, n in
                                      enumerate(["tmem_ld", "ps_store", "bar1", "depthwise", "bar2"])})
    print(f"pw_dw 32->{cout} {act} legacy={os.environ.get('WM_PW_DW_LEGACY', '0')}: {ms:.3f} ms, "
          f"{gb / ms * 1e3:.0f} GB/s algorithmic ({gb / ms * 1e3 / 6553 * 100:.1f}% of the measured HBM peak)")
