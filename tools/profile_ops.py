"""Run one hot-path op a few times at a 4K level size -- the command ncu wraps (GPU box only).

    python tools/profile_ops.py ss2d 1080 1920 [iters]
    python tools/profile_ops.py dwt 2160 3840
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops  # noqa: E402


def main():
    what = sys.argv[1]
    h, w = int(sys.argv[2]), int(sys.argv[3])
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    dev = torch.device("cuda:0")
    sd = torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_UHDLOL4K.pth"), map_location="cpu")["params"]
    pre = "restoration_network.down_group1.l_blk.0.self_attention."
    prm = [sd[pre + k].to(dev) for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")]
    torch.manual_seed(0)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if what == "ss2d":
        x = torch.nn.functional.silu(0.5 * torch.randn(1, 64, h, w, device=dev))
        fn = lambda: ops.ss2d_core(x, *prm)
        nbytes = 2 * x.numel() * 4
    elif what == "dwt":
        x = torch.randn(1, 32, h, w, device=dev)
        fn = lambda: ops.dwt_haar(x)
        nbytes = 2 * x.numel() * 4
    elif what == "iwt":
        lo, hi = torch.randn(1, 32, h, w, device=dev), torch.randn(1, 96, h, w, device=dev)
        fn = lambda: ops.iwt_haar(lo, hi)
        nbytes = 2 * (lo.numel() + hi.numel()) * 4
    else:
        raise SystemExit(what)
    fn()
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"{what} {h}x{w}: {ms:.3f} ms/call, {nbytes / ms / 1e6:.1f} GB/s algorithmic")


if __name__ == "__main__":
    main()
