"""Run one hot-path op a few times at a 4K level size -- the command ncu wraps (GPU box only).

    python tools/profile_ops.py ss2d 1080 1920 [iters]
    python tools/profile_ops.py dwt 2160 3840
    python tools/profile_ops.py conv_gate|conv_k4|pw_dw|ss2d_bwd 1080 1920
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
    elif what in ("conv_gate", "conv_k4"):
        x = torch.randn(1, 64, h, w, device=dev)
        w3 = torch.randn(64, 64, 3, 3, device=dev) * 0.1
        w1, b1 = torch.randn(64, 64, 1, 1, device=dev) * 0.1, torch.zeros(64, device=dev)
        w4 = torch.randn(32, 64, 3, 3, device=dev) * 0.1
        t = ops.conv3x3(x, w3, gate_w=w1, gate_b=b1, out_c4=True)
        if what == "conv_gate":
            fn = lambda: ops.conv3x3(x, w3, gate_w=w1, gate_b=b1, out_c4=True)
            nbytes = 2 * x.numel() * 4
        else:
            fn = lambda: ops.conv3x3(t, w4, in_c4=True)
            nbytes = int(1.5 * x.numel() * 4)
    elif what == "pw_dw":
        x = torch.randn(1, 32, h, w, device=dev)
        pw_w = torch.randn(64, 32, device=dev) * 0.2
        dw_w, dw_b = torch.randn(64, 1, 3, 3, device=dev) * 0.3, torch.zeros(64, device=dev)
        ln_w, ln_b = torch.ones(32, device=dev), torch.zeros(32, device=dev)
        fn = lambda: ops.pw_dw(x, pw_w, None, dw_w, dw_b, ln_w, ln_b, 1e-6, act="silu")
        nbytes = 3 * x.numel() * 4
    elif what == "ss2d_bwd":
        x = torch.nn.functional.silu(0.5 * torch.randn(8, 64, h, w, device=dev))
        gy = torch.randn_like(x)
        fn = lambda: ops.ss2d_core_bwd(x, *prm, gy)
        nbytes = 3 * x.numel() * 4
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
