"""Per-phase cycle split of the SS2D pass kernels at a 4K level size (GPU box only)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops, _cabi
lib = _cabi.load()
dev = torch.device("cuda:0")
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
sd = torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_UHDLOL4K.pth"), map_location="cpu")["params"]
pre = "restoration_network.down_group1.l_blk.0.self_attention."
prm = [sd[pre + k].to(dev) for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")]
x = torch.nn.functional.silu(0.5 * torch.randn(1, 64, h, w, device=dev))
ops.ss2d_dirs(x, *prm); torch.cuda.synchronize()
for one_cta in (0, 1):
    dbg = torch.zeros(4096 * 6, dtype=torch.int64, device=dev)
    lib.wm_ss2d_debug_timing(dbg.data_ptr() | one_cta)   # bit 0: pad smem -> one CTA per SM
    ops.ss2d_dirs(x, *prm); torch.cuda.synchronize()     # pass 2 is the last writer
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.ss2d_dirs(x, *prm); e1.record(); torch.cuda.synchronize()
    lib.wm_ss2d_debug_timing(None)
    d = dbg.view(4096, 6).double()
    d = d[d[:, 5] > 0]
    col = (w + 3) // 4
    row = (d.shape[0] - 2 * col) // 2
    lo = 0
    for name, n in (("k0 row fwd", row), ("k1 col fwd", col), ("k2 row bwd", row), ("k3 col bwd", col)):
        part = d[lo:lo + n]; lo += n
        pt = (part[:, :5].sum(0) / part[:, 5].sum()).tolist()
        print(f"   {name}: {[int(v) for v in pt]} sum {int(sum(pt))}")
    per_tile = (d[:, :5].sum(0) / d[:, 5].sum()).tolist()
    print(f"pass 2, {h}x{w}, {2 - one_cta} CTA/SM: cycles per tile per CTA:",
          dict(zip(["wait_x", "projection", "delta", "scan", "store"], [int(v) for v in per_tile])),
          "sum", int(sum(per_tile)), f"| dirs call {e0.elapsed_time(e1):.3f} ms")
