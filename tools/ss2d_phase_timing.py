"""Per-phase cycle split of the SS2D pass-2 kernel and the time of one dirs call at the three 4K
level sizes (GPU box only)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops, _cabi
lib = _cabi.load()
dev = torch.device("cuda:0")
sizes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(1080, 1920), (540, 960), (270, 480)]
sd = torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_UHDLOL4K.pth"), map_location="cpu")["params"]
pre = "restoration_network.down_group1.l_blk.0.self_attention."
prm = [sd[pre + k].to(dev) for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")]
for h, w in sizes:
    x = torch.nn.functional.silu(0.5 * torch.randn(1, 64, h, w, device=dev))
    for _ in range(3):
        ops.ss2d_dirs(x, *prm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.ss2d_dirs(x, *prm)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    geo = (ctypes.c_int * 6)()
    lib.wm_ss2d_debug_geometry(1, h, w, geo)
    dbg = torch.zeros(8192 * 6, dtype=torch.int64, device=dev)
    lib.wm_ss2d_debug_timing(dbg.data_ptr())
    ops.ss2d_dirs(x, *prm); torch.cuda.synchronize()     # pass 2 is the last writer
    lib.wm_ss2d_debug_timing(None)
    d1 = dbg.view(8192, 6)[4096:].double()
    d1 = d1[d1[:, 5] > 0]
    pass1 = (d1[:, :5].sum(0) / d1[:, 5].sum()).tolist()
    d = dbg.view(8192, 6)[:4096].double()
    d = d[d[:, 5] > 0]
    per_tile = (d[:, :5].sum(0) / d[:, 5].sum()).tolist()
    upd = h * w * 4096 * 2
    print(f"{h}x{w}: dirs call {ms:.3f} ms ({upd / ms / 1e6 / (148 * 16 * 1.965):.1%} of the MUFU roofline, two passes) "
          f"plan row_T={geo[0]} row_ctas={geo[1]} col_seg={geo[2]}x{geo[3]} col_ctas={geo[4]} cols_first={geo[5]} | "
          f"pass-2 cycles per tile per CTA:", dict(zip(["wait_x", "projection", "delta", "scan", "store"], [int(v) for v in per_tile])),
          "sum", int(sum(per_tile)), "| pass 1:", dict(zip(["wait_x", "projection", "delta", "scan"], [int(v) for v in pass1[:4]])),
          "sum", int(sum(pass1[:4])))
