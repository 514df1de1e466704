"""Per-kernel time table of one 4K forward (GPU box only), via torch.profiler (kineto).

    python tools/profile_model.py [--tf32 0|1] [--height H --width W] > gpurun_out/model_kernels.txt
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wave_mamba_b200 as wm  # noqa: E402
from oracle import model as om  # noqa: E402  (synthetic input generator only)

ap = argparse.ArgumentParser()
ap.add_argument("--tf32", type=int, default=0)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--rows", type=int, default=45)
args = ap.parse_args()

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = bool(args.tf32)
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
params = torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_UHDLL.pth"), map_location="cpu")["params"]
net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
net.load_state_dict(params, strict=True)
net = net.to(dev).eval()
x = om.synth_lowlight(1, args.height, args.width, 1234)[0].to(dev)
with torch.no_grad():
    for _ in range(3):
        net.restoration_network(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        net.restoration_network(x)
    e.record()
    torch.cuda.synchronize()
    print(f"tf32={args.tf32}: {s.elapsed_time(e) / 3:.2f} ms per {args.width}x{args.height} forward, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        net.restoration_network(x)
        torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", None)
    if t is None:
        t = getattr(ev, "cuda_time_total", 0)
    if ev.device_type.name == "CUDA" or (t and ev.key.startswith(("void", "wm::", "ampere", "sm", "cutlass", "cudnn", "Memcpy", "Memset"))):
        rows.append((t / 1e3, ev.count, ev.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"device kernels: {len(rows)} distinct, {sum(r[1] for r in rows)} launches, {tot:.2f} ms total")
for t, c, k in rows[:args.rows]:
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {k[:110]}")
