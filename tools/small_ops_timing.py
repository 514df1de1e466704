"""Developer aid: times of the small memory-bound kernels at the 4K sizes, as a fraction of the measured
HBM peak (algorithmic bytes).  GPU box only."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_mamba_b200 import ops  # noqa: E402

dev = "cuda"
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs_sustained"]
except Exception:
    PEAK = 6553.0


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def report(name, ms, planes, h, w):
    gb = planes * h * w * 4 / 1e9
    print(f"{name:28s} {ms:7.3f} ms  {gb / ms * 1e3:6.0f} GB/s  {gb / ms * 1e3 / PEAK * 100:5.1f} % of {PEAK:.0f}")


g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
h, w = 1080, 1920
x = r(1, 32, h, w)
res = r(1, 32, h, w)
dw_w, dw_b, pw_w, pw_b = r(32, 1, 3, 3) * 0.3, r(32) * 0.1, r(32, 32, 1, 1) * 0.2, r(32) * 0.1
report("dw_act_pw gelu+res L1", timeit(lambda: ops.dw_act_pw(x, dw_w, dw_b, pw_w, pw_b, "gelu", res)), 96, h, w)
H, W = 2160, 3840
img = r(1, 3, H, W)
sw, sb = r(32, 3, 3, 3) * 0.2, r(32) * 0.1
report("stem 3->32 4K", timeit(lambda: ops.stem_conv3x3(img, sw, sb)), 35, H, W)
f = r(1, 32, H, W)
hw_, hb = r(3, 32, 3, 3) * 0.1, r(3) * 0.1
report("head 32->3 + res 4K", timeit(lambda: ops.head_conv3x3(f, hw_, hb, img)), 38, H, W)
low, high = r(1, 32, h, w), r(1, 96, h, w)
report("iwt L1", timeit(lambda: ops.iwt_haar(low, high)), 256, h, w)
big = r(1, 32, H, W)
report("dwt L1", timeit(lambda: ops.dwt_haar(big)), 256, h, w)
ipw = r(128, 32) * 0.2
lw, lb = torch.ones(32, device=dev), torch.zeros(32, device=dev)
report("lfss_z L1", timeit(lambda: ops.lfss_z(x, lw, lb, 1e-6, ipw)), 96, h, w)
for rr in (2, 4, 8):
    pw_, pb_ = r(32, 3 * rr * rr, 1, 1) * 0.2, r(32) * 0.1
    ms = timeit(lambda: ops.ps_down(img, pw_, pb_, rr))
    gb = (3 + 32.0 / (rr * rr)) * H * W * 4 / 1e9
    print(f"{'ps_down r=%d 4K' % rr:28s} {ms:7.3f} ms  {gb / ms * 1e3:6.0f} GB/s  {gb / ms * 1e3 / PEAK * 100:5.1f} % of {PEAK:.0f}")
