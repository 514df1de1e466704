"""Golden vectors for the per-image metrics, produced by the reference's OWN functions
(comput_psnr_ssim.py calculate_psnr / calculate_ssim, called as inference_wavemamba.py:117-118 calls them).
Build container only (needs /root/reference and cv2):  python tools/make_golden_metrics.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ref_shims  # noqa: E402

ref_shims.install()
from comput_psnr_ssim import calculate_psnr, calculate_ssim  # noqa: E402  (the reference's code)

rng = np.random.default_rng(11)


def smooth(h, w, seed):
    r = np.random.default_rng(seed)
    small = r.random((max(h // 8, 2), max(w // 8, 2), 3)).astype(np.float32)
    big = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(big, 0, 1)


cases = {}
# (a) noise against noise; (b) a smooth image against a slightly noisy copy (the inference regime);
# (c) ragged size, strong distortion; (d) identical images
a = rng.integers(0, 256, size=(37, 50, 3), dtype=np.uint8)
b = rng.integers(0, 256, size=(37, 50, 3), dtype=np.uint8)
cases["noise"] = (a, b)
s = smooth(96, 160, 3)
cases["smooth"] = ((s * 255).round().astype(np.uint8),
                   (np.clip(s + 0.02 * rng.standard_normal(s.shape), 0, 1) * 255).round().astype(np.uint8))
s = smooth(135, 241, 4)
cases["ragged"] = ((s * 255).round().astype(np.uint8),
                   (np.clip(0.6 * s + 0.1, 0, 1) * 255).round().astype(np.uint8))
cases["same"] = (cases["smooth"][0], cases["smooth"][0].copy())

out = {"window": cv2.getGaussianKernel(11, 1.5)[:, 0]}
for name, (x, y) in cases.items():
    out[f"{name}_a"], out[f"{name}_b"] = x, y
    for crop in (1, 0, 4):
        out[f"{name}_psnr_c{crop}"] = np.float64(calculate_psnr(x, y, crop_border=crop))
        out[f"{name}_ssim_c{crop}"] = np.float64(calculate_ssim(x, y, crop_border=crop))
        print(name, crop, out[f"{name}_psnr_c{crop}"], out[f"{name}_ssim_c{crop}"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)
