"""Golden vectors for the image I/O edges, produced by the reference's OWN functions
(basicsr/utils/img_util.py img2tensor / tensor2img and inference_wavemamba.py check_image_size).
Build container only (needs /root/reference and cv2):  python tools/make_golden_imgio.py"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ref_shims  # noqa: E402

ref_shims.install()
from basicsr.utils.img_util import img2tensor, tensor2img  # noqa: E402  (the reference's code)

# check_image_size lives in the inference script, which cannot be imported (it parses argv and pins
# CUDA devices at import time): take just that function's source from the unmodified file
src = open(os.path.join(ref_shims.REFERENCE_ROOT, "inference_wavemamba.py")).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "check_image_size")
ns = {"F": torch.nn.functional}
exec(compile(ast.Module([fn], []), "inference_wavemamba.py", "exec"), ns)
check_image_size = ns["check_image_size"]

rng = np.random.default_rng(7)
img = rng.integers(0, 256, size=(100, 150, 3), dtype=np.uint8)           # a cv2-style BGR image
x = check_image_size((img2tensor(img) / 255.).unsqueeze(0))               # inference_wavemamba.py:102-106
y = torch.rand(1, 3, 128, 256, generator=torch.Generator().manual_seed(8)) * 1.4 - 0.2
y.view(-1)[:256] = (torch.arange(256, dtype=torch.float32) + 0.5) / 255.0
out_img = tensor2img(y[:, :, :100, :150].clone())                         # inference_wavemamba.py:112-113
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "imgio.npz"), img=img, x=x.numpy(), y=y.numpy(),
                    out_img=out_img)
print("imgio:", img.shape, tuple(x.shape), out_img.shape, out_img.dtype)
