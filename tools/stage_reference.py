"""Stage the UNMODIFIED reference tree for the GPU box:  python tools/stage_reference.py

/root/reference exists only in the build container.  This copies the part of it that the
reference's own drivers need (basicsr/, inference_wavemamba.py, comput_psnr_ssim.py, options/)
verbatim into baseline/_ref/ -- git-ignored, so no reference source enters the history, but not
gpurun-ignored, so it travels to the GPU box where tests/test_reference_scripts_gpu.py runs the
reference's own scripts against this repo's plugin.  Nothing in the product reads baseline/_ref.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
WANT = ["basicsr", "inference_wavemamba.py", "comput_psnr_ssim.py", "options", "VERSION"]


def stage() -> str:
    if not os.path.isdir(SRC):
        raise SystemExit(f"{SRC} is not mounted; nothing to stage")
    os.makedirs(DST, exist_ok=True)
    for name in WANT:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(s, d)
    return DST


if __name__ == "__main__":
    print(stage())
    sys.exit(0)
