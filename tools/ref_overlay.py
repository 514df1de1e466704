"""Overlay this repo's plugin on a reference checkout (test infrastructure).

``make_overlay(ref_root, dst)`` builds ``dst/`` = the reference tree with every entry symlinked,
except ``basicsr/archs/wavemamba_arch.py`` which points at plugin/basicsr/archs/wavemamba_arch.py
-- exactly what a user does when they drop the plugin file over the reference's.
``python tools/ref_overlay.py REF DST script.py [args...]`` then runs one of the reference's own
scripts, unmodified, from that overlay with stand-ins for the packages that are not installed in
this image (timm, lmdb, pyiqa, torchmetrics, skimage; tools/ref_shims.py).
"""
from __future__ import annotations

import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "plugin", "basicsr", "archs", "wavemamba_arch.py")


def make_overlay(ref_root: str, dst: str) -> str:
    arch_dst = os.path.join(dst, "basicsr", "archs")
    os.makedirs(arch_dst, exist_ok=True)
    for entry in os.listdir(ref_root):
        if entry not in ("basicsr", "__pycache__"):
            os.symlink(os.path.join(ref_root, entry), os.path.join(dst, entry))
    for entry in os.listdir(os.path.join(ref_root, "basicsr")):
        if entry not in ("archs", "__pycache__"):
            os.symlink(os.path.join(ref_root, "basicsr", entry), os.path.join(dst, "basicsr", entry))
    for entry in os.listdir(os.path.join(ref_root, "basicsr", "archs")):
        if entry not in ("wavemamba_arch.py", "__pycache__"):
            os.symlink(os.path.join(ref_root, "basicsr", "archs", entry), os.path.join(arch_dst, entry))
    os.symlink(PLUGIN, os.path.join(arch_dst, "wavemamba_arch.py"))
    return dst


def main(argv):
    ref_root, dst, script, args = argv[0], argv[1], argv[2], argv[3:]
    if not os.path.isdir(os.path.join(dst, "basicsr")):
        make_overlay(ref_root, dst)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from tools import ref_shims
    ref_shims.install(dst)
    os.chdir(dst)
    sys.argv = [script] + list(args)
    runpy.run_path(os.path.join(dst, script), run_name="__main__")


if __name__ == "__main__":
    main(sys.argv[1:])
