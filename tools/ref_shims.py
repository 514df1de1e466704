"""Import the UNMODIFIED reference (``/root/reference``) inside the build container.

Test tooling only: tools/make_golden.py (build container), the tests that are skipped when
the reference tree is absent, and tests/test_reference_scripts_gpu.py, which runs the reference's
own drivers from the staged copy baseline/_ref (tools/stage_reference.py).  Nothing here is part
of the product path.

The reference needs four packages that are not installed here:
``timm`` (DropPath/to_2tuple/trunc_normal_), ``lmdb``, ``pyiqa`` and ``mamba_ssm``.  Tiny
stand-ins are injected into ``sys.modules`` before ``basicsr`` is imported.  The
``mamba_ssm`` stand-in is the package's published ``selective_scan_ref`` recurrence as a
sequential float32 torch loop -- this is the only piece of the golden vectors that does
not come from the reference's own code (parity unpinned for the scan).
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def _scan_shim(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
               return_last_state=False):
    """selective_scan_ref semantics, sequential over L, fp32 (grouped B/C: (B,G,N,L))."""
    dtype_in = u.dtype
    u, delta = u.float(), delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, L = u.shape
    G = B.shape[1]
    Bf = B.float().repeat_interleave(dim // G, dim=1)
    Cf = C.float().repeat_interleave(dim // G, dim=1)
    state = u.new_zeros(batch, dim, A.shape[1])
    out = []
    for i in range(L):
        dA = torch.exp(delta[:, :, i, None] * A[None])
        dBu = delta[:, :, i, None] * Bf[:, :, :, i] * u[:, :, i, None]
        state = dA * state + dBu
        out.append((state * Cf[:, :, :, i]).sum(-1))
    y = torch.stack(out, dim=2)
    if D is not None:
        y = y + u * D[None, :, None].float()
    if z is not None:
        y = y * F.silu(z)
    return y.to(dtype_in)


class _Identity(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return x


class _Anything(types.ModuleType):
    """Module whose every attribute is a harmless callable returning another dummy."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _dummy(*a, **k):
            return _Anything(name)
        return _dummy


def install(reference_root: str = REFERENCE_ROOT):
    if "basicsr" in sys.modules and getattr(sys.modules["basicsr"], "__wm_shimmed__", False):
        return
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.DropPath = _Identity
    timm_layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
    timm_layers.trunc_normal_ = nn.init.trunc_normal_
    timm.models = timm_models
    timm_models.layers = timm_layers
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", timm_models)
    sys.modules.setdefault("timm.models.layers", timm_layers)

    mamba = types.ModuleType("mamba_ssm")
    mamba_ops = types.ModuleType("mamba_ssm.ops")
    mamba_if = types.ModuleType("mamba_ssm.ops.selective_scan_interface")
    mamba_if.selective_scan_fn = _scan_shim
    mamba_if.selective_scan_ref = _scan_shim
    mamba.ops = mamba_ops
    mamba_ops.selective_scan_interface = mamba_if
    sys.modules.setdefault("mamba_ssm", mamba)
    sys.modules.setdefault("mamba_ssm.ops", mamba_ops)
    sys.modules.setdefault("mamba_ssm.ops.selective_scan_interface", mamba_if)

    # inference_wavemamba.py:16-18 builds an LPIPS metric at import time (torchmetrics is absent and
    # its AlexNet weights would need the network); the stand-in reports 0
    tm = types.ModuleType("torchmetrics")
    tm_image = types.ModuleType("torchmetrics.image")
    tm_lpip = types.ModuleType("torchmetrics.image.lpip")

    class _ZeroLPIPS:
        def __init__(self, *a, **k):
            pass

        def __call__(self, a, b):
            return torch.zeros(())

    tm_lpip.LearnedPerceptualImagePatchSimilarity = _ZeroLPIPS
    tm.image = tm_image
    tm_image.lpip = tm_lpip
    sys.modules.setdefault("torchmetrics", tm)
    sys.modules.setdefault("torchmetrics.image", tm_image)
    sys.modules.setdefault("torchmetrics.image.lpip", tm_lpip)

    # comput_psnr_ssim.py:5 imports skimage.metrics but only mentions it in comments
    sk = types.ModuleType("skimage")
    sk.metrics = types.ModuleType("skimage.metrics")
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.metrics", sk.metrics)

    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    sys.modules.setdefault("pyiqa", _Anything("pyiqa"))
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import basicsr  # noqa: F401  (heavy: pulls torchvision + cv2)
    basicsr.__wm_shimmed__ = True


def reference_arch():
    install()
    from basicsr.archs import wavemamba_arch
    return wavemamba_arch


def build_reference_model(ckpt_path: str):
    arch = reference_arch()
    net = arch.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    sd = torch.load(ckpt_path, map_location="cpu")["params"]
    net.load_state_dict(sd, strict=True)
    return net.eval()
