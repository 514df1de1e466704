"""Drop-in replacement for the reference's basicsr/archs/wavemamba_arch.py.

Overlay this one file on a reference checkout (or put ``plugin/`` ahead of the reference on
``sys.path``): ``basicsr/archs/__init__.py`` auto-imports every ``*_arch.py`` and the class
below registers itself as ``WaveMamba`` in ARCH_REGISTRY, so ``inference_wavemamba.py``
(``from basicsr.archs.wavemamba_arch import WaveMamba``) and ``basicsr/train.py``
(``build_network`` -> ``ARCH_REGISTRY.get('WaveMamba')``) pick up the sm_100a path unchanged.
Unlike the file it replaces it does not import mamba_ssm, timm or scipy.
"""
from basicsr.utils.registry import ARCH_REGISTRY

from wave_mamba_b200.arch import (DWT, IWT, SS2D, DownFRG, HFEBlock, LFSSBlock, SKFF, UNet,  # noqa: F401
                                  upFRG)
from wave_mamba_b200.arch import WaveMamba as _WaveMamba


@ARCH_REGISTRY.register()
class WaveMamba(_WaveMamba):
    """Same constructor, attributes and state-dict keys as the reference class (:1066-1176)."""
