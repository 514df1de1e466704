"""wave_mamba_b200 -- B200-native (sm_100a) forward path for Wave-Mamba.

Public surface:
  * ``WaveMamba``  -- the reference's registry class (same constructor / state dict);
  * ``ops``        -- tensor-level wrappers of the C ABI in include/wavemamba_b200.h;
  * ``enhance_bgr_u8`` -- the per-image body of inference_wavemamba.py (uint8 image in, uint8 out);
  * ``EnhancePipeline`` -- the same for a stream of images, PCIe copies overlapped with the forward;
  * ``GraphedForward`` -- CUDA-graph replay of a fixed-shape forward (the launch-bound LOL sizes);
  * ``metrics``    -- ``calculate_psnr`` / ``calculate_ssim`` of the inference loop, on the device;
  * ``build``      -- compiles csrc/*.cu into libwavemamba_b200.so (nvcc, sm_100a).
Importing the package does not need a GPU; calling any op without one raises.
"""
from ._cabi import WaveMambaNativeError, LIB_PATH  # noqa: F401
from . import ops  # noqa: F401
from .arch import WaveMamba, UNet  # noqa: F401
from .imageio import enhance_bgr_u8, EnhancePipeline  # noqa: F401
from .graph import GraphedForward  # noqa: F401
from . import metrics  # noqa: F401

__version__ = "0.1.0"
