"""Device-side mirror of the two metric functions the reference's inference loop calls per image
(inference_wavemamba.py:23-24, 117-118 -> comput_psnr_ssim.py ``calculate_psnr`` :387-438 and
``calculate_ssim`` :596-668): same names, same arguments, same defaults, the same numbers -- computed by
``wm_psnr_ssim_y_u8`` on the B200 instead of numpy + five ``cv2.filter2D`` passes on the host.

Only the path the inference script takes is built: uint8 images (a cv2 BGR image or a uint8 tensor),
``input_order='HWC'``, ``test_y_channel=True``.  Anything else raises -- there is no host fallback.
"""
from typing import Tuple, Union

import numpy as np
import torch

from . import ops

ImageLike = Union[np.ndarray, torch.Tensor]


def _to_device_u8(img: ImageLike, device) -> torch.Tensor:
    if isinstance(img, np.ndarray):
        if img.dtype != np.uint8:
            raise NotImplementedError(f"only uint8 images are supported on the device path, got {img.dtype}")
        t = torch.from_numpy(np.ascontiguousarray(img)).to(device, non_blocking=True)
    elif isinstance(img, torch.Tensor):
        if img.dtype != torch.uint8:
            raise NotImplementedError(f"only uint8 images are supported on the device path, got {img.dtype}")
        t = img.to(device).contiguous()
    else:
        raise TypeError(f"expected a numpy array or a torch tensor, got {type(img)}")
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.shape[-1] != 3:
        raise NotImplementedError(f"expected (H,W,3) or (B,H,W,3) BGR images, got {tuple(t.shape)}")
    return t


def _check_options(input_order: str, test_y_channel: bool) -> None:
    if input_order not in ("HWC", "CHW"):
        raise ValueError(f'Wrong input_order {input_order}. Supported input_orders are "HWC" and "CHW"')
    if input_order != "HWC" or not test_y_channel:
        raise NotImplementedError("the device path implements the inference defaults only: "
                                  "input_order='HWC', test_y_channel=True")


def calculate_psnr_ssim(img1: ImageLike, img2: ImageLike, crop_border: int = 1, input_order: str = "HWC",
                        test_y_channel: bool = True, device=None) -> Tuple[float, float]:
    """Both metrics from one pass over the image pair.  Returns python floats for one image pair, two
    lists for a batch."""
    _check_options(input_order, test_y_channel)
    if device is None:
        device = next((t.device for t in (img1, img2) if isinstance(t, torch.Tensor) and t.is_cuda),
                      torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None)
    if device is None:
        raise ops._cabi.WaveMambaNativeError("calculate_psnr / calculate_ssim: no CUDA device (there is no host fallback)")
    a, b = _to_device_u8(img1, device), _to_device_u8(img2, device)
    if a.shape != b.shape:
        raise AssertionError(f"Image shapes are differnet: {tuple(a.shape[1:])}, {tuple(b.shape[1:])}.")
    res = ops.psnr_ssim_y(a, b, crop_border).cpu()
    single = (isinstance(img1, np.ndarray) and img1.ndim == 3) or (isinstance(img1, torch.Tensor) and img1.dim() == 3)
    if single:
        return float(res[0, 0]), float(res[0, 1])
    return res[:, 0].tolist(), res[:, 1].tolist()


def calculate_psnr(img1: ImageLike, img2: ImageLike, crop_border: int = 1, input_order: str = "HWC",
                   test_y_channel: bool = True) -> float:
    """comput_psnr_ssim.py:387-438 (defaults of the inference loop)."""
    return calculate_psnr_ssim(img1, img2, crop_border, input_order, test_y_channel)[0]


def calculate_ssim(img1: ImageLike, img2: ImageLike, crop_border: int = 1, input_order: str = "HWC",
                   test_y_channel: bool = True) -> float:
    """comput_psnr_ssim.py:596-668 (defaults of the inference loop)."""
    return calculate_psnr_ssim(img1, img2, crop_border, input_order, test_y_channel)[1]
