"""PSNR / SSIM on the Y channel: the oracle against golden values produced by the
reference's own calculate_psnr / calculate_ssim (tools/make_golden_metrics.py), and the device kernel
(wm_psnr_ssim_y_u8) against both.

Tolerances.  The reference's PSNR is a float32 number (numpy forms the mean of float32 squares in float32,
pairwise): 2e-5 dB absolute covers its own rounding at 40 dB.  Its SSIM is fp64 throughout (cv2.filter2D on
float64 planes; OpenCV evaluates the 11x11 window by DFT): 1e-9 absolute.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import model as om

CASES = ["noise", "smooth", "ragged", "same"]
CROPS = [1, 0, 4]
PSNR_TOL, SSIM_TOL = 2e-5, 1e-9


@pytest.fixture(scope="module")
def gold():
    return load_golden("metrics")


def _close_psnr(got, want, tol=PSNR_TOL):
    if math.isinf(want):
        return math.isinf(got) and got > 0
    return abs(got - want) <= tol


def test_oracle_window_matches_cv2(gold):
    """cv2.getGaussianKernel(11, 1.5) as stored by the golden script: 1 ulp of the largest tap (5.6e-17)."""
    np.testing.assert_allclose(om.gaussian_window_11(), gold["window"].numpy(), rtol=0, atol=6e-17)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("crop", CROPS)
def test_oracle_metrics_vs_reference_golden(gold, case, crop):
    a, b = gold[f"{case}_a"].numpy(), gold[f"{case}_b"].numpy()
    assert _close_psnr(om.psnr_y(a, b, crop), float(gold[f"{case}_psnr_c{crop}"]))
    assert abs(om.ssim_y(a, b, crop) - float(gold[f"{case}_ssim_c{crop}"])) <= SSIM_TOL


def test_host_mirror_rejects_what_it_does_not_implement():
    """Same argument checks as the reference; the non-default paths fail loudly (no host fallback)."""
    from wave_mamba_b200 import metrics
    img = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(ValueError):
        metrics.calculate_psnr(img, img, input_order="WHC")
    with pytest.raises(NotImplementedError):
        metrics.calculate_psnr(img, img, input_order="CHW")
    with pytest.raises(NotImplementedError):
        metrics.calculate_ssim(img, img, test_y_channel=False)
    if not torch.cuda.is_available():
        with pytest.raises(Exception) as ei:
            metrics.calculate_psnr(img, img)
        assert "CUDA" in str(ei.value)


# ------------------------------------------------------------------------------------------- device
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("crop", CROPS)
def test_device_metrics_vs_reference_golden_and_oracle(gold, dev, case, crop):
    from wave_mamba_b200 import ops
    a, b = gold[f"{case}_a"], gold[f"{case}_b"]
    got = ops.psnr_ssim_y(a[None].to(dev), b[None].to(dev), crop).cpu()
    psnr, ssim = float(got[0, 0]), float(got[0, 1])
    assert _close_psnr(psnr, float(gold[f"{case}_psnr_c{crop}"]))
    assert abs(ssim - float(gold[f"{case}_ssim_c{crop}"])) <= SSIM_TOL
    # the oracle carries the mean in float32 as numpy does; the kernel in fp64
    assert _close_psnr(psnr, om.psnr_y(a.numpy(), b.numpy(), crop))
    assert abs(ssim - om.ssim_y(a.numpy(), b.numpy(), crop)) <= 1e-11


@pytest.mark.gpu
def test_device_metrics_host_mirror_and_batch(gold, dev):
    """calculate_psnr / calculate_ssim with the reference's call signature on cv2-style arrays, and a
    batch of two pairs in one launch."""
    from wave_mamba_b200 import metrics, ops
    a, b = gold["smooth_a"].numpy(), gold["smooth_b"].numpy()
    assert _close_psnr(metrics.calculate_psnr(a, b), float(gold["smooth_psnr_c1"]))
    assert abs(metrics.calculate_ssim(a, b) - float(gold["smooth_ssim_c1"])) <= SSIM_TOL
    with pytest.raises(AssertionError):
        metrics.calculate_psnr(a, b[:-1])
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    both = ops.psnr_ssim_y(torch.stack([ta, tb]), torch.stack([tb, tb]), 1).cpu()
    assert _close_psnr(float(both[0, 0]), float(gold["smooth_psnr_c1"]))
    assert math.isinf(float(both[1, 0])) and abs(float(both[1, 1]) - 1.0) <= 1e-12


@pytest.mark.gpu
def test_device_metrics_4k_properties(dev):
    """At the headline size (3840x2160) the oracle's tap-by-tap window takes minutes; the properties the
    definition gives: symmetric in its arguments (bit-exact), bit-reproducible, inf / 1 on identical
    images, and equal to the oracle on a 2160x64 strip cut from the same pair (every row tile, both
    vertical borders)."""
    from wave_mamba_b200 import ops
    g = torch.Generator().manual_seed(5)
    base = torch.rand(1, 3, 135, 240, generator=g)
    img = torch.nn.functional.interpolate(base, size=(2160, 3840), mode="bicubic", align_corners=False).clamp(0, 1)
    a = (img[0].permute(1, 2, 0) * 255).round().to(torch.uint8).contiguous()
    noise = torch.randint(-3, 4, a.shape, generator=g)
    b = (a.to(torch.int64) + noise).clamp(0, 255).to(torch.uint8)
    da, db = a[None].to(dev), b[None].to(dev)
    r1 = ops.psnr_ssim_y(da, db, 1)
    r2 = ops.psnr_ssim_y(db, da, 1)
    r3 = ops.psnr_ssim_y(da, db, 1)
    assert torch.equal(r1, r2) and torch.equal(r1, r3)
    assert 30.0 < float(r1[0, 0]) < 60.0 and 0.5 < float(r1[0, 1]) < 1.0
    same = ops.psnr_ssim_y(da, da.clone(), 1).cpu()
    assert math.isinf(float(same[0, 0])) and abs(float(same[0, 1]) - 1.0) <= 1e-12
    sa, sb = a[:, 1000:1064].contiguous(), b[:, 1000:1064].contiguous()
    got = ops.psnr_ssim_y(sa[None].to(dev), sb[None].to(dev), 1).cpu()
    assert _close_psnr(float(got[0, 0]), om.psnr_y(sa.numpy(), sb.numpy(), 1))
    assert abs(float(got[0, 1]) - om.ssim_y(sa.numpy(), sb.numpy(), 1)) <= 1e-11
