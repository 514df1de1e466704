"""The reference's OWN drivers, unmodified, driving this repo's plugin on the GPU.

baseline/_ref is a verbatim copy of the reference tree staged by tools/stage_reference.py
(git-ignored; it travels to the GPU box).  tools/ref_overlay.py overlays the one plugin file
(plugin/basicsr/archs/wavemamba_arch.py) on it and runs:

  * inference_wavemamba.py -i low -g high -w ckpt/WaveMamba_LOLv1.pth -o out   (reference
    inference_wavemamba.py:71-131: build, load, pad to x128, restoration_network, crop, tensor2img,
    PSNR/SSIM, imwrite) on two synthetic PNG pairs; the written images must equal
    wave_mamba_b200.enhance_bgr_u8 byte for byte and the printed PSNRs the oracle's within 1e-3 dB;
  * build_model(FeMaSRModel) -> feed_data -> test()   (basicsr/models/femasr_model.py:27,187-199);
  * build_model(FeMaSRModel, is_train=True) -> feed_data -> optimize_parameters()  (the body of
    basicsr/train.py's loop, femasr_model.py:157-185) with the optimiser / scheduler / loss options
    of options/train_wavemamba_lol.yml.

Stand-ins only for packages that are not installed in this image (timm, lmdb, pyiqa, torchmetrics, skimage).
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
CKPT = os.path.join(ROOT, "ckpt", "WaveMamba_LOLv1.pth")

from oracle import model as om  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "basicsr")),
                               reason="baseline/_ref not staged (python tools/stage_reference.py)")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@needs_ref
def test_inference_script_runs_unmodified_on_the_plugin(tmp_path, dev, params_cache):
    import cv2
    low, high, outd = tmp_path / "low", tmp_path / "high", tmp_path / "out"
    low.mkdir(); high.mkdir()
    cases = [("a.png", 256, 256, 0), ("b.png", 400, 600, 1)]
    for name, H, W, seed in cases:
        x, gt = om.synth_lowlight(1, H, W, seed)
        assert cv2.imwrite(str(low / name), om.to_uint8_bgr(x))
        assert cv2.imwrite(str(high / name), om.to_uint8_bgr(gt))
    cmd = [sys.executable, os.path.join(ROOT, "tools", "ref_overlay.py"), REF, str(tmp_path / "overlay"),
           "inference_wavemamba.py", "-i", str(low), "-g", str(high), "-w", CKPT, "-o", str(outd)]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-3000:]
    printed = [float(m) for m in re.findall(r"^psnr ([-0-9.einf]+)$", run.stdout, flags=re.M)]
    assert len(printed) == len(cases), run.stdout[-2000:]
    printed_ssim = [float(m) for m in re.findall(r"^ssim ([-0-9.einf]+)$", run.stdout, flags=re.M)]
    assert len(printed_ssim) == len(cases), run.stdout[-2000:]
    assert "avg_psnr" in run.stdout

    import wave_mamba_b200 as wm
    params = params_cache("LOLv1")
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    net.load_state_dict(params, strict=True)
    net = net.to(dev).eval()
    for (name, H, W, _), psnr_script, ssim_script in zip(cases, printed, printed_ssim):
        img = cv2.imread(str(low / name), cv2.IMREAD_UNCHANGED)
        gt_img = cv2.imread(str(high / name), cv2.IMREAD_UNCHANGED)
        written = cv2.imread(str(outd / name), cv2.IMREAD_UNCHANGED)
        # the script evaluates `img2tensor(img).to(device) / 255.` on a CUDA tensor, which torch
        # computes as a multiplication by the rounded reciprocal: cuda_division=True is that run
        ours = wm.enhance_bgr_u8(net, torch.from_numpy(img), window=128, cuda_division=True).cpu().numpy()
        assert written.shape == ours.shape == (H, W, 3)
        assert np.array_equal(written, ours), f"{name}: {(written != ours).sum()} bytes differ"
        ieee = wm.enhance_bgr_u8(net, torch.from_numpy(img), window=128).cpu().numpy()
        print(f"{name}: IEEE-division input path differs from the script's GPU run in "
              f"{(ieee != written).sum()} of {written.size} bytes")
        assert (ieee != written).sum() <= 1e-4 * written.size
        # the oracle on the same file
        xo = om.img_u8_to_f32(torch.from_numpy(img)[None], 128)
        yo = om.img_f32_to_u8(om.unet_forward(params, xo), H, W)[0].numpy()
        psnr_oracle = om.psnr_y(yo, gt_img)
        print(f"{name}: script psnr {psnr_script:.5f} dB, oracle {psnr_oracle:.5f} dB, "
              f"differing bytes vs oracle {(yo != written).sum()} of {written.size}")
        assert abs(psnr_script - psnr_oracle) <= 1e-3
        # the script's two host-side metric calls (:117-118), on the device (wm_psnr_ssim_y_u8)
        psnr_dev, ssim_dev = wm.metrics.calculate_psnr_ssim(written, gt_img)
        print(f"{name}: device psnr {psnr_dev:.6f} dB / ssim {ssim_dev:.9f}, script {psnr_script:.6f} / {ssim_script:.9f}")
        assert abs(psnr_dev - psnr_script) <= 2e-5 and abs(ssim_dev - ssim_script) <= 1e-9


@needs_ref
def test_femasr_model_validation_path_on_the_plugin(tmp_path, dev, params_cache):
    """build_model -> FeMaSRModel.__init__ (build_network, model_to_device, load_network strict,
    deepcopy) -> feed_data -> test() (net_g.test under no_grad), femasr_model.py:27,64-76,187-199."""
    x, _ = om.synth_lowlight(1, 128, 192, seed=4)
    torch.save(x, tmp_path / "x.pt")
    code = f"""
import sys, torch
sys.path.insert(0, {ROOT!r})
from tools import ref_overlay, ref_shims
ref_overlay.make_overlay({REF!r}, {str(tmp_path / 'overlay')!r})
ref_shims.install({str(tmp_path / 'overlay')!r})
from basicsr.models import build_model
opt = dict(model_type='FeMaSRModel', num_gpu=1, is_train=False, dist=False, val=dict(),
           network_g=dict(type='WaveMamba', in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0),
           path=dict(pretrain_network_g={CKPT!r}, strict_load=True))
model = build_model(opt)
assert type(model.net_g).__mro__[1].__module__ == 'wave_mamba_b200.arch'
model.feed_data(dict(lq=torch.load({str(tmp_path / 'x.pt')!r})))
model.test()
torch.save(model.output.cpu(), {str(tmp_path / 'y.pt')!r})
print('validation path ok', tuple(model.output.shape))
"""
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-3000:]
    y = torch.load(tmp_path / "y.pt")
    want = om.unet_forward(params_cache("LOLv1"), x)
    err = (y - want).abs().max().item()
    print(f"FeMaSRModel.test(): max abs err vs oracle {err:.3e}")
    assert err <= 2e-4


@needs_ref
def test_femasr_model_training_step_on_the_plugin(tmp_path, dev):
    """The reference's own training step driving the plugin: init_training_settings (losses, AdamW,
    CosineAnnealingRestartCyclicLR as options/train_wavemamba_lol.yml:73-97), feed_data,
    optimize_parameters twice (zero_grad -> net_g(lq) -> L1 (+FFT loss) -> backward -> step),
    update_learning_rate; the loss is finite and the weights move."""
    code = f"""
import sys, torch
sys.path.insert(0, {ROOT!r})
from tools import ref_overlay, ref_shims
from tools.synth import synth_lowlight
ref_overlay.make_overlay({REF!r}, {str(tmp_path / 'overlay')!r})
ref_shims.install({str(tmp_path / 'overlay')!r})
from basicsr.models import build_model
train = dict(optim_g=dict(type='AdamW', lr=1e-4, weight_decay=1e-3, betas=[0.9, 0.99]),
             scheduler=dict(type='CosineAnnealingRestartCyclicLR', periods=[100, 100000], restart_weights=[1, 1],
                            eta_mins=[0.0001, 0.0000001]),
             total_iter=101000, warmup_iter=-1,
             pixel_opt=dict(type='L1Loss', loss_weight=1.0, reduction='mean'),
             fft_opt=dict(type='FFTLoss', loss_weight=0.1, reduction='mean'))
opt = dict(model_type='FeMaSRModel', num_gpu=1, is_train=True, dist=False, val=dict(), train=train,
           network_g=dict(type='WaveMamba', in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0),
           path=dict(pretrain_network_g={CKPT!r}, strict_load=True))
model = build_model(opt)
assert type(model.net_g).__mro__[1].__module__ == 'wave_mamba_b200.arch'
before = [p.detach().clone() for p in model.net_g.parameters()]
x, gt = synth_lowlight(2, 64, 64, 11)
losses = []
for it in range(2):
    model.update_learning_rate(it + 1, warmup_iter=-1)
    model.feed_data(dict(lq=x, gt=gt))
    model.optimize_parameters(it + 1)
    losses.append({{k: float(v) for k, v in model.log_dict.items()}})
moved = sum(int(not torch.equal(a, b)) for a, b in zip(before, model.net_g.parameters()))
print('losses', losses, 'moved', moved, 'of', len(before))
assert all(v == v and abs(v) < 1e6 for d in losses for v in d.values())
assert moved == len(before)
"""
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-1500:] + run.stderr[-3000:]
    print(run.stdout.strip().splitlines()[-1])
