"""GPU parity of the whole forward through the reference-facing class (WaveMamba /
restoration_network) against the committed golden vectors (reference run) and the oracle.

Criteria: |PSNR_product - PSNR_reference| <= 1e-3 dB (BASELINE.json), raw fp32 max-abs error
<= 2e-4 on outputs in [0,1] with cuDNN/cuBLAS TF32 disabled on the product side, and equal
Matching argmin indices (SURVEY.md section 7 hazard)."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from oracle import model as om  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _net(params, dev):
    import wave_mamba_b200 as wm
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    net.load_state_dict(params, strict=True)
    return net.to(dev).eval()


@pytest.mark.parametrize("name", ["e2e_rand64_LOLv1", "e2e_synth48x80_UHDLL", "e2e_synth64_UHDLOL4K"])
def test_forward_matches_reference_golden(dev, params_cache, name):
    g = load_golden(name)
    net = _net(params_cache(g["ckpt"]), dev)
    with torch.no_grad():
        y = net.restoration_network(g["x"].to(dev)).cpu()
    err = (y - g["y"]).abs().max().item()
    print(f"{name}: max abs err vs reference {err:.3e}")
    assert err <= 2e-4
    if "gt" in g:
        for i in range(y.shape[0]):
            gt = om.to_uint8_bgr(g["gt"][i])
            d = abs(om.psnr_y(om.to_uint8_bgr(y[i]), gt) - om.psnr_y(om.to_uint8_bgr(g["y"][i]), gt))
            assert d <= 1e-3, d


def test_lol_400x600_batch_vs_oracle_with_matching_indices(dev, params_cache):
    """configs[1] shape family (400x600 padded to 8: 400x600 is already a multiple of 8) at
    batch 2 to keep the CPU oracle to seconds; argmin indices of every Matching call compared."""
    params = params_cache("LOLv1")
    x, gt = om.synth_lowlight(2, 200, 304, seed=0)
    trace = {}
    want = om.unet_forward(params, x, trace=trace)
    net = _net(params, dev)
    with torch.no_grad():
        y = net.restoration_network(x.to(dev)).cpu()
    err = (y - want).abs().max().item()
    # collect product-side indices in module order
    got_idx = {}
    for gname in ("down_group1", "down_group2", "down_group3", "up_group3", "up_group2", "up_group1"):
        grp = getattr(net.restoration_network, gname)
        for i, blk in enumerate(grp.h_blk):
            got_idx[f"{gname}.h{i}.attn"] = blk.attn.matching_transformation.last_index.cpu()
            got_idx[f"{gname}.h{i}.ffn"] = blk.ffn.matching_transformation.last_index.cpu()
    flips = {k: int((got_idx[k] != v).sum()) for k, v in trace["match_idx"].items()}
    print(f"max abs err {err:.3e}; argmin flips per call: {flips}")
    assert sum(flips.values()) == 0, flips
    assert err <= 2e-4
    for i in range(2):
        g8 = om.to_uint8_bgr(gt[i])
        d = abs(om.psnr_y(om.to_uint8_bgr(y[i]), g8) - om.psnr_y(om.to_uint8_bgr(want[i]), g8))
        assert d <= 1e-3, d


def test_registry_class_api(dev, params_cache):
    net = _net(params_cache("UHDLL"), dev)
    x, _ = om.synth_lowlight(1, 64, 64, seed=3)
    xd = x.to(dev)
    a = net(xd)
    b = net.test(xd)
    c = net.restoration_network(xd)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert not a.requires_grad
    padded = net.check_image_size(torch.rand(1, 3, 61, 70, device=dev))
    assert padded.shape[2] % 8 == 0 and padded.shape[3] % 8 == 0
    with pytest.raises(ValueError):
        net.restoration_network(torch.rand(1, 3, 60, 64, device=dev))
    net.train()
    with pytest.raises(NotImplementedError):
        net(xd)


def test_full_4k_forward_is_finite_and_deterministic(dev, params_cache):
    """BASELINE.json configs[2] shape: one 3840x2160 image.  The oracle cannot run this in
    seconds, so the checks are size-independent: finite output, run-to-run bit equality, and
    agreement of a 256x256 crop's *interior statistics* is not expected (global reductions),
    hence only invariants here; numerical parity at this size is covered kernel by kernel."""
    net = _net(params_cache("UHDLL"), dev)
    x, _ = om.synth_lowlight(1, 2160, 3840, seed=1234)
    xd = x.to(dev)
    with torch.no_grad():
        a = net.restoration_network(xd)
        b = net.restoration_network(xd)
    assert a.shape == xd.shape and torch.isfinite(a).all()
    assert torch.equal(a, b)


def test_cudnn_tf32_for_library_convs_is_measured_not_assumed(dev, params_cache):
    """The dense 3x3 convs are still library calls.  cuDNN's TF32 mode (the reference's own default
    on GPU) was measured to move the PSNR by 2.8e-3 dB on this input -- outside the 1e-3 dB budget
    -- so bench.py defaults to strict fp32 (--tf32 0).  This test pins both facts: fp32 mode meets
    the budget; TF32 mode stays a small, bounded deviation (and is reported, not hidden)."""
    params = params_cache("UHDLL")
    x, gt = om.synth_lowlight(1, 256, 384, seed=7)
    want = om.unet_forward(params, x)
    g8 = om.to_uint8_bgr(gt[0])
    deltas = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            net = _net(params, dev)
            with torch.no_grad():
                y = net.restoration_network(x.to(dev)).cpu()
        finally:
            torch.backends.cudnn.allow_tf32 = False
        deltas[tf32] = abs(om.psnr_y(om.to_uint8_bgr(y[0]), g8) - om.psnr_y(om.to_uint8_bgr(want[0]), g8))
        print(f"cudnn tf32={tf32}: |dPSNR| {deltas[tf32]:.2e} dB, max abs {(y - want).abs().max().item():.3e}, "
              f"differing uint8 values {int((om.to_uint8_bgr(y[0]) != om.to_uint8_bgr(want[0])).sum())}")
    assert deltas[False] <= 1e-3, deltas
    assert deltas[True] <= 5e-2, deltas


def test_lol_400x600_batch4_matches_oracle(dev, params_cache):
    """BASELINE configs[1] exactly: LOLv1 weights, 400x600, batch 4, fp32, against the CPU oracle:
    raw max-abs error <= 2e-4 and |dPSNR| <= 1e-3 dB per image."""
    params = params_cache("LOLv1")
    x, gt = om.synth_lowlight(4, 400, 600, seed=0)
    want = om.unet_forward(params, x)
    net = _net(params, dev)
    with torch.no_grad():
        y = net.restoration_network(x.to(dev)).cpu()
    err = (y - want).abs().max().item()
    print(f"400x600 x4: max abs err vs oracle {err:.3e}")
    assert err <= 2e-4
    for i in range(4):
        g8 = om.to_uint8_bgr(gt[i])
        d = abs(om.psnr_y(om.to_uint8_bgr(y[i]), g8) - om.psnr_y(om.to_uint8_bgr(want[i]), g8))
        assert d <= 1e-3, d
