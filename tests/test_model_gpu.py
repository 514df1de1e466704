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
    got_idx = _match_indices(net)
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
    b = net.test(xd)                                   # no_grad inside, as the reference's test()
    with torch.no_grad():
        a = net(xd)
        c = net.restoration_network(xd)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert not b.requires_grad
    padded = net.check_image_size(torch.rand(1, 3, 61, 70, device=dev))
    assert padded.shape[2] % 8 == 0 and padded.shape[3] % 8 == 0
    with pytest.raises(ValueError):
        net.restoration_network(torch.rand(1, 3, 60, 64, device=dev))
    # grad mode + trainable parameters (eval() or train(), as in PyTorch): the differentiable path
    d = net(xd)
    assert d.requires_grad and (d - a).abs().max().item() <= 2e-5
    for p in net.parameters():
        p.requires_grad_(False)
    e = net(xd)                                        # nothing to differentiate: the fused path
    assert not e.requires_grad and torch.equal(e, a)


class _Sink(dict):
    """Swallows the oracle's bulky trace entries (scan operands: tens of GB at 4K)."""

    def __setitem__(self, k, v):
        pass

    def setdefault(self, k, d=None):
        return _Sink()


class _MatchIdxOnly(dict):
    """Oracle trace that keeps only the Matching argmin indices."""

    def setdefault(self, k, d=None):
        return super().setdefault(k, d) if k == "match_idx" else _Sink()


def _match_indices(net):
    got = {}
    for gname in ("down_group1", "down_group2", "down_group3", "up_group3", "up_group2", "up_group1"):
        grp = getattr(net.restoration_network, gname)
        for i, blk in enumerate(grp.h_blk):
            got[f"{gname}.h{i}.attn"] = blk.attn.matching_transformation.last_index.cpu()
            got[f"{gname}.h{i}.ffn"] = blk.ffn.matching_transformation.last_index.cpu()
    return got


def test_full_4k_forward_matches_oracle(dev, params_cache):
    """BASELINE.json configs[2] exactly: one 3840x2160 image, UHD-LL weights, against the CPU oracle
    (~20 s on the box's host cores): raw fp32 max-abs error <= 2e-4, |dPSNR| <= 1e-3 dB, all 16
    Matching argmin index sets equal, the count of differing uint8 values reported; plus run-to-run
    bit equality of the product."""
    params = params_cache("UHDLL")
    x, gt = om.synth_lowlight(1, 2160, 3840, seed=0)
    net = _net(params, dev)
    xd = x.to(dev)
    with torch.no_grad():
        a = net.restoration_network(xd)
        got_idx = _match_indices(net)
        b = net.restoration_network(xd)
    assert a.shape == xd.shape and torch.isfinite(a).all()
    assert torch.equal(a, b)
    y = a.cpu()
    del a, b
    torch.cuda.empty_cache()
    trace = _MatchIdxOnly()
    want = om.unet_forward(params, x, trace=trace)
    assert len(trace["match_idx"]) == 16
    flips = {k: int((got_idx[k] != v).sum()) for k, v in trace["match_idx"].items()}
    diff = (y - want).abs()
    err = diff.max().item()
    y8, w8, g8 = om.to_uint8_bgr(y[0]), om.to_uint8_bgr(want[0]), om.to_uint8_bgr(gt[0])
    dpsnr = abs(om.psnr_y(y8, g8) - om.psnr_y(w8, g8))
    print(f"3840x2160: max abs err vs oracle {err:.3e}, rms {diff.pow(2).mean().sqrt().item():.3e}, "
          f"|dPSNR| {dpsnr:.2e} dB, differing uint8 values {int((y8 != w8).sum())} of {y8.size}, "
          f"argmin flips {sum(flips.values())}")
    assert sum(flips.values()) == 0, flips
    assert err <= 2e-4
    assert dpsnr <= 1e-3


def test_library_tf32_switches_do_not_touch_the_path(dev, params_cache):
    """No cuDNN / cuBLAS convolution or GEMM is left on the forward path (the dense 3x3 convs are
    the 3xTF32 tcgen05 kernel, fp32-accurate by construction), so torch's TF32 switches -- which the
    reference's GPU default leaves on for cuDNN -- must not change a single bit of the output."""
    params = params_cache("UHDLL")
    x, _ = om.synth_lowlight(1, 128, 192, seed=7)
    net = _net(params, dev)
    outs = []
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            with torch.no_grad():
                outs.append(net.restoration_network(x.to(dev)).cpu())
        finally:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
    assert torch.equal(outs[0], outs[1])


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


def test_cuda_graph_replay_equals_eager_400x600_batch4(dev, params_cache):
    """BASELINE configs[1] (400x600, batch 4) through wave_mamba_b200.GraphedForward: the replayed graph
    gives the eager result bit for bit, on a second input too (buffers and TMA descriptors are frozen at
    capture, the data is not), and the per-step time of both is reported."""
    import wave_mamba_b200 as wm
    net = _net(params_cache("LOLv1"), dev)
    x0, _ = om.synth_lowlight(4, 400, 600, seed=0)
    x1, _ = om.synth_lowlight(4, 400, 600, seed=1)
    x0, x1 = x0.to(dev), x1.to(dev)
    with torch.no_grad():
        want0 = net.restoration_network(x0).clone()
        want1 = net.restoration_network(x1).clone()
    g = wm.GraphedForward(net, x0.shape)
    assert torch.equal(g(x0), want0)
    assert torch.equal(g(x1), want1)
    assert torch.equal(g(x0), want0)

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    with torch.no_grad():
        t_eager = timed(lambda: net.restoration_network(x0))
    t_graph = timed(lambda: g(x0))
    print(f"400x600 batch 4: eager {t_eager:.2f} ms, CUDA graph {t_graph:.2f} ms per step")
    with pytest.raises(ValueError):
        g(x0[:2])


def test_enhance_pipeline_equals_enhance_bgr_u8(dev, params_cache):
    """EnhancePipeline (uploads / downloads on side streams, two staging slots) returns, image by image,
    the bytes of the one-stream enhance_bgr_u8 -- five different images through two slots, two shapes."""
    import wave_mamba_b200 as wm
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    net.load_state_dict(params_cache("LOLv1"), strict=True)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(9)
    shapes = [(200, 304)] * 3 + [(136, 248)] * 2
    imgs = [(torch.rand(h, w, 3, generator=g) * 60).to(torch.uint8).pin_memory() for h, w in shapes]
    outs = [torch.empty_like(i).pin_memory() for i in imgs]
    pipe = wm.EnhancePipeline(net, window=8)
    events = [pipe.submit(i, o) for i, o in zip(imgs, outs)]
    pipe.flush()
    assert all(e.query() for e in events)
    for i, o in zip(imgs, outs):
        want = wm.enhance_bgr_u8(net, i, window=8, sync=True).cpu()
        assert torch.equal(o, want)
    with pytest.raises(ValueError):
        pipe.submit(imgs[0].to(dev), outs[0])
