"""Training path (BASELINE configs[4]): the differentiable composition of this repo's
kernels (wave_mamba_b200.autograd) against torch autograd / the fp64 oracle.

Reference: the training step of basicsr/models/femasr_model.py:157-185 (net_g(lq) -> L1 -> backward ->
optimizer step) through wavemamba_arch.py's autograd graph and mamba_ssm's selective_scan_fn.
Tolerances: gradients within 1e-3 of the largest entry of each gradient tensor for whole-network checks
(fp32 kernels, 3xTF32 tensor-core products, sums over all pixels) and 1e-4 for single operators."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import model as om  # noqa: E402
from oracle import scan as oscan  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _rel(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _check_fn(fn, ref, inputs, dev, tol=1e-4):
    """fn: product Function on CUDA fp32 leaves; ref: the same math in torch on fp64 CPU leaves."""
    g = torch.Generator().manual_seed(99)
    leaves64 = [t.double().clone().requires_grad_(True) for t in inputs]
    leaves32 = [t.to(dev).clone().requires_grad_(True) for t in inputs]
    y64, y32 = ref(*leaves64), fn(*leaves32)
    assert _rel(y32, y64) <= tol, ("forward", _rel(y32, y64))
    gy = torch.randn(y64.shape, generator=g)
    want = torch.autograd.grad(y64, leaves64, gy.double())
    got = torch.autograd.grad(y32, leaves32, gy.to(dev))
    for i, (a, b) in enumerate(zip(got, want)):
        assert _rel(a, b) <= tol, (f"grad of input {i}", _rel(a, b))


def _r(*shape, g, s=1.0):
    return s * torch.randn(*shape, generator=g)


def test_pointwise_depthwise_layernorm_functions(dev):
    from wave_mamba_b200 import autograd as ag
    g = torch.Generator().manual_seed(1)
    for cin, cout in ((32, 32), (32, 64), (64, 32), (32, 96), (64, 64)):
        _check_fn(lambda x, w, b: ag.PW.apply(x, w, b), lambda x, w, b: F.conv2d(x, w, b),
                  [_r(2, cin, 9, 14, g=g), _r(cout, cin, 1, 1, g=g, s=0.2), _r(cout, g=g, s=0.1)], dev)
    for c in (32, 64, 96):
        _check_fn(lambda x, w, b: ag.DW.apply(x, w, b), lambda x, w, b: F.conv2d(x, w, b, padding=1, groups=x.shape[1]),
                  [_r(2, c, 9, 14, g=g), _r(c, 1, 3, 3, g=g, s=0.3), _r(c, g=g, s=0.1)], dev)
    for c, eps in ((32, 1e-6), (64, 1e-5)):
        def ref(x, w, b, eps=eps):
            mu = x.mean(1, keepdim=True)
            var = (x - mu).pow(2).mean(1, keepdim=True)
            return (x - mu) / (var + eps).sqrt() * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        _check_fn(lambda x, w, b, eps=eps: ag.LN2d.apply(x, w, b, eps), ref,
                  [_r(2, c, 9, 14, g=g), 1 + _r(c, g=g, s=0.1), _r(c, g=g, s=0.1)], dev)


def test_dense_conv_and_edge_functions(dev):
    from wave_mamba_b200 import autograd as ag
    g = torch.Generator().manual_seed(2)
    for cin, cout in ((64, 32), (64, 64), (32, 96)):
        _check_fn(lambda x, w, b: ag.Conv3.apply(x, w, b), lambda x, w, b: F.conv2d(x, w, b, padding=1),
                  [_r(2, cin, 11, 18, g=g), _r(cout, cin, 3, 3, g=g, s=0.1), _r(cout, g=g, s=0.1)], dev)
    # stem / head / ps_down: only the weights (and the head's feature input) need gradients
    img = torch.rand(2, 3, 16, 24, generator=g)
    _check_fn(lambda w, b: ag.StemConv.apply(img.to(dev), w, b), lambda w, b: F.conv2d(img.double(), w, b, padding=1),
              [_r(32, 3, 3, 3, g=g, s=0.2), _r(32, g=g, s=0.1)], dev)
    _check_fn(lambda x, w, b: ag.HeadConv.apply(x, w, b, img.to(dev)),
              lambda x, w, b: F.conv2d(x, w, b, padding=1) + img.double(),
              [_r(2, 32, 16, 24, g=g), _r(3, 32, 3, 3, g=g, s=0.1), _r(3, g=g, s=0.1)], dev)
    for r in (2, 4, 8):
        _check_fn(lambda w, b, r=r: ag.PSDown.apply(img.to(dev), w, b, r),
                  lambda w, b, r=r: F.conv2d(F.pixel_unshuffle(img.double(), r), w, b),
                  [_r(32, 3 * r * r, 1, 1, g=g, s=0.2), _r(32, g=g, s=0.1)], dev)


def test_gram_and_per_image_pointwise_functions(dev):
    from wave_mamba_b200 import autograd as ag
    g = torch.Generator().manual_seed(3)

    def gram_ref(x, y):
        xf, yf = x.flatten(2), y.flatten(2)
        return torch.cat([(xf @ yf.transpose(1, 2)).flatten(1), xf.pow(2).sum(-1), yf.pow(2).sum(-1)], dim=1)

    def gram_fn(x, y):
        gm, nx, ny = ag.Gram32.apply(x, y)
        return torch.cat([gm.flatten(1), nx, ny], dim=1)

    _check_fn(gram_fn, gram_ref, [_r(2, 32, 9, 14, g=g), _r(2, 32, 9, 14, g=g)], dev)
    _check_fn(lambda x, w, b, r: ag.PWPerImage.apply(x, w, b, r),
              lambda x, w, b, r: torch.einsum("boc,bchw->bohw", w, x) + b.view(1, -1, 1, 1) + r,
              [_r(2, 32, 9, 14, g=g), _r(2, 32, 32, g=g, s=0.2), _r(32, g=g, s=0.1), _r(2, 32, 9, 14, g=g)], dev)


def _net(params, dev):
    import wave_mamba_b200 as wm
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    net.load_state_dict(params, strict=True)
    return net.to(dev)


def test_training_forward_equals_inference_forward(dev, params_cache):
    params = params_cache("UHDLOL4K")
    x, _ = om.synth_lowlight(2, 64, 96, seed=5)
    net = _net(params, dev)
    net.eval()
    with torch.no_grad():
        want = net(x.to(dev))
    net.train()
    got = net(x.to(dev))
    assert got.requires_grad
    err = (got - want).abs().max().item()
    print(f"train-path forward vs fused inference forward: max abs diff {err:.3e}")
    assert err <= 2e-5


def test_whole_network_gradients_vs_oracle_autograd(dev, params_cache):
    """One fwd + bwd of the whole network (L1-like random upstream gradient) against autograd through the
    fp64 oracle (pure-torch sequential scan): all 591 parameter gradients."""
    params = params_cache("UHDLOL4K")
    x, gt = om.synth_lowlight(1, 32, 48, seed=6)
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in om.strip_prefix(params).items()}
    trace = {}
    y64 = om.unet_forward(p64, x.double(), scan_fn=oscan.selective_scan_loop, trace=trace)
    gy = torch.sign(y64.detach() - gt.double()) / y64.numel()          # dL/dy of the L1 loss (:173)
    names = sorted(p64)
    want = dict(zip(names, torch.autograd.grad(y64, [p64[n] for n in names], gy)))

    net = _net(params, dev).train()
    y = net(x.to(dev))
    idx_ok = all(torch.equal(getattr(net.restoration_network, gname).h_blk[i].__getattr__(part)
                             .matching_transformation.last_index.cpu(), trace["match_idx"][f"{gname}.h{i}.{part}"])
                 for gname in ("down_group1", "down_group2", "down_group3", "up_group3", "up_group2", "up_group1")
                 for i in range(len(getattr(net.restoration_network, gname).h_blk)) for part in ("attn", "ffn"))
    assert idx_ok, "Matching argmin differs between the product and the fp64 oracle on this input"
    assert (y.detach().cpu().double() - y64.detach()).abs().max().item() <= 2e-4
    y.backward(gy.float().to(dev))
    worst, worst_name, missing = 0.0, None, []
    for n, p in net.named_parameters():
        key = n[len("restoration_network."):]
        if p.grad is None:
            missing.append(n)
            continue
        e = _rel(p.grad, want[key])
        if e > worst:
            worst, worst_name = e, n
    print(f"whole-network gradients: worst relative error {worst:.2e} at {worst_name}; "
          f"{len(missing)} parameters without gradient")
    assert not missing, missing[:5]
    assert worst <= 1e-3


def test_one_optimizer_step_like_the_reference(dev, params_cache):
    """optimize_parameters (femasr_model.py:157-185): zero_grad -> net_g(lq) -> L1 -> backward -> AdamW
    step, twice; the loss must be finite and every parameter must move."""
    net = _net(params_cache("UHDLOL4K"), dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0.0, betas=(0.9, 0.99))
    x, gt = om.synth_lowlight(2, 64, 64, seed=7)
    x, gt = x.to(dev), gt.to(dev)
    before = [p.detach().clone() for p in net.parameters()]
    losses = []
    for _ in range(2):
        opt.zero_grad()
        loss = F.l1_loss(net(x), gt)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    print(f"two training steps at 2x3x64x64: L1 {losses[0]:.5f} -> {losses[1]:.5f}")
    assert all(torch.isfinite(torch.tensor(losses)))
    moved = sum(int(not torch.equal(a, b)) for a, b in zip(before, net.parameters()))
    assert moved == len(before), f"{len(before) - moved} parameters did not move"


def test_bf16_activation_storage_same_forward_close_gradients_less_memory(dev, params_cache):
    """SURVEY 8 f-3 'bf16 activation storage with fp32 scan state' (opt-in: net.activation_storage = 'bf16'
    or WM_ACT_STORAGE=bf16).  What is computed stays fp32 -- the loss is bit-identical -- only the feature
    maps autograd keeps are bf16: every parameter gradient must keep its direction (cosine >= 0.99 against
    the fp32-storage gradient, >= 0.999 over all parameters together) and the peak memory of the step must
    drop by at least 25 % (measured: 845 -> 559 MiB)."""
    params = params_cache("UHDLOL4K")
    x, gt = om.synth_lowlight(2, 256, 256, seed=8)
    x, gt = x.to(dev), gt.to(dev)

    def step(storage):
        net = _net(params, dev).train()
        net.restoration_network.activation_storage = storage
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        loss = F.l1_loss(net(x), gt)
        loss.backward()
        torch.cuda.synchronize()
        peak = torch.cuda.max_memory_allocated() - base
        return float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters()}, peak

    l32, g32, m32 = step(None)
    l16, g16, m16 = step("bf16")
    assert l32 == l16
    worst, worst_name = 1.0, None
    for n in g32:
        a, b = g32[n].flatten().double(), g16[n].flatten().double()
        if a.norm() == 0:
            continue
        c = float(torch.dot(a, b) / (a.norm() * b.norm()))
        if c < worst:
            worst, worst_name = c, n
    fa = torch.cat([g.flatten().double() for g in g32.values()])
    fb = torch.cat([g16[n].flatten().double() for n in g32])
    total = float(torch.dot(fa, fb) / (fa.norm() * fb.norm()))
    rel = float((fa - fb).norm() / fa.norm())
    print(f"bf16 activation storage at 2x3x256x256: peak {m32 / 2**20:.0f} -> {m16 / 2**20:.0f} MiB, "
          f"gradient cosine {total:.6f} overall (relative distance {rel:.2e}), worst parameter {worst:.4f} at {worst_name}")
    assert total >= 0.999 and worst >= 0.99
    assert m16 <= 0.75 * m32
