"""GPU parity tests: every kernel is called through the C ABI (wave_mamba_b200.ops -> ctypes ->
libwavemamba_b200.so) and compared with the CPU oracle on the same seeded inputs.

Tolerances (stated per test):
  * DWT / IWT: bit-exact.
  * SS2D core: |gpu - fp64 arbiter| <= 2e-5 * max(1, max|y|)  (fp32 recurrence; the CPU fp32
    oracle itself sits at ~1e-6 from the arbiter; the GPU uses ex2.approx and a chunked carry).
  * pointwise / depthwise groups: 2e-5 absolute on O(1) activations (different FMA order).
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

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


@pytest.fixture(scope="module")
def ops(dev):
    from wave_mamba_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------- DWT / IWT
@pytest.mark.parametrize("shape", [(2, 5, 6, 10), (1, 3, 16, 24), (2, 32, 40, 64), (1, 7, 2, 2),
                                   (1, 32, 270, 480), (3, 4, 10, 14)])
def test_dwt_bit_exact(ops, dev, shape):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g)
    want = om.haar_dwt(x)
    got = ops.dwt_haar(x.to(dev))
    for a, b, name in zip(got, want, ("ll", "hl", "lh", "hh")):
        assert torch.equal(a.cpu(), b), name


def test_dwt_golden(ops, dev):
    g = load_golden("dwt")
    got = ops.dwt_haar(g["x"].to(dev))
    for a, key in zip(got, ("ll", "hl", "lh", "hh")):
        assert torch.equal(a.cpu(), g[key])


@pytest.mark.parametrize("shape", [(2, 3, 3, 7), (1, 32, 20, 32), (2, 8, 5, 12), (1, 1, 1, 1),
                                   (1, 32, 135, 240)])
def test_iwt_bit_exact(ops, dev, shape):
    B, C, h, w = shape
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 4 * C, h, w, generator=g)
    want = om.haar_iwt(x)
    assert torch.equal(ops.iwt_haar_cat(x.to(dev)).cpu(), want)
    low, high = x[:, :C].contiguous(), x[:, C:].contiguous()
    assert torch.equal(ops.iwt_haar(low.to(dev), high.to(dev)).cpu(), want)


def test_iwt_golden(ops, dev):
    g = load_golden("iwt")
    assert torch.equal(ops.iwt_haar_cat(g["x"].to(dev)).cpu(), g["y"])


def test_dwt_iwt_round_trip_full_4k_level1(ops, dev):
    """Size-independent property at BASELINE.json's full size: IWT(DWT(x)) == x (to rounding)."""
    x = torch.randn(1, 32, 2160, 3840, device=dev)
    ll, hl, lh, hh = ops.dwt_haar(x)
    back = ops.iwt_haar(ll, torch.cat([hl, lh, hh], dim=1))
    assert (back - x).abs().max().item() <= 1e-6 * max(1.0, x.abs().max().item())
    # energy preservation (orthonormal transform)
    e_in = x.double().pow(2).sum()
    e_out = sum(t.double().pow(2).sum() for t in (ll, hl, lh, hh))
    assert abs((e_out / e_in).item() - 1.0) < 1e-6


def test_empty_inputs_are_noops(ops, dev):
    x = torch.empty(0, 4, 8, 8, device=dev)
    assert ops.dwt_haar(x)[0].shape == (0, 4, 4, 4)


# ------------------------------------------------------------------------------- SS2D core
def _ss_params(params_cache, ckpt="UHDLOL4K", block="down_group1.l_blk.0"):
    p = om.sub(om.strip_prefix(params_cache(ckpt)), block + ".self_attention")
    return [p[k] for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")]


def _arbiter(x, prm):
    """fp64 evaluation of the whole core (projections and recurrence in double)."""
    return om.ss2d_core(x.double(), *[t.double() for t in prm], scan_fn=oscan.selective_scan_c)


@pytest.mark.parametrize("shape", [(1, 64, 6, 10), (2, 64, 9, 5), (1, 64, 40, 56), (1, 64, 50, 75),
                                   (2, 64, 33, 64), (1, 64, 135, 240), (1, 64, 1, 1), (1, 64, 3, 130)])
def test_ss2d_core_vs_oracle(ops, dev, params_cache, shape):
    g = torch.Generator().manual_seed(0)
    x = F.silu(0.5 * torch.randn(*shape, generator=g))
    prm = _ss_params(params_cache)
    want64 = _arbiter(x, prm)
    want32 = om.ss2d_core(x, *prm)
    got = ops.ss2d_core(x.to(dev), *[t.to(dev) for t in prm]).cpu()
    scale = max(1.0, want64.abs().max().item())
    err_gpu = (got.double() - want64).abs().max().item()
    err_cpu = (want32.double() - want64).abs().max().item()
    print(f"{shape}: gpu-vs-f64 {err_gpu:.2e}  cpu32-vs-f64 {err_cpu:.2e}  scale {scale:.2f}")
    assert err_gpu <= 2e-5 * scale


@pytest.mark.parametrize("tag", ["a", "b"])
def test_ss2d_core_golden(ops, dev, params_cache, tag):
    g = load_golden(f"ss2d_core_{tag}")
    prm = _ss_params(params_cache, g["ckpt"], g["block"])
    got = ops.ss2d_core(g["x"].to(dev), *[t.to(dev) for t in prm]).cpu()
    torch.testing.assert_close(got, g["y"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("ckpt,block", [("LOLv1", "down_group3.l_blk.3"), ("UHDLL", "up_group2.l_blk.1")])
def test_ss2d_core_other_checkpoints(ops, dev, params_cache, ckpt, block):
    g = torch.Generator().manual_seed(5)
    x = F.silu(0.7 * torch.randn(1, 64, 24, 40, generator=g))
    prm = _ss_params(params_cache, ckpt, block)
    want64 = _arbiter(x, prm)
    got = ops.ss2d_core(x.to(dev), *[t.to(dev) for t in prm]).cpu()
    assert (got.double() - want64).abs().max().item() <= 2e-5 * max(1.0, want64.abs().max().item())


@pytest.mark.parametrize("shape", [(1, 64, 100, 37), (1, 64, 300, 8), (1, 64, 257, 9), (2, 64, 160, 20)])
def test_ss2d_core_with_segmented_columns(ops, dev, params_cache, shape):
    """Shapes for which the chunk planner cuts a column into 2..5 segments (the 4K level-2 map,
    540x960, runs with 2): the carry must chain the segments of a column before moving to the next
    column.  Same tolerance as test_ss2d_core_vs_oracle."""
    import ctypes
    from wave_mamba_b200 import _cabi
    geo = (ctypes.c_int * 6)()
    assert _cabi.load().wm_ss2d_debug_geometry(shape[0], shape[2], shape[3], geo) == 0
    assert geo[3] >= 2, f"planner no longer segments the columns of {shape}: pick another shape"
    g = torch.Generator().manual_seed(2)
    x = F.silu(0.5 * torch.randn(*shape, generator=g))
    prm = _ss_params(params_cache)
    want64 = _arbiter(x, prm)
    got = ops.ss2d_core(x.to(dev), *[t.to(dev) for t in prm]).cpu()
    scale = max(1.0, want64.abs().max().item())
    err = (got.double() - want64).abs().max().item()
    print(f"{shape}: {geo[3]} column segments, gpu-vs-f64 {err:.2e} (scale {scale:.2f})")
    assert err <= 2e-5 * scale


def _arbiter_streamed(x, prm):
    """The fp64 arbiter of ``_arbiter`` evaluated one direction at a time (same arithmetic; peak
    host memory ~5 GB instead of ~20 GB at the 4K level-1 size).  Index maps: SURVEY appendix A."""
    xp, dw, db, al, ds = [t.double() for t in prm]
    B, D, h, w = x.shape
    L = h * w
    A = -torch.exp(al)
    row = x.reshape(B, D, L)
    col = x.transpose(2, 3).reshape(B, D, L)
    y = torch.zeros(B, D, L, dtype=torch.float64)
    for k in range(4):
        seq = col if k % 2 else row
        seq = (seq.flip(-1) if k >= 2 else seq).double().contiguous()
        proj = torch.einsum("cd,bdl->bcl", xp[k], seq)
        dt_low, Bm, Cm = torch.split(proj, [2, 16, 16], dim=1)
        delta = torch.einsum("dr,brl->bdl", dw[k], dt_low)
        out = oscan.selective_scan_c(seq, delta, A[k * D:(k + 1) * D], Bm[:, None].contiguous(),
                                     Cm[:, None].contiguous(), ds[k * D:(k + 1) * D], db[k])
        del proj, dt_low, Bm, Cm, delta, seq
        if k >= 2:
            out = out.flip(-1)
        if k % 2:
            out = out.reshape(B, D, w, h).transpose(2, 3).reshape(B, D, L)
        y += out
    return y.reshape(B, D, h, w)


def test_streamed_arbiter_is_the_arbiter(params_cache):
    g = torch.Generator().manual_seed(11)
    x = F.silu(0.5 * torch.randn(2, 64, 9, 14, generator=g))
    prm = _ss_params(params_cache)
    a, b = _arbiter(x, prm), _arbiter_streamed(x, prm)
    assert (a - b).abs().max().item() <= 1e-12 * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("shape", [(1, 64, 540, 960), (1, 64, 1080, 1920)])
def test_ss2d_core_vs_arbiter_at_4k_level_sizes(ops, dev, params_cache, shape):
    """The SS2D core at the sizes bench.py runs (BASELINE configs[2]: 4K level-2 and level-1 maps,
    L = 518 400 / 2 073 600, the production chunk plans) against the fp64 arbiter; same tolerance
    as the small shapes: |gpu - f64| <= 2e-5 * max(1, max|y|)."""
    g = torch.Generator().manual_seed(0)
    x = F.silu(0.5 * torch.randn(*shape, generator=g))
    prm = _ss_params(params_cache)
    got = ops.ss2d_core(x.to(dev), *[t.to(dev) for t in prm]).cpu()
    want64 = _arbiter_streamed(x, prm)
    scale = max(1.0, want64.abs().max().item())
    diff = (got.double() - want64).abs()
    err = diff.max().item()
    print(f"{shape}: gpu-vs-f64 max {err:.2e} rms {diff.pow(2).mean().sqrt().item():.2e} (scale {scale:.2f})")
    assert err <= 2e-5 * scale


@pytest.mark.parametrize("shape", [(1, 64, 6, 10), (2, 64, 9, 5), (1, 64, 24, 40), (1, 64, 70, 9)])
def test_ss2d_core_backward_vs_oracle_autograd(ops, dev, params_cache, shape):
    """wm_ss2d_core_bwd against autograd through the fp64 oracle (pure-torch sequential scan,
    pinned by the finite-difference gradcheck in tests/test_oracle.py): gradients with respect to the
    input map and all five parameter tensors, for a random upstream gradient.  Tolerance: 2e-4 of
    the largest entry of each gradient (fp32 recurrences in both time directions, sums over L)."""
    g = torch.Generator().manual_seed(21)
    x = F.silu(0.5 * torch.randn(*shape, generator=g))
    gy = torch.randn(*shape, generator=g)
    prm = _ss_params(params_cache)
    leaves = [x.double().requires_grad_(True)] + [t.double().clone().requires_grad_(True) for t in prm]
    y = om.ss2d_core(*leaves, scan_fn=oscan.selective_scan_loop)
    want = torch.autograd.grad(y, leaves, gy.double())
    got = ops.ss2d_core_bwd(x.to(dev), *[t.to(dev) for t in prm], gy.to(dev))
    names = ("x", "x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")
    worst = 0.0
    for n, a, b in zip(names, got, want):
        scale = max(b.abs().max().item(), 1e-30)
        err = (a.cpu().double() - b).abs().max().item() / scale
        print(f"{shape} d{n}: rel err {err:.2e} (max |grad| {scale:.3e})")
        worst = max(worst, err)
    assert worst <= 2e-4
    again = ops.ss2d_core_bwd(x.to(dev), *[t.to(dev) for t in prm], gy.to(dev))
    assert all(torch.equal(a, b) for a, b in zip(got, again))


def test_ss2d_core_is_deterministic(ops, dev, params_cache):
    x = F.silu(torch.randn(2, 64, 48, 72, device=dev))
    prm = [t.to(dev) for t in _ss_params(params_cache)]
    a = ops.ss2d_core(x, *prm)
    b = ops.ss2d_core(x, *prm)
    assert torch.equal(a, b)


def test_ss2d_core_symmetries_at_full_4k_level1_size(ops, dev, params_cache):
    """Size-independent properties on the full (1,64,1080,1920) map: with the four directions
    sharing one weight set, the operator commutes with a 180-degree rotation (dir0<->dir2,
    dir1<->dir3) and with a transpose (dir0<->dir1, dir2<->dir3)."""
    xp, dw, db, al, ds = [t.to(dev) for t in _ss_params(params_cache)]
    xp = xp[:1].repeat(4, 1, 1).contiguous()
    dw = dw[:1].repeat(4, 1, 1).contiguous()
    db = db[:1].repeat(4, 1).contiguous()
    al = al[:64].repeat(4, 1).contiguous()
    ds = ds[:64].repeat(4).contiguous()
    x = F.silu(0.5 * torch.randn(1, 64, 1080, 1920, device=dev))
    y = ops.ss2d_core(x, xp, dw, db, al, ds)
    assert torch.isfinite(y).all()
    y_rot = ops.ss2d_core(x.flip(2, 3).contiguous(), xp, dw, db, al, ds).flip(2, 3)
    y_tr = ops.ss2d_core(x.transpose(2, 3).contiguous(), xp, dw, db, al, ds).transpose(2, 3)
    scale = y.abs().max().item()
    assert (y - y_rot).abs().max().item() <= 2e-5 * scale
    assert (y - y_tr).abs().max().item() <= 2e-5 * scale


# ------------------------------------------------------------------------------- pointwise / depthwise
def _rand(*shape, g, s=1.0):
    return s * torch.randn(*shape, generator=g)


@pytest.mark.parametrize("cout", [32, 64, 96])
@pytest.mark.parametrize("use_ln", [False, True])
@pytest.mark.parametrize("hw", [(13, 37), (8, 32), (24, 70), (40, 96)])
def test_pw_dw(ops, dev, cout, use_ln, hw):
    g = torch.Generator().manual_seed(7)
    h, w = hw
    x = _rand(2, 32, h, w, g=g)
    pw_w, pw_b = _rand(cout, 32, 1, 1, g=g, s=0.2), _rand(cout, g=g, s=0.1)
    dw_w, dw_b = _rand(cout, 1, 3, 3, g=g, s=0.3), _rand(cout, g=g, s=0.1)
    ln_w, ln_b = (1 + _rand(32, g=g, s=0.1), _rand(32, g=g, s=0.1)) if use_ln else (None, None)
    t = om.layer_norm_2d(x, ln_w, ln_b, 1e-6) if use_ln else x
    want = F.conv2d(F.conv2d(t, pw_w, pw_b), dw_w, dw_b, padding=1, groups=cout)
    d = lambda v: None if v is None else v.to(dev)
    got = ops.pw_dw(d(x), d(pw_w), d(pw_b), d(dw_w), d(dw_b), d(ln_w), d(ln_b), 1e-6).cpu()
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("cout,act", [(32, "none"), (64, "silu"), (64, "none"), (96, "none")])
def test_pw_dw_many_tiles_per_cta(ops, dev, cout, act):
    """The TMA + tcgen05 pipeline on a map where every persistent CTA walks several tiles (ragged
    bottom rows, batch 2: 650 tiles on 148 SMs), against the fp64 composition on the GPU; and the
    legacy cp.async / mma.sync kernel (w % 4 != 0) on the same kind of map."""
    g = torch.Generator().manual_seed(8)
    for h, w in ((203, 400), (61, 150)):
        x = _rand(2, 32, h, w, g=g).to(dev)
        pw_w, pw_b = _rand(cout, 32, 1, 1, g=g, s=0.2).to(dev), _rand(cout, g=g, s=0.1).to(dev)
        dw_w, dw_b = _rand(cout, 1, 3, 3, g=g, s=0.3).to(dev), _rand(cout, g=g, s=0.1).to(dev)
        ln_w, ln_b = (1 + _rand(32, g=g, s=0.1)).to(dev), _rand(32, g=g, s=0.1).to(dev)
        xd = x.double()
        mu = xd.mean(1, keepdim=True)
        t = (xd - mu) / ((xd - mu).pow(2).mean(1, keepdim=True) + 1e-6).sqrt()
        t = t * ln_w.double().view(1, -1, 1, 1) + ln_b.double().view(1, -1, 1, 1)
        want = F.conv2d(F.conv2d(t, pw_w.double(), pw_b.double()), dw_w.double(), dw_b.double(), padding=1,
                        groups=cout)
        if act == "silu":
            want = F.silu(want)
        got = ops.pw_dw(x, pw_w, pw_b, dw_w, dw_b, ln_w, ln_b, 1e-6, act=act)
        err = (got.double() - want).abs().max().item()
        assert err <= 2e-5 * max(1.0, want.abs().max().item()), (h, w, err)
        assert torch.equal(got, ops.pw_dw(x, pw_w, pw_b, dw_w, dw_b, ln_w, ln_b, 1e-6, act=act))


@pytest.mark.parametrize("hw", [(13, 37), (16, 64), (156, 160)])
@pytest.mark.parametrize("with_res", [False, True])
def test_dw_act_pw(ops, dev, hw, with_res):
    """(13, 37): the cp.async kernel (w % 4 != 0); (16, 64): the TMA pipeline, one tile per CTA;
    (156, 160): 200 tiles = several per persistent CTA, with a partial bottom row of tiles."""
    g = torch.Generator().manual_seed(8)
    h, w = hw
    x = _rand(2, 32, h, w, g=g)
    dw_w, dw_b = _rand(32, 1, 3, 3, g=g, s=0.3), _rand(32, g=g, s=0.1)
    pw_w, pw_b = _rand(32, 32, 1, 1, g=g, s=0.2), _rand(32, g=g, s=0.1)
    res = _rand(2, 32, h, w, g=g) if with_res else None
    want = F.conv2d(F.gelu(F.conv2d(x, dw_w, dw_b, padding=1, groups=32)), pw_w, pw_b)
    if with_res:
        want = want + res
    d = lambda v: None if v is None else v.to(dev)
    got = ops.dw_act_pw(d(x), d(dw_w), d(dw_b), d(pw_w), d(pw_b), "gelu", d(res)).cpu()
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)


def test_pw_plain_gate_and_residual(ops, dev):
    g = torch.Generator().manual_seed(9)
    d = lambda v: None if v is None else v.to(dev)
    for cin, cout in ((32, 32), (32, 64), (64, 32)):
        x = _rand(2, cin, 11, 29, g=g)
        w_, b_ = _rand(cout, cin, 1, 1, g=g, s=0.2), _rand(cout, g=g, s=0.1)
        res = _rand(2, cout, 11, 29, g=g)
        want = F.conv2d(x, w_, b_) + res
        torch.testing.assert_close(ops.pw(d(x), d(w_), d(b_), residual=d(res)).cpu(), want,
                                   rtol=2e-5, atol=2e-5)
    x = _rand(2, 64, 11, 29, g=g)
    w_, b_ = _rand(32, 32, 1, 1, g=g, s=0.2), _rand(32, g=g, s=0.1)
    a, b2 = x.chunk(2, dim=1)
    want = F.conv2d(F.gelu(a) * b2, w_, b_)
    torch.testing.assert_close(ops.pw(d(x), d(w_), d(b_), gate=True).cpu(), want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("hw", [(11, 29), (12, 30), (12, 32), (160, 200)])
def test_pw_per_image_weights_on_a_channel_slice(ops, dev, hw):
    """CMTAttention tail (reference :791-797, :849): v is the last third of the qkv tensor (batch
    stride 96*h*w) and every image has its own 32x32 matrix (project_out folded with the attention);
    one launch.  Odd and even pixel counts take the one- and the two-pixel kernels, multiples of 4 the
    TMA pipeline (128-pixel tiles): 384 pixels = 3 tiles per image, 160 x 200 x 3 images = 750 tiles =
    several per persistent CTA, which reloads the weights when it crosses into the next image."""
    g = torch.Generator().manual_seed(12)
    B, (h, w) = 3, hw
    qkv = _rand(B, 96, h, w, g=g).to(dev)
    wts = _rand(B, 32, 32, g=g, s=0.2).to(dev)
    bias, res = _rand(32, g=g, s=0.1).to(dev), _rand(B, 32, h, w, g=g).to(dev)
    v = qkv[:, 64:]
    want = torch.einsum("boc,bchw->bohw", wts.double(), v.double()) + bias.double().view(1, -1, 1, 1) + res.double()
    got = ops.pw(v, wts, bias, residual=res)
    assert (got.double() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_paconv_gate_and_layernorm2d(ops, dev):
    g = torch.Generator().manual_seed(10)
    x = _rand(2, 64, 9, 21, g=g)
    k2w, k2b = _rand(64, 64, 1, 1, g=g, s=0.2), _rand(64, g=g, s=0.1)
    k3 = _rand(2, 64, 9, 21, g=g)
    want = k3 * torch.sigmoid(F.conv2d(x, k2w, k2b))
    got = ops.paconv_gate(x.to(dev), k2w.to(dev), k2b.to(dev), k3.to(dev), inplace=True).cpu()
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)
    for c in (32, 64):
        xx = _rand(2, c, 7, 19, g=g, s=2.0)
        w_, b_ = 1 + _rand(c, g=g, s=0.1), _rand(c, g=g, s=0.1)
        torch.testing.assert_close(ops.layernorm2d(xx.to(dev), w_.to(dev), b_.to(dev), 1e-6).cpu(),
                                   om.layer_norm_2d(xx, w_, b_, 1e-6), rtol=2e-5, atol=2e-5)


# ------------------------------------------------------------------------------- error behaviour
def test_cpu_tensor_raises_no_fallback(ops):
    from wave_mamba_b200 import WaveMambaNativeError
    with pytest.raises(WaveMambaNativeError):
        ops.dwt_haar(torch.randn(1, 2, 4, 4))


def test_bad_arguments_are_rejected(ops, dev):
    with pytest.raises(ValueError):
        ops.dwt_haar(torch.randn(1, 2, 5, 4, device=dev))
    with pytest.raises(TypeError):
        ops.dwt_haar(torch.randn(1, 2, 4, 4, device=dev).double())
    from wave_mamba_b200 import _cabi
    lib = _cabi.load()
    x = torch.randn(1, 64, 8, 8, device=dev)
    rc = lib.wm_ss2d_core_fwd(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(),
                              x.data_ptr(), x.data_ptr(), None, 0, 1, 8, 8, None)
    assert rc == -1 and b"workspace" in lib.wm_last_error()


# ------------------------------------------------------------------------------- fused LFSS glue
def test_pw_dw_silu_without_bias(ops, dev):
    g = torch.Generator().manual_seed(11)
    x = _rand(2, 32, 13, 37, g=g)
    pw_w = _rand(64, 32, g=g, s=0.2)
    dw_w, dw_b = _rand(64, 1, 3, 3, g=g, s=0.3), _rand(64, g=g, s=0.1)
    ln_w, ln_b = 1 + _rand(32, g=g, s=0.1), _rand(32, g=g, s=0.1)
    t = om.layer_norm_2d(x, ln_w, ln_b, 1e-6)
    want = F.silu(F.conv2d(F.conv2d(t, pw_w[:, :, None, None]), dw_w, dw_b, padding=1, groups=64))
    got = ops.pw_dw(x.to(dev), pw_w.to(dev), None, dw_w.to(dev), dw_b.to(dev), ln_w.to(dev),
                    ln_b.to(dev), 1e-6, act="silu").cpu()
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)


def test_lfss_z_and_out(ops, dev):
    g = torch.Generator().manual_seed(12)
    B, h, w = 2, 9, 23
    x = _rand(B, 32, h, w, g=g)
    ln_w, ln_b = 1 + _rand(32, g=g, s=0.1), _rand(32, g=g, s=0.1)
    w_in = _rand(128, 32, g=g, s=0.2)
    t = om.layer_norm_2d(x, ln_w, ln_b, 1e-6)
    want_z = F.silu(F.conv2d(t, w_in[64:, :, None, None]))
    zs = ops.lfss_z(x.to(dev), ln_w.to(dev), ln_b.to(dev), 1e-6, w_in.to(dev))
    torch.testing.assert_close(zs.cpu(), want_z, rtol=2e-5, atol=2e-5)

    y, y2 = _rand(B, 64, h, w, g=g), _rand(B, 64, h, w, g=g)
    on_w, on_b = 1 + _rand(64, g=g, s=0.1), _rand(64, g=g, s=0.1)
    w_out = _rand(32, 64, g=g, s=0.2)
    skip = 1 + _rand(32, g=g, s=0.2)
    for second in (None, y2):
        ysum = y if second is None else y + second
        yn = F.layer_norm(ysum.permute(0, 2, 3, 1), (64,), on_w, on_b, 1e-5)
        want = x * skip.view(1, -1, 1, 1) + F.linear(yn * want_z.permute(0, 2, 3, 1), w_out).permute(0, 3, 1, 2)
        got = ops.lfss_out(y.to(dev), zs, on_w.to(dev), on_b.to(dev), 1e-5, w_out.to(dev), x.to(dev),
                           skip.to(dev), extra=() if second is None else (second.to(dev),)).cpu()
        torch.testing.assert_close(got, want, rtol=3e-5, atol=3e-5)


@pytest.mark.parametrize("hw", [(12, 20), (200, 160)])
def test_lfss_out_four_planes_tma(ops, dev, hw):
    """The production call: four direction planes summed in the order ((y0 + y2) + y1) + y3.  hw % 4 == 0
    takes the TMA pipeline (64-pixel tiles): 240 pixels leave a partial tile, 32000 x 2 images are 1000
    tiles = several per persistent CTA."""
    g = torch.Generator().manual_seed(112)
    B, (h, w) = 2, hw
    x = _rand(B, 32, h, w, g=g)
    zs = F.silu(_rand(B, 64, h, w, g=g))
    ys = [_rand(B, 64, h, w, g=g) for _ in range(4)]
    on_w, on_b = 1 + _rand(64, g=g, s=0.1), _rand(64, g=g, s=0.1)
    w_out = _rand(32, 64, g=g, s=0.2)
    skip = 1 + _rand(32, g=g, s=0.2)
    ysum = ((ys[0] + ys[1]) + ys[2]) + ys[3]
    yn = F.layer_norm(ysum.permute(0, 2, 3, 1), (64,), on_w, on_b, 1e-5)
    want = x * skip.view(1, -1, 1, 1) + F.linear(yn * zs.permute(0, 2, 3, 1), w_out).permute(0, 3, 1, 2)
    d = lambda v: v.to(dev)
    got = ops.lfss_out(d(ys[0]), d(zs), d(on_w), d(on_b), 1e-5, d(w_out), d(x), d(skip),
                       extra=(d(ys[1]), d(ys[2]), d(ys[3]))).cpu()
    torch.testing.assert_close(got, want, rtol=3e-5, atol=3e-5)


@pytest.mark.parametrize("hw", [(12, 20), (9, 23), (200, 160)])
def test_lfss_tail_gate_computed_in_kernel(ops, dev, hw):
    """wm_lfss_tail_fwd = lfss_z + lfss_out in one kernel (the production call of an LFSSBlock).  9 x 23
    (odd pixel count) runs as the two kernels through a scratch z; the others take the TMA pipeline:
    240 pixels leave a partial tile, 32000 x 2 images are 1000 tiles = several per persistent CTA."""
    g = torch.Generator().manual_seed(113)
    B, (h, w) = 2, hw
    x = _rand(B, 32, h, w, g=g)
    ln_w, ln_b = 1 + _rand(32, g=g, s=0.1), _rand(32, g=g, s=0.1)
    w_in = _rand(128, 32, g=g, s=0.2)
    ys = [_rand(B, 64, h, w, g=g) for _ in range(4)]
    on_w, on_b = 1 + _rand(64, g=g, s=0.1), _rand(64, g=g, s=0.1)
    w_out = _rand(32, 64, g=g, s=0.2)
    skip = 1 + _rand(32, g=g, s=0.2)
    zs = F.silu(F.conv2d(om.layer_norm_2d(x, ln_w, ln_b, 1e-6), w_in[64:, :, None, None]))
    ysum = ((ys[0] + ys[1]) + ys[2]) + ys[3]
    yn = F.layer_norm(ysum.permute(0, 2, 3, 1), (64,), on_w, on_b, 1e-5)
    want = x * skip.view(1, -1, 1, 1) + F.linear(yn * zs.permute(0, 2, 3, 1), w_out).permute(0, 3, 1, 2)
    d = lambda v: v.to(dev)
    got = ops.lfss_tail([d(t) for t in ys], d(x), d(ln_w), d(ln_b), 1e-6, d(w_in), d(on_w), d(on_b), 1e-5,
                        d(w_out), d(skip)).cpu()
    torch.testing.assert_close(got, want, rtol=3e-5, atol=3e-5)


@pytest.mark.parametrize("hw", [(11, 29), (12, 32), (10, 26), (100, 160)])
def test_pw_gate_with_scaled_residual(ops, dev, hw):
    """LFSSBlock.ffn tail.  11 x 29: register-staged kernel; the others (h*w % 4 == 0) the TMA pipeline:
    384 pixels = 3 full tiles, 260 pixels leave a partial tile, 16000 x 2 = 250 tiles (two per CTA)."""
    g = torch.Generator().manual_seed(13)
    h, w = hw
    x = _rand(2, 64, h, w, g=g)
    w_, b_ = _rand(32, 32, 1, 1, g=g, s=0.2), _rand(32, g=g, s=0.1)
    res, sc = _rand(2, 32, h, w, g=g), 1 + _rand(32, g=g, s=0.2)
    a, b2 = x.chunk(2, dim=1)
    want = res * sc.view(1, -1, 1, 1) + F.conv2d(F.gelu(a) * b2, w_, b_)
    got = ops.pw(x.to(dev), w_.to(dev), b_.to(dev), gate=True, residual=res.to(dev),
                 res_scale=sc.to(dev)).cpu()
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)


def test_lfss_block_fused_path_vs_golden_and_oracle(dev, params_cache):
    """LFSSBlock through the reference calling convention (B, L, C) -> fused NCHW kernels."""
    import wave_mamba_b200.arch as arch
    g = load_golden("lfss_block")
    p = om.sub(om.strip_prefix(params_cache(g["ckpt"])), g["block"])
    blk = arch.LFSSBlock(32)
    blk.load_state_dict(p, strict=True)
    blk = blk.to(dev).eval()
    h, w = int(g["h"]), int(g["w"])
    with torch.no_grad():
        got = blk(g["x"].to(dev), [h, w]).cpu()
    torch.testing.assert_close(got, g["y"], rtol=3e-5, atol=3e-5)


# ------------------------------------------------------------------------------- Gram / norms
@pytest.mark.parametrize("shape", [(1, 32, 8, 12), (2, 32, 33, 37), (1, 32, 135, 240), (2, 32, 64, 130)])
def test_gram32_vs_float64(ops, dev, shape):
    g = torch.Generator().manual_seed(21)
    B, C, h, w = shape
    wide = _rand(B, 96, h, w, g=g)                  # x is a channel slice (batch stride 96*h*w)
    y = _rand(B, 32, h, w, g=g)
    xd = wide.to(dev)[:, 32:64]
    G, nx, ny = ops.gram32(xd, y.to(dev))
    x64 = wide[:, 32:64].double().flatten(2, 3)
    y64 = y.double().flatten(2, 3)
    want = x64 @ y64.transpose(1, 2)
    scale = want.abs().max().item()
    # tolerance: fp32 products, 128-term fp32 tile sums, fp64 across tiles -> ~1e-6 relative
    assert (G.cpu().double() - want).abs().max().item() <= 2e-6 * scale + 1e-6
    torch.testing.assert_close(nx.cpu().double(), x64.pow(2).sum(-1), rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(ny.cpu().double(), y64.pow(2).sum(-1), rtol=2e-6, atol=1e-6)
    G2, _, _ = ops.gram32(xd, y.to(dev))
    assert torch.equal(G, G2)                       # deterministic


def test_gram32_argmin_matches_cdist_oracle(ops, dev):
    """The Matching decision (argmin over candidate channels) from the Gram pass equals the
    oracle's torch.cdist + topk on the same maps."""
    g = torch.Generator().manual_seed(22)
    x, p = _rand(2, 32, 40, 56, g=g), _rand(2, 32, 40, 56, g=g)
    p[:, 5] = x[:, 9] + 1e-3 * _rand(2, 40, 56, g=g)        # a near-duplicate pair
    G, nx, ny = ops.gram32(x.to(dev), p.to(dev))
    d2 = (nx[:, :, None] + ny[:, None, :] - 2 * G).cpu()
    want = torch.cdist(x.flatten(2, 3), p.flatten(2, 3)).topk(k=1, largest=False).indices.squeeze(-1)
    assert torch.equal(d2.argmin(-1), want)
    assert int(want[0, 9]) == 5


def test_match_index_and_attn_mixed_fold_the_32x32_tails(ops, dev):
    """wm_gram32_match_fwd / wm_gram32_attn_fwd = the Gram pass + the torch ops that followed it in an
    HFEBlock.  The argmin equals torch's on the same fp32 expression (and the cdist oracle's); the folded
    attention weights equal the torch formula to fp32 rounding."""
    g = torch.Generator().manual_seed(23)
    x, p = _rand(3, 32, 40, 56, g=g), _rand(3, 32, 40, 56, g=g)
    p[:, 5] = x[:, 9] + 1e-3 * _rand(3, 40, 56, g=g)        # a near-duplicate pair
    xd, pd = x.to(dev), p.to(dev)
    G, nx, ny = ops.gram32(xd, pd)
    idx = ops.match_index(xd, pd)
    assert idx.dtype == torch.int32 and idx.shape == (3, 32)
    assert torch.equal(idx.long(), (nx[:, :, None] + ny[:, None, :] - 2.0 * G).topk(k=1, largest=False).indices.squeeze(-1))
    want = torch.cdist(x.flatten(2, 3), p.flatten(2, 3)).topk(k=1, largest=False).indices.squeeze(-1)
    assert torch.equal(idx.long().cpu(), want) and int(want[0, 9]) == 5
    # attention tail on a channel slice (q, k are thirds of the qkv tensor in the model)
    qkv = _rand(3, 96, 24, 40, g=g).to(dev)
    q, k = qkv[:, :32], qkv[:, 32:64]
    temp = torch.tensor([[[1.7]]], device=dev)
    w_po = _rand(32, 32, 1, 1, g=g, s=0.2).to(dev)
    Gq, nq2, nk2 = ops.gram32(q, k)
    nq, nk = nq2.sqrt().clamp_min(1e-12), nk2.sqrt().clamp_min(1e-12)
    attn = (Gq / (nq[:, :, None] * nk[:, None, :]) * temp).softmax(dim=-1)
    want_mixed = (w_po.view(1, 32, 32, 1) * attn.unsqueeze(1)).sum(2)
    got = ops.attn_mixed(q, k, temp, w_po)
    torch.testing.assert_close(got, want_mixed, rtol=1e-5, atol=1e-6)
    assert torch.equal(got, ops.attn_mixed(q, k, temp, w_po))      # deterministic


# ------------------------------------------------------------------------------- dense 3x3 conv
@pytest.mark.parametrize("cin,cout", [(64, 32), (64, 64), (32, 96), (32, 32)])
@pytest.mark.parametrize("hw", [(13, 37), (8, 32), (40, 70)])
def test_conv3x3_plain(ops, dev, cin, cout, hw):
    """3xTF32 tensor-core conv vs float64 F.conv2d: fp32-level accuracy (not TF32-level)."""
    g = torch.Generator().manual_seed(31)
    h, w = hw
    x = _rand(2, cin, h, w, g=g)
    wt = _rand(cout, cin, 3, 3, g=g, s=0.1)
    bias = _rand(cout, g=g, s=0.1)
    want = F.conv2d(x.double(), wt.double(), bias.double(), padding=1)
    got = ops.conv3x3(x.to(dev), wt.to(dev), bias.to(dev)).cpu().double()
    err = (got - want).abs().max().item()
    tf32_like = (F.conv2d(x.bfloat16().float(), wt, bias, padding=1).double() - want).abs().max().item()
    print(f"{cin}->{cout} {hw}: max err {err:.2e} (scale {want.abs().max().item():.2f}; bf16-input err {tf32_like:.2e})")
    assert err <= 2e-5 * max(1.0, want.abs().max().item())


def test_conv3x3_two_inputs_with_channel_gather(ops, dev):
    g = torch.Generator().manual_seed(32)
    B, h, w = 2, 19, 45
    xa = _rand(B, 32, h, w, g=g)
    wide = _rand(B, 40, h, w, g=g)
    idx = torch.stack([torch.randperm(40, generator=g)[:32] for _ in range(B)]).to(torch.int32)
    wt = _rand(32, 64, 3, 3, g=g, s=0.1)
    gathered = torch.stack([wide[b, idx[b].long()] for b in range(B)])
    want = F.conv2d(torch.cat([xa, gathered], 1).double(), wt.double(), None, padding=1)
    got = ops.conv3x3(xa.to(dev), wt.to(dev), x_b=wide.to(dev), chan_map=idx.to(dev)).cpu().double()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
    # identity second input (the cat([LL, x_d]) of DownFRG)
    xb = _rand(B, 32, h, w, g=g)
    want = F.conv2d(torch.cat([xa, xb], 1).double(), wt.double(), None, padding=1)
    got = ops.conv3x3(xa.to(dev), wt.to(dev), x_b=xb.to(dev)).cpu().double()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_conv3x3_paconv_gate(ops, dev):
    """PAConv stage A: k3(x) * sigmoid(k2(x) + b) in one kernel (reference :694-697)."""
    g = torch.Generator().manual_seed(33)
    x = _rand(2, 64, 21, 50, g=g)
    k3 = _rand(64, 64, 3, 3, g=g, s=0.1)
    k2w, k2b = _rand(64, 64, 1, 1, g=g, s=0.2), _rand(64, g=g, s=0.1)
    want = F.conv2d(x.double(), k3.double(), None, padding=1) * torch.sigmoid(
        F.conv2d(x.double(), k2w.double(), k2b.double()))
    got = ops.conv3x3(x.to(dev), k3.to(dev), gate_w=k2w.to(dev), gate_b=k2b.to(dev)).cpu().double()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("cin,cout,gate", [(64, 64, True), (64, 32, False), (64, 64, False),
                                           (32, 96, False), (32, 32, False)])
def test_conv3x3_many_tiles_per_cta(ops, dev, cin, cout, gate):
    """A map large enough that every persistent CTA walks several tiles (638 tiles on 148 SMs): the
    weight ring, the two X slabs and the TMEM accumulator buffers all wrap their mbarrier phases more
    than once.  Reference: fp64 convolution on the GPU; ragged right/bottom tiles included."""
    g = torch.Generator().manual_seed(36)
    B, h, w = 2, 200, 330
    x = _rand(B, cin, h, w, g=g).to(dev)
    wt = _rand(cout, cin, 3, 3, g=g, s=0.1).to(dev)
    if gate:
        k2w, k2b = _rand(cout, cin, 1, 1, g=g, s=0.2).to(dev), _rand(cout, g=g, s=0.1).to(dev)
        want = F.conv2d(x.double(), wt.double(), None, padding=1) * torch.sigmoid(
            F.conv2d(x.double(), k2w.double(), k2b.double()))
        got = ops.conv3x3(x, wt, gate_w=k2w, gate_b=k2b)
    else:
        bias = _rand(cout, g=g, s=0.1).to(dev)
        want = F.conv2d(x.double(), wt.double(), bias.double(), padding=1)
        got = ops.conv3x3(x, wt, bias)
    err = (got.double() - want).abs().max().item()
    assert err <= 2e-5 * max(1.0, want.abs().max().item()), err
    assert torch.equal(got, ops.conv3x3(x, wt, gate_w=k2w, gate_b=k2b) if gate else ops.conv3x3(x, wt, bias))


def test_conv_pack_cache_sees_data_writes(ops, dev):
    """The pre-pack cache must not serve stale weights after a write through ``.data`` (which does
    not bump ``_version``) once ``clear_pack_cache`` is called, nor after an in-place ``copy_``."""
    g = torch.Generator().manual_seed(37)
    x = _rand(1, 32, 16, 40, g=g).to(dev)
    w = torch.nn.Parameter(_rand(32, 32, 3, 3, g=g, s=0.1).to(dev))
    w2 = _rand(32, 32, 3, 3, g=g, s=0.1).to(dev)
    y1 = ops.conv3x3(x, w)
    with torch.no_grad():
        w.copy_(w2)                                   # bumps _version
    y2 = ops.conv3x3(x, w)
    torch.testing.assert_close(y2, F.conv2d(x, w2, None, padding=1), rtol=2e-5, atol=2e-5)
    w.data.mul_(2.0)                                  # behind autograd's back
    ops.clear_pack_cache()
    y3 = ops.conv3x3(x, w)
    torch.testing.assert_close(y3, 2.0 * y2, rtol=2e-5, atol=2e-5)
    assert not torch.equal(y1, y2)


@pytest.mark.parametrize("hw", [(21, 45), (24, 136), (9, 70), (8, 64), (200, 1280)])
def test_stem_and_head_conv(ops, dev, hw):
    """Odd widths take the scalar stores; w % 4 == 0 takes the stem's 16-byte rows and the head's TMA
    pipeline (tiles of 32 x 64 starting one column left of a multiple of 64); 136 and 70 leave partial
    tiles; 200 x 1280 x 2 images = 294 tiles, two per persistent CTA."""
    g = torch.Generator().manual_seed(41)
    h, w = hw
    x = torch.rand(2, 3, h, w, generator=g)
    w1, b1 = _rand(32, 3, 3, 3, g=g, s=0.2), _rand(32, g=g, s=0.1)
    want = F.conv2d(x, w1, b1, padding=1)
    torch.testing.assert_close(ops.stem_conv3x3(x.to(dev), w1.to(dev), b1.to(dev)).cpu(), want,
                               rtol=2e-5, atol=2e-5)
    f = _rand(2, 32, h, w, g=g)
    w2, b2 = _rand(3, 32, 3, 3, g=g, s=0.1), _rand(3, g=g, s=0.1)
    want = F.conv2d(f, w2, b2, padding=1) + x
    got = ops.head_conv3x3(f.to(dev), w2.to(dev), b2.to(dev), residual=x.to(dev)).cpu()
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("shape", [(2, 32, 48, 64), (1, 32, 20, 24), (1, 32, 270, 480)])
def test_dwt_with_skff_pool_in_its_epilogue(ops, dev, shape):
    """wm_dwt_haar_pool_fwd: the bands are bit-exact with the plain DWT, and SKFF through the pooled buffer
    (one streaming pass) equals SKFF through its own pool pass to fp32 rounding of the mean.  20 x 24
    (W % 16 != 0) takes the plain DWT + the separate pool kernel inside the same entry point."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g).to(dev)
    plain = ops.dwt_haar(x)
    ll, hl, lh, hh, pool = ops.dwt_haar_pool(x)
    for a, b in zip(plain, (ll, hl, lh, hh)):
        assert torch.equal(a, b)
    w_du, pr = _rand(4, 32, 1, 1, g=g, s=0.3).to(dev), torch.tensor([0.25], device=dev)
    fcs = [_rand(32, 4, 1, 1, g=g, s=0.5).to(dev) for _ in range(3)]
    want = ops.skff(hl, lh, hh, w_du, pr, *fcs)
    got = ops.skff(hl, lh, hh, w_du, pr, *fcs, pool=pool)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------- SKFF / ps_down
def test_skff_golden(ops, dev):
    """SKFF (ref:939-959) against the reference's own output; 2e-6 abs (pool in fp64 partials)."""
    from conftest import load_params
    g = load_golden("skff")
    sd = om.strip_prefix(load_params(str(g["ckpt"])))
    pre = str(g["block"]) + "."
    p = {k[len(pre):]: v.to(dev) for k, v in sd.items() if k.startswith(pre)}
    got = ops.skff(g["a"].to(dev), g["b"].to(dev), g["c"].to(dev), p["conv_du.0.weight"],
                   p["conv_du.1.weight"], p["fcs.0.weight"], p["fcs.1.weight"], p["fcs.2.weight"])
    torch.testing.assert_close(got.cpu(), g["y"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("shape", [(1, 32, 8, 8), (2, 32, 25, 37), (1, 32, 135, 240), (3, 32, 1, 1)])
def test_skff_vs_oracle(ops, dev, shape):
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(*shape, generator=g) for _ in range(3)]
    p = {"conv_du.0.weight": 0.3 * torch.randn(4, 32, 1, 1, generator=g),
         "conv_du.1.weight": torch.tensor([0.25]),
         "fcs.0.weight": torch.randn(32, 4, 1, 1, generator=g),
         "fcs.1.weight": torch.randn(32, 4, 1, 1, generator=g),
         "fcs.2.weight": torch.randn(32, 4, 1, 1, generator=g)}
    want = om.skff(p, feats)
    got = ops.skff(*[f.to(dev) for f in feats], p["conv_du.0.weight"].to(dev), p["conv_du.1.weight"].to(dev),
                   p["fcs.0.weight"].to(dev), p["fcs.1.weight"].to(dev), p["fcs.2.weight"].to(dev))
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("r", [2, 4, 8])
@pytest.mark.parametrize("shape", [(1, 3, 16, 24), (2, 3, 64, 40), (1, 3, 200, 304)])
def test_ps_down_vs_oracle(ops, dev, r, shape):
    """PixelUnshuffle(r) + 1x1 conv (ref:1014-1025): 2e-5 abs (different summation order)."""
    g = torch.Generator().manual_seed(r)
    x = torch.rand(*shape, generator=g)
    w = torch.randn(32, 3 * r * r, 1, 1, generator=g) * 0.2
    b = torch.randn(32, generator=g)
    want = F.conv2d(F.pixel_unshuffle(x, r), w, b)
    got = ops.ps_down(x.to(dev), w.to(dev), b.to(dev), r)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=2e-5)


# ------------------------------------------------------------------------------- image I/O edges
@pytest.mark.parametrize("shape,window", [((1, 16, 24), 8), ((2, 50, 37), 8), ((1, 200, 300), 128),
                                          ((1, 135, 241), 128), ((1, 256, 256), 128)])
def test_img_u8_to_f32_bit_exact(ops, dev, shape, window):
    """img2tensor + /255. + reflect pad (inference_wavemamba.py:28-36,101-106): bit-exact."""
    g = torch.Generator().manual_seed(3)
    B, H, W = shape
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    want = om.img_u8_to_f32(img, window)
    got = ops.img_u8_to_f32(img.to(dev), window)
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)
    # the same reference line evaluated on a CUDA tensor (torch multiplies by the reciprocal there)
    x = img.to(dev).flip(-1).permute(0, 3, 1, 2).float() / 255.
    want_cuda = F.pad(x, (0, want.shape[3] - W, 0, want.shape[2] - H), "reflect")
    assert torch.equal(ops.img_u8_to_f32(img.to(dev), window, cuda_division=True), want_cuda)


@pytest.mark.parametrize("shape,crop", [((1, 16, 24), (16, 24)), ((2, 56, 40), (50, 37)),
                                        ((1, 256, 384), (200, 300)), ((1, 128, 256), (128, 256))])
def test_img_f32_to_u8_bit_exact(ops, dev, shape, crop):
    """crop + tensor2img (img_util.py:36-98): clamp, *255, round half to even, BGR: bit-exact."""
    g = torch.Generator().manual_seed(4)
    B, Hs, Ws = shape
    x = torch.rand(B, 3, Hs, Ws, generator=g) * 1.4 - 0.2          # exercises both clamps
    x.view(-1)[:256] = (torch.arange(256, dtype=torch.float32) + 0.5) / 255.0   # exact .5 ties
    want = om.img_f32_to_u8(x, *crop)
    got = ops.img_f32_to_u8(x.to(dev), *crop)
    assert torch.equal(got.cpu(), want)


def test_enhance_bgr_u8_matches_float_path(dev):
    """The uint8 entry point = the reference loop body: same uint8 image as converting on the host."""
    import wave_mamba_b200 as wm
    from conftest import load_params
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    net.load_state_dict(load_params("LOLv1"), strict=True)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(9)
    img = (torch.rand(100, 150, 3, generator=g) * 60).to(torch.uint8)
    got = wm.enhance_bgr_u8(net, img, window=128).cpu()
    with torch.no_grad():
        y = net.restoration_network(om.img_u8_to_f32(img[None], 128).to(dev)).cpu()
    want = om.img_f32_to_u8(y, 100, 150)[0]
    assert got.shape == (100, 150, 3) and torch.equal(got, want)


def test_img_io_edges_golden(ops, dev):
    """The device conversions against the reference's own img2tensor / check_image_size / tensor2img."""
    g = load_golden("imgio")
    assert torch.equal(ops.img_u8_to_f32(g["img"][None].to(dev), 128).cpu(), g["x"])
    assert torch.equal(ops.img_f32_to_u8(g["y"].to(dev), 100, 150).cpu()[0], g["out_img"])


@pytest.mark.parametrize("shape", [(2, 21, 50), (1, 7, 32), (1, 40, 96)])
def test_conv3x3_channel_quad_layout(ops, dev, shape):
    """PAConv k3 (+k2 gate, gathered second input) -> k4 through the (B, C/4, h, w, 4) intermediate:
    identical to the NCHW path bit for bit (same arithmetic, different addressing) and within 2e-5
    of the fp64 reference (reference :694-698)."""
    g = torch.Generator().manual_seed(35)
    B, h, w = shape
    x, per = _rand(B, 32, h, w, g=g), _rand(B, 32, h, w, g=g)
    idx = torch.stack([torch.randperm(32, generator=g) for _ in range(B)]).to(torch.int32)
    k3, k4 = _rand(64, 64, 3, 3, g=g, s=0.1), _rand(32, 64, 3, 3, g=g, s=0.1)
    k2w, k2b = _rand(64, 64, 1, 1, g=g, s=0.2), _rand(64, g=g, s=0.1)
    a = dict(x_b=per.to(dev), chan_map=idx.to(dev), gate_w=k2w.to(dev), gate_b=k2b.to(dev))
    t_nchw = ops.conv3x3(x.to(dev), k3.to(dev), **a)
    t_c4 = ops.conv3x3(x.to(dev), k3.to(dev), out_c4=True, **a)
    assert t_c4.shape == (B, 16, h, w, 4)
    assert torch.equal(t_c4.permute(0, 1, 4, 2, 3).reshape(B, 64, h, w), t_nchw)
    y_nchw = ops.conv3x3(t_nchw, k4.to(dev))
    y_c4 = ops.conv3x3(t_c4, k4.to(dev), in_c4=True)
    assert torch.equal(y_c4, y_nchw)
    cat = torch.cat([x, torch.stack([per[b][idx[b].long()] for b in range(B)])], dim=1).double()
    t = F.conv2d(cat, k3.double(), None, padding=1) * torch.sigmoid(F.conv2d(cat, k2w.double(), k2b.double()))
    want = F.conv2d(t, k4.double(), None, padding=1)
    assert (y_c4.cpu().double() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


# ------------------------------------------------------------------------------- DWT / IWT backward
@pytest.mark.parametrize("shape", [(2, 5, 6, 10), (1, 32, 40, 64)])
def test_dwt_iwt_backward_match_autograd_of_the_oracle(dev, shape):
    """The Haar pair is orthonormal: DWT.backward = IWT and IWT.backward = DWT (same kernels).
    Compared with torch autograd through the CPU oracle: 1e-6 abs on O(1) gradients."""
    from wave_mamba_b200.arch import DWT, IWT
    g = torch.Generator().manual_seed(11)
    B, C, H, W = shape
    x = torch.randn(B, C, H, W, generator=g)
    wts = [torch.randn(B, C, H // 2, W // 2, generator=g) for _ in range(4)]
    xr = x.clone().requires_grad_(True)
    sum((b * w).sum() for b, w in zip(om.haar_dwt(xr), wts)).backward()
    xg = x.to(dev).requires_grad_(True)
    sum((b * w.to(dev)).sum() for b, w in zip(DWT()(xg), wts)).backward()
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=0, atol=1e-6)

    low, high = torch.randn(B, C, H // 2, W // 2, generator=g), torch.randn(B, 3 * C, H // 2, W // 2, generator=g)
    wy = torch.randn(B, C, H, W, generator=g)
    lr, hr = low.clone().requires_grad_(True), high.clone().requires_grad_(True)
    (om.haar_iwt(torch.cat([lr, hr], dim=1)) * wy).sum().backward()
    lg, hg = low.to(dev).requires_grad_(True), high.to(dev).requires_grad_(True)
    (IWT()(lg, hg) * wy.to(dev)).sum().backward()
    torch.testing.assert_close(lg.grad.cpu(), lr.grad, rtol=0, atol=1e-6)
    torch.testing.assert_close(hg.grad.cpu(), hr.grad, rtol=0, atol=1e-6)
