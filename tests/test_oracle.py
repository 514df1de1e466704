"""CPU tests: pin the oracle against the golden vectors produced by the UNMODIFIED
reference arch file (tools/make_golden.py), and cross-check the scan implementations."""
import torch
import pytest

from conftest import load_golden
from oracle import model as om
from oracle import scan as oscan


def _sub(params, block):
    return om.sub(om.strip_prefix(params), block)


def test_dwt_matches_reference_bit_exact():
    g = load_golden("dwt")
    ll, hl, lh, hh = om.haar_dwt(g["x"])
    for got, key in ((ll, "ll"), (hl, "hl"), (lh, "lh"), (hh, "hh")):
        assert torch.equal(got, g[key]), key


def test_iwt_matches_reference_bit_exact():
    g = load_golden("iwt")
    assert torch.equal(om.haar_iwt(g["x"]), g["y"])


def test_dwt_iwt_round_trip():
    x = torch.randn(1, 3, 8, 12)
    ll, hl, lh, hh = om.haar_dwt(x)
    back = om.haar_iwt(torch.cat([ll, hl, lh, hh], dim=1))
    torch.testing.assert_close(back, x, rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_ss2d_core_matches_reference(params_cache, tag):
    g = load_golden(f"ss2d_core_{tag}")
    p = _sub(params_cache(g["ckpt"]), g["block"] + ".self_attention")
    for scan_fn in (oscan.selective_scan_c, oscan.selective_scan_loop):
        y = om.ss2d_core(g["x"], p["x_proj_weight"], p["dt_projs_weight"], p["dt_projs_bias"],
                         p["A_logs"], p["Ds"], scan_fn)
        # fp32 tolerance: the reference einsum (bmm) and ours differ in summation order
        torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=2e-6)


def test_ss2d_full_and_lfss_block_match_reference(params_cache):
    g = load_golden("ss2d_full")
    p = _sub(params_cache(g["ckpt"]), g["block"])
    y = om.ss2d(om.sub(p, "self_attention"), g["x"])
    torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=2e-6)
    g = load_golden("lfss_block")
    x = g["x"].reshape(2, int(g["h"]), int(g["w"]), 32)
    y = om.lfss_block(p, x).reshape(2, -1, 32)
    torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=2e-6)


def test_hfe_block_matches_reference_including_argmin(params_cache):
    g = load_golden("hfe_block")
    p = _sub(params_cache(g["ckpt"]), g["block"])
    trace = {}
    y = om.hfe_block(p, g["x"], g["perception"], trace, "t.")
    assert torch.equal(trace["match_idx"]["t.attn"], g["idx_attn"])
    assert torch.equal(trace["match_idx"]["t.ffn"], g["idx_ffn"])
    torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=2e-6)


def test_skff_matches_reference(params_cache):
    g = load_golden("skff")
    p = _sub(params_cache(g["ckpt"]), g["block"])
    torch.testing.assert_close(om.skff(p, [g["a"], g["b"], g["c"]]), g["y"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["e2e_rand64_LOLv1", "e2e_synth48x80_UHDLL", "e2e_synth64_UHDLOL4K"])
def test_end_to_end_matches_reference(params_cache, name):
    g = load_golden(name)
    y = om.unet_forward(params_cache(g["ckpt"]), g["x"])
    err = (y - g["y"]).abs().max().item()
    assert err < 2e-5, err
    if "gt" in g:
        a, b = om.to_uint8_bgr(y[:1]), om.to_uint8_bgr(g["y"][:1])
        gt = om.to_uint8_bgr(g["gt"][:1])
        assert abs(om.psnr_y(a, gt) - om.psnr_y(b, gt)) <= 1e-3


def test_kat_stats_from_survey(params_cache):
    """SURVEY.md section 4 known-answer stats for rand(1,3,64,64), LOLv1 weights."""
    torch.manual_seed(0)
    x = torch.rand(1, 3, 64, 64)
    y = om.unet_forward(params_cache("LOLv1"), x)
    assert abs(y.mean().item() - 0.7356967) < 2e-6
    assert abs(y.std().item() - 0.2477267) < 2e-6


def test_synth_generator_is_the_one_the_goldens_used():
    g = load_golden("e2e_synth64_UHDLOL4K")
    x, gt = om.synth_lowlight(1, 64, 64, seed=0)
    assert torch.equal(x, g["x"]) and torch.equal(gt, g["gt"])


def test_scan_c_vs_loop_vs_float64():
    torch.manual_seed(3)
    B, G, D, N, L = 2, 4, 8, 16, 300
    u = torch.nn.functional.silu(torch.randn(B, G * D, L))
    delta = 0.5 * torch.randn(B, G * D, L) - 2
    A = -torch.exp(torch.randn(G * D, N) * 1.5)
    Bm, Cm = torch.randn(B, G, N, L), torch.randn(B, G, N, L)
    Dv, bias = torch.rand(G * D) + 0.1, torch.randn(G * D)
    y_c = oscan.selective_scan_c(u, delta, A, Bm, Cm, Dv, bias)
    y_l = oscan.selective_scan_loop(u, delta, A, Bm, Cm, Dv, bias)
    y_64 = oscan.selective_scan_c(u, delta, A, Bm, Cm, Dv, bias, arbiter64=True)
    y_d = oscan.selective_scan_c(*[t.double() for t in (u, delta, A, Bm, Cm, Dv, bias)])
    torch.testing.assert_close(y_64, y_d, rtol=1e-12, atol=1e-12)
    assert (y_c.double() - y_64).abs().max() < 2e-5
    assert (y_l.double() - y_64).abs().max() < 2e-5
    # softplus threshold branch (x > 20 -> x)
    big = torch.full((1, 4, 3), 25.0)
    y = oscan.selective_scan_c(torch.ones(1, 4, 3), big, -torch.ones(4, 2), torch.ones(1, 1, 2, 3),
                               torch.ones(1, 1, 2, 3))
    y2 = oscan.selective_scan_loop(torch.ones(1, 4, 3), big, -torch.ones(4, 2),
                                   torch.ones(1, 1, 2, 3), torch.ones(1, 1, 2, 3))
    torch.testing.assert_close(y, y2)


def test_float64_arbiter_model_agrees_with_fp32(params_cache):
    g = load_golden("e2e_synth64_UHDLOL4K")
    p64 = om.cast_params(params_cache(g["ckpt"]), torch.float64)
    y64 = om.unet_forward(p64, g["x"].double())
    assert (y64 - g["y"].double()).abs().max() < 5e-5


def test_img_io_edges_match_reference():
    """oracle img_u8_to_f32 / img_f32_to_u8 vs the reference's img2tensor + /255. + check_image_size
    and crop + tensor2img (golden from tools/make_golden_imgio.py): bit-exact."""
    g = load_golden("imgio")
    assert torch.equal(om.img_u8_to_f32(g["img"][None], 128), g["x"])
    assert torch.equal(om.img_f32_to_u8(g["y"], 100, 150)[0], g["out_img"])


def test_ss2d_core_backward_oracle_gradcheck(params_cache):
    """The backward oracle for the next round: autograd through the pure-torch
    sequential scan (oracle.scan.selective_scan_loop, fp64) is what a CUDA SS2D backward will be
    compared against.  Pinned here with central finite differences on a tiny map: gradients with
    respect to the input map, the decay parameters, the skip weights and the dt bias."""
    from oracle import scan as oscan
    p = om.sub(om.strip_prefix(params_cache("UHDLOL4K")), "down_group1.l_blk.0.self_attention")
    prm = [p[k].double() for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")]
    g = torch.Generator().manual_seed(13)
    x = torch.nn.functional.silu(0.5 * torch.randn(1, 64, 2, 3, generator=g)).double().requires_grad_(True)
    xp, dw = prm[0], prm[1]
    db = prm[2].clone().requires_grad_(True)
    al = prm[3].clone().requires_grad_(True)
    ds = prm[4].clone().requires_grad_(True)

    def fn(x_, db_, al_, ds_):
        return om.ss2d_core(x_, xp, dw, db_, al_, ds_, scan_fn=oscan.selective_scan_loop)

    assert torch.autograd.gradcheck(fn, (x, db, al, ds), eps=1e-6, atol=1e-6, rtol=1e-4, nondet_tol=0.0,
                                    fast_mode=True)
