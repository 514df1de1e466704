"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/*.h declares, the host-side mirror has the reference's constructor / state dict, and
the product refuses to run without a GPU (no fallback).  No compute calls here."""
import ctypes
import glob
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_params


def _declared_symbols():
    names = set()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(hdr).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(wm_[a-z0-9_]+)\s*\(", text))
    return names


@pytest.fixture(scope="module")
def lib_path():
    from wave_mamba_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_binding_table_matches_header(lib_path):
    from wave_mamba_b200 import _cabi
    assert set(_cabi.SIGNATURES) == _declared_symbols()
    lib = _cabi.load()
    assert lib.wm_abi_version() == _cabi.ABI_VERSION


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure(lib_path):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wave_mamba_b200 import _cabi, ops, WaveMambaNativeError
    lib = _cabi.load()
    assert lib.wm_device_check() == -3
    assert b"no CPU fallback" in lib.wm_last_error()
    with pytest.raises(WaveMambaNativeError):
        ops.dwt_haar(torch.randn(1, 1, 4, 4))


def test_size_queries_need_no_gpu(lib_path):
    from wave_mamba_b200 import _cabi
    lib = _cabi.load()
    assert lib.wm_ss2d_core_workspace_bytes(0, 8, 8) == 0
    n = lib.wm_ss2d_core_workspace_bytes(1, 1080, 1920)
    plane = 64 * 1080 * 1920 * 4
    # four direction planes + the chunk aggregates + the replay tiles pass 1 hands to pass 2
    # (43 KB per 64 positions and direction = 10.5 planes, plus per-CTA slot padding)
    assert 14.5 * plane < n < 17 * plane
    assert lib.wm_ss2d_core_workspace_bytes(2, 135, 240) > 2 * 4 * 64 * 135 * 240 * 4


@pytest.mark.parametrize("ckpt", ["LOLv1", "UHDLL", "UHDLOL4K"])
def test_shipped_checkpoints_load_strict(ckpt):
    import wave_mamba_b200 as wm
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0,
                       some_ignored_yaml_key=1)
    sd = load_params(ckpt)
    assert len(sd) == 591
    res = net.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sum(p.numel() for p in net.parameters()) == 1512718
    assert all(k.startswith("restoration_network.") for k in net.state_dict())
    ss = net.restoration_network.down_group1.l_blk[0].self_attention
    assert getattr(ss.A_logs, "_no_weight_decay", False) and getattr(ss.Ds, "_no_weight_decay", False)


def test_cpu_forward_raises_instead_of_falling_back():
    import wave_mamba_b200 as wm
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 1, 1], n_h_blocks=[1, 1, 1]).eval()
    with pytest.raises(wm.WaveMambaNativeError):
        net(torch.rand(1, 3, 16, 16))


def test_unsupported_width_is_rejected():
    import wave_mamba_b200 as wm
    with pytest.raises(NotImplementedError):
        wm.WaveMamba(in_chn=3, wf=48)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wave_mamba_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True) + \
            glob.glob(os.path.join(ROOT, "plugin", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
        assert "/root/reference" not in src, path


@pytest.mark.skipif(not os.path.isdir("/root/reference/basicsr"), reason="reference tree not mounted")
def test_plugin_overlays_the_reference_registry(tmp_path):
    """Overlay plugin/basicsr/archs/wavemamba_arch.py on a (symlinked) reference checkout and
    build the network through the reference's own build_network / ARCH_REGISTRY."""
    sys.path.insert(0, ROOT)
    from tools.ref_overlay import make_overlay
    over = tmp_path / "overlay"
    make_overlay("/root/reference", str(over))
    code = f"""
import sys
sys.path.insert(0, {str(ROOT)!r})
from tools import ref_shims
ref_shims.install({str(over)!r})
import torch
from basicsr.archs import build_network
from basicsr.utils.registry import ARCH_REGISTRY
opt = dict(type='WaveMamba', in_chn=3, wf=32, n_l_blocks=[1,2,4], n_h_blocks=[1,1,2], ffn_scale=2.0)
net = build_network(opt)
assert type(net).__mro__[1].__module__ == 'wave_mamba_b200.arch', type(net).__mro__
from basicsr.archs.wavemamba_arch import WaveMamba
assert ARCH_REGISTRY.get('WaveMamba') is WaveMamba
sd = torch.load({os.path.join(ROOT, 'ckpt', 'WaveMamba_LOLv1.pth')!r}, map_location='cpu')['params']
print(net.load_state_dict(sd, strict=True))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "All keys matched" in out.stdout


def test_ss2d_chunk_plan_is_consistent(lib_path):
    """The SS2D chunk planner (host logic, no GPU): chunk lengths are whole tiles, the chunks cover
    every position of both scan orientations, and the workspace holds the planes + aggregates."""
    from wave_mamba_b200 import _cabi
    lib = _cabi.load()
    for B, h, w in [(1, 1080, 1920), (1, 540, 960), (1, 270, 480), (4, 200, 300), (4, 50, 75),
                    (8, 256, 256), (2, 135, 240), (1, 9, 13), (1, 1, 1), (3, 17, 64)]:
        out = (ctypes.c_int * 6)()
        assert lib.wm_ss2d_debug_geometry(B, h, w, out) == 0
        row_T, row_ctas, col_seg, ncolseg, col_ctas, cols_first = list(out)
        L = h * w
        assert row_T % 16 == 0 and row_T >= 64
        assert col_seg % 16 == 0 and col_seg * ncolseg >= h and col_seg * (ncolseg - 1) < h
        row_chunks = -(-L // row_T)
        assert row_ctas == -(-row_chunks // 8)          # the forward walks 8 strands per CTA
        assert col_ctas == -(-w // 8) * ncolseg
        assert cols_first in (0, 1)
        max_chunks = max(row_chunks, w * ncolseg)
        need = 4 * B * 64 * L * 4 + 2 * B * 4 * max_chunks * 1024 * 4
        assert lib.wm_ss2d_core_workspace_bytes(B, h, w) >= need
    assert lib.wm_ss2d_debug_geometry(0, 8, 8, (ctypes.c_int * 6)()) != 0


def test_uint8_entry_point_validates_and_refuses_cpu(lib_path):
    """wave_mamba_b200.enhance_bgr_u8: shape/dtype errors are Python errors, a CPU module is refused
    loudly (no fallback)."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import wave_mamba_b200 as wm
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0).eval()
    with pytest.raises(ValueError):
        wm.enhance_bgr_u8(net, torch.zeros(16, 16, 3))                      # not uint8
    with pytest.raises(ValueError):
        wm.enhance_bgr_u8(net, torch.zeros(16, 16, 4, dtype=torch.uint8))   # not 3 channels
    with pytest.raises((wm.WaveMambaNativeError, RuntimeError)):
        wm.enhance_bgr_u8(net, torch.zeros(16, 16, 3, dtype=torch.uint8), device=torch.device("cpu"))


def test_pipelines_and_metrics_refuse_cpu(lib_path):
    """EnhancePipeline, ShardedEnhancePipeline and the device metrics fail loudly without a CUDA device /
    process group -- none of them has a host path; the metric size query needs no GPU."""
    from wave_mamba_b200 import _cabi, ops, parallel
    import wave_mamba_b200 as wm
    lib = _cabi.load()
    tiles = ((2160 - 2 + 31) // 32) * ((3840 - 2 + 31) // 32)
    assert lib.wm_psnr_ssim_y_workspace_bytes(1, 2160, 3840, 1) == tiles * 2 * 8
    assert lib.wm_psnr_ssim_y_workspace_bytes(3, 2, 2, 1) == 0          # nothing left after the crop
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0).eval()
    with pytest.raises(wm.WaveMambaNativeError):
        wm.EnhancePipeline(net)
    with pytest.raises(RuntimeError):
        parallel.ShardedEnhancePipeline(net, torch.device("cpu"))
    img = torch.zeros(1, 8, 8, 3, dtype=torch.uint8)
    with pytest.raises(wm.WaveMambaNativeError):
        ops.psnr_ssim_y(img, img)
