"""ctypes binding of include/wavemamba_b200.h -- the stub a reference maintainer would add.

The library is built in-tree by wave_mamba_b200/build.py.  There is NO fallback: if the
shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwavemamba_b200.so")

ABI_VERSION = 15

# name -> (restype, argtypes); mirrors include/wavemamba_b200.h one to one
SIGNATURES = {
    "wm_abi_version": (c_int, []),
    "wm_last_error": (c_char_p, []),
    "wm_device_check": (c_int, []),
    "wm_dwt_haar_fwd": (c_int, [c_void_p] * 5 + [c_int64] * 3 + [c_void_p]),
    "wm_dwt_haar_pool_fwd": (c_int, [c_void_p] * 6 + [c_size_t] + [c_int64] * 3 + [c_void_p]),
    "wm_iwt_haar_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p] + [c_int64] * 4 + [c_void_p]),
    "wm_ss2d_core_workspace_bytes": (c_size_t, [c_int64] * 3),
    "wm_ss2d_debug_timing": (c_int, [c_void_p]),
    "wm_ss2d_debug_geometry": (c_int, [c_int64] * 3 + [c_void_p]),
    "wm_ss2d_dirs_fwd": (c_int, [c_void_p] * 7 + [c_size_t] + [c_int64] * 3 + [c_void_p]),
    "wm_ss2d_core_fwd": (c_int, [c_void_p] * 8 + [c_size_t] + [c_int64] * 3 + [c_void_p]),
    "wm_ss2d_core_bwd_workspace_bytes": (c_size_t, [c_int64] * 3),
    "wm_ss2d_core_bwd": (c_int, [c_void_p] * 14 + [c_size_t] + [c_int64] * 3 + [c_void_p]),
    "wm_layernorm2d_fwd": (c_int, [c_void_p] * 3 + [c_float, c_void_p] + [c_int64] * 4 + [c_void_p]),
    "wm_pw_dw_fwd": (c_int, [c_void_p] * 3 + [c_float] + [c_void_p] * 4 + [c_int, c_void_p] + [c_int64] * 5 + [c_void_p]),
    "wm_dw_act_pw_fwd": (c_int, [c_void_p] * 5 + [c_int] + [c_void_p] * 2 + [c_int64] * 4 + [c_void_p]),
    "wm_pw_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int] + [c_void_p] * 3 + [c_int64] * 5 + [c_void_p]),
    "wm_lfss_z_fwd": (c_int, [c_void_p] * 3 + [c_float] + [c_void_p] * 2 + [c_int64] * 3 + [c_void_p]),
    "wm_lfss_out_fwd": (c_int, [c_void_p] * 7 + [c_float] + [c_void_p] * 4 + [c_int64] * 3 + [c_void_p]),
    "wm_lfss_tail_fwd": (c_int, [c_void_p] * 7 + [c_float] + [c_void_p] * 3 + [c_float] + [c_void_p] * 4 +
                         [c_int64] * 3 + [c_void_p]),
    "wm_gram32_workspace_bytes": (c_size_t, [c_int64] * 2),
    "wm_gram32_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t,
                              c_int64, c_int64, c_void_p]),
    "wm_gram32_match_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_int64, c_int64, c_void_p]),
    "wm_gram32_attn_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_int64, c_int64, c_void_p]),
    "wm_conv3x3_packed_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "wm_conv3x3_debug_timing": (c_int, [c_void_p]),
    "wm_debug_pipeline_error": (c_int, [c_void_p]),
    "wm_pw_dw_debug_timing": (c_int, [c_void_p]),
    "wm_conv3x3_prepack": (c_int, [c_void_p] * 3 + [c_int64] * 2 + [c_void_p]),
    "wm_conv3x3_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64] + [c_void_p] * 5 +
                       [c_int64] * 5 + [c_void_p]),
    "wm_conv3x3_ex_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64] + [c_void_p] * 5 +
                          [c_int64] * 5 + [c_int, c_int, c_void_p]),
    "wm_stem_conv3x3_fwd": (c_int, [c_void_p] * 4 + [c_int64] * 3 + [c_void_p]),
    "wm_head_conv3x3_fwd": (c_int, [c_void_p] * 5 + [c_int64] * 3 + [c_void_p]),
    "wm_paconv_gate_fwd": (c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_void_p]),
    "wm_skff_workspace_bytes": (c_size_t, [c_int64] * 3),
    "wm_skff_fwd": (c_int, [c_void_p] * 10 + [c_size_t] + [c_int64] * 4 + [c_void_p]),
    "wm_skff_apply_fwd": (c_int, [c_void_p] * 10 + [c_size_t] + [c_int64] * 4 + [c_void_p]),
    "wm_ps_down_fwd": (c_int, [c_void_p] * 4 + [c_int64] * 3 + [c_int, c_void_p]),
    "wm_dw3x3_fwd": (c_int, [c_void_p] * 4 + [c_int64] * 4 + [c_int, c_void_p]),
    "wm_train_workspace_bytes": (c_size_t, [c_int64] * 4),
    "wm_dw3x3_wgrad": (c_int, [c_void_p] * 5 + [c_size_t] + [c_int64] * 4 + [c_void_p]),
    "wm_layernorm2d_bwd": (c_int, [c_void_p] * 3 + [c_float] + [c_void_p] * 4 + [c_size_t] + [c_int64] * 4 + [c_void_p]),
    "wm_img_u8_to_f32_fwd": (c_int, [c_void_p] * 2 + [c_int64] * 5 + [c_int, c_void_p]),
    "wm_img_f32_to_u8_fwd": (c_int, [c_void_p] * 2 + [c_int64] * 5 + [c_void_p]),
    "wm_psnr_ssim_y_workspace_bytes": (c_size_t, [c_int64] * 3 + [c_int]),
    "wm_psnr_ssim_y_u8": (c_int, [c_void_p] * 4 + [c_size_t] + [c_int64] * 3 + [c_int, c_void_p]),
}


class WaveMambaNativeError(RuntimeError):
    """Raised when the CUDA library is missing or one of its entry points fails."""


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WM_B200_LIB", LIB_PATH)   # developer override: an A/B build variant
    if not os.path.exists(path):
        raise WaveMambaNativeError(
            f"{path} is missing. Build it with `python -m wave_mamba_b200.build` "
            "(needs nvcc). wave_mamba_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise WaveMambaNativeError(f"{path} does not export {name}; rebuild it") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.wm_abi_version() != ABI_VERSION:
        raise WaveMambaNativeError(
            f"ABI mismatch: library {lib.wm_abi_version()} vs binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().wm_last_error()
        raise WaveMambaNativeError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown'}")
