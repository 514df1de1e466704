"""Selective-scan oracle front end -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the call the reference makes at basicsr/archs/wavemamba_arch.py:465-471
(``selective_scan_fn(xs, dts, As, Bs, Cs, Ds, z=None, delta_bias, delta_softplus=True)``)
with mamba_ssm's published ``selective_scan_ref`` semantics.  PARITY UNPINNED against the
mamba_ssm CUDA kernel (package absent); see selective_scan_ref.c for the recurrence.

Two implementations of the same recurrence:
  * ``selective_scan_loop``  -- pure torch, one Python step per position; small L only.
  * ``selective_scan_c``     -- the C/OpenMP streaming loop in selective_scan_ref.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwm_oracle_scan.so")
_lib = None


def build_c_oracle(force: bool = False) -> str:
    """Compile selective_scan_ref.c with the committed Makefile (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = ctypes.CDLL(_LIB_PATH)
        i64 = ctypes.c_int64
        for name in ("wm_oracle_selective_scan_f32", "wm_oracle_selective_scan_f64",
                     "wm_oracle_selective_scan_d64"):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p] * 8 + [i64] * 5
    return _lib


def set_threads(n: int) -> int:
    """Set the OpenMP thread count of the C scan (torchrun exports OMP_NUM_THREADS=1)."""
    lib = _load()
    lib.wm_oracle_set_threads.restype = ctypes.c_int
    lib.wm_oracle_set_threads.argtypes = [ctypes.c_int]
    return int(lib.wm_oracle_set_threads(int(n)))


def selective_scan_loop(u, delta, A, Bm, Cm, D=None, delta_bias=None):
    """Sequential recurrence in the dtype of ``u`` (fp32 or fp64).

    u, delta: (B, DIM, L); A: (DIM, N); Bm, Cm: (B, G, N, L); D, delta_bias: (DIM).
    """
    batch, dim, L = u.shape
    G, N = Bm.shape[1], Bm.shape[2]
    rep = dim // G
    if delta_bias is not None:
        delta = delta + delta_bias[None, :, None]
    delta = F.softplus(delta)  # threshold 20, as torch / the CUDA kernel
    Bfull = Bm.repeat_interleave(rep, dim=1)  # (B, DIM, N, L): group g -> channels g*rep..
    Cfull = Cm.repeat_interleave(rep, dim=1)
    h = torch.zeros(batch, dim, N, dtype=u.dtype)
    ys = []
    for l in range(L):
        dt = delta[:, :, l]
        decay = torch.exp(dt[:, :, None] * A[None])
        h = decay * h + (dt * u[:, :, l])[:, :, None] * Bfull[:, :, :, l]
        ys.append((h * Cfull[:, :, :, l]).sum(-1))
    y = torch.stack(ys, dim=-1)
    if D is not None:
        y = y + u * D[None, :, None]
    return y


def selective_scan_c(u, delta, A, Bm, Cm, D=None, delta_bias=None, arbiter64: bool = False):
    """C/OpenMP streaming scan.  dtype follows ``u`` (fp32 -> fp32 arithmetic, fp64 -> fp64).

    ``arbiter64=True`` with fp32 inputs evaluates the same recurrence in double and
    returns a float64 tensor (the arbiter named in SURVEY.md section 8c).
    """
    lib = _load()
    batch, dim, L = u.shape
    G, N = Bm.shape[1], Bm.shape[2]
    dt = u.dtype
    args = [t.contiguous().to(dt) if t is not None else None
            for t in (u, delta, A, Bm, Cm, D, delta_bias)]
    ptr = [ctypes.c_void_p(t.data_ptr()) if t is not None else None for t in args]
    if dt == torch.float32 and not arbiter64:
        out = torch.empty(batch, dim, L, dtype=torch.float32)
        fn = lib.wm_oracle_selective_scan_f32
    elif dt == torch.float32:
        out = torch.empty(batch, dim, L, dtype=torch.float64)
        fn = lib.wm_oracle_selective_scan_f64
    elif dt == torch.float64:
        out = torch.empty(batch, dim, L, dtype=torch.float64)
        fn = lib.wm_oracle_selective_scan_d64
    else:
        raise TypeError(f"oracle scan supports fp32/fp64, got {dt}")
    rc = fn(*ptr, ctypes.c_void_p(out.data_ptr()), batch, dim, G, N, L)
    if rc != 0:
        raise RuntimeError("oracle selective scan rejected its arguments")
    return out
