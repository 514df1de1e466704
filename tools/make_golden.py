"""Generate tests/golden/*.npz by running the UNMODIFIED reference arch file on CPU.

Run in the build container only (needs /root/reference):

    python tools/make_golden.py

Each fixture stores the exact inputs and the reference's outputs for one function of the
hot path, plus two small end-to-end forwards.  The selective
scan inside SS2D is the shim in tools/ref_shims.py (mamba_ssm is absent), everything else
is the reference's own code.  Fixtures are float32, compressed, a few hundred KB total.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CKPT = os.path.join(ROOT, "ckpt")


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
           for k, v in arrays.items()}
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  " +
          ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def synth_lowlight(bsz, H, W, seed):
    """Same generator as oracle.model.synth_lowlight; duplicated here so the
    golden script does not depend on the oracle it is pinning."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(bsz, 3, max(H // 16, 1), max(W // 16, 1), generator=g)
    base = F.interpolate(base, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    x = (base * 0.15 + 0.02 * torch.rand(bsz, 3, H, W, generator=g)).clamp(0, 1)
    return x.contiguous(), base.contiguous()


@torch.no_grad()
def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    arch = ref_shims.reference_arch()
    g = torch.Generator().manual_seed(20240717)

    # ---- rows 1-2: DWT / IWT ------------------------------------------------------
    x = torch.randn(2, 5, 6, 10, generator=g)
    ll, hl, lh, hh = arch.dwt_init(x)
    save("dwt", x=x, ll=ll, hl=hl, lh=lh, hh=hh)
    xi = torch.randn(2, 12, 3, 7, generator=g)
    save("iwt", x=xi, y=arch.iwt_init(xi))

    nets = {n: ref_shims.build_reference_model(os.path.join(CKPT, f"WaveMamba_{n}.pth"))
            for n in ("LOLv1", "UHDLL", "UHDLOL4K")}

    # ---- rows 3-6: SS2D.forward_core (+ the 4-way sum) -----------------------------
    ss = nets["UHDLOL4K"].restoration_network.down_group1.l_blk[0].self_attention
    for tag, shape in (("a", (1, 64, 6, 10)), ("b", (2, 64, 9, 5))):
        xc = torch.nn.functional.silu(0.5 * torch.randn(*shape, generator=g))
        y1, y2, y3, y4 = ss.forward_core(xc)
        y = (y1 + y2 + y3 + y4).view(shape[0], 64, shape[2], shape[3])
        save(f"ss2d_core_{tag}", x=xc, y=y, ckpt="UHDLOL4K", block="down_group1.l_blk.0")

    # ---- rows 7-8: SS2D.forward and LFSSBlock --------------------------------------
    blk = nets["UHDLL"].restoration_network.down_group2.l_blk[1]
    xin = 0.5 * torch.randn(1, 8, 12, 32, generator=g)
    save("ss2d_full", x=xin, y=blk.self_attention(xin), ckpt="UHDLL", block="down_group2.l_blk.1")
    xin = 0.5 * torch.randn(2, 6 * 10, 32, generator=g)
    save("lfss_block", x=xin, h=6, w=10, y=blk(xin, [6, 10]), ckpt="UHDLL",
         block="down_group2.l_blk.1")

    # ---- rows 9-10: HFEBlock (with the Matching argmin indices) --------------------
    hb = nets["UHDLL"].restoration_network.up_group3.h_blk[1]
    idx_log = []
    orig_nn = arch.neirest_neighbores

    def spy(input_maps, candidate_maps, distances, num_matches):
        idx_log.append(distances.topk(k=1, largest=False).indices.squeeze(-1).clone())
        return orig_nn(input_maps, candidate_maps, distances, num_matches)

    arch.neirest_neighbores = spy
    xh = 0.3 * torch.randn(2, 32, 8, 12, generator=g)
    per = 0.5 * torch.randn(2, 32, 8, 12, generator=g)
    yh = hb(xh, per)
    arch.neirest_neighbores = orig_nn
    save("hfe_block", x=xh, perception=per, y=yh, idx_attn=idx_log[0], idx_ffn=idx_log[1],
         ckpt="UHDLL", block="up_group3.h_blk.1")

    # ---- SKFF ("next" row 4, needed by the end-to-end path) ------------------------
    sk = nets["UHDLL"].restoration_network.down_group1.h_fusion
    feats = [torch.randn(2, 32, 5, 7, generator=g) for _ in range(3)]
    save("skff", a=feats[0], b=feats[1], c=feats[2], y=sk(feats), ckpt="UHDLL",
         block="down_group1.h_fusion")

    # ---- end to end ----------------------------------------------------------------
    torch.manual_seed(0)
    x64 = torch.rand(1, 3, 64, 64)
    y64 = nets["LOLv1"].restoration_network(x64)
    print("KAT 64x64 rand/LOLv1: mean %.7f std %.7f (SURVEY: 0.7356967 / 0.2477267)"
          % (y64.mean().item(), y64.std().item()))
    save("e2e_rand64_LOLv1", x=x64, y=y64, ckpt="LOLv1")

    xs, gt = synth_lowlight(2, 48, 80, seed=0)
    ys = nets["UHDLL"].restoration_network(xs)
    save("e2e_synth48x80_UHDLL", x=xs, gt=gt, y=ys, ckpt="UHDLL")

    xs, gt = synth_lowlight(1, 64, 64, seed=0)
    ys = nets["UHDLOL4K"](xs)   # through WaveMamba.forward (registry class)
    save("e2e_synth64_UHDLOL4K", x=xs, gt=gt, y=ys, ckpt="UHDLOL4K")


if __name__ == "__main__":
    main()
