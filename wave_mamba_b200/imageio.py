"""The per-image body of the reference's inference loop with uint8 edges.

``inference_wavemamba.py:101-113`` does, per image: cv2 image (H,W,3 uint8 BGR) -> img2tensor ->
/255. -> reflect-pad to a multiple of 128 -> ``restoration_network`` -> crop -> tensor2img (clamp,
*255, round, uint8, RGB->BGR).  Here the uint8 image is what crosses PCIe (3 bytes per pixel each
way instead of 12) and the conversions run on the device (wm_img_u8_to_f32_fwd /
wm_img_f32_to_u8_fwd, bit-exact with img2tensor / tensor2img).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


@torch.no_grad()
def enhance_bgr_u8(net, img: torch.Tensor, window: int = 128, out: Optional[torch.Tensor] = None,
                   device: Optional[torch.device] = None, sync: bool = False,
                   cuda_division: bool = False) -> torch.Tensor:
    """img: (H,W,3) or (B,H,W,3) uint8 BGR, on the host (ideally pinned) or already on the GPU.
    Returns the enhanced uint8 BGR image(s) with the input's leading shape: on the GPU, or copied
    into ``out`` (a host uint8 tensor of the same shape, ideally pinned) when given.
    ``net`` is a ``WaveMamba`` (its ``restoration_network`` is used, as the reference does).

    ``cuda_division=True`` reproduces the reference script's GPU run bit for bit (its ``/ 255.`` on a
    CUDA tensor is a multiplication by the rounded reciprocal); the default is the IEEE division the
    CPU reference and the oracle compute.

    Stream contract: everything is enqueued on the current CUDA stream.  With a pinned ``img`` /
    ``out`` the two PCIe copies are asynchronous: do not overwrite ``img`` or read ``out`` before the
    stream has been synchronised -- pass ``sync=True`` for the reference loop's blocking behaviour
    (``tensor2img`` returns finished data)."""
    if img.dtype != torch.uint8 or img.dim() not in (3, 4) or img.shape[-1] != 3:
        raise ValueError(f"expected a (H,W,3) or (B,H,W,3) uint8 image, got {tuple(img.shape)} {img.dtype}")
    squeeze = img.dim() == 3
    if squeeze:
        img = img.unsqueeze(0)
    if not img.is_cuda:
        if device is None:
            device = next(net.parameters()).device
        img = img.to(device, non_blocking=True)
    img = img.contiguous()
    _, H, W, _ = img.shape
    fwd = getattr(net, "restoration_network", net)
    x = ops.img_u8_to_f32(img, window, cuda_division)                 # img2tensor, /255., check_image_size
    y = fwd(x)
    res = ops.img_f32_to_u8(y, H, W)                   # [:, :, :h, :w] + tensor2img
    if squeeze:
        res = res[0]
    if out is not None:
        out.copy_(res, non_blocking=True)
        res = out
    if sync:
        torch.cuda.current_stream(img.device).synchronize()
    return res


class EnhancePipeline:
    """The same per-image body for a stream of images (the reference loop walks a folder,
    inference_wavemamba.py:92-131): the upload of image i+1 and the download of result i-1 run on their own
    CUDA streams while image i is in the network -- the side-stream protocol of the reference's
    ``CUDAPrefetcher`` (basicsr/data/prefetch_dataloader.py:101-118) applied to both PCIe directions.

        pipe = EnhancePipeline(net, window=128)
        for img, out in zip(pinned_inputs, pinned_outputs):
            done = pipe.submit(img, out)          # returns at once; ``done`` is a CUDA event
        pipe.flush()                              # every ``out`` is complete

    ``img`` / ``out``: (H,W,3) uint8 BGR host tensors, pinned for the copies to be asynchronous.  ``img`` may
    be reused by the caller once the event returned ``depth`` submits later has fired (or after ``flush``);
    results are bit-identical to ``enhance_bgr_u8``.  Device staging buffers are kept per slot."""

    def __init__(self, net, window: int = 128, cuda_division: bool = False,
                 device: Optional[torch.device] = None, depth: int = 2):
        self.fwd = getattr(net, "restoration_network", net)
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise ops._cabi.WaveMambaNativeError("EnhancePipeline: the network must live on a CUDA device")
        self.window, self.cuda_division, self.depth = int(window), bool(cuda_division), max(int(depth), 2)
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._in = [None] * self.depth          # device uint8 staging, one per slot
        self._in_free = [None] * self.depth     # fired when the network no longer reads the staging buffer
        self._done = [None] * self.depth        # fired when the slot's result has reached the host
        self._n = 0

    @torch.no_grad()
    def submit(self, img: torch.Tensor, out: torch.Tensor) -> torch.cuda.Event:
        if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[-1] != 3 or img.is_cuda:
            raise ValueError(f"img: expected a (H,W,3) uint8 host tensor, got {tuple(img.shape)} {img.dtype} on {img.device}")
        if out.shape != img.shape or out.dtype != torch.uint8 or out.is_cuda:
            raise ValueError("out: expected a host uint8 tensor of the input's shape")
        slot = self._n % self.depth
        self._n += 1
        H, W, _ = img.shape
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            if self._in[slot] is None or self._in[slot].shape != (1, H, W, 3):
                self._in[slot] = torch.empty(1, H, W, 3, dtype=torch.uint8, device=self.device)
                self.h2d.wait_stream(cur)                       # the allocation belongs to the current stream
            with torch.cuda.stream(self.h2d):
                if self._in_free[slot] is not None:
                    self.h2d.wait_event(self._in_free[slot])    # the previous user of this slot has been converted
                self._in[slot][0].copy_(img, non_blocking=True)
                up = self.h2d.record_event()
            cur.wait_event(up)
            x = ops.img_u8_to_f32(self._in[slot], self.window, self.cuda_division)
            self._in_free[slot] = cur.record_event()
            res = ops.img_f32_to_u8(self.fwd(x), H, W)[0]
            ready = cur.record_event()
            res.record_stream(self.d2h)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(ready)
                out.copy_(res, non_blocking=True)
                self._done[slot] = self.d2h.record_event()
        return self._done[slot]

    def flush(self, block: bool = True) -> None:
        """Make the current stream wait for every pending download; ``block`` also waits on the host."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self._done:
            if ev is not None:
                cur.wait_event(ev)
        if block:
            for ev in self._done:
                if ev is not None:
                    ev.synchronize()
