"""What compute-sanitizer wraps (GPU box): inference forwards at ragged sizes, the uint8 edges and pipeline,
the device metrics, one training step.  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wave_mamba_b200 as wm  # noqa: E402
from tools.synth import synth_lowlight  # noqa: E402

dev = torch.device("cuda:0")
params = torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_LOLv1.pth"), map_location="cpu")["params"]
net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
net.load_state_dict(params, strict=True)
net = net.to(dev).eval()
with torch.no_grad():
    for shape in ((1, 136, 248), (2, 200, 304), (1, 400, 600)):
        x, _ = synth_lowlight(shape[0], shape[1], shape[2], seed=1)
        y = net(x.to(dev))
        print("forward", tuple(y.shape), float(y.mean()))
g = torch.Generator().manual_seed(0)
imgs = [(torch.rand(135, 241, 3, generator=g) * 80).to(torch.uint8).pin_memory() for _ in range(3)]
outs = [torch.empty_like(i).pin_memory() for i in imgs]
pipe = wm.EnhancePipeline(net, window=128)
for i, o in zip(imgs, outs):
    pipe.submit(i, o)
pipe.flush()
print("pipeline", [int(o.sum()) for o in outs])
print("metrics", wm.metrics.calculate_psnr_ssim(outs[0].numpy(), imgs[0].numpy()))
net.train()
x, gt = synth_lowlight(1, 64, 96, seed=2)
loss = F.l1_loss(net(x.to(dev)), gt.to(dev))
loss.backward()
torch.cuda.synchronize()
print("train step", float(loss.detach()), sum(int(p.grad is not None) for p in net.parameters()), "gradients")
