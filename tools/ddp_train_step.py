"""BASELINE configs[4] smoke / timing: the reference's training step (basicsr/models/femasr_model.py:157-185:
zero_grad -> net_g(lq) -> L1 -> backward -> AdamW step) on N GPUs under DistributedDataParallel, the
wrap basicsr/models/base_model.py:111-114 applies.

    torchrun --standalone --nnodes=1 --nproc-per-node N tools/ddp_train_step.py [--batch 8 --size 512 --steps 3 --act-storage bf16]

fp32 compute; --act-storage bf16 keeps the saved feature maps in bf16.  Prints one JSON line on rank 0: step time (CUDA events,
max over ranks), images/s, and whether the ranks hold identical parameters after the steps (they start
from the same checkpoint and see different data, so equality proves the gradient all-reduce ran)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--act-storage", choices=["fp32", "bf16"], default="fp32")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import wave_mamba_b200 as wm
    from tools.synth import synth_lowlight

    params = torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_UHDLOL4K.pth"), map_location="cpu")["params"]
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    net.load_state_dict(params, strict=True)
    net = net.to(dev).train()
    net.restoration_network.activation_storage = args.act_storage
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.0, betas=(0.9, 0.99))
    x, gt = synth_lowlight(args.batch, args.size, args.size, seed=100 + rank)     # a different shard per rank
    x, gt = x.to(dev), gt.to(dev)

    def step():
        opt.zero_grad()
        loss = F.l1_loss(model(x), gt)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / args.steps], device=dev, dtype=torch.float64)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([int(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({
            "what": "training step (fwd + bwd + AdamW), fp32, DDP" if world > 1 else "training step, fp32, 1 GPU",
            "n_gpus": world, "batch_per_gpu": args.batch, "size": args.size, "steps": args.steps,
            "activation_storage": args.act_storage,
            "ms_per_step": float(ms.item()), "images_per_s": world * args.batch / (float(ms.item()) * 1e-3),
            "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "ranks_hold_identical_parameters": same}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
