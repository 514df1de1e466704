"""Multi-GPU correctness of the configs[3] edge path: a batch of uint8 images held by rank 0 is scattered
over NCCL, enhanced on every rank, gathered back, and must equal rank 0 enhancing the whole batch alone.

    torchrun --standalone --nnodes=1 --nproc-per-node 2 tools/sharded_check.py [--batch 3 --size 128]"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=3)
    ap.add_argument("--size", type=int, default=128)
    args = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import wave_mamba_b200 as wm
    from tools.synth import f32_to_u8_bgr, synth_lowlight
    from wave_mamba_b200 import parallel

    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    if rank == 0:
        net.load_state_dict(torch.load(os.path.join(ROOT, "ckpt", "WaveMamba_UHDLL.pth"), map_location="cpu")["params"],
                            strict=True)
    net = net.to(dev).eval()
    parallel.broadcast_parameters(net, src=0)
    batch = out = None
    if rank == 0:
        x, _ = synth_lowlight(args.batch, args.size, args.size + 64, seed=9)
        batch = f32_to_u8_bgr(x).pin_memory()
        out = torch.empty_like(batch).pin_memory()
    res = parallel.sharded_enhance_u8(net, batch, dev, window=8, out=out)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        alone = wm.enhance_bgr_u8(net, batch, window=8).cpu()
        ok = torch.equal(res, alone)
        print(f"sharded_enhance_u8 over {dist.get_world_size()} ranks, batch {args.batch}: "
              f"{'identical to' if ok else 'DIFFERS from'} the single-GPU result")
    # the same edge path as a pipeline: three different batches through two staging slots
    shape = (args.batch, args.size, args.size + 64, 3)
    batches, outs = [None] * 3, [None] * 3
    if rank == 0:
        batches = [f32_to_u8_bgr(synth_lowlight(args.batch, args.size, args.size + 64, seed=20 + i)[0]).pin_memory()
                   for i in range(3)]
        outs = [torch.empty_like(b).pin_memory() for b in batches]
    pipe = parallel.ShardedEnhancePipeline(net, dev, window=8)
    for b, o in zip(batches, outs):
        pipe.submit(b, o, shape)
    pipe.flush()
    if rank == 0:
        ok2 = all(torch.equal(o, wm.enhance_bgr_u8(net, b, window=8).cpu()) for b, o in zip(batches, outs))
        print(f"ShardedEnhancePipeline, 3 batches: {'identical to' if ok2 else 'DIFFERS from'} the single-GPU results (pipelined)")
        ok = ok and ok2
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
