#!/usr/bin/env python
"""Benchmark of the Wave-Mamba forward hot path (BASELINE.json: 3840x2160 images/s, forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one forward of one synthetic 3840x2160 low-light image per GPU through the
reference-facing class (``WaveMamba.restoration_network``), UHD-LL weights.  N>1 is launched by
torchrun, one rank per GPU; images are independent, so ranks share nothing on the data path
(weak scaling, NCCL only for the one-off weight broadcast and the timing reduction).  A second
multi-GPU figure (``e2e.root_batch``, BASELINE configs[3]) times a batch of N uint8 images held by
rank 0: host -> rank 0 -> NCCL scatter -> forward on every rank -> NCCL gather -> host.

One JSON line on stdout (rank 0):
  value        images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same through the public API (the body of inference_wavemamba.py's loop): every step copies
               its uint8 BGR image from pinned host memory, converts / pads on the device, runs the
               forward, converts back and copies the uint8 result to pinned host memory.  `value` is
               wave_mamba_b200.EnhancePipeline (the copies of neighbouring images overlap the forward),
               `serial` the one-stream wave_mamba_b200.enhance_bgr_u8, `float32_edges` the serial form
               with the reference loop's 12-byte pixels
  roofline     the dominant hand-written kernel group (SS2D core): algorithmic bytes / measured
               duration vs the measured HBM peak (MEASURED_PEAKS.json), plus per-kernel rows
  cpu_baseline the CPU oracle (port of the reference forward) timed on this box's host cores: ONE
               forward of the same 3840x2160 image (about 20 s)
``--impl reference`` times the reference's CPU implementation of the path (the oracle port:
the reference has no compilable native sources and mamba_ssm is absent) on the SAME 3840x2160
workload, one image per step, real (not extrapolated) times.  One step is ~20 s on 16 host cores, so
the arm is time-boxed (--ref-budget-s): it runs one warm-up and as many of the K steps as fit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "3840x2160 images/sec fwd"
UNIT = "images/s"
H4K, W4K = 2160, 3840
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--ref-budget-s", type=float, default=270.0,
                    help="--impl reference: wall-clock budget; timed steps stop once it is spent")
    ap.add_argument("--ckpt", default="UHDLL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profiler-window", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop "
                         "(for `ncu --profile-from-start off`)")
    ap.add_argument("--no-clocks", action="store_true", help="do not spawn the nvidia-smi sampler "
                    "(use under ncu, which waits for child processes)")
    return ap.parse_args()


def load_params(name):
    return torch.load(os.path.join(ROOT, "ckpt", f"WaveMamba_{name}.pth"), map_location="cpu")["params"]


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference forward), bounded sample
# ----------------------------------------------------------------------------------------------
def workload_config(H, W, ckpt):
    """The ``config`` object is the same for both arms: it names the workload, not the engine."""
    return {"workload": f"UHD-LL {W}x{H} batch=1 per GPU forward (BASELINE configs[2])",
            "weights": f"WaveMamba_{ckpt}.pth", "input": "synth_lowlight(seed 1234 + rank), fp32 NCHW",
            "l2": "inputs and activations (>=1 GB per level-1 tensor) exceed the 126 MB L2"}


def cpu_forward_once(params, x):
    from oracle import model as om          # the CPU arm is the one place bench.py runs the oracle
    t0 = time.perf_counter()
    om.unet_forward(params, x)
    return time.perf_counter() - t0


def cpu_setup():
    from oracle import scan as oscan
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    oscan.set_threads(n)                    # torchrun exports OMP_NUM_THREADS=1
    return torch.get_num_threads()


def cpu_baseline(params, H, W):
    """One forward of the oracle on the full HxW image (no scaling, no extrapolation)."""
    from tools.synth import synth_lowlight
    cores = cpu_setup()
    x, _ = synth_lowlight(1, H, W, 1234)
    sec = cpu_forward_once(params, x)
    return {
        "value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"one {W}x{H} synthetic image, one forward ({sec:.2f} s), no warm-up; oracle = functional "
                  "torch-CPU restatement of the reference forward + C/OpenMP sequential selective scan",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools.synth import synth_lowlight
    params = load_params(args.ckpt)
    H, W = args.height, args.width
    cores = cpu_setup()
    x, _ = synth_lowlight(1, H, W, 1234)
    t_start = time.perf_counter()
    warm = min(args.warmup, 1)
    for _ in range(warm):
        cpu_forward_once(params, x)
    steps, total = 0, 0.0
    while steps < max(args.steps, 1):
        total += cpu_forward_once(params, x)
        steps += 1
        per = total / steps
        if time.perf_counter() - t_start + per > args.ref_budget_s:
            break
    sec = total / steps
    value = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "time_box_s": args.ref_budget_s,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(H, W, args.ckpt),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} timed forwards of one {W}x{H} image after {warm} warm-up "
                                   f"(time box {args.ref_budget_s:.0f} s); oracle port of the reference "
                                   "forward + C/OpenMP sequential selective scan, all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                # `timeout` bounds the sampler's life even if this process dies before stop()
                ["timeout", "600", "nvidia-smi", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# per-kernel CUDA-event instrumentation of wave_mamba_b200.ops
# ----------------------------------------------------------------------------------------------
class OpTimer:
    """Wraps the ops entry points with CUDA events on the current stream."""

    def __init__(self, ops):
        self.ops = ops
        self.records = []   # (name, bytes, start_event, end_event)
        self.enabled = False
        self._orig = {}
        self.conv_flops = 0.0

    @staticmethod
    def _numel_bytes(*tensors):
        return sum(t.numel() * t.element_size() for t in tensors if isinstance(t, torch.Tensor))

    def install(self):
        ops = self.ops

        def wrap(name, bytes_fn):
            orig = getattr(ops, name)
            self._orig[name] = orig

            def inner(*a, **k):
                if not self.enabled:
                    return orig(*a, **k)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = orig(*a, **k)
                e.record()
                self.records.append((name, bytes_fn(a, k, out), s, e))
                return out
            setattr(ops, name, inner)

        nb = self._numel_bytes
        timer = self

        def conv_bytes(a, k, o):
            w3 = a[1]
            cout, cin = w3.shape[0], w3.shape[1]
            px = o.shape[0] * o.shape[2] * o.shape[3]
            timer.conv_flops += 2.0 * px * cout * cin * (9 + (1 if k.get("gate_w") is not None else 0))
            return (cin + cout) * px * 4
        # algorithmic bytes: every input activation read once, every output written once
        wrap("dwt_haar", lambda a, k, o: nb(a[0]) + nb(*o))
        wrap("dwt_haar_pool", lambda a, k, o: nb(a[0]) + nb(*o[:4]))
        wrap("iwt_haar", lambda a, k, o: nb(a[0], a[1]) + nb(o))
        wrap("iwt_haar_cat", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("ss2d_core", lambda a, k, o: nb(a[0]) + nb(o))      # 512*B*L  
        wrap("layernorm2d", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("pw_dw", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("dw_act_pw", lambda a, k, o: nb(a[0]) + nb(o) + nb(k.get("residual"),
                                                                  a[6] if len(a) > 6 else None))
        wrap("pw", lambda a, k, o: nb(a[0]) + nb(o) + nb(k.get("residual")))
        wrap("paconv_gate", lambda a, k, o: nb(a[0]) + 2 * nb(o))
        wrap("conv3x3", conv_bytes)
        wrap("stem_conv3x3", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("head_conv3x3", lambda a, k, o: nb(a[0]) + nb(o) + nb(k.get("residual")))
        gram_bytes = lambda a, k, o: 2 * 32 * a[0].shape[0] * a[0].shape[2] * a[0].shape[3] * 4
        wrap("gram32", gram_bytes)
        wrap("match_index", gram_bytes)      # Gram pass + Matching argmin
        wrap("attn_mixed", gram_bytes)       # Gram pass + CMTAttention 32x32 tail
        wrap("lfss_z", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("lfss_out", lambda a, k, o: nb(a[0], a[1], a[6]) + nb(o) + nb(*k.get("extra", ())))
        wrap("lfss_tail", lambda a, k, o: nb(*a[0]) + nb(a[1]) + nb(o))   # 4 planes + x read, out written
        wrap("ss2d_dirs", lambda a, k, o: 2 * nb(a[0]))           # 512*B*L
        # pool pass reads 3 (inside the DWT epilogue when `pool` is given), apply reads 3 + writes 1
        wrap("skff", lambda a, k, o: (1 if k.get("pool") is not None else 2) * nb(a[0], a[1], a[2]) + nb(o))
        wrap("ps_down", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("img_u8_to_f32", lambda a, k, o: nb(a[0]) + nb(o))
        wrap("img_f32_to_u8", lambda a, k, o: nb(a[0]) + nb(o))

    def summary(self, peak_gbs):
        agg = {}
        for name, nbytes, s, e in self.records:
            ms = s.elapsed_time(e)
            d = agg.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0})
            d["calls"] += 1; d["ms"] += ms; d["bytes"] += nbytes
        rows = []
        for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
            row = {"kernel": name, "calls": d["calls"], "ms_total": round(d["ms"], 4),
                   "algorithmic_gb": round(d["bytes"] / 1e9, 4), "achieved_gbs": round(gbs, 1),
                   "frac_of_hbm_peak": round(gbs / peak_gbs, 4)}
            if name == "conv3x3" and d["ms"] > 0:
                row["bound"] = "tensor (tcgen05 kind::tf32, 3xTF32 => 3 MMAs per fp32-accurate MAC)"
                row["achieved_tflops_fp32_equiv"] = round(self.conv_flops / (d["ms"] * 1e-3) / 1e12, 2)
            rows.append(row)
        return rows


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the real stdout; everything else (NCCL banners, warnings) was
    redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # Keep stdout clean for the driver: libraries (e.g. NCCL's version banner) print to fd 1.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, (world, args.gpus)

    torch.backends.cudnn.benchmark = True          # as the reference drivers set it
    torch.backends.cudnn.allow_tf32 = False        # (no library conv / GEMM is on the path anyway)
    torch.backends.cuda.matmul.allow_tf32 = False

    import wave_mamba_b200 as wm
    from wave_mamba_b200 import ops, parallel
    from tools.synth import f32_to_u8_bgr, synth_lowlight

    H, W = args.height, args.width
    params = load_params(args.ckpt) if rank == 0 else None
    net = wm.WaveMamba(in_chn=3, wf=32, n_l_blocks=[1, 2, 4], n_h_blocks=[1, 1, 2], ffn_scale=2.0)
    if rank == 0:
        net.load_state_dict(params, strict=True)
    net = net.to(dev).eval()
    if world > 1:
        parallel.broadcast_parameters(net, src=0)     # NCCL, once, outside the timed region

    x_host, _ = synth_lowlight(1, H, W, seed=1234 + rank)      # each rank owns one image (shard)
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    y_host = torch.empty_like(x_host).pin_memory()
    # the same image as a cv2-style uint8 BGR array (what inference_wavemamba.py reads from disk)
    img_host = f32_to_u8_bgr(x_host)[0].contiguous().pin_memory()
    out_host = torch.empty_like(img_host).pin_memory()
    torch.cuda.synchronize()

    timer = OpTimer(ops)
    timer.install()
    peak_gbs, peak_src = measured_hbm_peak()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fwd = net.restoration_network
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            y = fwd(x_dev)
        # ---- device-resident timing ------------------------------------------------------
        barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0 and not args.no_clocks:
            clocks.start()
        launches0 = ops.launch_count
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.profiler_window:
            torch.cuda.profiler.start()
        s.record()
        for _ in range(args.steps):
            y = fwd(x_dev)
        e.record()
        barrier()
        if args.profiler_window:
            torch.cuda.profiler.stop()
        launches = ops.launch_count - launches0
        ms_dev = max_over_ranks(s.elapsed_time(e))
        # ---- per-kernel table: the same K steps again with CUDA events around every op (kept out
        # of the timed region above: ~900 event records per step cost about 1 %) -----------------
        fwd(x_dev)      # un-instrumented: the device is still busy when the host starts the instrumented steps
                        # (after a sync the first op's event pair would also span the host's launch latency)
        timer.enabled = True
        for _ in range(args.steps):
            y = fwd(x_dev)
        barrier()
        timer.enabled = False
        # ---- end to end through the public API (wave_mamba_b200.enhance_bgr_u8 = the body of the
        # reference's inference loop): pinned uint8 BGR image -> device -> u8->f32 -> forward ->
        # f32->u8 -> pinned host.  window=8: 2160x3840 needs no padding, same workload as `value`
        # (inference_wavemamba.py pads to multiples of 128, i.e. 2176 rows, +0.7 % pixels).
        for _ in range(2):
            wm.enhance_bgr_u8(net, img_host, window=8, out=out_host)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(args.steps):
            wm.enhance_bgr_u8(net, img_host, window=8, out=out_host)
        e2.record()
        barrier()
        ms_e2e_serial = max_over_ranks(s2.elapsed_time(e2))
        # the same K images through wave_mamba_b200.EnhancePipeline: identical work per step (one H2D of the
        # pinned uint8 image, forward, one D2H of the uint8 result), the copies of neighbouring images on
        # their own streams; the timed region ends when the last result has reached the host
        pipe = wm.EnhancePipeline(net, window=8)
        outs = [out_host, torch.empty_like(out_host).pin_memory()]
        for i in range(2):
            pipe.submit(img_host, outs[i % 2])
        pipe.flush()
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for i in range(args.steps):
            pipe.submit(img_host, outs[i % 2])
        pipe.flush(block=False)
        e2.record()
        barrier()
        ms_e2e = max_over_ranks(s2.elapsed_time(e2))
        if not torch.equal(outs[0], outs[1]) or not torch.equal(outs[0], wm.enhance_bgr_u8(net, img_host, window=8).cpu()):
            raise RuntimeError("EnhancePipeline result differs from enhance_bgr_u8")
        # the float32 edges of the reference loop, for comparison (12 bytes per pixel each way)
        for _ in range(2):
            y_host.copy_(fwd(x_host.to(dev, non_blocking=True)), non_blocking=True)
        barrier()
        s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(args.steps):
            xd = x_host.to(dev, non_blocking=True)
            y_host.copy_(fwd(xd), non_blocking=True)
        e3.record()
        barrier()
        ms_e2e_f32 = max_over_ranks(s3.elapsed_time(e3))
        # ---- BASELINE configs[3]: a batch of `world` uint8 images held by rank 0's host; scattered
        # and gathered over NCCL/NVLink INSIDE the timed region (SURVEY 8d "inputs resident on root
        # -> outputs gathered on root") ---------------------------------------------------------
        root_batch = None
        if dist is not None:
            batch_host = out_batch = None
            if rank == 0:
                batch_host = img_host.unsqueeze(0).repeat(world, 1, 1, 1).contiguous().pin_memory()
                out_batch = torch.empty_like(batch_host).pin_memory()
            stats = {}
            for _ in range(2):
                parallel.sharded_enhance_u8(net, batch_host, dev, window=8, out=out_batch)
            barrier()
            s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s4.record()
            for _ in range(args.steps):
                parallel.sharded_enhance_u8(net, batch_host, dev, window=8, out=out_batch, stats=stats)
            e4.record()
            barrier()
            ms_root = max_over_ranks(s4.elapsed_time(e4))
            # the same batches through parallel.ShardedEnhancePipeline: upload + scatter of batch i+1 and
            # gather + download of batch i-1 on side streams / their own NCCL communicators
            spipe = parallel.ShardedEnhancePipeline(net, dev, window=8)
            bshape = (world,) + tuple(img_host.shape)
            out_b2 = torch.empty_like(out_batch).pin_memory() if rank == 0 else None
            for i in range(2):
                spipe.submit(batch_host, out_batch if i % 2 == 0 else out_b2, bshape)
            spipe.flush()
            barrier()
            s5, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s5.record()
            for i in range(args.steps):
                spipe.submit(batch_host, out_batch if i % 2 == 0 else out_b2, bshape)
            spipe.flush(block=False)
            e5.record()
            barrier()
            ms_root_pipe = max_over_ranks(s5.elapsed_time(e5))
            if rank == 0 and not (torch.equal(out_batch, out_b2) and torch.equal(out_batch[0], out_host)):
                raise RuntimeError("ShardedEnhancePipeline result differs from enhance_bgr_u8")
            if rank == 0:
                torch.cuda.synchronize()
                coll = {k: round(sum(a.elapsed_time(b) for a, b in v) / args.steps, 4)
                        for k, v in stats.items()}
                root_batch = {
                    "value": world * args.steps / (ms_root * 1e-3), "unit": UNIT,
                    "ms_per_step": ms_root / args.steps, "images_per_step": world,
                    "h2d_bytes_per_step": batch_host.numel(), "d2h_bytes_per_step": out_batch.numel(),
                    "collectives": "ncclScatter of uint8 inputs + ncclGather of uint8 outputs "
                                   "(torch.distributed scatter/gather, NCCL over NVLink)",
                    "rank0_ms_per_step": coll,
                    "api": "wave_mamba_b200.parallel.sharded_enhance_u8(net, pinned uint8 batch on rank 0)",
                    "pipelined": {"value": world * args.steps / (ms_root_pipe * 1e-3), "unit": UNIT,
                                  "ms_per_step": ms_root_pipe / args.steps,
                                  "api": "wave_mamba_b200.parallel.ShardedEnhancePipeline: the same copies and "
                                         "collectives per step, those of neighbouring batches on side streams"}}
        clk = clocks.stop() if rank == 0 else None
    checksum = float(y_host.double().mean())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    rows = timer.summary(peak_gbs)
    ss = next((r for r in rows if r["kernel"] in ("ss2d_dirs", "ss2d_core")), None)
    L_total = sum((H // (2 ** l)) * (W // (2 ** l)) * n for l, n in ((1, 2), (2, 4), (3, 8)))
    roofline = None
    if ss:
        roofline = {
            "kernel": "ss2d_dirs (pass 1 + carry + replaying pass 2; the 4-way sum is fused into lfss_out)", "bound": "hbm",
            "achieved": ss["achieved_gbs"], "peak": peak_gbs, "unit": "GB/s",
            "frac": ss["frac_of_hbm_peak"],
            # ncu --set full, profiles/r2_ncu_ss2d_level1.txt: dram__bytes_read.sum + dram__bytes_write.sum
            # of pass 1 (2.269 + 5.663 GB) and the replaying pass 2 (5.633 + 2.105 GB) of ONE 4K level-1
            # call (1 x 64 x 1080 x 1920), whose algorithmic bytes are 1.062 GB
            "traffic": 15.670e9, "traffic_unit": "bytes per level-1 call (algorithmic: 1.062e9)",
            "traffic_note": "deliberate: pass 1 reads x once per direction and streams out the projected "
                            "tiles (43 KB per 64 positions), pass 2 reads them back instead of recomputing "
                            "the projection + softplus, and writes four direction planes that lfss_out sums; "
                            "bytes are spent on idle bandwidth to buy back instructions of a MUFU-bound kernel "
                            "(round 1: 6.6 GB traffic, 26.4 ms; now 15.7 GB, 21.9 ms)",
            "peak_source": peak_src,
            "bytes_definition": "512*B*L per call (x read once + merged y written once), SURVEY 8d",
            "ms_per_image": round(ss["ms_total"] / args.steps, 4),
            "scan_operand_bytes_frac": round(ss["frac_of_hbm_peak"] * 7.0, 4),
            "note": "MUFU-bound (one ex2 per state update, evaluated in both passes), not HBM-bound: the "
                    "binding roofline is `binding` (ncu: XU 67-69 % in pass 1, 76 % in pass 2); see DESIGN.md 4.2",
            "state_updates_per_s": round(L_total * 4096 * args.steps / (ss["ms_total"] * 1e-3), 1),
            # the binding resource: one MUFU ex2 per state update and per pass (two passes)
            # + 2 per (position, channel) for softplus; MUFU peak = 16 lanes/clk/SM
            "binding": {"resource": "MUFU (XU pipe)", "unit": "G exp/s",
                        "achieved": round(L_total * (4096 * 2 + 256 * 2 * 2) * args.steps
                                          / (ss["ms_total"] * 1e-3) / 1e9, 1),
                        "peak": round(16 * 148 * ((clk or {}).get("sm_mhz") or 1965.0) * 1e6 / 1e9, 1)},
        }
        b_ = roofline["binding"]
        b_["frac"] = round(b_["achieved"] / b_["peak"], 4) if b_["peak"] else None
    n_img = world * args.steps
    line = {
        "metric": METRIC, "value": n_img / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(H, W, args.ckpt),
        "parallelism": f"batch-sharded x{world}: one image per rank, no data-path collective in "
                       "`value`/`e2e`; `e2e.root_batch` adds the NCCL scatter/gather edges",
        "e2e": {"value": n_img / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": img_host.numel(), "d2h_bytes_per_step": out_host.numel(),
                "api": "wave_mamba_b200.EnhancePipeline(net, window=8).submit(pinned uint8 BGR image, pinned out) per "
                       "step, flush at the end: the copies of images i+1 / i-1 overlap the forward of image i",
                "serial": {"value": n_img / (ms_e2e_serial * 1e-3), "unit": UNIT,
                           "api": "wave_mamba_b200.enhance_bgr_u8(net, pinned uint8 BGR image, window=8, out=pinned): "
                                  "copy in, forward, copy out on one stream"},
                "float32_edges": {"value": n_img / (ms_e2e_f32 * 1e-3), "unit": UNIT,
                                  "h2d_bytes_per_step": x_host.numel() * 4,
                                  "d2h_bytes_per_step": y_host.numel() * 4},
                "root_batch": root_batch},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "kernels": rows,
        "output_mean": checksum,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(params, H, W)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
