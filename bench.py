#!/usr/bin/env python
"""Headline benchmark: attention fwd+bwd tokens/s (BASELINE.json metric) on the 3-D ViT-B configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload = "cfg3"): BASELINE.json configs[2] — ViT-B 3D encoder, 96^3 volume, patch 8 ->
1728 tokens, 12 heads x 64, bf16 — the configuration the metric's tensor-pipe target is quoted on. One "step" =
one global-attention layer (the hot path: fused QK^T -> softmax -> PV forward and its backward) over a batch of
16 synthetic volumes per GPU = 27,648 layer-tokens per GPU per step. N > 1 shards the batch (independent
samples, no data-path collective): scaling "weak".

Printed JSON (one line, rank 0): value = whole-job tokens/s with inputs resident in HBM; e2e = same metric with the
step's inputs (qkv, dO) copied from pinned host memory and a result scalar read back inside the timed region;
roofline = the backward launch group (the dominant kernel) against the measured bf16 tensor peak;
cpu_baseline = the oracle's CPU restatement of the same attention timed on this box's host cores.

`--impl reference` times the reference algorithm's CPU path (oracle port: the reference is pure PyTorch and its
tree does not exist on the GPU box) on the same config, with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "attention fwd+bwd tokens/s"
UNIT = "tokens/s"
CFG3 = dict(B=16, N=1728, H=12, d=64, C=768)
N_BUFFERS = 4  # rotating input sets (each step's qkv+dO = 170 MB > the 126 MB L2)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": float(d["bf16_tflops"]), "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 0)),
                "hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def oracle_cpu_tokens_per_s(seconds_budget=12.0, sample_b=1):
    """fwd+bwd of the oracle's dense attention (reference algorithm, backbone_vit.py:191-201) on host cores."""
    from oracle import attention_oracle as ao

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    N, H, d = CFG3["N"], CFG3["H"], CFG3["d"]
    g = torch.Generator().manual_seed(0)
    q, k, v = [torch.randn(sample_b, H, N, d, generator=g, requires_grad=True) for _ in range(3)]
    d_o = torch.randn(sample_b, H, N, d, generator=g)

    def step():
        o = ao.dense_attention(q, k, v, d ** -0.5)
        torch.autograd.grad(o, [q, k, v], d_o)

    step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < seconds_budget and len(times) < 30):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return sample_b * N / best, threads, f"oracle attention core fp32, B={sample_b} x {H} heads x {N} tokens, best of {len(times)}", times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" = one fwd+bwd of a 1-sample slice of the cfg3 batch on all host threads
    from oracle import attention_oracle as ao

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    N, H, d = CFG3["N"], CFG3["H"], CFG3["d"]
    g = torch.Generator().manual_seed(0)
    q, k, v = [torch.randn(1, H, N, d, generator=g, requires_grad=True) for _ in range(3)]
    d_o = torch.randn(1, H, N, d, generator=g)

    def step():
        o = ao.dense_attention(q, k, v, d ** -0.5)
        torch.autograd.grad(o, [q, k, v], d_o)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * N / dt
    sample = f"1 of the {CFG3['B']} volumes per step (fp32, {H} heads x {N} tokens), oracle port of backbone_vit.py:191-201"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: ViT-B 3D 96^3 patch 8, 1728 tokens, 12 heads x 64 (CPU sample: 1 volume/step)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_cfg5(args, dist, rank, world, local_rank):
    """configs[4]: one global-attention layer over 262,144 tokens (B=1, 12 heads x 64), sequence-sharded over the
    ranks with ring K/V exchange (ring.py). Strong scaling: the total sequence is fixed."""
    from long_context_biomedical_imaging_b200 import ring

    if dist is None:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    N_total, H, d, C = 262144, 12, 64, 768
    n_local = N_total // world
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(rank)
    qkv = torch.randn(1, n_local, 3, H, d, device=dev).to(torch.bfloat16)
    d_o = torch.randn(1, n_local, H, d, device=dev).to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    comm = ring.RingComm()
    steps = args.steps if args.steps is not None else 2
    warmup = max(1, args.warmup if args.warmup is not None else 1)

    def step():
        acc, lse = ring.ring_attention_forward(q, k, v, d ** -0.5, comm)
        o = acc.to(torch.bfloat16)
        ring.ring_attention_backward(q, k, v, o, d_o, lse, d ** -0.5, comm)

    for _ in range(warmup):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        peaks = measured_peaks()
        flops = 12.0 * N_total * N_total * C          # per step, whole job
        tf_per_gpu = flops * steps / (ms * 1e-3) / 1e12 / world
        print(json.dumps({
            "metric": METRIC, "value": N_total * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg5: ViT-B 2D 1024^2 patch 2 -> 262,144 tokens, one global-attention layer fwd+bwd",
                       "parallelism": f"sequence-sharded x{world}, ring K/V over NCCL", "tokens_per_gpu": n_local},
            "tflops_algorithmic_per_gpu": tf_per_gpu, "tensor_frac_of_measured_peak": tf_per_gpu / peaks["bf16_tflops"],
            "gpu_launches": steps * world * 6}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_swin(args, dist, rank, world, local_rank):
    """configs[1] / configs[3]: the window attention of one stage-1 Swin block (SW-MSA, shift = window // 2), fwd+bwd.
    cfg2 (Swin-T 2D, 512^2, patch 4 -> 128^2 tokens x 96 channels, 3 heads x 32, window 7, batch 16 per GPU) shards
    the batch: no data-path collective, weak scaling. cfg4 (Swin 3D 'unetr', 128^3, patch 2 -> 64^3 tokens x 48
    channels, 3 heads x 16, window 7^3 = 343 tokens, batch 1) shards the 1000 windows over the ranks
    (window_parallel.py: one all-reduce forward, gradient all-reduces backward): strong scaling."""
    from long_context_biomedical_imaging_b200 import ops, window_parallel

    if dist is None and args.workload == "cfg4":
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29534")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    if args.workload == "cfg2":
        B, grid, C, H, window, what = 16, (128, 128), 96, 3, (7, 7), "cfg2: Swin-T 2D 512^2 patch 4, stage-1 SW-MSA block attention, window 7, batch 16 per GPU"
        torch.manual_seed(rank)
    else:
        B, grid, C, H, window, what = 1, (64, 64, 64), 48, 3, (7, 7, 7), "cfg4: Swin 3D 'unetr' 128^3 patch 2, stage-1 SW-MSA block attention, window 7^3, batch 1"
        torch.manual_seed(0)                  # replicated tokens
    shift = tuple(w // 2 for w in window)
    n_tab = 1
    for w in window:
        n_tab *= 2 * w - 1
    qkv = torch.randn(B, *grid, 3 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
    bias = torch.randn(3 * C, device=dev).requires_grad_(True)
    table = (torch.randn(n_tab, H, device=dev) * 0.5).requires_grad_(True)
    d_out = torch.randn(B, *grid, C, device=dev).to(torch.bfloat16)
    steps = args.steps if args.steps is not None else 50
    warmup = max(3, args.warmup if args.warmup is not None else 5)

    def step():
        if args.workload == "cfg4":
            out = window_parallel.window_attention_sharded(qkv, bias, table, grid, window, shift, H)
        else:
            out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
        out.backward(d_out)
        qkv.grad = bias.grad = table.grad = None

    for _ in range(warmup):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        peaks = measured_peaks()
        n_tok = B
        for g in grid:
            n_tok *= g
        weak = args.workload == "cfg2"
        tokens_job = n_tok * (world if weak else 1)
        ms_step = ms / steps
        alg_bytes = 24.0 * C * n_tok                   # per GPU-sized problem: q,k,v,o read/written fwd + bwd
        gbs = alg_bytes * (1 if weak else 1.0 / world) / (ms_step * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": tokens_job / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": what,
                       "parallelism": (f"batch-sharded x{world} (no data-path collective)" if weak else
                                       f"windows sharded x{world} (all-reduce of outputs and of replicated-input gradients)"),
                       "tokens_per_step": tokens_job},
            "roofline": {"bound": "hbm", "kernel": "window attention fwd + bwd launch group", "achieved": gbs,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                         "algorithmic_bytes_per_step_per_gpu": alg_bytes * (1 if weak else 1.0 / world),
                         "note": "includes the autograd wrapper and, for cfg4, the collectives"},
            "gpu_launches": steps * 4}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of dense_attn_bwd_kernel at 16 volumes (265.9 MB + 125.7 MB), per volume
BWD_DRAM_BYTES_PER_VOLUME = (265_883_904 + 125_711_872) // 16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=CFG3["B"], help="volumes per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg5", "cfg2", "cfg4"],
                    help="cfg3 (default, headline); cfg5: ViT-B 2D 1024^2 patch 2 = 262,144 tokens, ring K/V "
                         "sequence-parallel over the ranks (strong scaling); cfg2 / cfg4: stage-1 Swin window attention, "
                         "batch-sharded (2-D) / window-sharded (3-D)")
    args = ap.parse_args()

    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 10
        args.warmup = args.warmup if args.warmup is not None else 3
        run_reference_arm(args)
        return
    if args.workload == "cfg3":
        args.steps = args.steps if args.steps is not None else 300
        args.warmup = max(3, args.warmup if args.warmup is not None else 20)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from long_context_biomedical_imaging_b200 import ops

    if args.workload == "cfg5":
        run_cfg5(args, dist, rank, world, local_rank)
        return
    if args.workload in ("cfg2", "cfg4"):
        run_swin(args, dist, rank, world, local_rank)
        return

    B, N, H, d, C = args.batch, CFG3["N"], CFG3["H"], CFG3["d"], CFG3["C"]
    scale = d ** -0.5
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(rank)
    qkvs = [torch.randn(B, N, 3, H, d, device=dev).to(torch.bfloat16) for _ in range(N_BUFFERS)]
    d_os = [torch.randn(B, N, H, d, device=dev).to(torch.bfloat16) for _ in range(N_BUFFERS)]
    o = torch.empty(B, N, H, d, device=dev, dtype=torch.bfloat16)
    dqkv = torch.empty_like(qkvs[0])

    def step(i, ev=None):
        qkv, d_o = qkvs[i % N_BUFFERS], d_os[i % N_BUFFERS]
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        _, lse = ops.dense_attn_fwd(q, k, v, scale, out=o)
        if ev is not None:
            ev[0].record()
        ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
        if ev is not None:
            ev[1].record()
        return lse

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()

    # ---------------- timed region 1: inputs resident in HBM
    bwd_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i, bwd_events[i])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    bwd_ms = statistics.mean(a.elapsed_time(b) for a, b in bwd_events)

    # ---------------- timed region 2 (e2e): pinned host inputs -> H2D -> fwd+bwd -> scalar D2H, every step.
    # The copies of step i+1 run on a side stream into the other half of a double buffer while step i computes
    # (what a user-level input pipeline does); every byte still crosses PCIe inside the timed region.
    host_qkv = [t.cpu().pin_memory() for t in qkvs[:2]]
    host_do = [t.cpu().pin_memory() for t in d_os[:2]]
    dev_qkv = [torch.empty_like(qkvs[0]) for _ in range(2)]
    dev_do = [torch.empty_like(d_os[0]) for _ in range(2)]
    result_host = torch.empty(1, dtype=torch.float32).pin_memory()
    e2e_steps = max(10, args.steps // 10)
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def enqueue_copy(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])          # the compute that last read this buffer has finished
            dev_qkv[s].copy_(host_qkv[s], non_blocking=True)
            dev_do[s].copy_(host_do[s], non_blocking=True)
            copied[s].record(copy_stream)

    def e2e_step(i, last):
        s = i % 2
        if not last:
            enqueue_copy(i + 1)
        torch.cuda.current_stream().wait_event(copied[s])
        q, k, v = dev_qkv[s][:, :, 0], dev_qkv[s][:, :, 1], dev_qkv[s][:, :, 2]
        _, lse = ops.dense_attn_fwd(q, k, v, scale, out=o)
        ops.dense_attn_bwd(q, k, v, o, dev_do[s], lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
        consumed[s].record()
        result_host.copy_(dqkv.view(-1)[:1].float(), non_blocking=True)

    for ev in consumed:
        ev.record()
    enqueue_copy(0)
    for i in range(3):
        e2e_step(i, False)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    # the copy for step 3 is already in flight from the warm-up; it is re-issued here so that every timed step's
    # bytes are copied inside the timed region
    enqueue_copy(3)
    for i in range(3, 3 + e2e_steps):
        e2e_step(i, i == 2 + e2e_steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    clocks = sampler.stop() if rank == 0 else None

    if dist is not None:
        t = torch.tensor([ms_total, e2e_ms, bwd_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, bwd_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        peaks = measured_peaks()
        tokens_per_step = world * B * N
        value = tokens_per_step * args.steps / (ms_total * 1e-3)
        e2e_value = tokens_per_step * e2e_steps / (e2e_ms * 1e-3)
        alg_flops_step = 12.0 * B * N * N * C          # per GPU, fwd 4 + bwd 8
        bwd_flops = 8.0 * B * N * N * C
        bwd_tflops = bwd_flops / (bwd_ms * 1e-3) / 1e12
        step_tflops = alg_flops_step / (ms_total / args.steps * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg3: ViT-B 3D encoder attention layer, 96^3 volume, patch 8 -> 1728 tokens, "
                                   "12 heads x 64, fwd+bwd", "volumes_per_gpu": B, "tokens_per_step": tokens_per_step,
                       "parallelism": f"batch-sharded x{world} (no data-path collective)",
                       "l2": f"inputs rotate over {N_BUFFERS} buffer sets of 170 MB (> 126 MB L2)"},
            "tflops_algorithmic_per_gpu": step_tflops,
            "tensor_frac_of_measured_peak": step_tflops / peaks["bf16_tflops"],
            "roofline": {"bound": "tensor", "kernel": "dense_attn_bwd launch group (prep + bwd_main + finish)",
                         "achieved": bwd_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": bwd_tflops / peaks["bf16_tflops"], "traffic": BWD_DRAM_BYTES_PER_VOLUME * B,
                         "traffic_unit": "bytes per launch (dram read + write of dense_attn_bwd_kernel, one ncu --set full "
                                         "capture at 16 volumes: profiles/r01_ncu_dense_final_summary.csv, scaled by volumes)",
                         "peak_source": peaks["source"] + " (burst)", "peak_sustained": peaks["bf16_tflops_sustained"],
                         "algorithmic_flops_per_launch": bwd_flops, "avg_launch_ms": bwd_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(dev_qkv[0].numel() * 2 + dev_do[0].numel() * 2), "d2h_bytes_per_step": 4,
                    "overlap": "H2D of step i+1 on a side stream (double buffer) while step i computes"},
            "gpu_launches": 4 * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            v, cores, sample, _ = oracle_cpu_tokens_per_s()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
