#!/usr/bin/env python
"""Headline benchmark: attention fwd+bwd tokens/s (BASELINE.json metric) on the 3-D ViT-B configuration, plus the
sharded workloads of BASELINE.json (cfg5 ring, cfg4 window-sharded, cfg2 batch-sharded) in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3|cfg5|cfg4|cfg2] [--no-extras]

Headline (config.workload = "cfg3"): BASELINE.json configs[2] — ViT-B 3D encoder, 96^3 volume, patch 8 -> 1728 tokens,
12 heads x 64, bf16 — the configuration the metric's tensor-pipe target is quoted on. One "step" = one
global-attention layer (the hot path: fused QK^T -> softmax -> PV forward and its backward) over a batch of 16 synthetic
volumes per GPU = 27,648 layer-tokens per GPU per step. N > 1 shards the batch (independent samples, no data-path
collective): scaling "weak".

Printed JSON (one line, rank 0):
  value         whole-job tokens/s with inputs resident in HBM
  e2e           same metric with the step's inputs (qkv, dO) copied from pinned host memory and a result scalar read
                back inside the timed region; e2e_full_d2h: the same with o + dqkv (170 MB) copied back every step
  roofline      the backward launch group (the dominant kernel) against the measured bf16 tensor peak
  cpu_baseline  the reference's attention core timed on this box's host cores (bounded sample)
  parity_check  BEFORE anything is timed: the ring (sequence-parallel) attention over the real process group (NCCL at
                N > 1) and the window-sharded Swin attention, fwd + bwd, against the CPU oracle on rank 0; the run
                aborts non-zero when a max-rel error exceeds 2e-2
  workloads     cfg5 (262,144 tokens, ring K/V over the N ranks, strong scaling, with the 1-GPU time of the same total
                sequence measured in the same run), cfg4 (3-D Swin stage-1 block, windows sharded over the N ranks),
                cfg2 (2-D Swin stage-1 block, batch-sharded), each with its own roofline

`--impl reference` times the reference's own CPU attention (the unmodified SABlock class from baseline/_ref when
__graft_entry__.build() staged it, else the oracle port) on the same config — all 16 volumes per step — with all host
threads.
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

# A ring hop moves 100-300 MB inside a multi-millisecond attention step: a few NCCL point-to-point channels (= SMs taken
# from the persistent attention kernels) are plenty. Must be set before the NCCL communicator is created.
os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "4")
os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "1")

import torch  # noqa: E402

METRIC = "attention fwd+bwd tokens/s"
UNIT = "tokens/s"
CFG3 = dict(B=16, N=1728, H=12, d=64, C=768)
N_BUFFERS = 4  # rotating input sets (each step's qkv+dO = 170 MB > the 126 MB L2)
PARITY_TOL = 2e-2


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


# --------------------------------------------------------------------------------------------------
# CPU legs: the reference's attention core on the host cores
# --------------------------------------------------------------------------------------------------
def _cpu_attention_step_fn(batch):
    """Returns (step, kind, what): one fwd+bwd of the reference's ViT attention core (backbone_vit.py:191-201) over
    `batch` volumes of cfg3 on the CPU. kind "reference": the UNMODIFIED reference SABlock class (files staged under
    baseline/_ref by __graft_entry__.build(), imported through oracle/ref_shim.py) with its two Linear layers replaced by
    nn.Identity, so that exactly the attention core the GPU step computes is timed; kind "port": the oracle."""
    N, H, d, C = CFG3["N"], CFG3["H"], CFG3["d"], CFG3["C"]
    g = torch.Generator().manual_seed(0)
    try:
        from oracle import ref_shim

        if not ref_shim.reference_available():
            raise RuntimeError("reference files not staged")
        ref_vit, _ = ref_shim.load_reference()
        blk = ref_vit.SABlock(False, False, C, H)
        blk.qkv = torch.nn.Identity()
        blk.out_proj = torch.nn.Identity()
        x = torch.randn(batch, N, 3 * C, generator=g, requires_grad=True)
        d_o = torch.randn(batch, N, C, generator=g)

        def step():
            y = blk(x)
            torch.autograd.grad(y, [x], d_o)

        return step, "reference", "reference SABlock.forward attention core (backbone_vit.py:191-201; qkv/out_proj = Identity)"
    except Exception:  # noqa: BLE001 - the staged files are optional on the GPU box
        from oracle import attention_oracle as ao

        q, k, v = [torch.randn(batch, H, N, d, generator=g, requires_grad=True) for _ in range(3)]
        d_o = torch.randn(batch, H, N, d, generator=g)

        def step():
            o = ao.dense_attention(q, k, v, d ** -0.5)
            torch.autograd.grad(o, [q, k, v], d_o)

        return step, "port", "oracle port of backbone_vit.py:191-201"


def cpu_baseline_tokens_per_s(seconds_budget=12.0, sample_b=1):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, kind, what = _cpu_attention_step_fn(sample_b)
    step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < seconds_budget and len(times) < 30):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    best = min(times)
    sample = (f"{what}, fp32, {sample_b} of the {CFG3['B']} volumes x {CFG3['H']} heads x {CFG3['N']} tokens, "
              f"best of {len(times)}")
    return sample_b * CFG3["N"] / best, threads, kind, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B, N, H = CFG3["B"], CFG3["N"], CFG3["H"]
    step, kind, what = _cpu_attention_step_fn(B)      # every step = all 16 volumes of the GPU arm's step
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * B * N / dt
    sample = f"{what}; all {B} volumes per step (fp32, {H} heads x {N} tokens), {threads} host threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: ViT-B 3D encoder attention layer, 96^3 volume, patch 8 -> 1728 tokens, "
                                   "12 heads x 64, fwd+bwd", "volumes_per_gpu": B, "tokens_per_step": B * N,
                       "note": "CPU arm: rank 0 only, one box's host cores regardless of --gpus"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _max_rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _barrier(dist):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def _timed(dist, dev, fn, steps):
    """K steps bracketed by barrier + synchronize on both sides; device time, max over ranks. Returns ms total."""
    _barrier(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    _barrier(dist)
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# --------------------------------------------------------------------------------------------------
# parity check over the real process group (runs before anything is timed)
# --------------------------------------------------------------------------------------------------
def parity_check(dist, rank, world, dev):
    """Ring attention (fwd + bwd through ring.ring_attention_qkv over the ranks' process group) at N = 256*P and 1024*P
    and window-sharded Swin attention (window_parallel, a 3-D shifted block) against the CPU oracle. Every rank builds
    the same full inputs from one seed, takes its shard, and the shards' results are gathered to rank 0."""
    from long_context_biomedical_imaging_b200 import ring, window_parallel
    from oracle import attention_oracle as ao

    res = {"backend": "nccl" if world > 1 else "single process", "ranks": world, "tolerance": PARITY_TOL}

    def gather_seq(x):                       # (B, n_local, F) -> (B, N, F) on every rank
        if world == 1:
            return x
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous())
        return torch.cat(parts, dim=1)

    H, d = 2, 64
    comm = ring.RingComm() if dist is not None else None
    worst = 0.0
    ring_res = {}
    for n_local in (256, 1024):
        N = n_local * world
        g = torch.Generator().manual_seed(1000 + n_local)
        qkv_full = torch.randn(1, N, 3 * H * d, generator=g).to(torch.bfloat16)
        dout_full = torch.randn(1, N, H * d, generator=g).to(torch.bfloat16)
        sl = slice(rank * n_local, (rank + 1) * n_local)
        local = qkv_full[:, sl].to(dev).requires_grad_(True)
        if comm is not None:
            out = ring.ring_attention_qkv(local, H, comm)
        else:
            from long_context_biomedical_imaging_b200 import ops

            out = ops.dense_attention_qkv(local, H)
        out.backward(dout_full[:, sl].to(dev))
        o_all, g_all = gather_seq(out.detach()), gather_seq(local.grad)
        if rank == 0:
            full = qkv_full.float().requires_grad_(True)
            qf, kf, vf = ao.split_qkv_vit(full, H)
            ref = ao.dense_attention(qf, kf, vf, d ** -0.5).permute(0, 2, 1, 3).reshape(1, N, H * d)
            ref.backward(dout_full.float())
            e_o, e_g = _max_rel(o_all.float().cpu(), ref.detach()), _max_rel(g_all.float().cpu(), full.grad)
            ring_res[f"N={N}"] = {"o": e_o, "dqkv": e_g}
            worst = max(worst, e_o, e_g)
    res["ring"] = ring_res

    # window-sharded 3-D shifted block: 2 x 2 x 4 = 16 windows of 343 tokens (padded grid 14 x 14 x 28), d = 16
    grid, window, shift, heads, dh = (12, 13, 26), (7, 7, 7), (3, 3, 3), 2, 16
    C = heads * dh
    g = torch.Generator().manual_seed(77)
    qkv = (torch.randn(1, *grid, 3 * C, generator=g) * 0.7).to(torch.bfloat16)
    bias = torch.randn(3 * C, generator=g)
    table = torch.randn(13 * 13 * 13, heads, generator=g) * 0.5
    d_out = torch.randn(1, *grid, C, generator=g).to(torch.bfloat16)
    qkv_d = qkv.to(dev).requires_grad_(True)
    bias_d, table_d = bias.to(dev).requires_grad_(True), table.to(dev).requires_grad_(True)
    out = window_parallel.window_attention_sharded(qkv_d, bias_d, table_d, grid, window, shift, heads)
    out.backward(d_out.to(dev))
    if rank == 0:
        qr, br, tr = qkv.float().requires_grad_(True), bias.clone().requires_grad_(True), table.clone().requires_grad_(True)
        ref = ao.window_attention_core(qr, br, tr, grid, window, shift, heads)
        ref.backward(d_out.float())
        errs = {"o": _max_rel(out.detach().float().cpu(), ref.detach()), "dqkv": _max_rel(qkv_d.grad.float().cpu(), qr.grad),
                "dtable": _max_rel(table_d.grad.cpu(), tr.grad), "dbias": _max_rel(bias_d.grad.cpu(), br.grad)}
        res["window_sharded"] = errs
        worst = max(worst, *errs.values())
    res["max"] = worst
    ok = torch.tensor([1 if worst <= PARITY_TOL else 0], device=dev)
    if dist is not None:
        dist.broadcast(ok, 0)
    res["ok"] = bool(int(ok.item()))
    return res


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def workload_cfg5(dist, rank, world, dev, steps=2, warmup=1):
    """configs[4]: one global-attention layer over 262,144 tokens (B=1, 12 heads x 64), sequence-sharded over the
    ranks with ring K/V exchange (ring.py). Strong scaling: the total sequence is fixed; at N > 1 rank 0 also times
    the SAME total sequence on one GPU with the plain fused kernels (the other ranks wait), so the strong-scaling
    efficiency is a same-run, same-kernel number."""
    from long_context_biomedical_imaging_b200 import ops, ring

    N_total, H, d, C = 262144, 12, 64, 768
    scale = d ** -0.5
    n_local = N_total // world
    torch.manual_seed(rank)
    qkv = torch.randn(1, n_local, 3, H, d, device=dev).to(torch.bfloat16)
    d_o = torch.randn(1, n_local, H, d, device=dev).to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    peaks = measured_peaks()
    flops = 12.0 * N_total * N_total * C          # per step, whole job

    if world > 1:
        comm = ring.RingComm()

        def step(_i):
            o, lse = ring.ring_attention_forward(q, k, v, scale, comm)
            ring.ring_attention_backward(q, k, v, o, d_o, lse, scale, comm)
    else:
        dqkv = torch.empty_like(qkv)

        def step(_i):
            o, lse = ops.dense_attn_fwd(q, k, v, scale)
            ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])

    for i in range(warmup):
        step(i)
    ms = _timed(dist, dev, step, steps) / steps
    del qkv, d_o, q, k, v
    torch.cuda.empty_cache()

    n1_ms = ms if world == 1 else None
    if world > 1:
        if rank == 0:
            qkv1 = torch.randn(1, N_total, 3, H, d, device=dev).to(torch.bfloat16)
            do1 = torch.randn(1, N_total, H, d, device=dev).to(torch.bfloat16)
            dqkv1 = torch.empty_like(qkv1)
            q1, k1, v1 = qkv1[:, :, 0], qkv1[:, :, 1], qkv1[:, :, 2]

            def step1():
                o, lse = ops.dense_attn_fwd(q1, k1, v1, scale)
                ops.dense_attn_bwd(q1, k1, v1, o, do1, lse, scale, dq=dqkv1[:, :, 0], dk=dqkv1[:, :, 1], dv=dqkv1[:, :, 2])

            step1()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step1()
            e1.record()
            torch.cuda.synchronize()
            n1_ms = e0.elapsed_time(e1)
            del qkv1, do1, dqkv1, q1, k1, v1
            torch.cuda.empty_cache()
        _barrier(dist)
    tf_gpu = flops / (ms * 1e-3) / 1e12 / world
    out = {"workload": "cfg5: ViT-B 2D 1024^2 patch 2 -> 262,144 tokens, one global-attention layer fwd+bwd, B=1, 12 heads x 64",
           "parallelism": (f"sequence-sharded x{world}, ring K/V exchange over NCCL (carried softmax state forward, "
                           f"dK/dV sums added on arrival backward)") if world > 1 else "one GPU, plain fused kernels",
           "scaling": "strong", "steps": steps, "warmup": warmup, "tokens_per_gpu": n_local, "ms_per_step": ms,
           "tokens_per_s": N_total / (ms * 1e-3), "tflops_per_gpu": tf_gpu,
           "roofline": {"bound": "tensor", "achieved": tf_gpu, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": tf_gpu / peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst)",
                        "algorithmic_flops_per_step_per_gpu": flops / world},
           "gpu_launches_per_step_per_rank": 4 * world if world > 1 else 4}
    if n1_ms is not None:
        out["n1_ms_per_step_same_run"] = n1_ms
        out["strong_eff_vs_n1_ms"] = n1_ms / (world * ms)
    return out


# dram__bytes_read + dram__bytes_write per step of the window-attention kernels (ncu --set full, one capture per kernel):
# cfg2 (batch 16): forward 257.3 + 46.5 MB, backward 348.4 + 126.9 MB; cfg4: tcgen05 forward 89.6 + 10.8 MB, dK/dV
# 107.1 + 44.0 MB, dQ 107.1 + 18.5 MB. Algorithmic bytes: 604 MB / 302 MB.
SWIN_DRAM_BYTES_PER_STEP = {"cfg2": 779.1e6, "cfg4": 377.1e6}


def workload_swin(name, dist, rank, world, dev, steps=30, warmup=5):
    """configs[1] / configs[3]: the window attention of one stage-1 Swin block (SW-MSA, shift = window // 2), fwd+bwd.
    cfg2 (Swin-T 2D, 512^2, patch 4 -> 128^2 tokens x 96 channels, 3 heads x 32, window 7, batch 16 per GPU) shards
    the batch: no data-path collective, weak scaling. cfg4 (Swin 3D 'unetr', 128^3, patch 2 -> 64^3 tokens x 48
    channels, 3 heads x 16, window 7^3 = 343 tokens, batch 1) shards the 1000 windows over the ranks
    (window_parallel.py): strong scaling."""
    from long_context_biomedical_imaging_b200 import ops, window_parallel

    if name == "cfg2":
        B, grid, C, H, window = 16, (128, 128), 96, 3, (7, 7)
        what = "cfg2: Swin-T 2D 512^2 patch 4, stage-1 SW-MSA block attention, window 7, batch 16 per GPU"
        torch.manual_seed(rank)
    else:
        B, grid, C, H, window = 1, (64, 64, 64), 48, 3, (7, 7, 7)
        what = "cfg4: Swin 3D 'unetr' 128^3 patch 2, stage-1 SW-MSA block attention, window 7^3, batch 1"
        torch.manual_seed(0)                  # replicated tokens
    shift = tuple(w // 2 for w in window)
    n_tab = 1
    for w in window:
        n_tab *= 2 * w - 1
    qkv = torch.randn(B, *grid, 3 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
    bias = torch.randn(3 * C, device=dev).requires_grad_(True)
    table = (torch.randn(n_tab, H, device=dev) * 0.5).requires_grad_(True)
    d_out = torch.randn(B, *grid, C, device=dev).to(torch.bfloat16)
    sharded = name == "cfg4" and world > 1

    def step(_i):
        if sharded:
            out = window_parallel.window_attention_sharded(qkv, bias, table, grid, window, shift, H)
        else:
            out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
        out.backward(d_out)
        qkv.grad = bias.grad = table.grad = None

    for i in range(warmup):
        step(i)
    ms = _timed(dist, dev, step, steps) / steps
    n1_ms = None
    if sharded:                               # the un-sharded block on one GPU, same run (the other ranks wait)
        if rank == 0:
            def step1():
                out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
                out.backward(d_out)
                qkv.grad = bias.grad = table.grad = None

            for _ in range(3):
                step1()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step1()
            e1.record()
            torch.cuda.synchronize()
            n1_ms = e0.elapsed_time(e1) / 10
        _barrier(dist)
    peaks = measured_peaks()
    n_tok = B
    for gdim in grid:
        n_tok *= gdim
    weak = name == "cfg2"
    tokens_job = n_tok * (world if weak else 1)
    per_gpu_share = 1.0 if (weak or not sharded) else 1.0 / world
    alg_bytes = 24.0 * C * n_tok * per_gpu_share       # q,k,v,o read/written fwd + bwd, this GPU's share
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    # second floor: the exponentials. One exp per (query, key) pair and head in the forward, and one per recompute in the
    # backward (the generic large-window backward recomputes P in both of its kernels, the small-window one once), on the
    # MUFU pipe at 16 per clock and SM (sm_100a). At n = 343, d = 16 this floor is above the HBM one.
    n_win_tok = 1
    for w_ in window:
        n_win_tok *= w_
    props = torch.cuda.get_device_properties(dev)
    exps = float(n_tok) * n_win_tok * H * (3 if name == "cfg4" else 2) * per_gpu_share
    clock_khz = float(getattr(props, "clock_rate", 1965000) or 1965000)
    exp_floor_ms = exps / (props.multi_processor_count * 16.0 * clock_khz * 1e3) * 1e3
    out = {"workload": what,
           "parallelism": (f"batch-sharded x{world} (no data-path collective)" if weak else
                           (f"windows sharded x{world} (window_parallel.py)" if sharded else "one GPU")),
           "scaling": "weak" if weak else "strong", "steps": steps, "warmup": warmup, "ms_per_step": ms,
           "tokens_per_step": tokens_job, "tokens_per_s": tokens_job / (ms * 1e-3),
           "roofline": {"bound": "hbm", "kernel": "window attention fwd + bwd launch group (through the autograd wrapper"
                                                  + (", incl. the collectives" if sharded else "") + ")",
                        "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                        "traffic": None if sharded else SWIN_DRAM_BYTES_PER_STEP[name],
                        "traffic_unit": "bytes per step on one GPU (dram read + write of the fwd + bwd window kernels, one ncu "
                                        "--set full capture each: profiles/r02_ncu_cfg2_small_final_summary.csv, "
                                        "r02_ncu_cfg4_final_summary.csv; the dK/dV kernel after the heads-fastest CTA order)",
                        "algorithmic_bytes_per_step_per_gpu": alg_bytes,
                        "exp_pipe_floor": {"exps_per_step_per_gpu": exps, "floor_ms": exp_floor_ms,
                                           "frac": exp_floor_ms / ms,
                                           "note": "MUFU ex2 at 16 / clk / SM and the device's max SM clock; the "
                                                   "binding floor when it exceeds algorithmic_bytes / peak"}},
           "gpu_launches_per_step_per_rank": 4}
    if n1_ms is not None:
        out["n1_ms_per_step_same_run"] = n1_ms
        out["strong_eff_vs_n1_ms"] = n1_ms / (world * ms)
    return out


# dram__bytes_read.sum + dram__bytes_write.sum of dense_attn_bwd_kernel at 16 volumes (265.9 MB + 125.7 MB), per volume
BWD_DRAM_BYTES_PER_VOLUME = (265_883_904 + 125_711_872) // 16


def run_cfg3(args, dist, rank, world, local_rank, dev):
    from long_context_biomedical_imaging_b200 import ops

    B, N, H, d, C = args.batch, CFG3["N"], CFG3["H"], CFG3["d"], CFG3["C"]
    scale = d ** -0.5
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

    for i in range(args.warmup):
        step(i)
    _barrier(dist)

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()

    # ---------------- timed region 1: inputs resident in HBM
    bwd_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(dist)
    e0.record()
    for i in range(args.steps):
        step(i, bwd_events[i])
    e1.record()
    _barrier(dist)
    ms_total = e0.elapsed_time(e1)
    bwd_ms = statistics.mean(a.elapsed_time(b) for a, b in bwd_events)

    # ---------------- timed region 2 (e2e): pinned host inputs -> H2D -> fwd+bwd -> D2H, every step.
    # The copies of step i+1 run on a side stream into the other half of a double buffer while step i computes
    # (what a user-level input pipeline does); every byte still crosses PCIe inside the timed region.
    host_qkv = [t.cpu().pin_memory() for t in qkvs[:2]]
    host_do = [t.cpu().pin_memory() for t in d_os[:2]]
    dev_qkv = [torch.empty_like(qkvs[0]) for _ in range(2)]
    dev_do = [torch.empty_like(d_os[0]) for _ in range(2)]
    result_host = torch.empty(1, dtype=torch.float32).pin_memory()
    host_o = torch.empty(o.shape, dtype=o.dtype).pin_memory()
    host_dqkv = torch.empty(dqkv.shape, dtype=dqkv.dtype).pin_memory()
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

    def e2e_step(i, last, full_d2h):
        s = i % 2
        if not last:
            enqueue_copy(i + 1)
        torch.cuda.current_stream().wait_event(copied[s])
        q, k, v = dev_qkv[s][:, :, 0], dev_qkv[s][:, :, 1], dev_qkv[s][:, :, 2]
        _, lse = ops.dense_attn_fwd(q, k, v, scale, out=o)
        ops.dense_attn_bwd(q, k, v, o, dev_do[s], lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
        consumed[s].record()
        if full_d2h:
            host_o.copy_(o, non_blocking=True)
            host_dqkv.copy_(dqkv, non_blocking=True)
        else:
            result_host.copy_(dqkv.view(-1)[:1].float(), non_blocking=True)

    def e2e_region(full_d2h):
        for ev in consumed:
            ev.record()
        enqueue_copy(0)
        for i in range(3):
            e2e_step(i, False, full_d2h)
        _barrier(dist)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        # the copy for step 3 is already in flight from the warm-up; it is re-issued here so that every timed step's
        # bytes are copied inside the timed region
        enqueue_copy(3)
        for i in range(3, 3 + e2e_steps):
            e2e_step(i, i == 2 + e2e_steps, full_d2h)
        f1.record()
        _barrier(dist)
        return f0.elapsed_time(f1)

    e2e_ms = e2e_region(False)
    e2e_full_ms = e2e_region(True)
    clocks = sampler.stop() if rank == 0 else None

    if dist is not None:
        t = torch.tensor([ms_total, e2e_ms, bwd_ms, e2e_full_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, bwd_ms, e2e_full_ms = [float(x) for x in t.tolist()]
    del host_qkv, host_do, host_o, host_dqkv, qkvs, d_os, dev_qkv, dev_do
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks = measured_peaks()
    tokens_per_step = world * B * N
    value = tokens_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = tokens_per_step * e2e_steps / (e2e_ms * 1e-3)
    e2e_full_value = tokens_per_step * e2e_steps / (e2e_full_ms * 1e-3)
    alg_flops_step = 12.0 * B * N * N * C          # per GPU, fwd 4 + bwd 8
    bwd_flops = 8.0 * B * N * N * C
    bwd_tflops = bwd_flops / (bwd_ms * 1e-3) / 1e12
    step_tflops = alg_flops_step / (ms_total / args.steps * 1e-3) / 1e12
    h2d = int(B * N * 3 * H * d * 2 + B * N * H * d * 2)
    d2h_full = int(B * N * H * d * 2 + B * N * 3 * H * d * 2)
    return {
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
                     "frac_of_sustained": bwd_tflops / peaks["bf16_tflops_sustained"] if peaks["bf16_tflops_sustained"] else None,
                     "algorithmic_flops_per_launch": bwd_flops, "avg_launch_ms": bwd_ms},
        "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "overlap": "H2D of step i+1 on a side stream (double buffer) while step i computes",
                "h2d_gb_per_s_per_gpu": h2d * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                "bound": "PCIe / host memory: 170 MB of inputs per 0.7 ms step is ~240 GB/s of demand per GPU against a "
                         "~55 GB/s Gen5 x16 link; with N ranks on one host the links share the host's memory and root "
                         "complexes, so this number scales with the host, not with the kernels"},
        "e2e_full_d2h": {"value": e2e_full_value, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h_full, "what": "o (B,N,C) and dqkv (B,N,3C) copied back every step"},
        "gpu_launches": 4 * args.steps,
        "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=CFG3["B"], help="volumes per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip parity_check and the extra workloads")
    ap.add_argument("--workload", default="all", choices=["all", "cfg3", "cfg5", "cfg2", "cfg4"],
                    help="all (default): cfg3 headline + parity_check + cfg5 / cfg4 / cfg2 in `workloads`; one name: that "
                         "workload alone as the line's value")
    args = ap.parse_args()

    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 5
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    try:
        if args.workload in ("cfg5", "cfg2", "cfg4"):
            steps = args.steps if args.steps is not None else (2 if args.workload == "cfg5" else 50)
            warmup = max(1 if args.workload == "cfg5" else 3, args.warmup if args.warmup is not None else 1)
            w = (workload_cfg5(dist, rank, world, dev, steps, warmup) if args.workload == "cfg5" else
                 workload_swin(args.workload, dist, rank, world, dev, steps, warmup))
            if rank == 0:
                line = {"metric": METRIC, "value": w["tokens_per_s"], "unit": UNIT, "n_gpus": world, "steps": steps,
                        "warmup": warmup, "ms_per_step": w["ms_per_step"], "higher_is_better": True,
                        "scaling": w["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                        "config": {"workload": w["workload"], "parallelism": w["parallelism"]},
                        "roofline": w["roofline"], "gpu_launches": steps * w["gpu_launches_per_step_per_rank"]}
                for key in ("n1_ms_per_step_same_run", "strong_eff_vs_n1_ms", "tflops_per_gpu"):
                    if key in w:
                        line[key] = w[key]
                print(json.dumps(line), flush=True)
            return

        args.steps = args.steps if args.steps is not None else 300
        args.warmup = max(3, args.warmup if args.warmup is not None else 20)
        extras = args.workload == "all" and not args.no_extras

        parity = None
        if extras:
            parity = parity_check(dist, rank, world, dev)
            if not parity["ok"]:
                if rank == 0:
                    print(json.dumps({"parity_check": parity, "error": "parity check failed"}), flush=True)
                raise SystemExit(3)

        line = run_cfg3(args, dist, rank, world, local_rank, dev)

        workloads = {}
        if extras:
            workloads["cfg5"] = workload_cfg5(dist, rank, world, dev)
            workloads["cfg4"] = workload_swin("cfg4", dist, rank, world, dev)
            workloads["cfg2"] = workload_swin("cfg2", dist, rank, world, dev)

        if rank == 0:
            if parity is not None:
                line["parity_check"] = parity
            if workloads:
                line["workloads"] = workloads
            if not args.no_cpu_baseline:
                v, cores, kind, sample = cpu_baseline_tokens_per_s()
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
            print(json.dumps(line), flush=True)
    finally:
        if dist is not None:
            try:
                dist.barrier()
            except Exception:  # noqa: BLE001
                pass
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
