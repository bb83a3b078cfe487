"""Window-sharded Swin attention over torch.distributed (SURVEY.md 8e, row "Swin windows").

The windows of one W-MSA / SW-MSA block are independent, so the flattened (batch, window) list is split evenly over
the ranks of a process group: every rank holds the full (replicated) qkv of the block, runs the fused kernel on its
window range only (`ops.window_attention(..., win_range=...)`, C ABI `lcbi_win_attn_{fwd,bwd}_range`) and the
per-rank outputs - disjoint token rows, zeros elsewhere - are summed with one all-reduce. Backward mirrors it: the
incoming gradient is replicated, every rank back-propagates through its windows, and the gradients of the replicated
inputs (qkv, qkv.bias, the relative-position table) are summed over the ranks. Consecutive blocks use different
(shifted) partitions, so tokens must be visible to every rank between blocks: this is the simple formulation of 8e
(a); per-GPU compute at cfg4 stage 1 is a few hundred microseconds, so the all-reduces (25 MB forward, 75 MB
backward) and not the kernels bound this configuration - bench.py --workload cfg4 reports it as measured.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def count_windows(grid, window):
    """Windows per image for a token grid and a CONSTRUCTOR window (reference get_window_size clamp,
    backbone_swin.py:200-224, then ceil division of the padded grid)."""
    total = 1
    for g, w in zip(grid, window):
        w_used = g if g <= w else w
        total *= -(-int(g) // int(w_used))
    return total


def shard_range(total, world, rank):
    """Contiguous, near-even split of `total` units; trailing ranks may get an empty range."""
    per = -(-total // world)
    begin = min(rank * per, total)
    return begin, min(per, total - begin)


class _ReplicatedInput(torch.autograd.Function):
    """Identity on a tensor every rank holds a copy of; its gradient is the sum of the ranks' partial gradients."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class _SumOutputs(torch.autograd.Function):
    """Sum of the ranks' partial outputs (disjoint rows); the incoming gradient is already replicated."""

    @staticmethod
    def forward(ctx, x, group):
        x = x.contiguous().clone()
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


def window_attention_sharded(qkv, qkv_bias, table, grid, window, shift, num_heads, scale=None, group=None, attn_fn=None):
    """Drop-in for `ops.window_attention` when the block's windows are split over the ranks of `group`.
    Every rank must pass the same (replicated) qkv / parameters and receives the same full output.
    `attn_fn` (default: the CUDA op) must accept `win_range=(begin, count)`; tests inject the CPU oracle."""
    if attn_fn is None:
        from . import ops
        attn_fn = ops.window_attention
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return attn_fn(qkv, qkv_bias, table, grid, window, shift, num_heads, scale)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = qkv.shape[0] * count_windows(grid, window)
    begin, count = shard_range(total, world, rank)
    qkv_r = _ReplicatedInput.apply(qkv, group)
    bias_r = _ReplicatedInput.apply(qkv_bias, group) if qkv_bias is not None else None
    table_r = _ReplicatedInput.apply(table, group)
    part = attn_fn(qkv_r, bias_r, table_r, grid, window, shift, num_heads, scale, win_range=(begin, count))
    return _SumOutputs.apply(part, group)
